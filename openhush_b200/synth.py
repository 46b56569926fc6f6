"""Deterministic synthetic inputs (model files + PCM clips) for tests and bench.py.

No checkpoint or dataset is reachable offline, so every model is random-init weights of the named Whisper
architecture written in the reference's model format (`ggml-<name>.bin`, the names of
/root/reference/src/engine/whisper.rs:71-79; container layout SURVEY.md Appendix A.1), and every clip is
synthetic 16 kHz mono f32 audio.  The generator itself is C (tools/ohgen.c, counter-based RNG); this module
is the thin ctypes front and the table of architectures.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_TOOLS = os.path.join(_HERE, "tools")
_LIB = None


class _HParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
        "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer", "n_mels", "ftype")]


class _Shaping(ctypes.Structure):
    _fields_ = [("emb_logit_std", ctypes.c_double), ("ts_gain", ctypes.c_double),
                ("eot_gain", ctypes.c_double), ("w_gain", ctypes.c_double),
                ("emb_norm", ctypes.c_double), ("qk_gain", ctypes.c_double), ("ts_slope", ctypes.c_double)]


def build_tools() -> str:
    so = os.path.join(_TOOLS, "libohgen.so")
    src = os.path.join(_TOOLS, "ohgen.c")
    if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-C", _TOOLS], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_tools())
        lib.ohgen_audio.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p]
        lib.ohgen_audio.restype = None
        lib.ohgen_mel_filters.argtypes = [ctypes.c_int, ctypes.c_void_p]
        lib.ohgen_mel_filters.restype = None
        lib.ohgen_write_model.argtypes = [ctypes.c_char_p, ctypes.POINTER(_HParams), ctypes.c_uint64, ctypes.POINTER(_Shaping)]
        lib.ohgen_write_model.restype = ctypes.c_int
        _LIB = lib
    return _LIB


@dataclass(frozen=True)
class Arch:
    """Whisper dims (d/h/L per size)."""
    name: str
    d: int
    heads: int
    n_enc: int
    n_dec: int
    n_mels: int
    n_vocab: int


ARCHS = {
    # "micro" is not a Whisper size: same structure, small enough for CPU tests (3 decoder layers so the
    # "distilled" heuristic of Appendix A.1 does not switch timestamps off)
    "micro": Arch("micro", 128, 2, 2, 3, 80, 51865),
    "micro-v3": Arch("micro-v3", 128, 2, 2, 3, 128, 51866),
    "tiny.en": Arch("tiny.en", 384, 6, 4, 4, 80, 51864),
    "tiny": Arch("tiny", 384, 6, 4, 4, 80, 51865),
    "base": Arch("base", 512, 8, 6, 6, 80, 51865),
    "small": Arch("small", 768, 12, 12, 12, 80, 51865),
    "medium": Arch("medium", 1024, 16, 24, 24, 80, 51865),
    "large-v3": Arch("large-v3", 1280, 20, 32, 32, 128, 51866),
}

# shaping defaults (see tools/ohgen.c header): chosen so that greedy decodes of random-init weights emit
# timestamp pairs, advance `seek`, and usually terminate without the temperature fallback
DEFAULT_SHAPING = dict(emb_logit_std=7.0, ts_gain=1.45, eot_gain=1.3, w_gain=1.0, emb_norm=2.0, qk_gain=2.0, ts_slope=0.3)


def write_model(path: str, arch: str | Arch, seed: int = 1234, ftype: int = 1, **shaping) -> str:
    a = ARCHS[arch] if isinstance(arch, str) else arch
    hp = _HParams(a.n_vocab, 1500, a.d, a.heads, a.n_enc, 448, a.d, a.heads, a.n_dec, a.n_mels, ftype)
    sh = dict(DEFAULT_SHAPING)
    sh.update(shaping)
    rc = _lib().ohgen_write_model(path.encode(), ctypes.byref(hp), seed, ctypes.byref(_Shaping(**sh)))
    if rc != 0:
        raise OSError(f"ohgen_write_model({path}) failed: {rc}")
    return path


def model_path(arch: str, cache_dir: str | None = None, seed: int = 1234, **shaping) -> str:
    """Return (generating on first use) `<cache>/ggml-<arch>.bin`."""
    cache_dir = cache_dir or os.environ.get("OH_MODEL_CACHE", "/tmp/openhush_b200_models")
    os.makedirs(cache_dir, exist_ok=True)
    tag = "" if not shaping and seed == 1234 else "-" + "-".join([str(seed)] + [f"{k}{v}" for k, v in sorted(shaping.items())])
    p = os.path.join(cache_dir, f"ggml-{arch}{tag}.bin")
    if not os.path.exists(p):
        tmp = p + f".tmp{os.getpid()}"
        write_model(tmp, arch, seed=seed, **shaping)
        os.replace(tmp, p)
    return p


def audio(seed: int, n_samples: int = 480000) -> np.ndarray:
    out = np.empty(n_samples, dtype=np.float32)
    _lib().ohgen_audio(seed, n_samples, out.ctypes.data)
    return out


def audio_batch(seeds, n_samples: int = 480000) -> np.ndarray:
    out = np.empty((len(seeds), n_samples), dtype=np.float32)
    lib = _lib()
    for i, s in enumerate(seeds):
        lib.ohgen_audio(int(s), n_samples, out[i].ctypes.data)
    return out


def mel_filters(n_mel: int) -> np.ndarray:
    out = np.empty((n_mel, 201), dtype=np.float32)
    _lib().ohgen_mel_filters(n_mel, out.ctypes.data)
    return out
