"""ctypes binding of libohwhisper.so (include/oh_whisper.h).  No torch types cross this boundary.

`load()` raises if the shared library has not been built or cannot be loaded — there is no CPU fallback; the
oracle under /oracle is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import _abi
from ._abi import ContextParams, FullParams, Timings

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(CSRC, "libohwhisper.so")
_LIB = None

# every symbol include/oh_whisper.h declares: name -> (restype, argtypes)
_vp, _ip, _fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
SYMBOLS = {
    "oh_context_default_params": (ContextParams, []),
    "oh_init_from_file": (_vp, [C.c_char_p, ContextParams]),
    "oh_init_state": (_vp, [_vp]),
    "oh_free_state": (None, [_vp]),
    "oh_free": (None, [_vp]),
    "oh_last_error": (C.c_char_p, []),
    "oh_full_default_params": (FullParams, [C.c_int]),
    "oh_full": (C.c_int, [_vp, _vp, FullParams, _vp, C.c_int]),
    "oh_full_batch": (C.c_int, [_vp, _vp, FullParams, _vp, _vp, C.c_int]),
    "oh_full_batch_device": (C.c_int, [_vp, _vp, FullParams, _vp, _vp, C.c_int]),
    "oh_stream_submit": (C.c_int, [_vp, _vp, FullParams, _vp, _vp, C.c_int, C.c_int, _ip]),
    "oh_stream_wait": (C.c_int, [_vp, _vp, C.c_int]),
    "oh_stream_pending": (C.c_int, [_vp]),
    "oh_full_item_rc": (C.c_int, [_vp, C.c_int]),
    "oh_full_n_segments": (C.c_int, [_vp]),
    "oh_full_get_segment_text": (C.c_char_p, [_vp, C.c_int]),
    "oh_full_get_segment_t0": (C.c_int64, [_vp, C.c_int]),
    "oh_full_get_segment_t1": (C.c_int64, [_vp, C.c_int]),
    "oh_full_n_tokens": (C.c_int, [_vp, C.c_int]),
    "oh_full_get_token_id": (C.c_int, [_vp, C.c_int, C.c_int]),
    "oh_full_get_token_p": (C.c_float, [_vp, C.c_int, C.c_int]),
    "oh_full_lang_id": (C.c_int, [_vp]),
    "oh_full_n_segments_item": (C.c_int, [_vp, C.c_int]),
    "oh_full_get_segment_text_item": (C.c_char_p, [_vp, C.c_int, C.c_int]),
    "oh_full_get_segment_t0_item": (C.c_int64, [_vp, C.c_int, C.c_int]),
    "oh_full_get_segment_t1_item": (C.c_int64, [_vp, C.c_int, C.c_int]),
    "oh_full_n_tokens_item": (C.c_int, [_vp, C.c_int, C.c_int]),
    "oh_full_get_token_id_item": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "oh_full_get_token_plog_item": (C.c_float, [_vp, C.c_int, C.c_int, C.c_int]),
    "oh_full_get_token_tid_item": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "oh_full_pack_records": (C.c_int64, [_vp, _vp, C.c_int64]),
    "oh_n_vocab": (C.c_int, [_vp]),
    "oh_n_audio_ctx": (C.c_int, [_vp]),
    "oh_n_audio_state": (C.c_int, [_vp]),
    "oh_n_audio_layer": (C.c_int, [_vp]),
    "oh_n_text_ctx": (C.c_int, [_vp]),
    "oh_n_text_state": (C.c_int, [_vp]),
    "oh_n_text_layer": (C.c_int, [_vp]),
    "oh_n_mels": (C.c_int, [_vp]),
    "oh_is_multilingual": (C.c_int, [_vp]),
    "oh_token_eot": (C.c_int, [_vp]),
    "oh_token_sot": (C.c_int, [_vp]),
    "oh_token_beg": (C.c_int, [_vp]),
    "oh_token_not": (C.c_int, [_vp]),
    "oh_token_to_str": (C.c_char_p, [_vp, C.c_int]),
    "oh_lang_id": (C.c_int, [C.c_char_p]),
    "oh_model_load_ms": (C.c_float, [_vp]),
    "oh_model_n_quantised": (C.c_int, [_vp]),
    "oh_pcm_16k_len": (C.c_longlong, [C.c_int, C.c_longlong]),
    "oh_load_pcm_16k": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, _vp, C.c_longlong, C.POINTER(C.c_longlong)]),
    "oh_load_pcm_16k_device": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, _vp, C.c_longlong, C.POINTER(C.c_longlong)]),
    "oh_chunk_rms_db": (C.c_int, [_vp, _vp, _vp, C.c_longlong, C.c_int, _vp, C.c_int]),
    "oh_preprocess_default_params": (_abi.PreprocessParams, []),
    "oh_preprocess_pcm": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(_abi.PreprocessParams), _vp]),
    "oh_preprocess_pcm_device": (C.c_int, [_vp, _vp, _vp, C.c_longlong, _vp, C.c_int, C.POINTER(_abi.PreprocessParams), _vp]),
    "oh_pcm_to_mel_batch": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "oh_pcm_to_mel_device": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "oh_encode_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "oh_encode_device": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "oh_decode_batch": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "oh_gemm_backend": (C.c_char_p, [_vp]),
    "oh_gemm_selftest_error": (C.c_float, []),
    "oh_profile": (C.c_int, [_vp, C.c_int]),
    "oh_profile_report": (C.c_int, [_vp, _vp, C.c_int]),
    "oh_get_timings": (Timings, [_vp]),
    "oh_state_set_stream": (C.c_int, [_vp, _vp]),
}


def build(force: bool = False) -> str:
    """Compile libohwhisper.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "oh_whisper.h"))
    stale = not os.path.exists(SO_PATH) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libohwhisper.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return SO_PATH


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        # OH_LIB_PATH: the diagnostic twin (`make -C openhush_b200/csrc diag`) for scripts/bench_decode_step.py; tests and bench.py never set it
        lib = C.CDLL(os.environ.get("OH_LIB_PATH") or SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # raises AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


class OhError(RuntimeError):
    pass


class Context:
    """oh_context: model weights on one GPU (mirrors whisper-rs WhisperContext, whisper.rs:159)."""

    def __init__(self, model_path: str, compute_type: int = 1, gpu_device: int = 0, max_batch: int = 1, use_gpu: bool = True):
        self.lib = load()
        p = self.lib.oh_context_default_params()
        p.use_gpu = 1 if use_gpu else 0
        p.gpu_device = gpu_device
        p.compute_type = compute_type
        p.max_batch = max_batch
        self.h = self.lib.oh_init_from_file(model_path.encode(), p)
        if not self.h:
            raise OhError("oh_init_from_file failed: " + self.lib.oh_last_error().decode(errors="replace"))
        L, h = self.lib, self.h
        self.n_vocab, self.n_mels = L.oh_n_vocab(h), L.oh_n_mels(h)
        self.n_audio_ctx, self.n_audio_state, self.n_audio_layer = L.oh_n_audio_ctx(h), L.oh_n_audio_state(h), L.oh_n_audio_layer(h)
        self.n_text_ctx, self.n_text_state, self.n_text_layer = L.oh_n_text_ctx(h), L.oh_n_text_state(h), L.oh_n_text_layer(h)
        self.eot, self.sot, self.beg, self.not_ = L.oh_token_eot(h), L.oh_token_sot(h), L.oh_token_beg(h), L.oh_token_not(h)
        self.max_batch = max_batch
        self.load_ms, self.n_quantised = L.oh_model_load_ms(h), L.oh_model_n_quantised(h)

    def token_to_str(self, i: int) -> bytes:
        return self.lib.oh_token_to_str(self.h, i)

    def close(self):
        if self.h:
            self.lib.oh_free(self.h)
            self.h = None


class State:
    """oh_state: buffers + results (mirrors whisper-rs WhisperState, whisper.rs:167)."""

    def __init__(self, ctx: Context):
        self.ctx, self.lib = ctx, ctx.lib
        self.h = self.lib.oh_init_state(ctx.h)
        if not self.h:
            raise OhError("oh_init_state failed: " + self.lib.oh_last_error().decode(errors="replace"))

    # ---- the hot call
    def full(self, pcm: np.ndarray, params: FullParams | None = None) -> int:
        params = params or self.lib.oh_full_default_params(0)
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self.lib.oh_full(self.ctx.h, self.h, params, pcm.ctypes.data, len(pcm))

    def full_batch(self, clips, params: FullParams | None = None) -> int:
        params = params or self.lib.oh_full_default_params(0)
        clips = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
        ptrs = (C.c_void_p * len(clips))(*[c.ctypes.data for c in clips])
        ns = (C.c_int * len(clips))(*[len(c) for c in clips])
        return self.lib.oh_full_batch(self.ctx.h, self.h, params, ptrs, ns, len(clips))

    def full_batch_ptrs(self, ptrs, n_samples, params: FullParams, on_device: bool) -> int:
        """ptrs: iterable of raw addresses (host or device), n_samples: iterable of ints."""
        ptrs = list(ptrs)
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        ns = (C.c_int * len(ptrs))(*[int(n) for n in n_samples])
        fn = self.lib.oh_full_batch_device if on_device else self.lib.oh_full_batch
        return fn(self.ctx.h, self.h, params, arr, ns, len(ptrs))

    # ---- continuous batching across calls
    def stream_submit(self, ptrs, n_samples, params: FullParams, on_device: bool = False) -> int:
        """Register jobs (raw host/device addresses of 16 kHz f32 PCM); returns the ticket."""
        ptrs = list(ptrs)
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        ns = (C.c_int * len(ptrs))(*[int(n) for n in n_samples])
        ticket = C.c_int(0)
        rc = self.lib.oh_stream_submit(self.ctx.h, self.h, params, arr, ns, len(ptrs), 1 if on_device else 0, C.byref(ticket))
        if rc != 0:
            raise OhError(f"oh_stream_submit rc={rc}: " + self.lib.oh_last_error().decode(errors="replace"))
        return ticket.value

    def stream_submit_clips(self, clips, params: FullParams) -> int:
        clips = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
        return self.stream_submit([c.ctypes.data for c in clips], [len(c) for c in clips], params, on_device=False)

    def stream_wait(self, ticket: int) -> int:
        return self.lib.oh_stream_wait(self.ctx.h, self.h, ticket)

    def item_rc(self, item: int = 0) -> int:
        return self.lib.oh_full_item_rc(self.h, item)

    def segments(self, item: int = 0):
        L, h = self.lib, self.h
        out = []
        for i in range(L.oh_full_n_segments_item(h, item)):
            n = L.oh_full_n_tokens_item(h, item, i)
            out.append(dict(
                t0=L.oh_full_get_segment_t0_item(h, item, i), t1=L.oh_full_get_segment_t1_item(h, item, i),
                text=L.oh_full_get_segment_text_item(h, item, i),
                tokens=[L.oh_full_get_token_id_item(h, item, i, j) for j in range(n)],
                tids=[L.oh_full_get_token_tid_item(h, item, i, j) for j in range(n)],
                plogs=[L.oh_full_get_token_plog_item(h, item, i, j) for j in range(n)]))
        return out

    def pack_records(self) -> np.ndarray:
        n = self.lib.oh_full_pack_records(self.h, None, 0)
        out = np.zeros(max(int(n), 1), dtype=np.int64)
        self.lib.oh_full_pack_records(self.h, out.ctypes.data, n)
        return out[:n]

    def timings(self) -> Timings:
        return self.lib.oh_get_timings(self.h)

    def profile(self, on: bool):
        return self.lib.oh_profile(self.h, 1 if on else 0)

    def profile_report(self):
        """[{name, launches, total_ms, work}] per kernel since profile(True)."""
        from ._abi import KernelStat
        arr = (KernelStat * 16)()
        n = self.lib.oh_profile_report(self.h, arr, 16)
        return [dict(name=arr[i].name.decode(), launches=arr[i].launches, total_ms=arr[i].total_ms, work=arr[i].work) for i in range(min(n, 16))]

    def gemm_backend(self) -> str:
        return self.lib.oh_gemm_backend(self.h).decode()

    def set_stream(self, cuda_stream: int | None):
        return self.lib.oh_state_set_stream(self.h, C.c_void_p(cuda_stream or 0))

    # ---- pre-engine DSP
    def load_pcm_16k(self, frames: np.ndarray, sample_rate: int, bits: int = 0) -> np.ndarray:
        """frames: [n_frames, channels] (or [n_frames]) f32 / int16 / int32 (then `bits` = the integer width, e.g. 24).
        Returns the 16 kHz mono f32 clip the reference's load_wav_file would hand to the engine."""
        a = np.ascontiguousarray(frames)
        if a.ndim == 1:
            a = a[:, None]
        fmt = {np.dtype(np.float32): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2}[a.dtype]
        n = self.lib.oh_pcm_16k_len(sample_rate, a.shape[0])
        out = np.empty(max(int(n), 1), dtype=np.float32)
        got = C.c_longlong(0)
        rc = self.lib.oh_load_pcm_16k(self.ctx.h, self.h, a.ctypes.data, fmt, bits or (16 if fmt == 1 else 32), a.shape[1], sample_rate, a.shape[0],
                                      out.ctypes.data, len(out), C.byref(got))
        if rc != 0:
            raise OhError(f"oh_load_pcm_16k rc={rc}: " + self.lib.oh_last_error().decode(errors="replace"))
        return out[:got.value]

    def preprocess(self, clips, params=None):
        """Level processing of 16 kHz f32 clips on the device (the worker's preprocess_audio hook: RMS normalisation,
        compression, limiter — each if enabled in `params`).  Returns (processed clips, [batch, 2] RMS dB before / after)."""
        params = params or _abi.default_preprocess_params()
        clips = [np.array(c, dtype=np.float32, copy=True) for c in clips]
        ptrs = (C.c_void_p * len(clips))(*[c.ctypes.data for c in clips])
        ns = (C.c_int * len(clips))(*[len(c) for c in clips])
        rms = np.zeros((len(clips), 2), dtype=np.float32)
        rc = self.lib.oh_preprocess_pcm(self.ctx.h, self.h, ptrs, ns, len(clips), C.byref(params), rms.ctypes.data)
        if rc != 0:
            raise OhError(f"oh_preprocess_pcm rc={rc}: " + self.lib.oh_last_error().decode(errors="replace"))
        return clips, rms

    def preprocess_device(self, d_pcm_ptr: int, stride: int, n_samples, params=None):
        """Same on a device-resident [batch][stride] f32 array (in place, on the state's stream)."""
        params = params or _abi.default_preprocess_params()
        ns = (C.c_int * len(n_samples))(*[int(n) for n in n_samples])
        rms = np.zeros((len(n_samples), 2), dtype=np.float32)
        rc = self.lib.oh_preprocess_pcm_device(self.ctx.h, self.h, C.c_void_p(d_pcm_ptr), stride, ns, len(n_samples), C.byref(params), rms.ctypes.data)
        if rc != 0:
            raise OhError(f"oh_preprocess_pcm_device rc={rc}: " + self.lib.oh_last_error().decode(errors="replace"))
        return rms

    # ---- stages
    def pcm_to_mel(self, clips):
        clips = [np.ascontiguousarray(c, dtype=np.float32) for c in clips]
        n_mel = self.ctx.n_mels
        outs = [np.empty((n_mel, (len(c) + 480000) // 160), dtype=np.float32) for c in clips]
        ptrs = (C.c_void_p * len(clips))(*[c.ctypes.data for c in clips])
        ns = (C.c_int * len(clips))(*[len(c) for c in clips])
        optrs = (C.c_void_p * len(clips))(*[o.ctypes.data for o in outs])
        nl = (C.c_int * len(clips))()
        rc = self.lib.oh_pcm_to_mel_batch(self.ctx.h, self.h, ptrs, ns, len(clips), optrs, nl)
        if rc != 0:
            raise OhError(f"oh_pcm_to_mel_batch rc={rc}: " + self.lib.oh_last_error().decode(errors="replace"))
        return outs

    def pcm_to_mel_device(self, d_pcm_ptr: int, n_samples: int, batch: int, d_out_ptr: int | None = None) -> int:
        return self.lib.oh_pcm_to_mel_device(self.ctx.h, self.h, C.c_void_p(d_pcm_ptr), n_samples, batch, C.c_void_p(d_out_ptr or 0))

    def encode(self, mel_windows: np.ndarray) -> np.ndarray:
        w = np.ascontiguousarray(mel_windows, dtype=np.float32)
        assert w.ndim == 3 and w.shape[1:] == (self.ctx.n_mels, 3000), w.shape
        out = np.empty((w.shape[0], self.ctx.n_audio_ctx, self.ctx.n_audio_state), dtype=np.float32)
        rc = self.lib.oh_encode_batch(self.ctx.h, self.h, w.ctypes.data, w.shape[0], out.ctypes.data)
        if rc != 0:
            raise OhError(f"oh_encode_batch rc={rc}: " + self.lib.oh_last_error().decode(errors="replace"))
        return out

    def encode_device(self, d_mel_ptr: int, batch: int) -> int:
        return self.lib.oh_encode_device(self.ctx.h, self.h, C.c_void_p(d_mel_ptr), batch)

    def decode(self, tokens: np.ndarray, n_past: int) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        if t.ndim == 1:
            t = t[None]
        out = np.empty((t.shape[0], self.ctx.n_vocab), dtype=np.float32)
        rc = self.lib.oh_decode_batch(self.ctx.h, self.h, t.ctypes.data, t.shape[1], n_past, t.shape[0], out.ctypes.data)
        if rc != 0:
            raise OhError(f"oh_decode_batch rc={rc}: " + self.lib.oh_last_error().decode(errors="replace"))
        return out

    def decode_noread(self, tokens: np.ndarray, n_past: int) -> int:
        """oh_decode_batch without the logit read-back (timing loops)."""
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        if t.ndim == 1:
            t = t[None]
        return self.lib.oh_decode_batch(self.ctx.h, self.h, t.ctypes.data, t.shape[1], n_past, t.shape[0], None)

    def close(self):
        if self.h:
            self.lib.oh_free_state(self.h)
            self.h = None
