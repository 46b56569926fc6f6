"""ORACLE — ctypes front of liboracle.so (test infrastructure; see orc_common.h for provenance).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from openhush_b200._abi import FullParams, default_full_params

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GELU_TANH_F16LUT, GELU_TANH, GELU_ERF = 0, 1, 2
ACT_NONE, ACT_F16, ACT_BF16 = 0, 1, 2


def build() -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, ip, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float)
        sig = {
            "orc_last_error": (C.c_char_p, []),
            "orc_model_load": (vp, [C.c_char_p]),
            "orc_model_free": (None, [vp]),
            "orc_model_hparams": (None, [vp, vp]),
            "orc_model_special_tokens": (None, [vp, vp]),
            "orc_token_to_str": (C.c_char_p, [vp, C.c_int]),
            "orc_model_mel_filters": (vp, [vp]),
            "orc_model_gelu_table": (vp, [vp]),
            "orc_state_new": (vp, [vp]),
            "orc_state_free": (None, [vp]),
            "orc_state_set_options": (None, [vp, C.c_int, C.c_int, C.c_int]),
            "orc_pcm_to_mel": (C.c_int, [vp, vp, C.c_int]),
            "orc_mel_dims": (None, [vp, ip, ip, ip]),
            "orc_mel_data": (vp, [vp]),
            "orc_encode": (C.c_int, [vp, vp, vp, vp]),
            "orc_cross_k": (vp, [vp]),
            "orc_cross_v": (vp, [vp]),
            "orc_decode": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
            "orc_process_logits": (C.c_int, [vp, C.POINTER(FullParams), vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, vp, vp,
                                             ip, ip, fp, fp, fp, fp]),
            "orc_full": (C.c_int, [vp, C.POINTER(FullParams), vp, C.c_int]),
            "orc_full_n_segments": (C.c_int, [vp]),
            "orc_full_get_segment_text": (C.c_char_p, [vp, C.c_int]),
            "orc_full_get_segment_t0": (C.c_int64, [vp, C.c_int]),
            "orc_full_get_segment_t1": (C.c_int64, [vp, C.c_int]),
            "orc_full_n_tokens": (C.c_int, [vp, C.c_int]),
            "orc_full_get_token_id": (C.c_int, [vp, C.c_int, C.c_int]),
            "orc_full_get_token_tid": (C.c_int, [vp, C.c_int, C.c_int]),
            "orc_full_get_token_p": (C.c_float, [vp, C.c_int, C.c_int]),
            "orc_full_get_token_plog": (C.c_float, [vp, C.c_int, C.c_int]),
            "orc_full_lang_id": (C.c_int, [vp]),
            "orc_full_stats": (None, [vp, vp]),
            "orc_full_window_seek": (C.c_int, [vp, C.c_int]),
            "orc_full_window_temperature": (C.c_float, [vp, C.c_int]),
            "orc_lang_id": (C.c_int, [C.c_char_p]),
            "orc_full_window_n_tokens": (C.c_int, [vp, C.c_int]),
            "orc_full_window_token": (C.c_int, [vp, C.c_int, C.c_int]),
            "orc_model_round_weights_bf16": (None, [vp]),
            "orc_resample_sinc_table": (None, [C.c_double, vp]),
            "orc_mono_mix": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_longlong, vp]),
            "orc_resample_sinc": (C.c_longlong, [vp, C.c_longlong, C.c_int, C.c_int, vp, C.c_longlong]),
            "orc_min_duration_padding": (C.c_longlong, [C.c_longlong]),
            "orc_rms_db": (C.c_float, [vp, C.c_longlong]),
            "orc_apply_gain": (None, [vp, C.c_longlong, C.c_float]),
            "orc_normalize_rms": (None, [vp, C.c_longlong, C.c_float]),
            "orc_compress": (None, [vp, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
            "orc_limit": (None, [vp, C.c_longlong, C.c_float, C.c_float]),
            "orc_preprocess_audio": (None, [vp, C.c_longlong, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def load_pcm_16k(frames: np.ndarray, sample_rate: int, bits: int = 0, pad: bool = True) -> np.ndarray:
    """The reference's load_wav_file steps after the WAV reader (src/input/audio.rs:368-424): to f32, mono mix, sinc
    resampling to 16 kHz, 1.1 s padding.  frames: [n_frames, channels] f32 / int16 / int32 (`bits`-bit integers)."""
    L = lib()
    a = np.ascontiguousarray(frames)
    if a.ndim == 1:
        a = a[:, None]
    fmt = {np.dtype(np.float32): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2}[a.dtype]
    mono = np.empty(a.shape[0], dtype=np.float32)
    L.orc_mono_mix(a.ctypes.data, fmt, bits or (16 if fmt == 1 else 32), a.shape[1], a.shape[0], mono.ctypes.data)
    n = L.orc_resample_sinc(mono.ctypes.data, len(mono), sample_rate, 16000, None, 0)
    out = np.empty(max(int(n), 1), dtype=np.float32)
    L.orc_resample_sinc(mono.ctypes.data, len(mono), sample_rate, 16000, out.ctypes.data, len(out))
    out = out[:n]
    if pad:
        out = np.concatenate([out, np.zeros(int(L.orc_min_duration_padding(len(out))), np.float32)])
    return out


# ---- the worker's optional level processing (orc_preprocess.cpp; src/input/audio.rs:89-239, src/queue/worker.rs:196-240)
def rms_db(x) -> float:
    a = np.ascontiguousarray(x, dtype=np.float32)
    return float(lib().orc_rms_db(a.ctypes.data, len(a)))


def apply_gain(x, gain_db: float) -> np.ndarray:
    a = np.array(x, dtype=np.float32, copy=True)
    lib().orc_apply_gain(a.ctypes.data, len(a), gain_db)
    return a


def normalize_rms(x, target_db: float) -> np.ndarray:
    a = np.array(x, dtype=np.float32, copy=True)
    lib().orc_normalize_rms(a.ctypes.data, len(a), target_db)
    return a


def compress(x, threshold_db: float, ratio: float, attack_ms: float, release_ms: float, makeup_gain_db: float) -> np.ndarray:
    a = np.array(x, dtype=np.float32, copy=True)
    lib().orc_compress(a.ctypes.data, len(a), threshold_db, ratio, attack_ms, release_ms, makeup_gain_db)
    return a


def limit(x, ceiling_db: float, release_ms: float) -> np.ndarray:
    a = np.array(x, dtype=np.float32, copy=True)
    lib().orc_limit(a.ctypes.data, len(a), ceiling_db, release_ms)
    return a


def preprocess_audio(x, params):
    """params: any object with the fields of oh_preprocess_params.  Returns (processed clip, (rms_db before, after))."""
    a = np.array(x, dtype=np.float32, copy=True)
    flags = np.array([params.normalize, params.compress, params.limit], dtype=np.int32)
    v = np.array([params.target_db, params.threshold_db, params.ratio, params.attack_ms, params.release_ms, params.makeup_gain_db,
                  params.ceiling_db, params.limiter_release_ms], dtype=np.float32)
    rms = np.zeros(2, dtype=np.float32)
    lib().orc_preprocess_audio(a.ctypes.data, len(a), flags.ctypes.data, v.ctypes.data, rms.ctypes.data)
    return a, (float(rms[0]), float(rms[1]))


class OracleModel:
    def __init__(self, path: str, bf16_weights: bool = False):
        self._l = lib()
        self.h = self._l.orc_model_load(path.encode())
        if not self.h:
            raise RuntimeError("oracle: " + self._l.orc_last_error().decode())
        if bf16_weights:  # model of the product's OH_COMPUTE_BF16 mode (weights re-rounded at load)
            self._l.orc_model_round_weights_bf16(self.h)
        hp = np.zeros(11, dtype=np.int32)
        self._l.orc_model_hparams(self.h, hp.ctypes.data)
        (self.n_vocab, self.n_audio_ctx, self.n_audio_state, self.n_audio_head, self.n_audio_layer, self.n_text_ctx,
         self.n_text_state, self.n_text_head, self.n_text_layer, self.n_mels, self.ftype) = [int(x) for x in hp]
        st = np.zeros(10, dtype=np.int32)
        self._l.orc_model_special_tokens(self.h, st.ctypes.data)
        (self.eot, self.sot, self.translate, self.transcribe, self.solm, self.prev, self.nosp, self.not_, self.beg,
         self.blank) = [int(x) for x in st]

    def token_to_str(self, i: int) -> bytes:
        return self._l.orc_token_to_str(self.h, i)

    def mel_filters(self) -> np.ndarray:
        p = self._l.orc_model_mel_filters(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(self.n_mels, 201)).copy()

    def gelu_table(self) -> np.ndarray:
        p = self._l.orc_model_gelu_table(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint16)), shape=(65536,)).copy()

    def process_logits(self, logits: np.ndarray, tokens_cur=(), has_ts=False, seek_delta=3000, temperature=0.0, params=None):
        """A.6 + greedy A.7 on one logit vector.  Returns dict(logits, logprobs, probs, id, tid, p, plog, pt, ptsum)."""
        params = params or default_full_params()
        lg = np.ascontiguousarray(logits, dtype=np.float32).copy()
        lp = np.empty_like(lg)
        pr = np.empty_like(lg)
        tk = np.ascontiguousarray(tokens_cur, dtype=np.int32)
        i_id, i_tid = C.c_int(), C.c_int()
        f = [C.c_float() for _ in range(4)]
        self._l.orc_process_logits(self.h, C.byref(params), tk.ctypes.data, len(tk), int(has_ts), int(seek_delta), float(temperature),
                                   lg.ctypes.data, lp.ctypes.data, pr.ctypes.data, C.byref(i_id), C.byref(i_tid),
                                   *[C.byref(x) for x in f])
        return dict(logits=lg, logprobs=lp, probs=pr, id=i_id.value, tid=i_tid.value, p=f[0].value, plog=f[1].value,
                    pt=f[2].value, ptsum=f[3].value)

    def close(self):
        if self.h:
            self._l.orc_model_free(self.h)
            self.h = None


class OracleState:
    def __init__(self, model: OracleModel, gelu: int = GELU_TANH_F16LUT, act_round: int = ACT_F16, n_threads: int = 0):
        self._l = lib()
        self.model = model
        self.h = self._l.orc_state_new(model.h)
        self._l.orc_state_set_options(self.h, gelu, act_round, n_threads)

    def set_options(self, gelu: int, act_round: int, n_threads: int = 0):
        self._l.orc_state_set_options(self.h, gelu, act_round, n_threads)

    def pcm_to_mel(self, pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        rc = self._l.orc_pcm_to_mel(self.h, pcm.ctypes.data, len(pcm))
        if rc != 0:
            raise RuntimeError(f"orc_pcm_to_mel rc={rc}")
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._l.orc_mel_dims(self.h, C.byref(a), C.byref(b), C.byref(c))
        p = self._l.orc_mel_data(self.h)
        mel = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(a.value, b.value)).copy()
        return mel, c.value

    def encode(self, mel_window: np.ndarray, want_conv: bool = False):
        m = self.model
        w = np.ascontiguousarray(mel_window, dtype=np.float32)
        assert w.shape == (m.n_mels, 2 * m.n_audio_ctx)
        out = np.empty((m.n_audio_ctx, m.n_audio_state), dtype=np.float32)
        conv = np.empty_like(out) if want_conv else None
        self._l.orc_encode(self.h, w.ctypes.data, out.ctypes.data, conv.ctypes.data if want_conv else None)
        return (out, conv) if want_conv else out

    def cross_kv(self):
        m = self.model
        shp = (m.n_text_layer, m.n_audio_ctx, m.n_text_state)
        k = np.ctypeslib.as_array(C.cast(self._l.orc_cross_k(self.h), C.POINTER(C.c_float)), shape=shp).copy()
        v = np.ctypeslib.as_array(C.cast(self._l.orc_cross_v(self.h), C.POINTER(C.c_float)), shape=shp).copy()
        return k, v

    def decode(self, tokens, n_past: int) -> np.ndarray:
        tk = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(self.model.n_vocab, dtype=np.float32)
        rc = self._l.orc_decode(self.h, tk.ctypes.data, len(tk), n_past, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"orc_decode rc={rc}")
        return out

    def full(self, pcm: np.ndarray, params: FullParams | None = None) -> int:
        params = params or default_full_params()
        pcm = np.ascontiguousarray(pcm, dtype=np.float32)
        return self._l.orc_full(self.h, C.byref(params), pcm.ctypes.data, len(pcm))

    def segments(self):
        L, h = self._l, self.h
        out = []
        for i in range(L.orc_full_n_segments(h)):
            n = L.orc_full_n_tokens(h, i)
            out.append(dict(
                t0=L.orc_full_get_segment_t0(h, i), t1=L.orc_full_get_segment_t1(h, i), text=L.orc_full_get_segment_text(h, i),
                tokens=[L.orc_full_get_token_id(h, i, j) for j in range(n)],
                tids=[L.orc_full_get_token_tid(h, i, j) for j in range(n)],
                plogs=[L.orc_full_get_token_plog(h, i, j) for j in range(n)]))
        return out

    def stats(self):
        s = np.zeros(4, dtype=np.int32)
        self._l.orc_full_stats(self.h, s.ctypes.data)
        nw = int(s[3])
        return dict(n_encode=int(s[0]), n_decode_tokens=int(s[1]), n_fallbacks=int(s[2]),
                    seeks=[self._l.orc_full_window_seek(self.h, w) for w in range(nw)],
                    temps=[self._l.orc_full_window_temperature(self.h, w) for w in range(nw)])

    def window_tokens(self, w: int):
        return [self._l.orc_full_window_token(self.h, w, i) for i in range(self._l.orc_full_window_n_tokens(self.h, w))]

    def lang_id(self) -> int:
        return self._l.orc_full_lang_id(self.h)

    def close(self):
        if self.h:
            self._l.orc_state_free(self.h)
            self.h = None
