"""ctypes mirrors of the PODs in include/oh_whisper.h (kept field-for-field in header order)."""
import ctypes as C


class ContextParams(C.Structure):
    _fields_ = [("use_gpu", C.c_int), ("flash_attn", C.c_int), ("gpu_device", C.c_int),
                ("compute_type", C.c_int), ("max_batch", C.c_int)]


class FullParams(C.Structure):
    _fields_ = [
        ("strategy", C.c_int), ("n_threads", C.c_int), ("n_max_text_ctx", C.c_int),
        ("offset_ms", C.c_int), ("duration_ms", C.c_int),
        ("translate", C.c_int), ("no_context", C.c_int), ("no_timestamps", C.c_int), ("single_segment", C.c_int),
        ("print_special", C.c_int), ("print_progress", C.c_int), ("print_realtime", C.c_int), ("print_timestamps", C.c_int),
        ("max_tokens", C.c_int), ("suppress_blank", C.c_int), ("suppress_non_speech_tokens", C.c_int),
        ("temperature", C.c_float), ("max_initial_ts", C.c_float), ("length_penalty", C.c_float),
        ("temperature_inc", C.c_float), ("entropy_thold", C.c_float), ("logprob_thold", C.c_float), ("no_speech_thold", C.c_float),
        ("greedy_best_of", C.c_int), ("beam_size", C.c_int), ("beam_patience", C.c_float),
        ("language", C.c_char_p),
    ]


class Timings(C.Structure):
    _fields_ = [("mel_ms", C.c_float), ("encode_ms", C.c_float), ("decode_ms", C.c_float), ("total_ms", C.c_float),
                ("n_windows", C.c_int), ("n_decode_steps", C.c_int), ("n_tokens", C.c_int), ("n_fallbacks", C.c_int),
                ("n_kernel_launches", C.c_longlong), ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong),
                ("n_prompt_steps", C.c_longlong), ("n_row_steps", C.c_longlong), ("n_armed_row_steps", C.c_longlong)]


class PreprocessParams(C.Structure):
    _fields_ = [("normalize", C.c_int), ("target_db", C.c_float),
                ("compress", C.c_int), ("threshold_db", C.c_float), ("ratio", C.c_float), ("attack_ms", C.c_float),
                ("release_ms", C.c_float), ("makeup_gain_db", C.c_float),
                ("limit", C.c_int), ("ceiling_db", C.c_float), ("limiter_release_ms", C.c_float)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int), ("total_ms", C.c_float), ("work", C.c_double)]


COMPUTE_F32, COMPUTE_F16, COMPUTE_BF16 = 0, 1, 2
SAMPLING_GREEDY, SAMPLING_BEAM_SEARCH = 0, 1


def default_full_params(strategy: int = SAMPLING_GREEDY) -> FullParams:
    """SURVEY.md Appendix A.2 — the values whisper_full_default_params() yields for the reference's call
    (/root/reference/src/engine/whisper.rs:243-263).  The product's oh_full_default_params must agree;
    tests/test_abi.py checks that field by field."""
    p = FullParams()
    p.strategy = strategy
    p.n_threads = 4
    p.n_max_text_ctx = 16384
    p.no_context = 1
    p.suppress_blank = 1
    p.temperature = 0.0
    p.max_initial_ts = 1.0
    p.length_penalty = -1.0
    p.temperature_inc = 0.2
    p.entropy_thold = 2.4
    p.logprob_thold = -1.0
    p.no_speech_thold = 0.6
    p.greedy_best_of = 1 if strategy == SAMPLING_GREEDY else -1
    p.beam_size = -1 if strategy == SAMPLING_GREEDY else 5
    p.beam_patience = -1.0
    p.language = b"en"
    return p


def default_preprocess_params() -> PreprocessParams:
    """The reference's NormalizationConfig / CompressionConfig / LimiterConfig defaults (/root/reference/src/config.rs:
    1129-1159 values, 1239-1269 `enabled: true` once `preprocessing` is on); oh_preprocess_default_params must agree."""
    return PreprocessParams(1, -18.0, 1, -24.0, 4.0, 5.0, 50.0, 6.0, 1, -1.0, 50.0)
