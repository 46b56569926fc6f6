"""openhush_b200 — B200-native transcription hot path of OpenHush behind a whisper.h-shaped C ABI.

Import is light (no CUDA needed); `openhush_b200.lib()` loads libohwhisper.so and raises loudly if it is
missing — there is no CPU fallback on the product path.
"""
from ._abi import (COMPUTE_BF16, COMPUTE_F16, COMPUTE_F32, SAMPLING_BEAM_SEARCH, SAMPLING_GREEDY, ContextParams, FullParams,
                   PreprocessParams, Timings, default_full_params, default_preprocess_params)

__all__ = ["COMPUTE_BF16", "COMPUTE_F16", "COMPUTE_F32", "SAMPLING_BEAM_SEARCH", "SAMPLING_GREEDY", "ContextParams",
           "FullParams", "PreprocessParams", "Timings", "default_full_params", "default_preprocess_params"]
