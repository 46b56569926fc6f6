"""Frozen outputs of the oracle's level processing (oracle/orc_preprocess.cpp: the reference's normalize_rms / compress /
limit, /root/reference/src/input/audio.rs:108-239, in the worker's order, worker.rs:196-240) for the reference's default
configuration (config.rs:1129-1159) on two short clips — so that a drift of the restatement shows up without a GPU.
Run from the repo root: python tests/golden/make_oracle_levels_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openhush_b200 import _abi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

p = _abi.default_preprocess_params()
speech = synth.audio(11, 24000) * np.float32(0.3)
i = np.arange(16000, dtype=np.float32)
tone = (np.float32(1.4) * np.sin((np.float32(2.0 * np.pi) * np.float32(440.0) * i / np.float32(16000.0)).astype(np.float32))).astype(np.float32)
tone[8000:] *= np.float32(0.05)  # loud half (above the ceiling), then a quiet half
out = {}
for name, clip in (("speech", speech), ("tone", tone)):
    y, (before, after) = O.preprocess_audio(clip, p)
    out[name + "_in"] = clip
    out[name + "_out"] = y
    out[name + "_rms"] = np.array([before, after], np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_levels.npz"), **out)
print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
