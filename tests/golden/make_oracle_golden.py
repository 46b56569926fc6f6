"""Freezes the oracle's own full() results on the `micro` model (token ids, segment times) into
tests/golden/oracle_micro_full.npz so that (a) later edits of the oracle cannot drift silently and
(b) the GPU parity tests have a committed vector besides the live comparison.
fp32-parity numerics (tanh GELU in f32, no activation rounding: IEEE-f32-clean, no f16 table whose rounding
boundaries would amplify summation-order noise into token flips), temperature_inc = 0 (greedy only: the
t>0 path draws from std::mt19937 whose stream depends on call history).
Run from the repo root:  python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openhush_b200 import default_full_params, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

seeds = [1, 2, 3, 4, 5, 6, 7]
n_samples = [480000, 480000, 480000, 480000, 123457, 17600, 800000]
m = O.OracleModel(synth.model_path("micro"))
st = O.OracleState(m, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
tokens, times, tok_off, seg_off = [], [], [0], [0]
for seed, n in zip(seeds, n_samples):
    p = default_full_params()
    p.temperature_inc = 0.0
    assert st.full(synth.audio(seed, n), p) == 0
    segs = st.segments()
    for s in segs:
        tokens += s["tokens"]
        times += [s["t0"], s["t1"]]
    tok_off.append(len(tokens))
    seg_off.append(seg_off[-1] + len(segs))
    print(seed, n, "segments", len(segs), "tokens", tok_off[-1] - tok_off[-2], st.stats())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_micro_full.npz"), seeds=np.array(seeds), n_samples=np.array(n_samples),
                    temperature_inc=0.0, tokens=np.array(tokens, np.int32), times=np.array(times, np.int64),
                    tok_off=np.array(tok_off), seg_off=np.array(seg_off))
