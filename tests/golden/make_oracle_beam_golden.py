"""Freezes the oracle's beam-search results (beam_size 5, temperature_inc = 0) on the `micro` model
into tests/golden/oracle_micro_beam.npz — same purpose and numerics as make_oracle_golden.py.  The reference never
selects beam search, so like everything else in the oracle these vectors pin the restatement, not the reference.
Run from the repo root:  python tests/golden/make_oracle_beam_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openhush_b200 import default_full_params, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

seeds = [1, 2, 3, 5]
n_samples = [480000, 480000, 123456, 800000]
m = O.OracleModel(synth.model_path("micro"))
st = O.OracleState(m, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
tokens, times, tok_off, seg_off = [], [], [0], [0]
for seed, n in zip(seeds, n_samples):
    p = default_full_params(1)
    p.temperature_inc = 0.0
    assert st.full(synth.audio(seed, n), p) == 0
    segs = st.segments()
    for s in segs:
        tokens += s["tokens"]
        times += [s["t0"], s["t1"]]
    tok_off.append(len(tokens))
    seg_off.append(seg_off[-1] + len(segs))
    print(seed, n, "segments", len(segs), "tokens", tok_off[-1] - tok_off[-2], st.stats())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_micro_beam.npz"), seeds=np.array(seeds), n_samples=np.array(n_samples),
                    beam_size=5, temperature_inc=0.0, tokens=np.array(tokens, np.int32), times=np.array(times, np.int64),
                    tok_off=np.array(tok_off), seg_off=np.array(seg_off))
