"""Shared helpers of the GPU parity tests (the oracle is the checker, the CUDA path through the C ABI is the thing checked)."""
import json
import os

import numpy as np

from openhush_b200 import default_full_params


def flat(segs):
    return [t for s in segs for t in s["tokens"]], [x for s in segs for x in (s["t0"], s["t1"])]


def common_prefix(a, b):
    n = 0
    for x, y in zip(a, b):
        if x != y:
            break
        n += 1
    return n


def teacher_forced_agreement(st, om, ost, clip, prompt=None):
    """Greedy picks of the product's logits (A.6 rules + A.7 arg-max applied by the oracle's rule code) along the ORACLE's
    own token sequence for the first window of `clip`.  Returns (agreeing picks, picks, oracle tokens)."""
    p = default_full_params()
    p.temperature_inc = 0.0
    assert ost.full(clip, p) == 0
    mel, _ = ost.pcm_to_mel(clip)
    st.encode(np.ascontiguousarray(mel[None, :, :3000]))
    toks = ost.window_tokens(0)  # raw decoder tokens of the first window (segments drop the second timestamp of each pair)
    if prompt is None:
        prompt = [om.sot, om.sot + 1, om.transcribe] if om.n_vocab >= 51865 else [om.sot]
    agree = 0
    has_ts, seek_delta = False, 3000
    logits = st.decode(np.array([prompt], np.int32), 0)[0]
    for i, t in enumerate(toks):
        r = om.process_logits(logits, tokens_cur=toks[:i], has_ts=has_ts, seek_delta=seek_delta)
        agree += int(r["id"] == t)
        if t > om.beg:
            has_ts, seek_delta = True, 2 * (t - om.beg)
        if i + 1 < len(toks):
            logits = st.decode(np.array([[t]], np.int32), len(prompt) + i)[0]
    return agree, len(toks), toks


def record(name, **numbers):
    """Append measured parity numbers to gpurun_out/parity_numbers.jsonl (kept with the round's profiles)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_numbers.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **numbers)) + "\n")
    except OSError:
        pass
