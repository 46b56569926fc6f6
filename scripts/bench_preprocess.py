"""Timing of the optional level processing (normalise / compress / limit, preprocess.cu) on device-resident clips.
Usage: python scripts/bench_preprocess.py [batch] [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from openhush_b200 import COMPUTE_F32, _abi, lib, synth  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
n = int(secs * 16000)
ctx = lib.Context(synth.model_path("micro"), compute_type=COMPUTE_F32, max_batch=1)
st = lib.State(ctx)
host = torch.from_numpy(np.stack([synth.audio(100 + i, n) for i in range(min(batch, 8))])).repeat((batch + 7) // 8, 1)[:batch].contiguous()
p = _abi.default_preprocess_params()
for b in sorted({1, min(8, batch), batch}):
    best = 1e9
    for _ in range(3):
        d = host[:b].cuda()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.preprocess_device(d.data_ptr(), n, [n] * b, p)  # returns after the stream has drained (reads back the levels)
        best = min(best, time.perf_counter() - t0)
    print(f"level processing, {b} clips x {secs:.0f} s: {best * 1e3:.2f} ms = {b * secs / best:,.0f} audio-s/s", flush=True)
st.close()
ctx.close()
