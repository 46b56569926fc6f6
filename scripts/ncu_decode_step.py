"""A few decoder steps through the stage API with every row active (release library, graph replay) — the target of the
`ncu --set full` capture of the decode kernels.  Usage: python scripts/ncu_decode_step.py [arch] [rows] [n_past] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openhush_b200 import COMPUTE_F16, lib, synth  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n_past = int(sys.argv[3]) if len(sys.argv) > 3 else 100
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
ctx = lib.Context(synth.model_path(arch), compute_type=COMPUTE_F16, max_batch=rows)
st = lib.State(ctx)
mel = st.pcm_to_mel([synth.audio(1)])[0][:, :3000]
st.encode(np.repeat(mel[None], min(rows, 4), axis=0))
toks = np.full((rows, 1), ctx.sot, dtype=np.int32)
for i in range(steps):
    assert st.decode_noread(toks, n_past + i) == 0
print("done", flush=True)
st.close()
ctx.close()
