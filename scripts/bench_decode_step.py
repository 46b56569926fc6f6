"""Micro-benchmark of one decoder step (graph replay through the stage API, every row active, no logit read-back) and its
breakdown by elimination: the step is re-timed with one kernel category not launched (OH_SKIP_CATS bit mask;
results are meaningless in those runs, only the time matters).  The skip switch exists only in the diagnostic twin of the
library (`make -C openhush_b200/csrc diag` -> libohwhisper_diag.so, -DOH_DIAG), which this script loads.
Usage: python scripts/bench_decode_step.py [arch] [rows] [n_past] [quick]"""
import os
import sys

_CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "openhush_b200", "csrc")
if not os.path.exists(os.path.join(_CSRC, "libohwhisper_diag.so")):
    raise SystemExit("build the diagnostic library first: make -C openhush_b200/csrc diag")
os.environ["OH_LIB_PATH"] = os.path.join(_CSRC, "libohwhisper_diag.so")

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from openhush_b200 import COMPUTE_F16, lib, synth  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n_past0 = int(sys.argv[3]) if len(sys.argv) > 3 else 100
path = synth.model_path(arch)
ctx = lib.Context(path, compute_type=COMPUTE_F16, max_batch=rows)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
mel = None
CATS = {"gemm_decoder": 3, "layernorm": 4, "self_attention": 6, "cross_attention": 7, "embed": 8, "logits": 9}
variants = [("all kernels", 0)] + [("without " + k, 1 << v) for k, v in CATS.items()] + [("without all six", sum(1 << v for v in CATS.values()))]
if len(sys.argv) > 4 and sys.argv[4] == "quick":
    variants = [v for v in variants if v[0] in ("all kernels", "without gemm_decoder", "without self_attention")]
base = None
for name, mask in variants:
    os.environ["OH_SKIP_CATS"] = str(mask)
    st = lib.State(ctx)
    st.set_stream(stream.cuda_stream)
    if mel is None:
        mel = st.pcm_to_mel([synth.audio(1)])[0][:, :3000]
    st.encode(np.repeat(mel[None], min(rows, 4), axis=0))
    toks = np.full((rows, 1), ctx.sot, dtype=np.int32)
    for i in range(max(n_past0 - 4, 0), n_past0):  # capture + warm the graph (cache contents do not matter for timing)
        st.decode_noread(toks, i)
    torch.cuda.synchronize()
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(n):
        st.decode_noread(toks, n_past0 + i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if base is None:
        base = ms
    print(f"{name:28s} {ms:8.3f} ms/step   delta {base - ms:7.3f} ms", flush=True)
    st.close()
os.environ.pop("OH_SKIP_CATS", None)
