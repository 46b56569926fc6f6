"""Prints stage-by-stage numerical gaps between the CUDA path and the oracle (run on the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openhush_b200 import COMPUTE_BF16, COMPUTE_F16, COMPUTE_F32, default_full_params, lib, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "micro"
path = synth.model_path(arch)
om = O.OracleModel(path)
clip = synth.audio(1)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


for name, ctype, gelu, act in (("f32", COMPUTE_F32, O.GELU_TANH, O.ACT_NONE), ("f16", COMPUTE_F16, O.GELU_TANH_F16LUT, O.ACT_F16),
                               ("bf16", COMPUTE_BF16, O.GELU_TANH_F16LUT, O.ACT_BF16)):
    for gemm in (("tc", "simt") if ctype != COMPUTE_F32 else ("-",)):
        if gemm == "simt":
            os.environ["OH_GEMM"] = "simt"
        else:
            os.environ.pop("OH_GEMM", None)
        ost = O.OracleState(om, gelu=gelu, act_round=act)
        ctx = lib.Context(path, compute_type=ctype, max_batch=2)
        st = lib.State(ctx)
        mel, _ = ost.pcm_to_mel(clip)
        w = np.ascontiguousarray(mel[:, :3000])
        enc = st.encode(w[None])[0]
        ref = ost.encode(w)
        m = om
        prompt = np.array([m.sot, m.sot + 1, m.transcribe, 1541, 5163, m.beg + 40, m.beg + 40, 777], np.int32)
        out = [f"{name}/{gemm} backend={st.gemm_backend()} enc rel={rel(enc, ref):.3e} max={np.abs(enc - ref).max():.3e}"]
        for n in (1, 3, 8):
            lg = st.decode(prompt[None, :n], 0)[0]
            rl = ost.decode(prompt[:n], 0)
            out.append(f"logits[{n}] max|d|={np.abs(lg - rl).max():.3e} rel={rel(lg, rl):.3e} argmax {int(lg.argmax())}/{int(rl.argmax())}")
        print(" | ".join(out), flush=True)
        p = default_full_params()
        p.temperature_inc = 0.0
        for seed in (1, 2):
            c = synth.audio(seed)
            ost.full(c, p)
            st.full(c, p)
            a = [t for s in ost.segments() for t in s["tokens"]]
            b = [t for s in st.segments() for t in s["tokens"]]
            k = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
            print(f"   full seed {seed}: oracle {len(a)} tokens, gpu {len(b)} tokens, common prefix {k}", flush=True)
        st.close()
        ctx.close()
        ost.close()
os.environ.pop("OH_GEMM", None)
