"""Times the encoder pass (conv stem + layers + cross-K/V) for a batch of windows through the stage API.
Usage: python scripts/bench_encoder.py [arch] [windows]      (wrap in `ncu --metrics gpu__time_duration.sum` for a per-kernel list)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from openhush_b200 import COMPUTE_F16, lib, synth  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ctx = lib.Context(synth.model_path(arch), compute_type=COMPUTE_F16, max_batch=nw)
st = lib.State(ctx)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
st.set_stream(stream.cuda_stream)
mel = st.pcm_to_mel([synth.audio(1)])[0][:, :3000]
d_mel = torch.from_numpy(np.repeat(mel[None], nw, axis=0).copy()).cuda()
for _ in range(2):
    assert st.encode_device(d_mel.data_ptr(), nw) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record(stream)
for _ in range(n):
    st.encode_device(d_mel.data_ptr(), nw)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
L, d, T = ctx.n_audio_layer, ctx.n_audio_state, 1500
flops = nw * (2 * 3000 * 3 * ctx.n_mels * d + 2 * T * 3 * d * d + L * (8 * T * d * d + 4 * T * T * d + 16 * T * d * d) + ctx.n_text_layer * 4 * T * d * d)
print(f"encoder pass {arch} x{nw} windows: {ms:.2f} ms  = {ms / nw:.3f} ms/window, {flops / ms / 1e9:.1f} TFLOP/s overall")
st.profile(True)
st.encode_device(d_mel.data_ptr(), nw)
for r in sorted(st.profile_report(), key=lambda r: -r["total_ms"]):
    unit = "TFLOP/s" if r["name"].startswith("gemm") or r["name"] == "attention_encoder" else "GB/s"
    scale = 1e9 if unit == "TFLOP/s" else 1e6
    print(f"  {r['name']:22s} {r['total_ms']:9.3f} ms  {r['launches']:4d} launches  {r['work'] / r['total_ms'] / scale:9.1f} {unit}")
st.profile(False)
