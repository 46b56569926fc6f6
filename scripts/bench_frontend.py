"""Times the front-end kernels with inputs resident in HBM: K1 log-mel (mel_frames_kernel) over a batch of 30 s clips and the
pre-engine resampler (resample_sinc_kernel) over one long stereo 48 kHz int16 recording.  Prints achieved algorithmic GB/s
against the measured HBM peak.  Wrap in `ncu --set full -k regex:"mel_frames|resample_sinc"` for the counters.
Usage: python scripts/bench_frontend.py [arch] [clips]"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from openhush_b200 import COMPUTE_F16, lib, synth  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 256
peak = 6536.7
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
ctx = lib.Context(synth.model_path(arch), compute_type=COMPUTE_F16, max_batch=1)
st = lib.State(ctx)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
st.set_stream(stream.cuda_stream)
N = 480000


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- K1: log-mel of `clips` 30 s clips (PCM read once + raw log-mel written once)
pcm = torch.from_numpy(np.stack([synth.audio(i + 1) for i in range(min(clips, 16))])).cuda()
pcm = pcm.repeat((clips + pcm.shape[0] - 1) // pcm.shape[0], 1)[:clips].contiguous()
ms = timed(lambda: st.pcm_to_mel_device(pcm.data_ptr(), N, clips))
bytes_mel = clips * (4.0 * N + 4.0 * ctx.n_mels * 3001)
print(f"mel_frames  {arch} ({ctx.n_mels} mel bins) x{clips} clips: {ms:.3f} ms  {bytes_mel / ms / 1e6:.1f} GB/s algorithmic = {bytes_mel / ms / 1e6 / peak:.3f} of measured HBM peak")

# ---- N2: stereo 48 kHz int16, 10 minutes -> mono 16 kHz f32
rate, secs, ch = 48000, 600, 2
rng = np.random.RandomState(0)
rec = torch.from_numpy((rng.randn(rate * secs, ch) * 3000).astype(np.int16)).cuda()
n_out = st.lib.oh_pcm_16k_len(rate, rate * secs)
out = torch.empty(n_out, dtype=torch.float32, device="cuda")
got = C.c_longlong(0)
call = lambda: st.lib.oh_load_pcm_16k_device(ctx.h, st.h, C.c_void_p(rec.data_ptr()), 1, 16, ch, rate, rate * secs, C.c_void_p(out.data_ptr()), n_out, C.byref(got))  # noqa: E731
ms_call = timed(call)
# the call includes the host side of the resampler (its index recurrence: one double add per output sample, which fixes the
# output length exactly); the kernel alone is timed by the library's per-kernel events
st.profile(True)
for _ in range(3):
    call()
rep = [r for r in st.profile_report() if r["name"] == "resample_sinc"]
st.profile(False)
ms = rep[0]["total_ms"] / rep[0]["launches"]
bytes_rs = rate * secs * ch * 2.0 + got.value * 4.0
flops = got.value * 2.0 * 2 * 256
print(f"resample_sinc  {secs} s stereo 48 kHz i16 -> {got.value} samples: kernel {ms:.3f} ms  {bytes_rs / ms / 1e6:.1f} GB/s algorithmic = {bytes_rs / ms / 1e6 / peak:.3f} of measured HBM peak, "
      f"{flops / ms / 1e9:.2f} TFLOP/s f32; whole call (host recurrence + kernel) {ms_call:.3f} ms = {secs / (ms_call / 1e3):.0f} audio-s/s")
st.close()
ctx.close()
