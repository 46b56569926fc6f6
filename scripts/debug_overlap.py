"""Debug: the 11 s tail window of the overlap-streaming test, product (f32) vs oracle, alone and inside the batch, for several max_batch."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openhush_b200 import COMPUTE_F32, default_full_params, lib, synth, engine as E
from oracle import oracle as O

path = synth.model_path("micro")
om = O.OracleModel(path)
rec = np.concatenate([synth.audio(60), synth.audio(61), synth.audio(62, 160000)])
wins = [np.ascontiguousarray(w) for _, w in E.cut_overlapping_windows(rec)]
p = default_full_params()
p.temperature_inc = 0.0
p.translate = 1
ost = O.OracleState(om, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
refs = []
for w in wins:
    assert ost.full(w, p) == 0
    refs.append([(s["t0"], s["t1"], s["tokens"]) for s in ost.segments()])
    print("oracle seeks", ost.stats()["seeks"], "n_seg", len(refs[-1]))
for mb in (4, 8, 16, 32):
    ctx = lib.Context(path, compute_type=COMPUTE_F32, max_batch=mb)
    st = lib.State(ctx)
    assert st.full_batch(wins, p) == 0
    got = [[(s["t0"], s["t1"], s["tokens"]) for s in st.segments(i)] for i in range(3)]
    ok = [got[i] == refs[i] for i in range(3)]
    print("max_batch", mb, "batch equal:", ok, "windows", st.timings().n_windows)
    for i in range(3):
        if not ok[i]:
            print("  item", i, "\n   ref", refs[i][:6], "\n   got", got[i][:6])
    alone = []
    for i in range(3):
        assert st.full(wins[i], p) == 0
        alone.append([(s["t0"], s["t1"], s["tokens"]) for s in st.segments(0)] == refs[i])
    print("   alone equal:", alone)
    st.close(); ctx.close()
