"""The continuous-batching stream API (oh_stream_submit / oh_stream_wait): jobs of several live tickets share the decoder
rows, and every job's result is exactly what the blocking call gives (items are independent reference calls,
/root/reference/src/queue/mod.rs:18-27)."""
import numpy as np
import pytest

from _parity import flat
from openhush_b200 import COMPUTE_F16, COMPUTE_F32, default_full_params, lib, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ctype", [COMPUTE_F32, COMPUTE_F16])
def test_stream_tickets_equal_blocking_calls(micro_path, ctype):
    ctx = lib.Context(micro_path, compute_type=ctype, max_batch=4)  # 4 rows: 14 jobs queue, rows are re-armed across tickets
    st = lib.State(ctx)
    p = default_full_params()
    p.temperature_inc = 0.0
    sets = [[synth.audio(10 + i) for i in range(5)],
            [synth.audio(20, 300000), synth.audio(21, 17600), synth.audio(22), synth.audio(23, 100), np.zeros(48000, np.float32)],
            [synth.audio(30 + i, 200000 + 37000 * i) for i in range(4)]]
    ref = []
    for clips in sets:
        assert st.full_batch(clips, p) == 0
        ref.append([st.segments(i) for i in range(len(clips))])
    # all three tickets live at once, waited in submission order
    tickets = [st.stream_submit_clips(clips, p) for clips in sets]
    assert st.lib.oh_stream_pending(st.h) == 3
    for k, tk in enumerate(tickets):
        assert st.stream_wait(tk) == 0
        assert [st.segments(i) for i in range(len(sets[k]))] == ref[k], k
        assert st.timings().n_kernel_launches >= 0
    assert st.lib.oh_stream_pending(st.h) == 0
    assert st.item_rc(3) == 0 and st.stream_wait(12345) != 0  # unknown ticket
    # out-of-order wait: the later ticket first (the earlier one keeps decoding meanwhile and is complete or nearly so after)
    t0 = st.stream_submit_clips(sets[0], p)
    t1 = st.stream_submit_clips(sets[1], p)
    assert st.stream_wait(t1) == 0
    assert [st.segments(i) for i in range(len(sets[1]))] == ref[1]
    assert st.item_rc(3) == -1  # the 100-sample clip: mel failed, no segments
    assert st.stream_wait(t0) == 0
    assert [st.segments(i) for i in range(len(sets[0]))] == ref[0]
    # the blocking call still works with nothing pending, and a pipelined sequence (submit k+1, wait k) too
    tk = st.stream_submit_clips(sets[0], p)
    for k in range(3):
        nxt = st.stream_submit_clips(sets[(k + 1) % 3], p) if k < 2 else None
        assert st.stream_wait(tk) == 0
        assert [st.segments(i) for i in range(len(sets[k % 3]))] == ref[k % 3]
        tk = nxt
    t = st.timings()
    assert t.n_row_steps >= t.n_armed_row_steps >= 0
    st.close()
    ctx.close()


def test_stream_with_temperature_fallback_matches_blocking(micro_path):
    """Reference defaults (temperature_inc 0.2): rows that fail the t = 0 checks are parked and resolved by the ladder while
    other tickets are live.  Streamed jobs seed their generator per job, so they equal a blocking call on a fresh state."""
    ctx = lib.Context(micro_path, compute_type=COMPUTE_F32, max_batch=3)
    p = default_full_params()
    clips = [synth.audio(1), synth.audio(2), synth.audio(5, 250000)]
    ref = []
    for c in clips:
        f = lib.State(ctx)
        assert f.full(c, p) == 0
        ref.append(f.segments(0))
        f.close()
    st = lib.State(ctx)
    ta = st.stream_submit_clips(clips[:2], p)
    tb = st.stream_submit_clips(clips[2:], p)
    assert st.stream_wait(ta) == 0
    assert [st.segments(i) for i in range(2)] == ref[:2]
    assert st.stream_wait(tb) == 0
    assert st.segments(0) == ref[2]
    st.close()
    ctx.close()


def test_stream_policy_change_drains(micro_path):
    ctx = lib.Context(micro_path, compute_type=COMPUTE_F32, max_batch=2)
    st = lib.State(ctx)
    p = default_full_params()
    p.temperature_inc = 0.0
    q = default_full_params()
    q.temperature_inc = 0.0
    q.no_timestamps = 1
    c = [synth.audio(3), synth.audio(4)]
    assert st.full_batch(c, p) == 0
    rp = [st.segments(i) for i in range(2)]
    assert st.full_batch(c, q) == 0
    rq = [st.segments(i) for i in range(2)]
    assert rp != rq
    t0 = st.stream_submit_clips(c, p)
    t1 = st.stream_submit_clips(c, q)  # different decode policy: waits for ticket t0's jobs first
    assert st.stream_wait(t1) == 0 and [st.segments(i) for i in range(2)] == rq
    assert st.stream_wait(t0) == 0 and [st.segments(i) for i in range(2)] == rp
    st.close()
    ctx.close()


def test_batched_beam_search_groups_match_oracle(micro_path, micro_oracle):
    """BASELINE config 4's decode policy, batched: 16 rows = 3 groups of 5 beams, so three windows decode side by side, each
    group's beams sharing one cross-attention K/V slot; token ids and segment times must equal the oracle's beam search
    (which runs one window at a time) for every clip, through the blocking call and through two live stream tickets."""
    import os

    from oracle import oracle as O
    ctx = lib.Context(micro_path, compute_type=COMPUTE_F32, max_batch=16)
    st = lib.State(ctx)
    p = default_full_params(1)
    p.temperature_inc = 0.0
    clips = [synth.audio(1), synth.audio(3, 123456), synth.audio(7, 300000), synth.audio(8), synth.audio(9, 40000), synth.audio(2)]
    ref = []
    for c in clips:
        of = O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
        assert of.full(c, p) == 0
        ref.append(flat(of.segments()))
        of.close()
    assert st.full_batch(clips, p) == 0
    assert [flat(st.segments(i)) for i in range(len(clips))] == ref
    ta = st.stream_submit_clips(clips[:4], p)
    tb = st.stream_submit_clips(clips[4:], p)
    assert st.stream_wait(ta) == 0
    assert [flat(st.segments(i)) for i in range(4)] == ref[:4]
    assert st.stream_wait(tb) == 0
    assert [flat(st.segments(i)) for i in range(2)] == ref[4:]
    # reference defaults (temperature ladder on): fresh states on both sides so the t > 0 draws see the same generator
    q = default_full_params(1)
    for c in clips[:2]:
        sf, of = lib.State(ctx), O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
        assert sf.full(c, q) == 0 and of.full(c, q) == 0
        assert flat(sf.segments(0)) == flat(of.segments())
        assert sf.timings().n_fallbacks == of.stats()["n_fallbacks"]
        sf.close()
        of.close()
    # greedy after beam on the same state: the row-state block is clean again
    g = default_full_params()
    g.temperature_inc = 0.0
    of = O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
    assert st.full(clips[0], g) == 0 and of.full(clips[0], g) == 0
    assert flat(st.segments(0)) == flat(of.segments())
    of.close()
    st.close()
    ctx.close()
