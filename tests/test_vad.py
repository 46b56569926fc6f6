"""the VAD state machine mirror against the reference's own unit tests
(/root/reference/src/vad/mod.rs:246-316, same inputs and expectations), the job cut of the daemon's VAD path, and (GPU) the
chunk-level kernel of the stand-in engine."""
import numpy as np
import pytest

from openhush_b200.vad import DeviceEnergyVad, SpeechSegment, VadConfig, VadResult, VadState, speech_jobs


def test_vad_state_speech_detection():  # vad/mod.rs:250-285
    state = VadState(VadConfig(threshold=0.5, min_silence_ms=100, min_speech_ms=50), 16000)
    speech, silence = VadResult(0.8, True), VadResult(0.1, False)
    assert state.update(speech, 512) is None and state.is_speech()
    assert state.update(speech, 512) is None
    assert state.update(silence, 512) is None and state.is_speech()       # brief silence: not enough to end
    seg = state.update(silence, 1600)                                       # more silence: the segment is returned
    assert isinstance(seg, SpeechSegment) and not state.is_speech()
    assert (seg.start, seg.end) == (0, 1536) and abs(seg.avg_probability - (0.8 + 0.8 + 0.1 + 0.1) / 4) < 1e-6


def test_vad_state_too_short():  # vad/mod.rs:287-315
    state = VadState(VadConfig(threshold=0.5, min_silence_ms=100, min_speech_ms=500), 16000)
    state.update(VadResult(0.8, True), 512)
    assert state.update(VadResult(0.1, False), 3200) is None


def test_defaults_and_reset():
    c = VadConfig()
    assert (c.enabled, c.threshold, c.min_silence_ms, c.min_speech_ms, c.speech_pad_ms) == (False, 0.5, 700, 250, 30)  # vad/mod.rs:68-102
    s = VadState(c)
    s.update(VadResult(0.9, True), 512)
    s.reset()
    assert not s.is_speech() and s.total_samples == 0 and s.speech_start is None and s.probabilities == []


def test_speech_jobs_cut():
    """Two bursts of speech separated by a second of silence -> two jobs, each from its speech start to the point where the
    closing silence (700 ms) was noticed.  Faithful to the reference's arithmetic: a segment's length is measured up to the
    chunk in which the closing silence completed (`end = prev_total`, vad/mod.rs:189), so with 32 ms chunks even a 64 ms blip
    is reported (64 ms + ~700 ms >= min_speech 250 ms); only silence arriving in one large chunk drops it (the reference's
    own test_vad_state_too_short)."""
    sr, ch = 16000, 512
    x = np.zeros(sr * 8, np.float32)
    probs = np.full(len(x) // ch, 0.05)
    def burst(t0, t1):
        probs[int(t0 * sr) // ch:int(t1 * sr) // ch] = 0.9
    burst(0.5, 2.0)
    burst(3.5, 5.5)
    jobs = speech_jobs(x, probs, VadConfig())
    assert [j.chunk_id for j in jobs] == [0, 1] and not any(j.is_final for j in jobs)
    assert abs(jobs[0].buffer.duration_secs() - (2.0 - 0.5 + 0.7)) < 0.07
    assert abs(jobs[1].buffer.duration_secs() - (5.5 - 3.5 + 0.7)) < 0.07
    probs2 = np.full(len(x) // ch, 0.05)
    probs2[10:12] = 0.9                       # 64 ms blip at 0.32 s
    probs2[int(3.0 * sr) // ch:int(4.0 * sr) // ch] = 0.9
    jobs2 = speech_jobs(x, probs2, VadConfig())
    assert len(jobs2) == 2 and abs(jobs2[0].buffer.duration_secs() - (0.064 + 0.7)) < 0.07 and abs(jobs2[1].buffer.duration_secs() - 1.7) < 0.07


@pytest.mark.gpu
def test_device_chunk_levels_and_engine(micro_path):
    from openhush_b200 import COMPUTE_F32, lib, synth
    ctx = lib.Context(micro_path, compute_type=COMPUTE_F32, max_batch=1)
    st = lib.State(ctx)
    x = np.concatenate([np.zeros(16000, np.float32), synth.audio(5, 48000), np.zeros(24000, np.float32), 0.01 * synth.audio(6, 16000)])
    eng = DeviceEnergyVad(st, VadConfig())
    lv = eng.chunk_levels(x)
    n = len(x) // 512
    ref = 20 * np.log10(np.maximum(np.sqrt(np.mean(x[:n * 512].reshape(n, 512).astype(np.float64) ** 2, axis=1)), 1e-10))
    assert lv.shape == (n,) and np.abs(lv - ref).max() < 1e-3
    assert np.all(lv[:31] == -200.0)
    jobs = speech_jobs(x, eng.probabilities(x), VadConfig())
    assert len(jobs) == 1 and abs(jobs[0].buffer.duration_secs() - (3.0 + 0.7)) < 0.1   # the loud burst; the -40 dB tail stays below the level
    assert eng.process(x[16000:16512]).is_speech and not eng.process(x[:512]).is_speech
    st.close()
    ctx.close()
