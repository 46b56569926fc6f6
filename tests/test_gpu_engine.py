"""GPU tests of the engine boundary mirror (WhisperEngine / TranscriptionWorker) on top of the C ABI."""
import numpy as np
import pytest

from openhush_b200 import COMPUTE_F32, default_full_params, engine as E, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(micro_path):
    e = E.WhisperEngine.new(micro_path, "en", False, True, compute_type=COMPUTE_F32)
    yield e
    e.close()


def test_transcribe_text_matches_oracle(eng, micro_oracle):
    """text = concatenated segment texts, trimmed (whisper.rs:276-283); the engine passes translate INVERTED
    (whisper.rs:257), so config.translate = false reaches the library as translate = true."""
    clip = synth.audio(21, 200000)
    res = eng.transcribe(E.AudioBuffer(clip, 16000))
    ost = O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
    p = default_full_params()
    p.translate = 1
    assert ost.full(clip, p) == 0
    ref = b"".join(s["text"] for s in ost.segments()).decode("utf-8").strip()
    assert res.text == ref and res.language == "en" and res.duration_ms >= 0
    assert [(a, b) for a, b, _ in res.segments] == [(s["t0"], s["t1"]) for s in ost.segments()]
    ost.close()


def test_transcribe_validation_errors(eng):
    with pytest.raises(E.ValidationFailed) as ei:
        eng.transcribe(E.AudioBuffer(np.zeros(0, np.float32), 16000))
    assert isinstance(ei.value.source, E.Empty)
    x = synth.audio(1, 32000)
    x[10] = np.nan
    with pytest.raises(E.ValidationFailed) as ei:
        eng.transcribe(E.AudioBuffer(x, 16000))
    assert isinstance(ei.value.source, E.ContainsNaN)
    with pytest.raises(E.ValidationFailed):
        eng.transcribe(E.AudioBuffer(synth.audio(1, 32000), 48000))


def test_short_padded_clip_and_auto_language(micro_path):
    e = E.WhisperEngine.new(micro_path, "auto", False, True, compute_type=COMPUTE_F32)
    # the reference pads short input to 1.1 s before the engine (src/input/audio.rs:31-34, 415-424)
    clip = np.concatenate([synth.audio(5, 8000), np.zeros(17600 - 8000, np.float32)])
    res = e.transcribe(E.AudioBuffer(clip, 16000))
    assert res.language == "en"  # "auto" never runs detection: the library default "en" is used
    assert isinstance(res.text, str)
    # below 1 s the library returns no segments and no error
    res = e.transcribe(E.AudioBuffer(synth.audio(5, 12000), 16000))
    assert res.text == ""
    e.close()


def test_benchmark_and_worker(eng):
    b = eng.benchmark(0.2)
    assert b.test_audio_secs == 2.0 and b.overhead_secs >= 0 and abs(b.recommended_chunk_interval - b.overhead_secs * 1.2) < 1e-6
    w = E.TranscriptionWorker(eng)
    job = E.TranscriptionJob(E.AudioBuffer(synth.audio(22, 100000), 16000), sequence_id=3, chunk_id=1, is_final=True)
    r = w.process_job(job)
    assert r.sequence_id == 3 and r.chunk_id == 1 and r.is_final and r.text == eng.transcribe(job.buffer).text
    bad = E.TranscriptionJob(E.AudioBuffer(np.full(32000, np.inf, np.float32), 16000), sequence_id=4)
    assert w.process_job(bad).text == ""  # engine error -> logged + empty text, the worker keeps running (worker.rs:159-165)


def test_cpp_host_mirror_gpu(micro_path):
    import os
    import subprocess

    from openhush_b200 import lib
    exe = os.path.join(lib.CSRC, "host", "host_test")
    import torch
    n_dev = min(torch.cuda.device_count(), 2)
    r = subprocess.run([exe, micro_path, "200000", "1", str(n_dev)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST_TEST_OK" in r.stdout and "TEXT_LEN" in r.stdout and "PREPROCESS_OK" in r.stdout and f"POOL_OK workers={n_dev}" in r.stdout, r.stdout + r.stderr


def test_worker_pool_on_gpus_batches_jobs(micro_path):
    """One worker (thread + context + state) per visible GPU inside ONE process over one job queue
    (/root/reference/src/queue/worker.rs:255-267, config.gpu.devices /root/reference/src/config.rs:921-929): each worker
    drains up to max_batch queued jobs into one oh_full_batch; every answer equals the single transcribe() call.  With two
    or more GPUs this also exercises the per-device kernel set-up (contexts on different devices used from different threads)."""
    import torch

    from openhush_b200.tracker import BackpressureStrategy, TranscriptionTracker
    n_dev = min(torch.cuda.device_count(), 2)
    engines = [E.WhisperEngine.new(micro_path, "en", False, True, compute_type=COMPUTE_F32, gpu_device=d, max_batch=4) for d in range(n_dev)]
    jobs = [E.TranscriptionJob(E.AudioBuffer(synth.audio(40 + i, 60000 + 9000 * i), 16000), sequence_id=i) for i in range(10)]
    jobs.append(E.TranscriptionJob(E.AudioBuffer(np.full(32000, np.nan, np.float32), 16000), sequence_id=10))  # fails validation alone
    ref = []
    for j in jobs[:10]:  # a job of a batched call equals the single call on a fresh state (sampled fallback tokens: oh_whisper.h)
        fresh = E.WhisperEngine.new(micro_path, "en", False, True, compute_type=COMPUTE_F32)
        ref.append(fresh.transcribe(j.buffer).text)
        fresh.close()
    pool = E.WorkerPool(engines)
    tracker = TranscriptionTracker.new_ordered()
    for j in jobs:
        tracker.add_pending_with_config(j.sequence_id, 0, 0, 0, BackpressureStrategy.WARN)
        pool.submit(j)
    got = pool.drain(len(jobs), tracker, timeout=300.0)
    pool.shutdown()
    assert [r.sequence_id for r in got] == list(range(11))
    assert [r.text for r in got[:10]] == ref and got[10].text == ""
    assert sum(pool.jobs_done) == 11 and sum(pool.batches_done) < 11      # jobs really went through batched calls
    if n_dev == 2:
        assert all(n > 0 for n in pool.jobs_done) or max(pool.jobs_done) == 11
    for e in engines:
        e.close()


@pytest.mark.parametrize("kind", ["q8_0", "q5_0", "q5_1", "q4_0", "q4_1"])
def test_quantised_ggml_files_load_and_match_their_f16_expansion(micro_path, tmp_path, kind):
    """model files whose 2-D weights are ggml block-quantised (what whisper.cpp's quantize tool
    writes) load, and give exactly the results of the same file with those blocks expanded to f16 beforehand — i.e. the
    loader's dequantisation is the published x = q * d (+ m) per block of 32; also the load-time getter."""
    from _ggml_quant import requantise
    from openhush_b200 import lib
    qpath, fpath = str(tmp_path / f"ggml-micro-{kind}.bin"), str(tmp_path / "ggml-micro-expanded.bin")
    n_q = requantise(micro_path, qpath, fpath, kind)
    assert n_q > 10
    p = default_full_params()
    p.temperature_inc = 0.0
    clip = synth.audio(31, 250000)
    out = []
    for path in (qpath, fpath):
        ctx = lib.Context(path, compute_type=COMPUTE_F32, max_batch=1)
        st = lib.State(ctx)
        assert ctx.load_ms > 0.0 and ctx.n_quantised == (n_q if path == qpath else 0)
        assert st.full(clip, p) == 0
        out.append(st.segments(0))
        lg = st.decode(np.array([[ctx.sot, ctx.sot + 1, 50359, 1000]], np.int32), 0)[0]
        out.append(lg)
        st.close()
        ctx.close()
    assert out[0] == out[2] and len(out[0]) > 0
    assert np.array_equal(out[1], out[3])
    # a truncated quantised file is rejected with a load error, not a crash
    blob = open(qpath, "rb").read()
    bad = str(tmp_path / "ggml-micro-truncated.bin")
    open(bad, "wb").write(blob[:len(blob) - 1000])
    with pytest.raises(lib.OhError):
        lib.Context(bad, compute_type=COMPUTE_F32, max_batch=1)


@pytest.mark.parametrize("strategy", ["greedy", "beam"])
def test_overlap_streaming_recording_matches_oracle(micro_path, micro_oracle, strategy):
    """BASELINE config 4's host side: a 70 s recording cut into 30 s windows with 0.5 s overlap
    (/root/reference/src/recording.rs:28-35, :419-426), the windows decoded side by side (beam 5 or greedy), texts merged by the
    tracker's streaming hand-out with word-overlap de-duplication (/root/reference/src/queue/mod.rs:213-274) and the
    timed segments moved to the recording's clock.  Every window's text and times must equal the oracle's for that window."""
    from openhush_b200 import _abi
    from openhush_b200.tracker import TranscriptionTracker
    rec = np.concatenate([synth.audio(60), synth.audio(61), synth.audio(62, 160000)])
    e = E.WhisperEngine.new(micro_path, "en", False, True, compute_type=COMPUTE_F32, max_batch=16)
    p = default_full_params(1 if strategy == "beam" else 0)
    p.temperature_inc = 0.0
    p.translate = 1  # what the engine passes for config.translate = false (whisper.rs:257)
    e._params = lambda: p  # the reference's engine only ever builds Greedy (whisper.rs:243); config 4 selects beam search
    text, timed, results = E.transcribe_recording(e, rec)
    windows = E.cut_overlapping_windows(rec)
    assert [s for s, _ in windows] == [0, 472000, 944000] and len(results) == 3
    tracker = TranscriptionTracker.new()
    ref_texts, ref_timed = [], []
    for k, (start, w) in enumerate(windows):
        ost = O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
        assert ost.full(np.ascontiguousarray(w), p) == 0
        segs = ost.segments()
        wtext = ""
        for s in segs:  # whisper.rs:276-280: a segment that is not valid UTF-8 is skipped
            try:
                wtext += s["text"].decode("utf-8")
            except UnicodeDecodeError:
                pass
        wtext = wtext.strip()
        assert results[k].text == wtext, k
        tracker.add_result(E.JobResult(wtext, 0, k, k == 2, len(w) / 16000.0))
        ref_texts += [t.text for t in tracker.take_ready() if t.text]
        for s in segs:
            try:
                s["text"].decode("utf-8")
                ref_timed.append((start // 160 + s["t0"], start // 160 + s["t1"]))
            except UnicodeDecodeError:
                pass
        ost.close()
    assert text == " ".join(ref_texts) and len(text) > 0
    assert [(a, b) for a, b, _ in timed] == sorted(ref_timed)
    e.close()
