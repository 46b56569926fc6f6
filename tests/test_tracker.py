"""The step after the hot path: ordering, streaming merge and word-overlap dedup of results.
The first nine tests restate the reference's own unit tests (/root/reference/src/queue/mod.rs:324-466) against the
host mirror, with the same inputs and expected values; the rest cover what the mirror adds (timed segments)."""
from openhush_b200.engine import JobResult
from openhush_b200.tracker import BackpressureStrategy, TranscriptionTracker, merge_timed_segments


def result(seq, chunk, text, is_final):  # queue/mod.rs:313-322
    return JobResult(text, seq, chunk, is_final, 1.0)


def test_streaming_mode_outputs_immediately():
    tracker = TranscriptionTracker.new()
    tracker.add_pending(0, 0)
    tracker.add_pending(0, 1)
    tracker.add_result(result(0, 1, "world", True))  # out of order
    ready = tracker.take_ready()
    assert len(ready) == 1 and ready[0].text == "world"
    tracker.add_result(result(0, 0, "hello", False))
    ready = tracker.take_ready()
    assert len(ready) == 1 and ready[0].text == "hello"


def test_ordered_mode_waits():
    tracker = TranscriptionTracker.new_ordered()
    tracker.add_pending(0, 0)
    tracker.add_pending(1, 0)
    tracker.add_result(result(1, 0, "second", True))
    assert tracker.take_ready() == []
    tracker.add_result(result(0, 0, "first", True))
    ready = tracker.take_ready()
    assert [r.text for r in ready] == ["first", "second"]


def test_deduplication():
    tracker = TranscriptionTracker.new()
    tracker.add_pending(0, 0)
    tracker.add_result(result(0, 0, "hello world this is a test", False))
    assert tracker.take_ready()[0].text == "hello world this is a test"
    tracker.add_pending(0, 1)
    tracker.add_result(result(0, 1, "is a test and more words", True))
    assert tracker.take_ready()[0].text == "and more words"


def test_empty_tracker():
    tracker = TranscriptionTracker.new()
    assert tracker.is_empty() and tracker.pending_count() == 0 and tracker.waiting_count() == 0


def test_pending_count():
    tracker = TranscriptionTracker.new()
    tracker.add_pending(0, 0)
    tracker.add_pending(0, 1)
    assert tracker.pending_count() == 2
    tracker.add_result(result(0, 0, "test", False))
    assert tracker.pending_count() == 1 and tracker.waiting_count() == 1


def test_backpressure_drop_newest():
    tracker = TranscriptionTracker.new()
    for c in range(3):
        assert tracker.add_pending_with_config(0, c, 3, 2, BackpressureStrategy.DROP_NEWEST)
    assert not tracker.add_pending_with_config(0, 3, 3, 2, BackpressureStrategy.DROP_NEWEST)
    assert tracker.pending_count() == 3


def test_backpressure_drop_oldest():
    tracker = TranscriptionTracker.new()
    for c in range(3):
        assert tracker.add_pending_with_config(0, c, 3, 2, BackpressureStrategy.DROP_OLDEST)
    assert tracker.add_pending_with_config(0, 3, 3, 2, BackpressureStrategy.DROP_OLDEST)
    assert tracker.pending_count() == 3
    assert (0, 0) not in tracker.pending and (0, 3) in tracker.pending


def test_backpressure_warn_accepts():
    tracker = TranscriptionTracker.new()
    for c in range(4):
        assert tracker.add_pending_with_config(0, c, 3, 2, BackpressureStrategy.WARN)
    assert tracker.pending_count() == 4


def test_queue_stats():
    tracker = TranscriptionTracker.new()
    tracker.add_pending(0, 0)
    tracker.add_pending(0, 1)
    tracker.add_result(result(0, 0, "test", False))
    stats = tracker.stats()
    assert stats.pending_count == 1 and stats.waiting_count == 1


# ---- beyond the reference's tests ---------------------------------------------------------------------------
def test_dedup_details():
    tracker = TranscriptionTracker.new()
    tracker.add_result(result(0, 0, "short", False))          # <= 10 bytes: the remembered tail is not updated
    assert tracker.take_ready()[0].text == "short" and tracker.last_text_suffix == ""
    long = "x" * 40 + " the quick brown fox jumps"
    tracker.add_result(result(0, 1, long, False))
    tracker.take_ready()
    assert tracker.last_text_suffix == long[-50:]                # last 50 bytes
    tracker.add_result(result(0, 2, "fox jumps over", False))    # at most the first 10 words are candidates
    assert tracker.take_ready()[0].text == "over"
    tracker.reset_dedup()
    tracker.add_result(result(0, 3, "fox jumps again", True))
    assert tracker.take_ready()[0].text == "fox jumps again"
    # multi-byte text: the 50-byte cut never lands inside a character
    tracker.add_result(result(1, 0, "é" * 40, True))
    tracker.take_ready()
    assert tracker.last_text_suffix == "é" * 25
    # streaming hand-out is sorted by (sequence_id, chunk_id) whatever the completion order
    tracker.reset_dedup()
    for seq, chunk in [(2, 1), (1, 5), (2, 0)]:
        tracker.add_result(result(seq, chunk, f"{seq}.{chunk}", False))
    assert [(r.sequence_id, r.chunk_id) for r in tracker.take_ready()] == [(1, 5), (2, 0), (2, 1)]


def test_merge_timed_segments():
    a = JobResult("a b", 7, 0, False, 30.0, [(0, 250, "a"), (250, 2900, " b")])
    b = JobResult("c", 7, 1, True, 12.5, [(10, 400, "c")])
    assert merge_timed_segments([b, a]) == [(0, 250, "a"), (250, 2900, " b"), (3010, 3400, "c")]
    # overlapping capture: chunk 1 starts 0.5 s before chunk 0 ends
    assert merge_timed_segments([a, b], chunk_offsets_cs={0: 0, 1: 2950})[-1] == (2960, 3350, "c")


def test_worker_pool_shared_queue_and_ordered_handout():
    """one worker per GPU pulling from one job queue; the tracker restores the order.  Engines are stubs
    here (the GPU engines are exercised by the -m gpu tests).  A worker without an engine takes no jobs while another has
    one, so every text is a real transcript; workers drain up to max_batch jobs into one batched engine call."""
    import threading
    import time

    import numpy as np

    from openhush_b200.engine import AudioBuffer, TranscriptionFailed, TranscriptionJob, TranscriptionResult, WorkerPool

    class StubEngine:
        def __init__(self, delay):
            self.delay, self.batch_sizes = delay, []

        def transcribe(self, audio):
            time.sleep(self.delay)
            if len(audio.samples) == 16000 + 5:
                raise RuntimeError("not a WhisperError")                 # must not kill the worker thread
            return TranscriptionResult(f"n={len(audio.samples)}", "en", 1, [(0, 10, "x")])

        def transcribe_batch(self, audios):
            self.batch_sizes.append(len(audios))
            out = []
            for a in audios:
                try:
                    out.append(self.transcribe(a))
                except RuntimeError as e:
                    out.append(TranscriptionFailed(str(e)))
            return out

    engines = [StubEngine(0.02), StubEngine(0.001), None]
    pool = WorkerPool(engines, max_batch=4)
    tracker = TranscriptionTracker.new_ordered()
    n = 24
    for seq in range(n):
        tracker.add_pending_with_config(seq, 0, 0, 0, BackpressureStrategy.WARN)
        pool.submit(TranscriptionJob(AudioBuffer(np.zeros(16000 + seq, np.float32), 16000), seq))
    got = pool.drain(n, tracker, timeout=30.0)
    pool.shutdown()
    assert [r.sequence_id for r in got] == list(range(n))               # ordered hand-out despite out-of-order completion
    assert sum(pool.jobs_done) == n and pool.jobs_done[2] == 0           # every job taken exactly once, none by the engine-less worker
    assert all(r.text == f"n={16000 + r.sequence_id}" for r in got if r.sequence_id != 5)
    assert got[5].text == ""                                             # the failing job alone is answered with an empty text
    assert all(r.segments == [(0, 10, "x")] for r in got if r.text)
    sizes = engines[0].batch_sizes + engines[1].batch_sizes
    assert max(sizes) <= 4 and sum(sizes) + sum(1 for _ in []) <= n      # batched calls never exceed max_batch
    assert threading.active_count() < 20


def test_worker_pool_keeps_drains_in_flight_on_a_stream_engine():
    """An engine with the stream pair (submit_batch / collect over oh_stream_submit / oh_stream_wait) gets the next drain
    submitted before the previous one is collected — at most `depth` beyond the one collected next — handles are collected
    in submission order, a failing submit or collect costs only that drain's texts, and shutdown collects what is in flight."""
    import numpy as np

    from openhush_b200.engine import AudioBuffer, TranscriptionJob, TranscriptionResult, WorkerPool

    import threading

    class StreamEngine:
        def __init__(self):
            self.log, self.live, self.max_live, self.next_ticket = [], 0, 0, 0
            self.gate = threading.Event()  # the first submit waits until the producer has queued every job

        def transcribe(self, audio):  # the single-job form must not be used by the pipelined worker
            raise AssertionError("blocking path used")

        def submit_batch(self, audios):
            self.gate.wait(10.0)
            t = self.next_ticket
            self.next_ticket += 1
            self.live += 1
            self.max_live = max(self.max_live, self.live)
            self.log.append(("submit", t, len(audios)))
            if any(len(a.samples) == 16000 + 9 for a in audios):
                self.live -= 1
                raise RuntimeError("submit failed")
            return dict(ticket=t, audios=audios)

        def collect(self, handle):
            self.log.append(("collect", handle["ticket"]))
            self.live -= 1
            if any(len(a.samples) == 16000 + 17 for a in handle["audios"]):
                raise RuntimeError("collect failed")
            return [TranscriptionResult(f"n={len(a.samples)}", "en", 1, []) for a in handle["audios"]]

    for depth in (1, 2):
        eng = StreamEngine()
        pool = WorkerPool([eng], max_batch=4, depth=depth)
        tracker = TranscriptionTracker.new_ordered()
        n = 26
        for seq in range(n):
            tracker.add_pending_with_config(seq, 0, 0, 0, BackpressureStrategy.WARN)
            pool.submit(TranscriptionJob(AudioBuffer(np.zeros(16000 + seq, np.float32), 16000), seq))
        eng.gate.set()
        got = pool.drain(n, tracker, timeout=30.0)
        pool.shutdown()
        assert [r.sequence_id for r in got] == list(range(n)) and sum(pool.jobs_done) == n
        bad = {r.sequence_id for r in got if r.text == ""}
        assert {9, 17} <= bad and len(bad) <= 8                          # only the drains holding job 9 / job 17 lose their texts
        assert all(r.text == f"n={16000 + r.sequence_id}" for r in got if r.sequence_id not in bad)
        assert eng.max_live <= depth + 1 and eng.live == 0
        collected = [e[1] for e in eng.log if e[0] == "collect"]
        assert collected == sorted(collected)                            # oldest handle first
        assert eng.max_live == depth + 1                                 # drains were submitted while others were in flight


def test_worker_without_any_engine_still_answers_every_job():
    """The reference's single-worker rule (worker.rs:125-142): with no engine at all the tracker must not stall."""
    import numpy as np

    from openhush_b200.engine import AudioBuffer, TranscriptionJob, WorkerPool

    pool = WorkerPool([None, None])
    for seq in range(5):
        pool.submit(TranscriptionJob(AudioBuffer(np.zeros(16000, np.float32), 16000), seq))
    got = pool.drain(5, timeout=10.0)
    pool.shutdown()
    assert sorted(r.sequence_id for r in got) == list(range(5)) and all(r.text == "" for r in got)


def test_timed_output_formats_from_model_timestamps():
    """recording.rs:528-590 (the reference's own tests for the timestamp / segment formatters, same inputs and strings) and
    the N1 extension: segments carry the model's (t0, t1) instead of the capture loop's wall clock."""
    from types import SimpleNamespace

    import pytest

    from openhush_b200.tracker import (OutputFormat, TranscribedSegment, format_srt_timestamp, format_timestamp, format_vtt_timestamp,
                                       render_transcript, transcribed_segments)
    assert format_timestamp(0.0) == "00:00:00" and format_timestamp(61.0) == "00:01:01" and format_timestamp(3661.0) == "01:01:01"
    assert format_srt_timestamp(0.0) == "00:00:00,000" and format_srt_timestamp(1.5) == "00:00:01,500"
    assert format_srt_timestamp(3661.123) == "01:01:01,123"
    assert format_vtt_timestamp(0.0) == "00:00:00.000" and format_vtt_timestamp(1.5) == "00:00:01.500"
    assert OutputFormat.from_str("text") is OutputFormat.TEXT and OutputFormat.from_str("srt") is OutputFormat.SRT
    assert OutputFormat.from_str("VTT") is OutputFormat.VTT and OutputFormat.from_str("ts") is OutputFormat.TIMESTAMPED
    with pytest.raises(ValueError):
        OutputFormat.from_str("invalid")
    seg = TranscribedSegment(0.0, 5.0, "Hello world")
    assert "[00:00:00]" in seg.format_timestamped()
    assert "00:00:00,000 --> 00:00:05,000" in seg.format_srt(1) and "00:00:00.000 --> 00:00:05.000" in seg.format_vtt()
    assert TranscribedSegment(1.0, 2.0, "hi", speaker_id=3).format_srt(2) == "2\n00:00:01,000 --> 00:00:02,000\n<v Speaker 3>hi\n"
    # model timestamps (10 ms units, per chunk) -> recording clock -> subtitles
    a = SimpleNamespace(chunk_id=0, duration_secs=30.0, segments=[(0, 250, " first words"), (250, 2900, " and more")])
    b = SimpleNamespace(chunk_id=1, duration_secs=30.0, segments=[(10, 400, " next window"), (400, 400, "  ")])
    segs = transcribed_segments(merge_timed_segments([b, a]))
    assert [(s.start_secs, s.end_secs, s.text) for s in segs] == [(0.0, 2.5, "first words"), (2.5, 29.0, "and more"), (30.1, 34.0, "next window")]
    assert render_transcript(segs, OutputFormat.SRT).startswith("1\n00:00:00,000 --> 00:00:02,500\nfirst words\n2\n00:00:02,500 --> 00:00:29,000\n")
    assert render_transcript(segs, OutputFormat.VTT).startswith("WEBVTT\n\n00:00:00.000 --> 00:00:02.500\nfirst words\n")
    assert render_transcript(segs, OutputFormat.TEXT) == "first words\nand more\nnext window\n"
    assert render_transcript(segs, OutputFormat.TIMESTAMPED).splitlines()[2] == "[00:00:30] next window"
