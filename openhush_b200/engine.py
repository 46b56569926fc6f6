"""Host-side mirror of the reference's engine boundary, on top of the C ABI.

Same names, argument meaning and error behaviour as the Rust types it mirrors, so the tests read like the
reference's own (`src/engine/whisper.rs:741-997`, `src/engine/validation.rs:139-203`, `src/queue/worker.rs:319-538`):

    validate_audio / AudioValidationError / AudioValidationInfo   <- src/engine/validation.rs:9-136
    WhisperError, WhisperModel, TranscriptionResult, BenchmarkResult, WhisperEngine{new, transcribe, benchmark}
                                                                  <- src/engine/whisper.rs:14-387
    AudioBuffer                                                   <- src/input/audio.rs:55-61
    TranscriptionJob / JobResult / TranscriptionWorker.process_job <- src/queue/mod.rs:18-43, src/queue/worker.rs:119-193

The reference's toolchain (Rust) is not in this image; the Rust binding a maintainer would add is in INTEGRATION.md.
This module contains no arithmetic: everything goes through libohwhisper.so (there is no CPU fallback).
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass, field
from enum import Enum

import numpy as np

from . import lib as _lib
from ._abi import COMPUTE_F16

MAX_AUDIO_DURATION_SECS = 7200.0   # validation.rs:9
MIN_AUDIO_DURATION_SECS = 0.1      # validation.rs:12
EXPECTED_SAMPLE_RATE = 16000       # validation.rs:15


# ---------------------------------------------------------------------------------------------- validation.rs
class AudioValidationError(Exception):
    pass


class TooLong(AudioValidationError):
    def __init__(self, secs, limit):
        super().__init__(f"Audio too long: {secs:.1f}s exceeds maximum {limit:.1f}s")
        self.secs, self.limit = secs, limit


class TooShort(AudioValidationError):
    def __init__(self, secs, limit):
        super().__init__(f"Audio too short: {secs:.3f}s below minimum {limit:.3f}s")
        self.secs, self.limit = secs, limit


class Empty(AudioValidationError):
    def __init__(self):
        super().__init__("Audio is empty (no samples)")


class ContainsNaN(AudioValidationError):
    def __init__(self, count):
        super().__init__(f"Audio contains {count} NaN values")
        self.count = count


class ContainsInfinite(AudioValidationError):
    def __init__(self, count):
        super().__init__(f"Audio contains {count} infinite values")
        self.count = count


class InvalidSampleRate(AudioValidationError):
    def __init__(self, got, expected):
        super().__init__(f"Unexpected sample rate: {got}Hz (expected {expected}Hz)")
        self.got, self.expected = got, expected


@dataclass
class AudioValidationInfo:
    duration_secs: float
    sample_count: int
    min_value: float
    max_value: float
    rms: float


def validate_audio(samples, sample_rate: int) -> AudioValidationInfo:
    """Same checks in the same order as validation.rs:46-118: empty, sample rate, too long, too short, then one pass
    counting NaN / Inf and accumulating min / max / sum of squares (f64) over the finite samples."""
    x = np.asarray(samples, dtype=np.float32)
    if x.size == 0:
        raise Empty()
    if sample_rate != EXPECTED_SAMPLE_RATE:
        raise InvalidSampleRate(sample_rate, EXPECTED_SAMPLE_RATE)
    duration = np.float32(x.size) / np.float32(sample_rate)
    if duration > MAX_AUDIO_DURATION_SECS:
        raise TooLong(float(duration), MAX_AUDIO_DURATION_SECS)
    if duration < MIN_AUDIO_DURATION_SECS:
        raise TooShort(float(duration), MIN_AUDIO_DURATION_SECS)
    nan = np.isnan(x)
    inf = np.isinf(x)
    n_nan, n_inf = int(nan.sum()), int(inf.sum())
    if n_nan > 0:
        raise ContainsNaN(n_nan)
    if n_inf > 0:
        raise ContainsInfinite(n_inf)
    rms = math.sqrt(float(np.sum(x.astype(np.float64) ** 2)) / x.size)
    return AudioValidationInfo(float(duration), int(x.size), float(x.min()), float(x.max()), float(np.float32(rms)))


# ---------------------------------------------------------------------------------------------- input/audio.rs
@dataclass
class AudioBuffer:
    samples: np.ndarray
    sample_rate: int = 16000

    def duration_secs(self) -> float:
        return len(self.samples) / float(self.sample_rate)


# ---------------------------------------------------------------------------------------------- config.rs (audio preprocessing)
@dataclass
class NormalizationConfig:  # config.rs:1016-1023, 1239-1246
    enabled: bool = True
    target_db: float = -18.0


@dataclass
class CompressionConfig:  # config.rs:1025-1044, 1248-1259
    enabled: bool = True
    threshold_db: float = -24.0
    ratio: float = 4.0
    attack_ms: float = 5.0
    release_ms: float = 50.0
    makeup_gain_db: float = 6.0


@dataclass
class LimiterConfig:  # config.rs:1046-1056, 1261-1269
    enabled: bool = True
    ceiling_db: float = -1.0
    release_ms: float = 50.0


@dataclass
class NoiseReductionConfig:  # config.rs:993-1010 (RNNoise: a third-party model, not provided here)
    enabled: bool = False
    strength: float = 1.0


@dataclass
class AudioConfig:
    """The fields of config.rs:932-990 that the worker's preprocess_audio reads (worker.rs:196-240).  `preprocessing` is
    off by default, exactly as in the reference; the three stages are enabled by default once it is on."""
    preprocessing: bool = False
    normalization: NormalizationConfig = field(default_factory=NormalizationConfig)
    compression: CompressionConfig = field(default_factory=CompressionConfig)
    limiter: LimiterConfig = field(default_factory=LimiterConfig)
    noise_reduction: NoiseReductionConfig = field(default_factory=NoiseReductionConfig)

    def to_params(self):
        """The C-ABI struct of oh_preprocess_pcm for this configuration."""
        from ._abi import PreprocessParams
        n, c, l = self.normalization, self.compression, self.limiter
        return PreprocessParams(int(n.enabled), n.target_db, int(c.enabled), c.threshold_db, c.ratio, c.attack_ms, c.release_ms,
                                c.makeup_gain_db, int(l.enabled), l.ceiling_db, l.release_ms)


def preprocess_audio(state, buffers, config: AudioConfig):
    """worker.rs:196-240 for a batch of AudioBuffers, in place, on the device: RMS normalisation, compression and the
    limiter run as one call of oh_preprocess_pcm over all buffers (stages the configuration leaves disabled are skipped;
    nothing runs at all unless `config.preprocessing`).  Returns the [n, 2] RMS levels (dB) before / after, or None.
    Noise reduction (worker.rs:200-206) is the reference's RNNoise dependency and is not provided: asking for it raises."""
    if config.noise_reduction.enabled:
        raise NotImplementedError("noise_reduction (RNNoise) is not part of this library")
    if not config.preprocessing or not buffers:
        return None
    clips, rms = state.preprocess([b.samples for b in buffers], config.to_params())
    for b, c in zip(buffers, clips):
        b.samples = c
    return rms


# ---------------------------------------------------------------------------------------------- engine/whisper.rs
class WhisperError(Exception):
    pass


class ModelNotFound(WhisperError):
    def __init__(self, path, name):
        super().__init__(f"Model not found at {path}. Run 'openhush model download {name}'")
        self.path, self.name = path, name


class LoadFailed(WhisperError):
    def __init__(self, msg):
        super().__init__(f"Failed to load model: {msg}")


class TranscriptionFailed(WhisperError):
    def __init__(self, msg):
        super().__init__(f"Transcription failed: {msg}")


class ValidationFailed(WhisperError):
    def __init__(self, err: AudioValidationError):
        super().__init__(f"Audio validation failed: {err}")
        self.source = err


class WhisperModel(Enum):
    Tiny = "tiny"
    Base = "base"
    Small = "small"
    Medium = "medium"
    LargeV3 = "large-v3"

    @staticmethod
    def from_str(s: str) -> "WhisperModel":
        m = {"tiny": WhisperModel.Tiny, "base": WhisperModel.Base, "small": WhisperModel.Small, "medium": WhisperModel.Medium,
             "large": WhisperModel.LargeV3, "large-v3": WhisperModel.LargeV3, "largev3": WhisperModel.LargeV3}
        try:
            return m[s.lower()]
        except KeyError:
            raise ValueError(s) from None

    def filename(self) -> str:
        return f"ggml-{self.value}.bin"

    def size_bytes(self) -> int:
        return {"tiny": 75_000_000, "base": 142_000_000, "small": 466_000_000, "medium": 1_500_000_000, "large-v3": 3_000_000_000}[self.value]


@dataclass
class TranscriptionResult:
    text: str
    language: str
    duration_ms: int
    segments: list = field(default_factory=list)  # additive: (t0, t1, text) in 10 ms units — SURVEY.md §8f N1


@dataclass
class BenchmarkResult:
    overhead_secs: float
    recommended_chunk_interval: float
    test_audio_secs: float


# whisper-rs error names for the codes of whisper_full
_FULL_ERRORS = {-1: "UnableToCalculateSpectrogram", -6: "FailedToEncode", 7: "FailedToEncode", -7: "FailedToDecode", 8: "FailedToDecode"}


class WhisperEngine:
    """whisper.rs:110-310.  Owns the context and ONE cached state (buffers allocated once, whisper.rs:162-169)."""

    def __init__(self, model_path: str, language: str, translate: bool, use_gpu: bool, *, compute_type: int = COMPUTE_F16,
                 gpu_device: int = 0, max_batch: int = 1):
        if not os.path.exists(model_path):  # whisper.rs:141-154
            name = os.path.basename(model_path)
            name = name[5:] if name.startswith("ggml-") else name
            name = name[:-4] if name.endswith(".bin") else name
            raise ModelNotFound(model_path, name)
        try:
            self.context = _lib.Context(model_path, compute_type=compute_type, gpu_device=gpu_device, max_batch=max_batch, use_gpu=use_gpu)
            self.state = _lib.State(self.context)
        except _lib.OhError as e:
            raise LoadFailed(str(e)) from None
        self.language = language
        self.translate = translate

    @classmethod
    def new(cls, model_path, language, translate, use_gpu, **kw):
        return cls(model_path, language, translate, use_gpu, **kw)

    def _params(self):
        p = self.context.lib.oh_full_default_params(0)  # SamplingStrategy::Greedy { best_of: 1 }, whisper.rs:243
        p.greedy_best_of = 1
        if self.language != "auto":  # whisper.rs:246-248 — "auto" leaves the library default "en", no detection runs (A.2)
            p.language = self.language.encode()
        p.translate = 0 if self.translate else 1  # whisper.rs:251-257: the engine passes the INVERTED flag (SURVEY.md A.2 hazard)
        return p

    def transcribe(self, audio: AudioBuffer) -> TranscriptionResult:
        try:
            validate_audio(audio.samples, audio.sample_rate)  # whisper.rs:206
        except AudioValidationError as e:
            raise ValidationFailed(e) from None
        t0 = time.perf_counter()
        self._keepalive = self._params()
        rc = self.state.full(np.ascontiguousarray(audio.samples, dtype=np.float32), self._keepalive)
        if rc != 0:
            raise TranscriptionFailed(_FULL_ERRORS.get(rc, f"GenericError({rc})"))
        text, segs = "", []
        for s in self.state.segments(0):
            try:
                piece = s["text"].decode("utf-8")  # whisper.rs:276-280: a segment that is not valid UTF-8 is skipped
            except UnicodeDecodeError:
                continue
            text += piece
            segs.append((s["t0"], s["t1"], piece))
        lang = self.language
        if self.language == "auto":  # whisper.rs:288-296
            code = _lib.load().oh_full_lang_id(self.state.h)
            lang = LANG_CODES[code] if 0 <= code < len(LANG_CODES) else "auto"
        return TranscriptionResult(text.strip(), lang, int((time.perf_counter() - t0) * 1000), segs)

    def _result_of_item(self, item: int, t0: float) -> TranscriptionResult:
        text, segs = "", []
        for s in self.state.segments(item):
            try:
                piece = s["text"].decode("utf-8")  # whisper.rs:276-280: a segment that is not valid UTF-8 is skipped
            except UnicodeDecodeError:
                continue
            text += piece
            segs.append((s["t0"], s["t1"], piece))
        return TranscriptionResult(text.strip(), self.language if self.language != "auto" else "en", int((time.perf_counter() - t0) * 1000), segs)

    def transcribe_batch(self, audios) -> list:
        """Additive (the multi-GPU queue): `len(audios)` independent transcribe() calls as ONE oh_full_batch.
        Returns one entry per job: a TranscriptionResult, or the WhisperError transcribe() would have raised for that job
        (validation is per job, whisper.rs:206; a job whose log-mel cannot be computed fails alone)."""
        out = [None] * len(audios)
        clips, idx = [], []
        for i, a in enumerate(audios):
            try:
                validate_audio(a.samples, a.sample_rate)
            except AudioValidationError as e:
                out[i] = ValidationFailed(e)
                continue
            clips.append(np.ascontiguousarray(a.samples, dtype=np.float32))
            idx.append(i)
        if not clips:
            return out
        t0 = time.perf_counter()
        self._keepalive = self._params()
        rc = self.state.full_batch(clips, self._keepalive)
        for k, i in enumerate(idx):
            item_rc = rc if rc != 0 else self.state.item_rc(k)
            out[i] = TranscriptionFailed(_FULL_ERRORS.get(item_rc, f"GenericError({item_rc})")) if item_rc != 0 else self._result_of_item(k, t0)
        return out

    def submit_batch(self, audios):
        """Stream form of transcribe_batch (oh_stream_submit): registers the jobs and returns a handle at once; the jobs
        start sharing the decoder rows with those of earlier handles.  collect(handle) waits for them (oh_stream_wait) and
        returns what transcribe_batch would have.  Handles are collected in the order they were submitted."""
        out = [None] * len(audios)
        clips, idx = [], []
        for i, a in enumerate(audios):
            try:
                validate_audio(a.samples, a.sample_rate)
            except AudioValidationError as e:
                out[i] = ValidationFailed(e)
                continue
            clips.append(np.ascontiguousarray(a.samples, dtype=np.float32))
            idx.append(i)
        params = self._params()
        ticket = self.state.stream_submit_clips(clips, params) if clips else None
        return dict(ticket=ticket, out=out, idx=idx, params=params, t0=time.perf_counter())

    def collect(self, handle) -> list:
        out, idx = handle["out"], handle["idx"]
        if handle["ticket"] is not None:
            rc = self.state.stream_wait(handle["ticket"])
            for k, i in enumerate(idx):
                item_rc = rc if rc != 0 else self.state.item_rc(k)
                out[i] = TranscriptionFailed(_FULL_ERRORS.get(item_rc, f"GenericError({item_rc})")) if item_rc != 0 else self._result_of_item(k, handle["t0"])
        return out

    def benchmark(self, safety_margin: float) -> BenchmarkResult:
        """whisper.rs:334-387: 1 warm-up + 3 timed transcriptions of 2 s of silence."""
        audio = AudioBuffer(np.zeros(32000, np.float32), 16000)
        try:
            self.transcribe(audio)
        except WhisperError:
            pass
        total_ms = 0
        for _ in range(3):
            t0 = time.perf_counter()
            try:
                self.transcribe(audio)
            except WhisperError:
                pass
            total_ms += int((time.perf_counter() - t0) * 1000)
        overhead = (total_ms // 3) / 1000.0
        return BenchmarkResult(overhead, overhead * (1.0 + safety_margin), 2.0)

    def close(self):
        self.state.close()
        self.context.close()


# order of lang_id_to_code, whisper.rs:627-731
LANG_CODES = ["en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi", "he", "uk",
              "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk", "te", "fa", "lv", "bn",
              "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn",
              "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my",
              "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su"]


# ---------------------------------------------------------------------------------------------- queue/mod.rs, queue/worker.rs
@dataclass
class TranscriptionJob:
    buffer: AudioBuffer
    sequence_id: int
    chunk_id: int = 0
    is_final: bool = True


@dataclass
class JobResult:  # queue::TranscriptionResult, queue/mod.rs:31-43
    text: str
    sequence_id: int
    chunk_id: int
    is_final: bool
    duration_secs: float
    segments: list = field(default_factory=list)  # additive: (t0, t1, text) in 10 ms units, chunk-relative


class TranscriptionWorker:
    """queue/worker.rs:85-193 reduced to the hot-path caller: one engine, jobs in, one result per job out — always,
    so the tracker never stalls (engine missing → empty text, engine error → logged + empty text)."""

    def __init__(self, engine: WhisperEngine | None, audio_config: AudioConfig | None = None):
        self.engine = engine
        self.audio_config = audio_config or AudioConfig()  # worker.rs:52-53

    def _preprocess(self, jobs):
        """worker.rs:150-154: the buffers of the drained jobs go through preprocess_audio before the engine sees them (one
        device call for all of them); a failure here is an engine-side error like any other: logged, empty text."""
        if self.audio_config.preprocessing or self.audio_config.noise_reduction.enabled:
            preprocess_audio(self.engine.state, [j.buffer for j in jobs], self.audio_config)

    def load_engine(self, engine: WhisperEngine):  # WorkerCommand::LoadEngine, worker.rs:100-103
        self.engine = engine

    def unload_engine(self):  # WorkerCommand::UnloadEngine, worker.rs:104-111
        if self.engine is not None:
            self.engine.close()
        self.engine = None

    def process_job(self, job: TranscriptionJob) -> JobResult:
        dur = job.buffer.duration_secs() if self.engine is not None else 0.0  # worker.rs:136: no engine -> duration 0
        text, segs = "", []
        if self.engine is not None:
            try:
                self._preprocess([job])
                res = self.engine.transcribe(job.buffer)
                text, segs = res.text, res.segments
            except Exception:  # worker.rs:161-164: any engine error is logged and answered with an empty text
                text, segs = "", []
        return JobResult(text, job.sequence_id, job.chunk_id, job.is_final, dur, segs)

    def _results(self, jobs, res) -> list:
        out = []
        for j, r in zip(jobs, res):
            ok = isinstance(r, TranscriptionResult)
            out.append(JobResult(r.text if ok else "", j.sequence_id, j.chunk_id, j.is_final, j.buffer.duration_secs(), r.segments if ok else []))
        return out

    def submit_jobs(self, jobs):
        """Additive: process_jobs split in two for continuous batching across drains — the jobs are preprocessed and handed
        to the engine's stream (oh_stream_submit) and the worker is free to drain and submit the next ones before it calls
        collect_jobs for these, so the decoder rows do not run empty between drains.  Never raises: a failure becomes empty
        texts at collect time (worker.rs:161-164)."""
        try:
            self._preprocess(jobs)
            return self.engine.submit_batch([j.buffer for j in jobs])
        except Exception:
            return None

    def collect_jobs(self, jobs, handle) -> list:
        try:
            res = self.engine.collect(handle) if handle is not None else [None] * len(jobs)
        except Exception:
            res = [None] * len(jobs)
        return self._results(jobs, res)

    def process_jobs(self, jobs) -> list:
        """Additive: the jobs a worker drained from the queue (up to the engine's max_batch) as one batched engine call.
        One JobResult per job, always, in the order given."""
        if self.engine is None or len(jobs) == 1 or not hasattr(self.engine, "transcribe_batch"):
            return [self.process_job(j) for j in jobs]
        try:
            self._preprocess(jobs)
            res = self.engine.transcribe_batch([j.buffer for j in jobs])
        except Exception:
            res = [None] * len(jobs)
        return self._results(jobs, res)


def shard_jobs(n_jobs: int, rank: int, world: int):
    """Static round-robin assignment of independent jobs to GPU workers: job j goes to rank j % world."""
    return list(range(rank, n_jobs, world))


def merge_records(per_rank_records, world: int):
    """Inverse of the sharding after the gather of oh_full_pack_records blocks: rank r's local item i is global job
    i * world + r.  Returns {job: [(t0, t1, [token ids]), ...]}."""
    out = {}
    for r, rec in enumerate(per_rank_records):
        rec = [int(v) for v in rec]
        k = 0
        while k < len(rec):
            item, n_seg = rec[k], rec[k + 1]
            k += 2
            segs = []
            for _ in range(n_seg):
                t0, t1, n_tok = rec[k], rec[k + 1], rec[k + 2]
                k += 3
                segs.append((t0, t1, rec[k:k + n_tok]))
                k += n_tok
            out[item * world + r] = segs
    return out


class WorkerPool:
    """One worker per GPU over ONE shared job queue (the documented-but-unimplemented multi-GPU design of the
    reference: `config.gpu.devices`, /root/reference/src/config.rs:921-929, workers as in queue/worker.rs:255-267).
    Jobs are independent transcribe() calls, so a worker that finishes early simply pulls the next jobs (dynamic load
    balance: decode length is data-dependent); a worker drains up to its engine's `max_batch` queued jobs and hands them to
    the engine as ONE batched call (oh_full_batch); results carry (sequence_id, chunk_id) and go to one result queue, where
    a TranscriptionTracker restores the order.  Nothing is exchanged between GPUs.

    `engines`: one engine per device (anything with .transcribe(AudioBuffer), optionally .transcribe_batch(list)).  A None
    entry is a worker whose engine failed to load: it takes NO jobs while another worker has an engine (it would answer
    instantly with empty texts and empty the queue ahead of the working GPUs); only when no worker has an engine does one
    thread answer every job with an empty text, which is the reference's rule for its single worker (worker.rs:125-142:
    the tracker must never stall)."""

    def __init__(self, engines, max_batch: int | None = None, depth: int = 1):
        import queue
        import threading
        self.jobs: "queue.Queue" = queue.Queue()
        self.results: "queue.Queue" = queue.Queue()
        self.workers = [TranscriptionWorker(e) for e in engines]
        self.jobs_done = [0] * len(engines)
        self.batches_done = [0] * len(engines)
        self._max_batch = max_batch
        self._depth = max(0, depth)  # drains a worker may have in flight on its engine's stream beyond the one it collects next
        live = [i for i, e in enumerate(engines) if e is not None]
        consumers = live if live else list(range(min(1, len(engines))))
        self._threads = [threading.Thread(target=self._run, args=(i,), daemon=True) for i in consumers]
        for t in self._threads:
            t.start()

    def _batch_limit(self, i: int) -> int:
        if self._max_batch is not None:
            return max(1, self._max_batch)
        ctx = getattr(self.workers[i].engine, "context", None)
        return max(1, getattr(ctx, "max_batch", 1))

    def _run(self, i: int):
        """Drain up to `limit` jobs, hand them to the engine, answer.  With an engine that has the stream pair
        (submit_batch / collect) the worker keeps `depth` drains in flight: it submits the next drain before it collects
        the previous one (oh_stream_submit / oh_stream_wait), so the decoder rows stay busy across drains; it blocks on the
        queue only when nothing is in flight."""
        import collections
        import queue
        limit = self._batch_limit(i)
        w = self.workers[i]
        pipelined = self._depth > 0 and hasattr(w.engine, "submit_batch") and hasattr(w.engine, "collect")
        inflight = collections.deque()
        stop = False

        def answer(batch, results):
            for r in results:
                self.results.put(r)
            self.jobs_done[i] += len(batch)
            self.batches_done[i] += 1

        def empty(batch):  # a result per job, whatever happens: the tracker must never stall
            return [JobResult("", j.sequence_id, j.chunk_id, j.is_final, j.buffer.duration_secs(), []) for j in batch]

        while True:
            batch = []
            if not stop and len(inflight) <= self._depth:
                try:
                    job = self.jobs.get(block=not inflight)
                except queue.Empty:
                    job = False
                if job is None:  # WorkerCommand::Shutdown
                    stop = True
                elif job is not False:
                    batch = [job]
                    while len(batch) < limit:
                        try:
                            nxt = self.jobs.get_nowait()
                        except queue.Empty:
                            break
                        if nxt is None:
                            stop = True
                            break
                        batch.append(nxt)
            if batch:
                if pipelined:
                    inflight.append((batch, w.submit_jobs(batch)))
                else:
                    try:
                        results = w.process_jobs(batch)
                    except Exception:
                        results = empty(batch)
                    answer(batch, results)
                continue
            if inflight:
                batch, handle = inflight.popleft()
                try:
                    results = w.collect_jobs(batch, handle)
                except Exception:
                    results = empty(batch)
                answer(batch, results)
                continue
            if stop:
                return

    def submit(self, job: TranscriptionJob):
        self.jobs.put(job)

    def drain(self, n: int, tracker=None, timeout: float = 600.0):
        """Collects n results; with a tracker they come back in its hand-out order (streaming or ordered)."""
        out = []
        for _ in range(n):
            r = self.results.get(timeout=timeout)
            if tracker is None:
                out.append(r)
            else:
                tracker.add_result(r)
                out.extend(tracker.take_ready())
        return out

    def shutdown(self):
        for _ in self._threads:
            self.jobs.put(None)
        for t in self._threads:
            t.join()


# ---------------------------------------------------------------------------------------------- recording.rs: overlapped chunks
CHUNK_OVERLAP_SECS = 0.5  # /root/reference/src/recording.rs:32


def cut_overlapping_windows(samples: np.ndarray, window_secs: float = 30.0, overlap_secs: float = CHUNK_OVERLAP_SECS, sample_rate: int = 16000):
    """BASELINE config 4's windowing: consecutive windows of `window_secs` in which every window after the first starts
    with the last `overlap_secs` of the previous one (recording.rs:419-426 keeps that tail "for continuity").
    Returns [(start_sample, window_samples)]."""
    n_win, n_ov = int(round(window_secs * sample_rate)), int(overlap_secs * sample_rate)
    assert n_win > n_ov >= 0
    out, start = [], 0
    while True:
        out.append((start, samples[start:start + n_win]))
        if start + n_win >= len(samples):
            return out
        start += n_win - n_ov


def transcribe_recording(engine, samples: np.ndarray, *, sequence_id: int = 0, window_secs: float = 30.0, overlap_secs: float = CHUNK_OVERLAP_SECS,
                         separator: str = " "):
    """A long recording through the hot path the way config 4 states it: overlapped 30 s windows as independent jobs of one
    sequence (batched: the engine decodes up to its max_batch of them side by side), handed to the TranscriptionTracker in
    streaming mode so that the words repeated across a window boundary are dropped (queue/mod.rs:213-274), and the windows'
    timed segments put on the recording's clock (merge_timed_segments; chunk k starts at k * (window - overlap)).
    Returns (text, [(t0, t1, text)] in 10 ms units, [JobResult per window])."""
    from .tracker import BackpressureStrategy, TranscriptionTracker, merge_timed_segments
    windows = cut_overlapping_windows(samples, window_secs, overlap_secs)
    jobs = [TranscriptionJob(AudioBuffer(np.ascontiguousarray(w, dtype=np.float32), 16000), sequence_id, k, k == len(windows) - 1)
            for k, (_, w) in enumerate(windows)]
    tracker = TranscriptionTracker.new()  # streaming mode
    worker = TranscriptionWorker(engine)
    limit = max(1, getattr(getattr(engine, "context", None), "max_batch", 1))
    texts, results = [], []
    for j in jobs:
        tracker.add_pending_with_config(j.sequence_id, j.chunk_id, 0, 0, BackpressureStrategy.WARN)
    for b0 in range(0, len(jobs), limit):
        for r in worker.process_jobs(jobs[b0:b0 + limit]):
            results.append(r)
            tracker.add_result(JobResult(r.text, r.sequence_id, r.chunk_id, r.is_final, r.duration_secs, list(r.segments)))  # the tracker edits its copy
            texts.extend(t.text for t in tracker.take_ready() if t.text)
    offsets = {k: int(round(start / 160.0)) for k, (start, _) in enumerate(windows)}  # 10 ms units
    return separator.join(texts), merge_timed_segments(results, chunk_offsets_cs=offsets), results
