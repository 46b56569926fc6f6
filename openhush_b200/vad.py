"""VAD-driven windowing: host mirror of the reference's `VadEngine` trait, `VadConfig`, `VadResult`,
`SpeechSegment` and the streaming speech-segment state machine `VadState` (/root/reference/src/vad/mod.rs:34-245), plus the
cut of a recording into transcription jobs that the daemon's VAD path makes (/root/reference/src/daemon.rs:2062-2138).

The reference's engine behind the trait is Silero VAD, an ONNX model run through `ort` (/root/reference/src/vad/silero.rs:1-9);
its weights are not in the reference tree and no ONNX runtime exists here, so Silero is NOT restated.  What the trait makes
swappable is exactly the engine: `DeviceEnergyVad` below is a stand-in engine (per-chunk RMS level computed on the GPU for the
whole recording in one launch, mapped to a probability) that exercises the same state machine and job cut; it is an
energy detector, not a model of Silero's decisions."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np


@dataclass
class VadConfig:  # vad/mod.rs:58-103 (defaults :68-102)
    enabled: bool = False
    threshold: float = 0.5
    min_silence_ms: int = 700
    min_speech_ms: int = 250
    speech_pad_ms: int = 30


@dataclass
class VadResult:
    probability: float
    is_speech: bool


@dataclass
class SpeechSegment:  # vad/mod.rs:106-116
    start: int
    end: int
    avg_probability: float


@dataclass
class VadState:
    """vad/mod.rs:121-245: speech starts at the first speech chunk, ends after min_silence_ms of non-speech (at the start of
    that silence), and is reported only if it lasted min_speech_ms."""
    config: VadConfig
    sample_rate: int = 16000
    probabilities: list = field(default_factory=list)
    in_speech: bool = False
    speech_start: int | None = None
    silence_samples: int = 0
    total_samples: int = 0

    def update(self, result: VadResult, chunk_samples: int) -> SpeechSegment | None:
        self.probabilities.append(result.probability)
        prev_total = self.total_samples
        self.total_samples += chunk_samples
        min_silence = int(np.float32(self.config.min_silence_ms) / np.float32(1000.0) * np.float32(self.sample_rate))
        min_speech = int(np.float32(self.config.min_speech_ms) / np.float32(1000.0) * np.float32(self.sample_rate))
        if result.is_speech:
            self.silence_samples = 0
            if not self.in_speech:
                self.in_speech = True
                self.speech_start = prev_total
            return None
        self.silence_samples += chunk_samples
        if self.in_speech and self.silence_samples >= min_silence:
            self.in_speech = False
            start = self.speech_start if self.speech_start is not None else 0
            self.speech_start = None
            end = prev_total  # the segment ends where the silence that closed it was noticed
            if end - start >= min_speech:
                avg = float(np.sum(np.asarray(self.probabilities, np.float32), dtype=np.float32) / np.float32(len(self.probabilities))) if self.probabilities else 0.0
                self.probabilities.clear()
                return SpeechSegment(start, end, avg)
            self.probabilities.clear()
        return None

    def is_speech(self) -> bool:
        return self.in_speech

    def reset(self) -> None:
        self.probabilities.clear()
        self.in_speech = False
        self.speech_start = None
        self.silence_samples = 0
        self.total_samples = 0


class VadEngine:
    """vad/mod.rs:34-55.  process(samples) -> VadResult; reset(); chunk_size(); sample_rate()."""

    def process(self, samples: np.ndarray) -> VadResult:
        raise NotImplementedError

    def reset(self) -> None:
        pass

    def chunk_size(self) -> int:
        return 512  # 32 ms at 16 kHz, Silero's chunk (vad/silero.rs:1-9)

    def sample_rate(self) -> int:
        return 16000


class DeviceEnergyVad(VadEngine):
    """Stand-in engine (see the module docstring): the RMS level of every 512-sample chunk of a recording, computed by one
    kernel launch (`oh_chunk_rms_db`), mapped to a speech probability by a logistic curve around `level_db`."""

    def __init__(self, state, config: VadConfig, level_db: float = -40.0, slope_db: float = 3.0):
        self.state, self.config, self.level_db, self.slope_db = state, config, level_db, slope_db

    def chunk_levels(self, samples: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float32)
        n_chunks = len(x) // self.chunk_size()
        out = np.empty(max(n_chunks, 1), dtype=np.float32)
        rc = self.state.lib.oh_chunk_rms_db(self.state.ctx.h, self.state.h, x.ctypes.data, len(x), self.chunk_size(), out.ctypes.data, n_chunks)
        if rc != 0:
            raise RuntimeError(f"oh_chunk_rms_db rc={rc}")
        return out[:n_chunks]

    def probabilities(self, samples: np.ndarray) -> np.ndarray:
        return 1.0 / (1.0 + np.exp(-(self.chunk_levels(samples) - self.level_db) / self.slope_db))

    def process(self, samples: np.ndarray) -> VadResult:
        p = float(self.probabilities(samples)[0]) if len(samples) >= self.chunk_size() else 0.0
        return VadResult(p, p >= self.config.threshold)


def speech_jobs(samples: np.ndarray, probabilities, config: VadConfig, sequence_id: int = 0, chunk_size: int = 512, sample_rate: int = 16000):
    """The daemon's VAD path on a whole recording (daemon.rs:2062-2138): chunk probabilities -> VadState -> one
    TranscriptionJob per finished speech segment.  As there, a job starts where its speech started and ends at the position
    at which the segment's end was noticed (i.e. it includes the closing silence); chunk ids count the jobs."""
    from .engine import AudioBuffer, TranscriptionJob
    state = VadState(config, sample_rate)
    jobs, speech_start_pos, pos = [], None, 0
    for p in probabilities:
        r = VadResult(float(p), float(p) >= config.threshold)
        if r.is_speech and speech_start_pos is None:
            speech_start_pos = pos
        seg = state.update(r, chunk_size)
        cur = pos + chunk_size
        if seg is not None:
            start = speech_start_pos if speech_start_pos is not None else seg.start
            speech_start_pos = None
            jobs.append(TranscriptionJob(AudioBuffer(np.ascontiguousarray(samples[start:cur], dtype=np.float32), sample_rate), sequence_id, len(jobs), False))
        pos = cur
    return jobs
