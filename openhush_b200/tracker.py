"""Host mirror of the step right after the hot path: the ordering / streaming merge / word-overlap dedup of
transcription results.  Same names, arguments and behaviour as the reference's
`TranscriptionTracker` (/root/reference/src/queue/mod.rs:44-298) and `BackpressureStrategy`
(/root/reference/src/config.rs), so tests/test_tracker.py reads like the reference's own unit tests
(/root/reference/src/queue/mod.rs:300-466).  Pure host logic: no GPU, no oracle.

Additive over the reference: results may carry timed segments `(t0, t1, text)` in 10 ms units
(`JobResult.segments`, from oh_full_get_segment_t0/t1), and `merge_timed_segments` lays the chunks of one recording
on a common clock — what SRT/VTT writers need instead of wall-clock chunk times
(/root/reference/src/recording.rs:119-194 uses wall-clock)."""
from __future__ import annotations

import enum
import logging
from dataclasses import dataclass

log = logging.getLogger("openhush_b200.tracker")


class BackpressureStrategy(enum.Enum):
    WARN = "warn"
    DROP_OLDEST = "drop_oldest"
    DROP_NEWEST = "drop_newest"


@dataclass
class QueueStats:  # queue/mod.rs:300-308
    pending_count: int
    waiting_count: int


class TranscriptionTracker:
    """queue/mod.rs:57-298.  Streaming mode (default): every completed chunk is handed out at once, sorted by
    (sequence_id, chunk_id), with the words that repeat the tail of the previous output removed.  Ordered mode:
    results are held until every lower sequence id has been handed out."""

    def __init__(self, streaming: bool = True):
        self.pending: set[tuple[int, int]] = set()
        self.completed: dict[tuple[int, int], object] = {}
        self.next_output_id = 0
        self.streaming = streaming
        self.last_text_suffix = ""

    @classmethod
    def new(cls):  # queue/mod.rs:74-79
        return cls(True)

    @classmethod
    def new_ordered(cls):  # queue/mod.rs:83-88
        return cls(False)

    # ---- pending side -------------------------------------------------------------------------------------
    def add_pending(self, sequence_id: int, chunk_id: int) -> bool:  # queue/mod.rs:94-96: max 10, high water 8, Warn
        return self.add_pending_with_config(sequence_id, chunk_id, 10, 8, BackpressureStrategy.WARN)

    def add_pending_with_config(self, sequence_id: int, chunk_id: int, max_pending: int, high_water_mark: int,
                                strategy: BackpressureStrategy) -> bool:
        """queue/mod.rs:110-174.  max_pending 0 = unlimited.  Returns False only for DROP_NEWEST at capacity."""
        n = len(self.pending)
        if max_pending > 0 and n >= max_pending:
            if strategy is BackpressureStrategy.DROP_OLDEST:
                if self.pending:
                    oldest = min(self.pending)
                    self.pending.discard(oldest)
                    log.warning("Backpressure: dropped oldest job (seq %d.%d) to accept (seq %d.%d)", *oldest, sequence_id, chunk_id)
            elif strategy is BackpressureStrategy.DROP_NEWEST:
                log.warning("Backpressure: rejecting job (seq %d.%d) - queue full (%d/%d)", sequence_id, chunk_id, n, max_pending)
                return False
            else:
                log.warning("Queue at capacity (%d/%d) but accepting job anyway", n, max_pending)
        elif high_water_mark > 0 and n >= high_water_mark:
            log.warning("Queue depth %d approaching limit %d - transcription falling behind", n, max_pending)
        self.pending.add((sequence_id, chunk_id))
        return True

    def stats(self) -> QueueStats:
        return QueueStats(len(self.pending), len(self.completed))

    # ---- completed side -----------------------------------------------------------------------------------
    def add_result(self, result) -> None:  # queue/mod.rs:186-198
        key = (result.sequence_id, result.chunk_id)
        self.pending.discard(key)
        self.completed[key] = result

    def take_ready(self) -> list:  # queue/mod.rs:204-210
        return self._take_ready_streaming() if self.streaming else self._take_ready_ordered()

    def _take_ready_streaming(self) -> list:  # queue/mod.rs:213-233
        ready = [self.completed[k] for k in sorted(self.completed)]
        self.completed = {}
        for r in ready:
            if self.last_text_suffix and r.text:
                r.text = self.deduplicate_text(r.text)
            # the reference keeps the last 50 BYTES of the UTF-8 text when it is longer than 10 bytes (a byte offset
            # inside a multi-byte character would panic there; here the cut moves forward to the next boundary)
            raw = r.text.encode("utf-8")
            if len(raw) > 10:
                start = max(0, len(raw) - 50)
                while start < len(raw) and (raw[start] & 0xC0) == 0x80:
                    start += 1
                self.last_text_suffix = raw[start:].decode("utf-8")
        return ready

    def _take_ready_ordered(self) -> list:  # queue/mod.rs:236-247: by sequence id, chunk 0 only
        ready = []
        while (self.next_output_id, 0) in self.completed:
            ready.append(self.completed.pop((self.next_output_id, 0)))
            self.next_output_id += 1
        return ready

    def deduplicate_text(self, text: str) -> str:
        """queue/mod.rs:250-274: the longest prefix of up to 10 words of `text` that occurs (as a substring of the
        space-joined words) in the remembered tail of the previous output is dropped."""
        words = text.split()
        if not words:
            return text
        skip = 0
        for i in range(1, min(len(words), 10) + 1):
            if " ".join(words[:i]) in self.last_text_suffix:
                skip = i
        return " ".join(words[skip:]) if skip > 0 else text

    def reset_dedup(self) -> None:
        self.last_text_suffix = ""

    def is_empty(self) -> bool:
        return not self.pending and not self.completed

    def pending_count(self) -> int:
        return len(self.pending)

    def waiting_count(self) -> int:
        return len(self.completed)


def merge_timed_segments(results, chunk_offsets_cs=None):
    """Timed segments of one recording on a common clock.  `results`: JobResult-like objects of ONE sequence_id with
    `.chunk_id`, `.duration_secs` and `.segments` = [(t0, t1, text)] in 10 ms units relative to their own chunk (what
    oh_full_get_segment_t0/t1 return).  Chunk k starts where chunk k-1 ended unless `chunk_offsets_cs[k]` says otherwise
    (overlapping capture, /root/reference/src/recording.rs:32).  Returns [(t0, t1, text)] sorted by start time."""
    out, clock = [], 0
    for r in sorted(results, key=lambda r: r.chunk_id):
        base = chunk_offsets_cs[r.chunk_id] if chunk_offsets_cs is not None else clock
        for t0, t1, text in getattr(r, "segments", []) or []:
            out.append((base + t0, base + t1, text))
        clock = base + int(round(r.duration_secs * 100.0))
    out.sort(key=lambda s: (s[0], s[1]))
    return out


# ---------------------------------------------------------------------------------------------- recording.rs: timed output
def _ms_u32(secs: float) -> int:
    """`(secs * 1000.0) as u32` on an f32 (recording.rs:178, :188): f32 product, truncation towards zero, saturating."""
    import numpy as np
    v = float(np.float32(secs) * np.float32(1000.0))
    return 0 if not v > 0.0 else min(int(v), 0xFFFFFFFF)


def format_timestamp(secs: float) -> str:  # recording.rs:167-173, HH:MM:SS
    import numpy as np
    v = float(np.float32(secs))
    total = 0 if not v > 0.0 else min(int(v), 0xFFFFFFFF)
    return f"{total // 3600:02d}:{(total % 3600) // 60:02d}:{total % 60:02d}"


def format_srt_timestamp(secs: float) -> str:  # recording.rs:176-184, HH:MM:SS,mmm
    ms = _ms_u32(secs)
    return f"{ms // 3600000:02d}:{(ms % 3600000) // 60000:02d}:{(ms % 60000) // 1000:02d},{ms % 1000:03d}"


def format_vtt_timestamp(secs: float) -> str:  # recording.rs:186-194, HH:MM:SS.mmm
    ms = _ms_u32(secs)
    return f"{ms // 3600000:02d}:{(ms % 3600000) // 60000:02d}:{(ms % 60000) // 1000:02d}.{ms % 1000:03d}"


@dataclass
class TranscribedSegment:
    """recording.rs:117-165.  The reference fills start / end from the wall clock of the capture loop; here they come from
    the model's own timestamps (oh_full_get_segment_t0 / t1 through merge_timed_segments) — SURVEY.md section 8f N1."""
    start_secs: float
    end_secs: float
    text: str
    speaker_id: int | None = None

    def _spoken(self) -> str:
        return f"<v Speaker {self.speaker_id}>{self.text}" if self.speaker_id is not None else self.text

    def format_timestamped(self) -> str:
        start = format_timestamp(self.start_secs)
        return f"[{start}] Speaker {self.speaker_id}: {self.text}" if self.speaker_id is not None else f"[{start}] {self.text}"

    def format_srt(self, index: int) -> str:
        return f"{index}\n{format_srt_timestamp(self.start_secs)} --> {format_srt_timestamp(self.end_secs)}\n{self._spoken()}\n"

    def format_vtt(self) -> str:
        return f"{format_vtt_timestamp(self.start_secs)} --> {format_vtt_timestamp(self.end_secs)}\n{self._spoken()}\n"


def transcribed_segments(merged) -> list:
    """[(t0, t1, text)] in 10 ms units on the recording's clock (merge_timed_segments) -> TranscribedSegment list; empty
    pieces are dropped and texts stripped, as the capture loop does before it stores a segment."""
    out = []
    for t0, t1, text in merged:
        t = text.strip()
        if t:
            out.append(TranscribedSegment(t0 / 100.0, t1 / 100.0, t))
    return out


class OutputFormat(enum.Enum):  # recording.rs:72-101
    TEXT = "text"
    TIMESTAMPED = "timestamped"
    SRT = "srt"
    VTT = "vtt"

    @staticmethod
    def from_str(s: str) -> "OutputFormat":
        key = s.lower()
        for names, fmt in ((("text", "txt"), OutputFormat.TEXT), (("timestamped", "ts"), OutputFormat.TIMESTAMPED),
                           (("srt", "subrip"), OutputFormat.SRT), (("vtt", "webvtt"), OutputFormat.VTT)):
            if key in names:
                return fmt
        raise ValueError(f"Unknown format '{s}'. Use: text, timestamped, srt, vtt")


def render_transcript(segments, output_format: OutputFormat = OutputFormat.TEXT) -> str:
    """What RecordingSession::save_to_file writes (recording.rs:505-524) for a list of TranscribedSegments: plain text one
    line per segment, timestamped lines, SRT entries numbered from 1, or WebVTT cues under the "WEBVTT" header."""
    out = "WEBVTT\n\n" if output_format is OutputFormat.VTT else ""
    for i, seg in enumerate(segments):
        if output_format is OutputFormat.TEXT:
            out += f"{seg.text}\n"
        elif output_format is OutputFormat.TIMESTAMPED:
            out += f"{seg.format_timestamped()}\n"
        elif output_format is OutputFormat.SRT:
            out += seg.format_srt(i + 1)
        else:
            out += seg.format_vtt()
    return out
