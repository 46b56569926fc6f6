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
