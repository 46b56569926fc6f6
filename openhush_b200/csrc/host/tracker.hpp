// C++ host mirror of the reference's TranscriptionTracker (/root/reference/src/queue/mod.rs:44-298): the step right
// after the hot path — pending/completed bookkeeping with backpressure, streaming hand-out sorted by
// (sequence_id, chunk_id) with word-overlap de-duplication against the tail of the previous output, or strictly
// ordered hand-out by sequence id.  Same names, arguments and results as the reference.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <optional>
#include <set>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "engine.hpp"

namespace openhush {

enum class BackpressureStrategy { Warn, DropOldest, DropNewest };  // config.rs:868-879

struct QueueStats {
    size_t pending_count, waiting_count;
};

class TranscriptionTracker {
  public:
    using ChunkKey = std::pair<uint64_t, uint32_t>;
    explicit TranscriptionTracker(bool streaming = true) : streaming_(streaming) {}
    static TranscriptionTracker new_ordered() { return TranscriptionTracker(false); }

    bool add_pending(uint64_t sequence_id, uint32_t chunk_id) { return add_pending_with_config(sequence_id, chunk_id, 10, 8, BackpressureStrategy::Warn); }
    // max_pending 0 = unlimited; false only for DropNewest at capacity (queue/mod.rs:110-174)
    bool add_pending_with_config(uint64_t sequence_id, uint32_t chunk_id, uint32_t max_pending, uint32_t /*high_water_mark*/, BackpressureStrategy strategy) {
        if (max_pending > 0 && pending_.size() >= max_pending) {
            if (strategy == BackpressureStrategy::DropOldest) {
                if (!pending_.empty()) pending_.erase(pending_.begin());  // std::set is ordered: begin() is the oldest key
            } else if (strategy == BackpressureStrategy::DropNewest) {
                return false;
            }
        }
        pending_.insert({sequence_id, chunk_id});
        return true;
    }
    void add_result(JobResult r) {
        const ChunkKey k{r.sequence_id, r.chunk_id};
        pending_.erase(k);
        completed_.insert_or_assign(k, std::move(r));
    }
    std::vector<JobResult> take_ready() { return streaming_ ? take_ready_streaming() : take_ready_ordered(); }

    std::string deduplicate_text(const std::string& text) const {  // queue/mod.rs:250-274
        std::vector<std::string> words;
        std::istringstream is(text);
        for (std::string w; is >> w;) words.push_back(w);
        if (words.empty()) return text;
        size_t skip = 0;
        std::string prefix;
        for (size_t i = 1; i <= std::min<size_t>(words.size(), 10); ++i) {
            if (i > 1) prefix += ' ';
            prefix += words[i - 1];
            if (last_text_suffix_.find(prefix) != std::string::npos) skip = i;
        }
        if (skip == 0) return text;
        std::string out;
        for (size_t i = skip; i < words.size(); ++i) {
            if (i > skip) out += ' ';
            out += words[i];
        }
        return out;
    }
    void reset_dedup() { last_text_suffix_.clear(); }
    bool is_empty() const { return pending_.empty() && completed_.empty(); }
    size_t pending_count() const { return pending_.size(); }
    size_t waiting_count() const { return completed_.size(); }
    QueueStats stats() const { return {pending_.size(), completed_.size()}; }
    bool is_pending(uint64_t s, uint32_t c) const { return pending_.count({s, c}) != 0; }
    const std::string& last_text_suffix() const { return last_text_suffix_; }

  private:
    std::vector<JobResult> take_ready_streaming() {  // queue/mod.rs:213-233
        std::vector<JobResult> ready;
        for (auto& kv : completed_) ready.push_back(std::move(kv.second));  // std::map iterates in key order
        completed_.clear();
        for (JobResult& r : ready) {
            if (!last_text_suffix_.empty() && !r.text.empty()) r.text = deduplicate_text(r.text);
            if (r.text.size() > 10) {  // last 50 bytes; the cut moves forward to a UTF-8 boundary (the reference would panic inside a character)
                size_t start = r.text.size() > 50 ? r.text.size() - 50 : 0;
                while (start < r.text.size() && (static_cast<unsigned char>(r.text[start]) & 0xC0) == 0x80) ++start;
                last_text_suffix_ = r.text.substr(start);
            }
        }
        return ready;
    }
    std::vector<JobResult> take_ready_ordered() {  // queue/mod.rs:236-247
        std::vector<JobResult> ready;
        for (auto it = completed_.find({next_output_id_, 0}); it != completed_.end(); it = completed_.find({next_output_id_, 0})) {
            ready.push_back(std::move(it->second));
            completed_.erase(it);
            ++next_output_id_;
        }
        return ready;
    }

    std::set<ChunkKey> pending_;
    std::map<ChunkKey, JobResult> completed_;
    uint64_t next_output_id_ = 0;
    bool streaming_;
    std::string last_text_suffix_;
};

// ---- timed output ---------------------------------------------------------------------------
// Timed segments of one recording on a common clock: `results` are the JobResults of ONE sequence; chunk k starts where
// chunk k-1 ended unless chunk_offsets_cs has an entry for it (overlapping capture, recording.rs:32).  10 ms units.
inline std::vector<TimedSegment> merge_timed_segments(std::vector<JobResult> results, const std::map<uint32_t, int64_t>* chunk_offsets_cs = nullptr) {
    std::sort(results.begin(), results.end(), [](const JobResult& a, const JobResult& b) { return a.chunk_id < b.chunk_id; });
    std::vector<TimedSegment> out;
    int64_t clock = 0;
    for (const JobResult& r : results) {
        int64_t base = clock;
        if (chunk_offsets_cs) {
            auto it = chunk_offsets_cs->find(r.chunk_id);
            if (it != chunk_offsets_cs->end()) base = it->second;
        }
        for (const TimedSegment& s : r.segments) out.push_back(TimedSegment{base + s.t0, base + s.t1, s.text});
        clock = base + (int64_t)std::llround((double)r.duration_secs * 100.0);
    }
    std::stable_sort(out.begin(), out.end(), [](const TimedSegment& a, const TimedSegment& b) { return a.t0 != b.t0 ? a.t0 < b.t0 : a.t1 < b.t1; });
    return out;
}

// recording.rs:167-194: `secs as u32` / `(secs * 1000.0) as u32` on f32 (truncating, saturating at 0)
inline std::string format_timestamp(float secs) {
    const uint32_t total = secs > 0.0f ? (uint32_t)secs : 0u;
    char buf[32];
    snprintf(buf, sizeof buf, "%02u:%02u:%02u", total / 3600, (total % 3600) / 60, total % 60);
    return buf;
}
inline std::string format_ms_timestamp(float secs, char sep) {
    const float v = secs * 1000.0f;
    const uint32_t ms = v > 0.0f ? (uint32_t)v : 0u;
    char buf[40];
    snprintf(buf, sizeof buf, "%02u:%02u:%02u%c%03u", ms / 3600000, (ms % 3600000) / 60000, (ms % 60000) / 1000, sep, ms % 1000);
    return buf;
}
inline std::string format_srt_timestamp(float secs) { return format_ms_timestamp(secs, ','); }
inline std::string format_vtt_timestamp(float secs) { return format_ms_timestamp(secs, '.'); }

// recording.rs:117-165.  The reference fills start / end from the capture loop's wall clock (recording.rs:385-392); here
// they come from the model's timestamps through merge_timed_segments.
struct TranscribedSegment {
    float start_secs = 0.0f, end_secs = 0.0f;
    std::string text;
    std::optional<uint32_t> speaker_id;
    std::string spoken() const { return speaker_id ? "<v Speaker " + std::to_string(*speaker_id) + ">" + text : text; }
    std::string format_timestamped() const {
        const std::string start = format_timestamp(start_secs);
        return speaker_id ? "[" + start + "] Speaker " + std::to_string(*speaker_id) + ": " + text : "[" + start + "] " + text;
    }
    std::string format_srt(size_t index) const {
        return std::to_string(index) + "\n" + format_srt_timestamp(start_secs) + " --> " + format_srt_timestamp(end_secs) + "\n" + spoken() + "\n";
    }
    std::string format_vtt() const { return format_vtt_timestamp(start_secs) + " --> " + format_vtt_timestamp(end_secs) + "\n" + spoken() + "\n"; }
};

inline std::vector<TranscribedSegment> transcribed_segments(const std::vector<TimedSegment>& merged) {
    std::vector<TranscribedSegment> out;
    for (const TimedSegment& s : merged) {
        const size_t a = s.text.find_first_not_of(" \t\r\n"), b = s.text.find_last_not_of(" \t\r\n");
        if (a == std::string::npos) continue;  // the capture loop stores no empty segment (recording.rs:386)
        TranscribedSegment t;
        t.start_secs = (float)s.t0 / 100.0f;
        t.end_secs = (float)s.t1 / 100.0f;
        t.text = s.text.substr(a, b - a + 1);
        out.push_back(std::move(t));
    }
    return out;
}

enum class OutputFormat { Text, Timestamped, Srt, Vtt };  // recording.rs:72-101
inline std::optional<OutputFormat> output_format_from_str(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (s == "text" || s == "txt") return OutputFormat::Text;
    if (s == "timestamped" || s == "ts") return OutputFormat::Timestamped;
    if (s == "srt" || s == "subrip") return OutputFormat::Srt;
    if (s == "vtt" || s == "webvtt") return OutputFormat::Vtt;
    return std::nullopt;
}
// what RecordingSession::save_to_file writes (recording.rs:505-524)
inline std::string render_transcript(const std::vector<TranscribedSegment>& segments, OutputFormat format) {
    std::string out = format == OutputFormat::Vtt ? "WEBVTT\n\n" : "";
    for (size_t i = 0; i < segments.size(); ++i) {
        switch (format) {
            case OutputFormat::Text: out += segments[i].text + "\n"; break;
            case OutputFormat::Timestamped: out += segments[i].format_timestamped() + "\n"; break;
            case OutputFormat::Srt: out += segments[i].format_srt(i + 1); break;
            case OutputFormat::Vtt: out += segments[i].format_vtt(); break;
        }
    }
    return out;
}

}  // namespace openhush
