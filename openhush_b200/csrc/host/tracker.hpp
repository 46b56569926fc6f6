// C++ host mirror of the reference's TranscriptionTracker (/root/reference/src/queue/mod.rs:44-298): the step right
// after the hot path — pending/completed bookkeeping with backpressure, streaming hand-out sorted by
// (sequence_id, chunk_id) with word-overlap de-duplication against the tail of the previous output, or strictly
// ordered hand-out by sequence id.  Same names, arguments and results as the reference.
#pragma once
#include <algorithm>
#include <map>
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

}  // namespace openhush
