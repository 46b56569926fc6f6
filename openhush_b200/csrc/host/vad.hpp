// C++ host mirror of the reference's VAD boundary: the `VadEngine` trait, `VadConfig`, `VadResult`,
// `SpeechSegment` and the streaming speech-segment state machine `VadState` (/root/reference/src/vad/mod.rs:34-245), plus the
// cut of a recording into transcription jobs that the daemon's VAD path makes (/root/reference/src/daemon.rs:2062-2138).
// The reference's engine behind the trait is Silero VAD (an ONNX model, /root/reference/src/vad/silero.rs:1-9; its weights
// are not in the reference tree): it is NOT restated.  `DeviceEnergyVad` is a stand-in engine over the library's per-chunk
// level kernel (oh_chunk_rms_db) that drives the same state machine; an energy detector, not a model of Silero.
// Header-only, depends on engine.hpp (and through it on include/oh_whisper.h).  No arithmetic of the hot path lives here.
#pragma once
#include <cmath>
#include <memory>
#include <optional>
#include <vector>

#include "engine.hpp"

namespace openhush {

struct VadConfig {  // vad/mod.rs:58-103 (defaults :68-102)
    bool enabled = false;
    float threshold = 0.5f;
    uint32_t min_silence_ms = 700, min_speech_ms = 250, speech_pad_ms = 30;
};
struct VadResult {
    float probability;
    bool is_speech;
};
struct SpeechSegment {  // vad/mod.rs:106-116
    size_t start, end;
    float avg_probability;
};

// vad/mod.rs:34-55
class VadEngine {
  public:
    virtual ~VadEngine() = default;
    virtual VadResult process(const float* samples, size_t n) = 0;
    virtual void reset() {}
    virtual size_t chunk_size() const { return 512; }  // 32 ms at 16 kHz, Silero's chunk
    virtual uint32_t sample_rate() const { return 16000; }
};

// vad/mod.rs:121-245: speech starts at the first speech chunk, ends after min_silence_ms of non-speech (at the position where
// that silence was noticed), and is reported only if it lasted min_speech_ms
class VadState {
  public:
    explicit VadState(VadConfig config, uint32_t sample_rate = 16000) : config_(config), sample_rate_(sample_rate) {}
    std::optional<SpeechSegment> update(VadResult result, size_t chunk_samples) {
        probabilities_.push_back(result.probability);
        const size_t prev_total = total_samples_;
        total_samples_ += chunk_samples;
        const size_t min_silence = (size_t)((float)config_.min_silence_ms / 1000.0f * (float)sample_rate_);
        const size_t min_speech = (size_t)((float)config_.min_speech_ms / 1000.0f * (float)sample_rate_);
        if (result.is_speech) {
            silence_samples_ = 0;
            if (!in_speech_) {
                in_speech_ = true;
                speech_start_ = prev_total;
            }
            return std::nullopt;
        }
        silence_samples_ += chunk_samples;
        if (in_speech_ && silence_samples_ >= min_silence) {
            in_speech_ = false;
            const size_t start = speech_start_.value_or(0), end = prev_total;
            speech_start_.reset();
            std::optional<SpeechSegment> out;
            if (end - start >= min_speech) {
                float sum = 0.0f;
                for (float p : probabilities_) sum += p;
                out = SpeechSegment{start, end, probabilities_.empty() ? 0.0f : sum / (float)probabilities_.size()};
            }
            probabilities_.clear();
            return out;
        }
        return std::nullopt;
    }
    bool is_speech() const { return in_speech_; }
    size_t total_samples() const { return total_samples_; }
    void reset() {
        probabilities_.clear();
        in_speech_ = false;
        speech_start_.reset();
        silence_samples_ = 0;
        total_samples_ = 0;
    }

  private:
    VadConfig config_;
    uint32_t sample_rate_;
    std::vector<float> probabilities_;
    bool in_speech_ = false;
    std::optional<size_t> speech_start_;
    size_t silence_samples_ = 0, total_samples_ = 0;
};

// The daemon's VAD path over a whole recording (daemon.rs:2062-2138): chunk probabilities -> VadState -> one job per finished
// speech segment, from where its speech started to the position at which the segment's end was noticed; chunk ids count
// the jobs of the sequence.
inline std::vector<TranscriptionJob> speech_jobs(const std::vector<float>& samples, const std::vector<float>& probabilities, const VadConfig& config,
                                                 uint64_t sequence_id = 0, size_t chunk_size = 512, uint32_t sample_rate = 16000) {
    VadState state(config, sample_rate);
    std::vector<TranscriptionJob> jobs;
    std::optional<size_t> speech_start_pos;
    size_t pos = 0;
    for (float p : probabilities) {
        const VadResult r{p, p >= config.threshold};
        if (r.is_speech && !speech_start_pos) speech_start_pos = pos;
        const auto seg = state.update(r, chunk_size);
        const size_t cur = pos + chunk_size;
        if (seg) {
            const size_t start = speech_start_pos.value_or(seg->start), end = std::min(cur, samples.size());
            speech_start_pos.reset();
            TranscriptionJob job;
            job.buffer.samples.assign(samples.begin() + (std::ptrdiff_t)std::min(start, end), samples.begin() + (std::ptrdiff_t)end);
            job.buffer.sample_rate = sample_rate;
            job.sequence_id = sequence_id;
            job.chunk_id = (uint32_t)jobs.size();
            job.is_final = false;
            jobs.push_back(std::move(job));
        }
        pos = cur;
    }
    return jobs;
}

// Stand-in engine (see the header comment): RMS level of every 512-sample chunk of a recording from ONE kernel launch,
// mapped to a speech probability by a logistic curve around level_db.
class DeviceEnergyVad : public VadEngine {
  public:
    DeviceEnergyVad(oh_context* ctx, oh_state* state, VadConfig config, float level_db = -40.0f, float slope_db = 3.0f)
        : ctx_(ctx), state_(state), config_(config), level_db_(level_db), slope_db_(slope_db) {}
    std::vector<float> probabilities(const float* samples, size_t n) const {
        const int n_chunks = (int)(n / chunk_size());
        std::vector<float> lv((size_t)std::max(n_chunks, 0));
        if (n_chunks > 0 && oh_chunk_rms_db(ctx_, state_, samples, (long long)n, (int)chunk_size(), lv.data(), n_chunks) != OH_OK) return {};
        for (float& v : lv) v = 1.0f / (1.0f + std::exp(-(v - level_db_) / slope_db_));
        return lv;
    }
    VadResult process(const float* samples, size_t n) override {
        const auto p = probabilities(samples, n);
        const float v = p.empty() ? 0.0f : p[0];
        return VadResult{v, v >= config_.threshold};
    }

  private:
    oh_context* ctx_;
    oh_state* state_;
    VadConfig config_;
    float level_db_, slope_db_;
};

}  // namespace openhush
