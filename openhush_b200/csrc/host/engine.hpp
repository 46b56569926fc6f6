// C++ host side above the C ABI: the reference's engine boundary restated for a compiled-language caller
// (the reference is Rust; no Rust toolchain exists in this image — INTEGRATION.md has the Rust binding).
//
//   validate_audio / AudioValidationError      <- /root/reference/src/engine/validation.rs:9-136
//   WhisperError, TranscriptionResult, BenchmarkResult, WhisperEngine{create, transcribe, benchmark}
//                                               <- /root/reference/src/engine/whisper.rs:14-387
//   TranscriptionJob / JobResult / process_job  <- /root/reference/src/queue/mod.rs:18-43, src/queue/worker.rs:119-193
//
// Header-only, depends only on include/oh_whisper.h; link with libohwhisper.so.  No arithmetic lives here.
#pragma once
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <variant>
#include <vector>

#include "../../../include/oh_whisper.h"

namespace openhush {

constexpr float MAX_AUDIO_DURATION_SECS = 7200.0f;
constexpr float MIN_AUDIO_DURATION_SECS = 0.1f;
constexpr uint32_t EXPECTED_SAMPLE_RATE = 16000;

struct AudioValidationError {
    enum Kind { TooLong, TooShort, Empty, ContainsNaN, ContainsInfinite, InvalidSampleRate } kind;
    double a = 0, b = 0;  // payload: (secs, limit) | count | (got, expected)
    std::string message() const {
        char buf[128];
        switch (kind) {
            case TooLong: snprintf(buf, sizeof buf, "Audio too long: %.1fs exceeds maximum %.1fs", a, b); break;
            case TooShort: snprintf(buf, sizeof buf, "Audio too short: %.3fs below minimum %.3fs", a, b); break;
            case Empty: snprintf(buf, sizeof buf, "Audio is empty (no samples)"); break;
            case ContainsNaN: snprintf(buf, sizeof buf, "Audio contains %zu NaN values", (size_t)a); break;
            case ContainsInfinite: snprintf(buf, sizeof buf, "Audio contains %zu infinite values", (size_t)a); break;
            default: snprintf(buf, sizeof buf, "Unexpected sample rate: %uHz (expected %uHz)", (unsigned)a, (unsigned)b); break;
        }
        return buf;
    }
};
struct AudioValidationInfo {
    float duration_secs;
    size_t sample_count;
    float min_value, max_value, rms;
};

// same checks, same order as validation.rs:46-118
inline std::variant<AudioValidationInfo, AudioValidationError> validate_audio(const float* samples, size_t n, uint32_t sample_rate) {
    using E = AudioValidationError;
    if (n == 0) return E{E::Empty};
    if (sample_rate != EXPECTED_SAMPLE_RATE) return E{E::InvalidSampleRate, (double)sample_rate, (double)EXPECTED_SAMPLE_RATE};
    const float duration = (float)n / (float)sample_rate;
    if (duration > MAX_AUDIO_DURATION_SECS) return E{E::TooLong, duration, MAX_AUDIO_DURATION_SECS};
    if (duration < MIN_AUDIO_DURATION_SECS) return E{E::TooShort, duration, MIN_AUDIO_DURATION_SECS};
    size_t n_nan = 0, n_inf = 0;
    float mn = 3.402823466e38f, mx = -3.402823466e38f;
    double ss = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const float s = samples[i];
        if (std::isnan(s)) ++n_nan;
        else if (std::isinf(s)) ++n_inf;
        else {
            mn = s < mn ? s : mn;
            mx = s > mx ? s : mx;
            ss += (double)s * (double)s;
        }
    }
    if (n_nan) return E{E::ContainsNaN, (double)n_nan};
    if (n_inf) return E{E::ContainsInfinite, (double)n_inf};
    return AudioValidationInfo{duration, n, mn, mx, (float)std::sqrt(ss / (double)n)};
}

struct WhisperError {
    enum Kind { ModelNotFound, LoadFailed, TranscriptionFailed, ValidationFailed } kind;
    std::string detail;
    std::string message() const {
        switch (kind) {
            case ModelNotFound: return "Model not found at " + detail;
            case LoadFailed: return "Failed to load model: " + detail;
            case TranscriptionFailed: return "Transcription failed: " + detail;
            default: return "Audio validation failed: " + detail;
        }
    }
};
struct TimedSegment {  // additive: the model's own segment times, 10 ms units, relative to the clip
    int64_t t0 = 0, t1 = 0;
    std::string text;
};
struct TranscriptionResult {
    std::string text, language;
    uint64_t duration_ms = 0;
    std::vector<TimedSegment> segments;  // additive
};
struct BenchmarkResult {
    float overhead_secs, recommended_chunk_interval, test_audio_secs;
};
struct AudioBuffer {
    std::vector<float> samples;
    uint32_t sample_rate = 16000;
    float duration_secs() const { return (float)samples.size() / (float)sample_rate; }
};

// The fields of AudioConfig the worker's preprocess_audio reads (config.rs:956-975; defaults config.rs:984, 1129-1159,
// 1239-1269: preprocessing off, its three stages on).  RNNoise noise reduction is a third-party model and not provided.
struct AudioConfig {
    bool preprocessing = false;
    oh_preprocess_params levels = oh_preprocess_default_params();
};

template <class T>
using Result = std::variant<T, WhisperError>;

class WhisperEngine {
  public:
    // whisper.rs:129-179.  The engine may be created on one thread and used on another (daemon.rs:1803-1805): the
    // library re-selects its CUDA device at every entry point.
    static Result<std::unique_ptr<WhisperEngine>> create(const std::string& model_path, const std::string& language, bool translate,
                                                         bool use_gpu, int gpu_device = 0, int max_batch = 1) {
        struct stat sb;
        if (stat(model_path.c_str(), &sb) != 0) {
            std::string name = model_path.substr(model_path.find_last_of('/') + 1);
            if (name.rfind("ggml-", 0) == 0) name = name.substr(5);
            if (name.size() > 4 && name.substr(name.size() - 4) == ".bin") name.resize(name.size() - 4);
            return WhisperError{WhisperError::ModelNotFound, model_path + ". Run 'openhush model download " + name + "'"};
        }
        oh_context_params cp = oh_context_default_params();
        cp.use_gpu = use_gpu ? 1 : 0;
        cp.gpu_device = gpu_device;
        cp.max_batch = max_batch;
        oh_context* ctx = oh_init_from_file(model_path.c_str(), cp);
        if (!ctx) return WhisperError{WhisperError::LoadFailed, oh_last_error()};
        oh_state* st = oh_init_state(ctx);
        if (!st) {
            oh_free(ctx);
            return WhisperError{WhisperError::LoadFailed, oh_last_error()};
        }
        auto e = std::unique_ptr<WhisperEngine>(new WhisperEngine());
        e->ctx_ = ctx;
        e->state_ = st;
        e->language_ = language;
        e->translate_ = translate;
        e->max_batch_ = max_batch < 1 ? 1 : max_batch;
        return e;
    }
    ~WhisperEngine() {
        oh_free_state(state_);
        oh_free(ctx_);
    }
    // whisper.rs:204-310
    Result<TranscriptionResult> transcribe(const AudioBuffer& audio) const {
        auto v = validate_audio(audio.samples.data(), audio.samples.size(), audio.sample_rate);
        if (auto* err = std::get_if<AudioValidationError>(&v)) return WhisperError{WhisperError::ValidationFailed, err->message()};
        const auto t0 = std::chrono::steady_clock::now();
        oh_full_params p = oh_full_default_params(OH_SAMPLING_GREEDY);
        p.greedy_best_of = 1;
        if (language_ != "auto") p.language = language_.c_str();
        p.translate = translate_ ? 0 : 1;  // whisper.rs:251-257 passes the inverted flag
        const int rc = oh_full(ctx_, state_, p, audio.samples.data(), (int)audio.samples.size());
        if (rc != 0) {
            const char* name = rc == OH_ERR_MEL ? "UnableToCalculateSpectrogram" : rc == OH_ERR_ENCODE ? "FailedToEncode" : rc == OH_ERR_DECODE ? "FailedToDecode" : nullptr;
            return WhisperError{WhisperError::TranscriptionFailed, name ? std::string(name) : "GenericError(" + std::to_string(rc) + ")"};
        }
        std::string text;
        TranscriptionResult r;
        const int n = oh_full_n_segments(state_);
        for (int i = 0; i < n; ++i) {
            const char* s = oh_full_get_segment_text(state_, i);
            if (!s || !valid_utf8(s)) continue;  // whisper.rs:276-280: invalid UTF-8 segments are skipped
            text += s;
            r.segments.push_back(TimedSegment{oh_full_get_segment_t0(state_, i), oh_full_get_segment_t1(state_, i), s});
        }
        r.text = trim(text);
        r.language = language_ == "auto" ? "en" : language_;
        r.duration_ms = (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
        return r;
    }
    // additive (the multi-GPU queue): n independent transcribe() calls as ONE oh_full_batch; one entry
    // per job, the error a job would have got from transcribe() stays with that job alone
    std::vector<Result<TranscriptionResult>> transcribe_batch(const std::vector<const AudioBuffer*>& audios) const {
        std::vector<Result<TranscriptionResult>> out(audios.size(), WhisperError{WhisperError::TranscriptionFailed, "not run"});
        std::vector<const float*> ptrs;
        std::vector<int> lens;
        std::vector<size_t> idx;
        for (size_t i = 0; i < audios.size(); ++i) {
            auto v = validate_audio(audios[i]->samples.data(), audios[i]->samples.size(), audios[i]->sample_rate);
            if (auto* err = std::get_if<AudioValidationError>(&v)) {
                out[i] = WhisperError{WhisperError::ValidationFailed, err->message()};
                continue;
            }
            ptrs.push_back(audios[i]->samples.data());
            lens.push_back((int)audios[i]->samples.size());
            idx.push_back(i);
        }
        if (ptrs.empty()) return out;
        const auto t0 = std::chrono::steady_clock::now();
        oh_full_params p = oh_full_default_params(OH_SAMPLING_GREEDY);
        p.greedy_best_of = 1;
        if (language_ != "auto") p.language = language_.c_str();
        p.translate = translate_ ? 0 : 1;
        const int rc = oh_full_batch(ctx_, state_, p, ptrs.data(), lens.data(), (int)ptrs.size());
        fill_results(out, idx, rc, t0);
        return out;
    }
    // Stream form of transcribe_batch (oh_stream_submit / oh_stream_wait): submit_batch registers the jobs and returns at
    // once — they start sharing the decoder rows with the jobs of earlier tickets — and collect() waits for them and gives
    // what transcribe_batch would have.  Tickets are collected in the order they were submitted.
    struct Pending {
        int ticket = -1;
        std::vector<Result<TranscriptionResult>> out;
        std::vector<size_t> idx;
        std::chrono::steady_clock::time_point t0;
    };
    Pending submit_batch(const std::vector<const AudioBuffer*>& audios) const {
        Pending pd;
        pd.out.resize(audios.size());
        std::vector<const float*> ptrs;
        std::vector<int> lens;
        for (size_t i = 0; i < audios.size(); ++i) {
            auto v = validate_audio(audios[i]->samples.data(), audios[i]->samples.size(), audios[i]->sample_rate);
            if (auto* err = std::get_if<AudioValidationError>(&v)) {
                pd.out[i] = WhisperError{WhisperError::ValidationFailed, err->message()};
                continue;
            }
            ptrs.push_back(audios[i]->samples.data());
            lens.push_back((int)audios[i]->samples.size());
            pd.idx.push_back(i);
        }
        pd.t0 = std::chrono::steady_clock::now();
        if (ptrs.empty()) return pd;
        oh_full_params p = oh_full_default_params(OH_SAMPLING_GREEDY);
        p.greedy_best_of = 1;
        if (language_ != "auto") p.language = language_.c_str();
        p.translate = translate_ ? 0 : 1;
        int ticket = -1;
        const int rc = oh_stream_submit(ctx_, state_, p, ptrs.data(), lens.data(), (int)ptrs.size(), 0, &ticket);
        if (rc != 0) {
            for (size_t k : pd.idx) pd.out[k] = WhisperError{WhisperError::TranscriptionFailed, "GenericError(" + std::to_string(rc) + ")"};
            pd.idx.clear();
            return pd;
        }
        pd.ticket = ticket;
        return pd;
    }
    std::vector<Result<TranscriptionResult>> collect(Pending& pd) const {
        if (pd.ticket >= 0) fill_results(pd.out, pd.idx, oh_stream_wait(ctx_, state_, pd.ticket), pd.t0);
        pd.ticket = -1;
        return std::move(pd.out);
    }
    int max_batch() const { return max_batch_; }
    // worker.rs:196-240 for the buffers of the jobs a worker drained: normalise / compress / limit on the device, in place,
    // one oh_preprocess_pcm call for all of them.  Nothing runs unless config.preprocessing.
    bool preprocess_audio(const std::vector<AudioBuffer*>& buffers, const AudioConfig& config) const {
        if (!config.preprocessing || buffers.empty()) return true;
        std::vector<float*> ptrs;
        std::vector<int> ns;
        for (AudioBuffer* b : buffers) {
            ptrs.push_back(b->samples.data());
            ns.push_back((int)b->samples.size());
        }
        return oh_preprocess_pcm(ctx_, state_, ptrs.data(), ns.data(), (int)buffers.size(), &config.levels, nullptr) == OH_OK;
    }
    // whisper.rs:334-387
    BenchmarkResult benchmark(float safety_margin) const {
        AudioBuffer a;
        a.samples.assign(32000, 0.0f);
        (void)transcribe(a);
        uint64_t total = 0;
        for (int i = 0; i < 3; ++i) {
            const auto t0 = std::chrono::steady_clock::now();
            (void)transcribe(a);
            total += (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
        }
        const float overhead = (float)(total / 3) / 1000.0f;
        return BenchmarkResult{overhead, overhead * (1.0f + safety_margin), 2.0f};
    }

  private:
    WhisperEngine() = default;
    static bool valid_utf8(const char* s) {
        const unsigned char* p = (const unsigned char*)s;
        while (*p) {
            int n = *p < 0x80 ? 0 : (*p >> 5) == 6 ? 1 : (*p >> 4) == 14 ? 2 : (*p >> 3) == 30 ? 3 : -1;
            if (n < 0) return false;
            ++p;
            for (; n > 0; --n, ++p)
                if ((*p >> 6) != 2) return false;
        }
        return true;
    }
    static std::string trim(const std::string& s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    // per-job results of the batched call that just completed on the state (rc: the call's own return code)
    void fill_results(std::vector<Result<TranscriptionResult>>& out, const std::vector<size_t>& idx, int rc,
                      std::chrono::steady_clock::time_point t0) const {
        const uint64_t ms = (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
        for (size_t k = 0; k < idx.size(); ++k) {
            const int item_rc = rc != 0 ? rc : oh_full_item_rc(state_, (int)k);
            if (item_rc != 0) {
                const char* name = item_rc == OH_ERR_MEL ? "UnableToCalculateSpectrogram" : item_rc == OH_ERR_ENCODE ? "FailedToEncode" : item_rc == OH_ERR_DECODE ? "FailedToDecode" : nullptr;
                out[idx[k]] = WhisperError{WhisperError::TranscriptionFailed, name ? std::string(name) : "GenericError(" + std::to_string(item_rc) + ")"};
                continue;
            }
            std::string text;
            TranscriptionResult r;
            const int n = oh_full_n_segments_item(state_, (int)k);
            for (int i = 0; i < n; ++i) {
                const char* sgm = oh_full_get_segment_text_item(state_, (int)k, i);
                if (!sgm || !valid_utf8(sgm)) continue;
                text += sgm;
                r.segments.push_back(TimedSegment{oh_full_get_segment_t0_item(state_, (int)k, i), oh_full_get_segment_t1_item(state_, (int)k, i), sgm});
            }
            r.text = trim(text);
            r.language = language_ == "auto" ? "en" : language_;
            r.duration_ms = ms;
            out[idx[k]] = r;
        }
    }
    oh_context* ctx_ = nullptr;
    oh_state* state_ = nullptr;
    std::string language_;
    bool translate_ = false;
    int max_batch_ = 1;
};

// queue/mod.rs:18-43 and the hot-path caller queue/worker.rs:119-193
struct TranscriptionJob {
    AudioBuffer buffer;
    uint64_t sequence_id = 0;
    uint32_t chunk_id = 0;
    bool is_final = true;
};
struct JobResult {
    std::string text;
    uint64_t sequence_id;
    uint32_t chunk_id;
    bool is_final;
    float duration_secs;
    std::vector<TimedSegment> segments;  // additive: (t0, t1, text) of the chunk, 10 ms units, chunk-relative
};
inline JobResult process_job(const WhisperEngine* engine, const TranscriptionJob& job, const AudioConfig* audio_config = nullptr) {
    JobResult r{"", job.sequence_id, job.chunk_id, job.is_final, job.buffer.duration_secs()};
    if (!engine) {  // worker.rs:125-142: no engine -> empty result with duration 0 (:136), the tracker never stalls
        r.duration_secs = 0.0f;
        return r;
    }
    AudioBuffer buffer = job.buffer;  // worker.rs:150-154: the job's buffer goes through preprocess_audio first
    if (audio_config && !engine->preprocess_audio({&buffer}, *audio_config)) return r;
    auto res = engine->transcribe(buffer);
    if (auto* ok = std::get_if<TranscriptionResult>(&res)) {  // worker.rs:159-165: error -> logged + empty text
        r.text = ok->text;
        r.segments = ok->segments;
    }
    return r;
}


// One worker thread per GPU over ONE shared job queue (the reference's documented multi-GPU design: config.gpu.devices,
// /root/reference/src/config.rs:921-929; spawn_worker, /root/reference/src/queue/worker.rs:255-267).  A worker drains up to its
// engine's max_batch queued jobs and hands them to the engine as one oh_full_batch; results go to one queue, a
// TranscriptionTracker restores the order.  A worker without an engine takes no jobs while another worker has one; with no
// engine at all one thread answers every job with an empty text (worker.rs:125-142).  Nothing is exchanged between GPUs.
class WorkerPool {
  public:
    explicit WorkerPool(std::vector<std::unique_ptr<WhisperEngine>> engines) : engines_(std::move(engines)), jobs_done_(engines_.size(), 0) {
        bool any = false;
        for (auto& e : engines_) any |= (bool)e;
        for (size_t i = 0; i < engines_.size(); ++i)
            if (engines_[i] || (!any && i == 0)) threads_.emplace_back([this, i] { run(i); });
    }
    ~WorkerPool() { shutdown(); }
    void submit(TranscriptionJob job) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            jobs_.push_back(std::move(job));
        }
        cv_jobs_.notify_one();
    }
    // blocks until n results are available (in completion order)
    std::vector<JobResult> drain(size_t n) {
        std::vector<JobResult> out;
        std::unique_lock<std::mutex> lk(mu_);
        while (out.size() < n) {
            cv_results_.wait(lk, [&] { return !results_.empty(); });
            while (!results_.empty() && out.size() < n) {
                out.push_back(std::move(results_.front()));
                results_.pop_front();
            }
        }
        return out;
    }
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_jobs_.notify_all();
        for (auto& t : threads_)
            if (t.joinable()) t.join();
        threads_.clear();
    }
    const std::vector<size_t>& jobs_done() const { return jobs_done_; }

  private:
    // Drain up to max_batch jobs, hand them to the engine's stream, answer.  A worker keeps one drain in flight while it
    // collects the previous one (oh_stream_submit / oh_stream_wait), so its decoder rows stay busy across drains; it blocks
    // on the queue only when nothing is in flight.
    void run(size_t i) {
        const WhisperEngine* eng = engines_[i].get();
        const size_t limit = eng ? (size_t)eng->max_batch() : 1, depth = 1;
        std::deque<std::pair<std::vector<TranscriptionJob>, WhisperEngine::Pending>> inflight;
        auto answer = [&](std::vector<TranscriptionJob>& batch, const std::vector<Result<TranscriptionResult>>* out) {
            std::vector<JobResult> res;
            for (auto& j : batch) res.push_back(JobResult{"", j.sequence_id, j.chunk_id, j.is_final, eng ? j.buffer.duration_secs() : 0.0f});
            if (out)
                for (size_t k = 0; k < out->size() && k < res.size(); ++k)
                    if (auto* ok = std::get_if<TranscriptionResult>(&(*out)[k])) {  // error -> empty text (worker.rs:159-165)
                        res[k].text = ok->text;
                        res[k].segments = ok->segments;
                    }
            {
                std::lock_guard<std::mutex> lk(mu_);
                for (auto& r : res) results_.push_back(std::move(r));
                jobs_done_[i] += batch.size();
            }
            cv_results_.notify_all();
        };
        while (true) {
            std::vector<TranscriptionJob> batch;
            bool stopping = false;
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (inflight.empty()) cv_jobs_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (inflight.size() <= depth)
                    while (!jobs_.empty() && batch.size() < limit) {
                        batch.push_back(std::move(jobs_.front()));
                        jobs_.pop_front();
                    }
                stopping = stop_ && jobs_.empty();
            }
            if (!batch.empty()) {
                if (!eng) {
                    answer(batch, nullptr);
                    continue;
                }
                inflight.emplace_back(std::move(batch), WhisperEngine::Pending{});
                std::vector<const AudioBuffer*> bufs;
                for (auto& j : inflight.back().first) bufs.push_back(&j.buffer);
                inflight.back().second = eng->submit_batch(bufs);
                continue;
            }
            if (!inflight.empty()) {
                auto out = eng->collect(inflight.front().second);
                answer(inflight.front().first, &out);
                inflight.pop_front();
                continue;
            }
            if (stopping) return;
        }
    }
    std::vector<std::unique_ptr<WhisperEngine>> engines_;
    std::vector<size_t> jobs_done_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_jobs_, cv_results_;
    std::deque<TranscriptionJob> jobs_;
    std::deque<JobResult> results_;
    bool stop_ = false;
};

}  // namespace openhush
