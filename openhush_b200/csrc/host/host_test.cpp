// Native test driver of the C++ host mirror.  `host_test` alone: CPU checks (validation order / messages, missing model,
// no-GPU load failure, worker without engine).  `host_test <model> <n_samples>`: additionally transcribes silence and a
// ramp on the GPU and prints the text length (used by the -m gpu tests).
#include <cstdlib>
#include <cstring>

#include "engine.hpp"
#include "tracker.hpp"
#include "vad.hpp"

using namespace openhush;
#define CHECK(c)                                                          \
    do {                                                                  \
        if (!(c)) {                                                       \
            fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #c); \
            return 1;                                                     \
        }                                                                 \
    } while (0)

int main(int argc, char** argv) {
    {
        std::vector<float> x(16000, 0.5f);
        auto v = validate_audio(x.data(), x.size(), 16000);
        auto* info = std::get_if<AudioValidationInfo>(&v);
        CHECK(info && std::fabs(info->rms - 0.5f) < 1e-3f && info->sample_count == 16000);
        CHECK(std::get<AudioValidationError>(validate_audio(x.data(), 0, 16000)).kind == AudioValidationError::Empty);
        CHECK(std::get<AudioValidationError>(validate_audio(x.data(), 0, 8000)).kind == AudioValidationError::Empty);
        CHECK(std::get<AudioValidationError>(validate_audio(x.data(), 800, 16000)).kind == AudioValidationError::TooShort);
        auto e = std::get<AudioValidationError>(validate_audio(x.data(), x.size(), 44100));
        CHECK(e.kind == AudioValidationError::InvalidSampleRate && e.message() == "Unexpected sample rate: 44100Hz (expected 16000Hz)");
        x[3] = NAN;
        x[4] = INFINITY;
        auto en = std::get<AudioValidationError>(validate_audio(x.data(), x.size(), 16000));
        CHECK(en.kind == AudioValidationError::ContainsNaN && en.message() == "Audio contains 1 NaN values");
        x[3] = 0.f;
        CHECK(std::get<AudioValidationError>(validate_audio(x.data(), x.size(), 16000)).kind == AudioValidationError::ContainsInfinite);
    }
    {
        auto r = WhisperEngine::create("/nonexistent/ggml-base.bin", "en", false, true);
        auto* err = std::get_if<WhisperError>(&r);
        CHECK(err && err->kind == WhisperError::ModelNotFound && err->message().find("openhush model download base") != std::string::npos);
        TranscriptionJob job;
        job.buffer.samples.assign(16000, 0.f);
        job.sequence_id = 9;
        JobResult jr = process_job(nullptr, job);
        CHECK(jr.text.empty() && jr.sequence_id == 9 && jr.duration_secs == 0.0f);  // worker.rs:136
    }
    {  // the reference's tracker unit tests (queue/mod.rs:324-466), same inputs and expected values
        auto res = [](uint64_t seq, uint32_t chunk, const char* text, bool fin) { return JobResult{text, seq, chunk, fin, 1.0f}; };
        TranscriptionTracker t;
        t.add_pending(0, 0);
        t.add_pending(0, 1);
        t.add_result(res(0, 1, "world", true));
        auto ready = t.take_ready();
        CHECK(ready.size() == 1 && ready[0].text == "world");
        t.add_result(res(0, 0, "hello", false));
        ready = t.take_ready();
        CHECK(ready.size() == 1 && ready[0].text == "hello");
        TranscriptionTracker o = TranscriptionTracker::new_ordered();
        o.add_pending(0, 0);
        o.add_pending(1, 0);
        o.add_result(res(1, 0, "second", true));
        CHECK(o.take_ready().empty());
        o.add_result(res(0, 0, "first", true));
        ready = o.take_ready();
        CHECK(ready.size() == 2 && ready[0].text == "first" && ready[1].text == "second");
        TranscriptionTracker d;
        d.add_result(res(0, 0, "hello world this is a test", false));
        CHECK(d.take_ready()[0].text == "hello world this is a test");
        d.add_result(res(0, 1, "is a test and more words", true));
        CHECK(d.take_ready()[0].text == "and more words");
        TranscriptionTracker b;
        for (uint32_t c = 0; c < 3; ++c) CHECK(b.add_pending_with_config(0, c, 3, 2, BackpressureStrategy::DropNewest));
        CHECK(!b.add_pending_with_config(0, 3, 3, 2, BackpressureStrategy::DropNewest) && b.pending_count() == 3);
        TranscriptionTracker b2;
        for (uint32_t c = 0; c < 3; ++c) CHECK(b2.add_pending_with_config(0, c, 3, 2, BackpressureStrategy::DropOldest));
        CHECK(b2.add_pending_with_config(0, 3, 3, 2, BackpressureStrategy::DropOldest) && b2.pending_count() == 3);
        CHECK(!b2.is_pending(0, 0) && b2.is_pending(0, 3));
        TranscriptionTracker b3;
        for (uint32_t c = 0; c < 4; ++c) CHECK(b3.add_pending_with_config(0, c, 3, 2, BackpressureStrategy::Warn));
        CHECK(b3.pending_count() == 4 && b3.stats().waiting_count == 0 && !b3.is_empty());
    }
    {  // timed output: the reference's formatter tests (recording.rs:528-590) and the model-timestamp route (N1)
        CHECK(format_timestamp(0.0f) == "00:00:00" && format_timestamp(61.0f) == "00:01:01" && format_timestamp(3661.0f) == "01:01:01");
        CHECK(format_srt_timestamp(0.0f) == "00:00:00,000" && format_srt_timestamp(1.5f) == "00:00:01,500" && format_srt_timestamp(3661.123f) == "01:01:01,123");
        CHECK(format_vtt_timestamp(0.0f) == "00:00:00.000" && format_vtt_timestamp(1.5f) == "00:00:01.500");
        CHECK(output_format_from_str("text") == OutputFormat::Text && output_format_from_str("srt") == OutputFormat::Srt);
        CHECK(output_format_from_str("VTT") == OutputFormat::Vtt && !output_format_from_str("invalid"));
        TranscribedSegment seg;
        seg.end_secs = 5.0f;
        seg.text = "Hello world";
        CHECK(seg.format_timestamped().find("[00:00:00]") != std::string::npos);
        CHECK(seg.format_srt(1).find("00:00:00,000 --> 00:00:05,000") != std::string::npos);
        CHECK(seg.format_vtt().find("00:00:00.000 --> 00:00:05.000") != std::string::npos);
        JobResult a{"", 0, 0, false, 30.0f, {{0, 250, " first words"}, {250, 2900, " and more"}}};
        JobResult b{"", 0, 1, true, 30.0f, {{10, 400, " next window"}, {400, 400, "  "}}};
        auto segs = transcribed_segments(merge_timed_segments({b, a}));
        CHECK(segs.size() == 3 && segs[0].text == "first words" && segs[1].start_secs == 2.5f && segs[1].end_secs == 29.0f);
        CHECK(segs[2].text == "next window" && std::fabs(segs[2].start_secs - 30.1f) < 1e-4f && segs[2].end_secs == 34.0f);
        CHECK(render_transcript(segs, OutputFormat::Srt).rfind("1\n00:00:00,000 --> 00:00:02,500\nfirst words\n2\n00:00:02,500 --> 00:00:29,000\n", 0) == 0);
        CHECK(render_transcript(segs, OutputFormat::Vtt).rfind("WEBVTT\n\n00:00:00.000 --> 00:00:02.500\nfirst words\n", 0) == 0);
        CHECK(render_transcript(segs, OutputFormat::Text) == "first words\nand more\nnext window\n");
        std::map<uint32_t, int64_t> off{{0, 0}, {1, 2950}};
        CHECK(merge_timed_segments({a, b}, &off)[2].t0 == 2960);
    }
    {  // VAD state machine: the reference's unit tests (vad/mod.rs:250-315), defaults, the daemon's job cut
        VadConfig c1;
        c1.min_silence_ms = 100;
        c1.min_speech_ms = 50;
        VadState s1(c1, 16000);
        const VadResult speech{0.8f, true}, silence{0.1f, false};
        CHECK(!s1.update(speech, 512) && s1.is_speech());
        CHECK(!s1.update(speech, 512));
        CHECK(!s1.update(silence, 512) && s1.is_speech());  // brief silence: not enough to end
        auto seg = s1.update(silence, 1600);                // more silence: the segment is returned
        CHECK(seg && !s1.is_speech() && seg->start == 0 && seg->end == 1536 && std::fabs(seg->avg_probability - 0.45f) < 1e-6f);
        VadConfig c2;
        c2.min_silence_ms = 100;
        c2.min_speech_ms = 500;
        VadState s2(c2, 16000);
        s2.update(speech, 512);
        CHECK(!s2.update(silence, 3200));                   // too short: dropped
        VadConfig d;
        CHECK(!d.enabled && d.threshold == 0.5f && d.min_silence_ms == 700 && d.min_speech_ms == 250 && d.speech_pad_ms == 30);
        VadState s3(d);
        s3.update(VadResult{0.9f, true}, 512);
        s3.reset();
        CHECK(!s3.is_speech() && s3.total_samples() == 0);
        // two bursts of speech separated by silence -> two jobs, each up to where the closing 700 ms of silence was noticed
        const size_t sr = 16000, ch = 512;
        std::vector<float> x(sr * 8, 0.0f), probs(x.size() / ch, 0.05f);
        auto burst = [&](double t0, double t1) {
            for (size_t k = (size_t)(t0 * sr) / ch; k < (size_t)(t1 * sr) / ch; ++k) probs[k] = 0.9f;
        };
        burst(0.5, 2.0);
        burst(3.5, 5.5);
        auto jobs = speech_jobs(x, probs, VadConfig());
        CHECK(jobs.size() == 2 && jobs[0].chunk_id == 0 && jobs[1].chunk_id == 1 && !jobs[0].is_final && !jobs[1].is_final);
        CHECK(std::fabs(jobs[0].buffer.duration_secs() - 2.2f) < 0.07f && std::fabs(jobs[1].buffer.duration_secs() - 2.7f) < 0.07f);
    }
    if (argc >= 2) {
        auto r = WhisperEngine::create(argv[1], "en", false, argc >= 4 ? atoi(argv[3]) != 0 : true);
        if (auto* err = std::get_if<WhisperError>(&r)) {
            printf("LOAD_FAILED %s\n", err->message().c_str());
            return argc >= 4 ? 0 : 2;
        }
        auto& eng = std::get<std::unique_ptr<WhisperEngine>>(r);
        AudioBuffer a;
        const int n = argc >= 3 ? atoi(argv[2]) : 32000;
        a.samples.resize(n);
        for (int i = 0; i < n; ++i) a.samples[i] = 0.1f * std::sin(0.05f * i) * (float)((i / 4000) % 2);
        auto tr = eng->transcribe(a);
        auto* ok = std::get_if<TranscriptionResult>(&tr);
        CHECK(ok != nullptr);
        printf("TEXT_LEN %zu LANG %s MS %llu SEGMENTS %zu\n", ok->text.size(), ok->language.c_str(), (unsigned long long)ok->duration_ms, ok->segments.size());
        CHECK((ok->text.empty() && ok->segments.empty()) || !ok->segments.empty());  // the text is the segments' texts, trimmed
        for (size_t i = 0; i < ok->segments.size(); ++i) CHECK(ok->segments[i].t0 >= 0);  // the model's own times, 10 ms units
        BenchmarkResult b = eng->benchmark(0.2f);
        CHECK(b.test_audio_secs == 2.0f && b.recommended_chunk_interval >= b.overhead_secs);
        {   // the worker's level-processing hook (worker.rs:196-240) through the mirror: off by default, on -> the engine
            // sees a normalised / compressed / limited buffer (peak under the -1 dB ceiling, level up from -26 dB to about -19.6 dB)
            AudioBuffer q = a;
            AudioConfig off, on;
            on.preprocessing = true;
            CHECK(eng->preprocess_audio({&q}, off) && q.samples == a.samples);
            CHECK(eng->preprocess_audio({&q}, on));
            double ss0 = 0.0, ss1 = 0.0;
            float peak = 0.0f;
            for (int i = 0; i < n; ++i) {
                ss0 += (double)a.samples[i] * a.samples[i];
                ss1 += (double)q.samples[i] * q.samples[i];
                peak = std::max(peak, std::fabs(q.samples[i]));
            }
            CHECK(peak <= std::pow(10.0f, -1.0f / 20.0f) + 1e-6f && ss1 > 3.5 * ss0 && ss1 < 5.5 * ss0);  // -26.0 dB -> -19.6 dB (oracle: x 4.38)
            // two fresh engines (a state's first call is what sampled fallback tokens are defined against, oh_whisper.h)
            auto f1 = WhisperEngine::create(argv[1], "en", false, true), f2 = WhisperEngine::create(argv[1], "en", false, true);
            CHECK(std::get_if<std::unique_ptr<WhisperEngine>>(&f1) && std::get_if<std::unique_ptr<WhisperEngine>>(&f2));
            TranscriptionJob job;
            job.buffer = a;
            JobResult with = process_job(std::get<std::unique_ptr<WhisperEngine>>(f1).get(), job, &on);
            auto direct = std::get<std::unique_ptr<WhisperEngine>>(f2)->transcribe(q);
            CHECK(std::get_if<TranscriptionResult>(&direct) && with.text == std::get<TranscriptionResult>(direct).text);
            printf("PREPROCESS_OK peak %.4f\n", peak);
        }
        a.samples[5] = NAN;
        auto bad = eng->transcribe(a);
        CHECK(std::get_if<WhisperError>(&bad) && std::get<WhisperError>(bad).kind == WhisperError::ValidationFailed);
        a.samples[5] = 0.0f;
        // one worker thread per visible GPU (up to 2) over one job queue, each draining up to max_batch jobs into one
        // oh_full_batch (queue/worker.rs:255-267; config.gpu.devices): every answer equals the single transcribe() call
        std::string ref_text;
        {   // a fresh engine: the first call on a fresh state is what every job of a batched call equals (oh_whisper.h)
            auto fr = WhisperEngine::create(argv[1], "en", false, true);
            CHECK(std::get_if<std::unique_ptr<WhisperEngine>>(&fr) != nullptr);
            ref_text = std::get<TranscriptionResult>(std::get<std::unique_ptr<WhisperEngine>>(fr)->transcribe(a)).text;
        }
        const int n_dev = argc >= 5 ? atoi(argv[4]) : 1;
        std::vector<std::unique_ptr<WhisperEngine>> engines;
        for (int d = 0; d < n_dev; ++d) {
            auto er = WhisperEngine::create(argv[1], "en", false, true, d, 4);
            CHECK(std::get_if<std::unique_ptr<WhisperEngine>>(&er) != nullptr);
            engines.push_back(std::move(std::get<std::unique_ptr<WhisperEngine>>(er)));
        }
        engines.push_back(nullptr);  // a worker whose engine failed to load takes no jobs
        WorkerPool pool(std::move(engines));
        const int n_jobs = 9;
        for (int j = 0; j < n_jobs; ++j) {
            TranscriptionJob job;
            job.buffer = a;
            if (j == 4) job.buffer.samples[7] = INFINITY;  // fails validation alone
            job.sequence_id = (uint64_t)j;
            pool.submit(std::move(job));
        }
        auto got = pool.drain(n_jobs);
        pool.shutdown();
        CHECK((int)got.size() == n_jobs);
        size_t taken = 0;
        for (size_t w = 0; w < pool.jobs_done().size(); ++w) taken += pool.jobs_done()[w];
        CHECK((int)taken == n_jobs && pool.jobs_done().back() == 0);
        for (auto& r : got) CHECK(r.sequence_id == 4 ? r.text.empty() : r.text == ref_text);
        printf("POOL_OK workers=%d\n", n_dev);
    }
    printf("HOST_TEST_OK\n");
    return 0;
}
