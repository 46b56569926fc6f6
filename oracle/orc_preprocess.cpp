// ORACLE (test infrastructure) — the worker's optional level processing between capture and the engine, restated from
//   /root/reference/src/queue/worker.rs:196-240   preprocess_audio: normalise, then compress, then limit (each if enabled)
//   /root/reference/src/input/audio.rs:89-104     AudioBuffer::rms_db      sequential f32 sum of squares, 20 log10(sqrt(mean))
//   audio.rs:108-136                               normalize_rms / apply_gain   samples *= 10^((target_db - rms_db) / 20)
//   audio.rs:139-198                               compress                 attack / release envelope, gain curve, make-up gain
//   audio.rs:200-239                               limit                    instant attack, smoothed release
// Operation order and rounding follow the Rust source statement by statement (f32 throughout, no fused multiply-add: rustc
// never contracts, so contraction is switched off for this file).  Pinned by the reference's own unit tests for these
// functions (audio.rs:1189-1330), restated in tests/test_preprocess.py; what cannot be pinned offline is the last bit of
// log10f / powf / expf (Rust calls the platform libm too).
#pragma GCC optimize("fp-contract=off")
#include <cmath>
#include <cstdint>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr float kRate = 16000.0f;

float rms_db(const float* x, long long n) {  // audio.rs:89-104
    if (n <= 0) return -INFINITY;
    float sum_squares = 0.0f;
    for (long long i = 0; i < n; ++i) sum_squares += x[i] * x[i];
    const float rms = std::sqrt(sum_squares / (float)n);
    return rms > 0.0f ? 20.0f * std::log10(rms) : -INFINITY;
}

void apply_gain(float* x, long long n, float gain_db) {  // audio.rs:124-130
    const float gain_linear = std::pow(10.0f, gain_db / 20.0f);
    for (long long i = 0; i < n; ++i) x[i] *= gain_linear;
}

void normalize_rms(float* x, long long n, float target_db) {  // audio.rs:108-122
    const float current = rms_db(x, n);
    if (std::isfinite(current)) apply_gain(x, n, target_db - current);
}

void compress(float* x, long long n, float threshold_db, float ratio, float attack_ms, float release_ms, float makeup_gain_db) {  // audio.rs:139-198
    if (n <= 0 || ratio <= 1.0f) return;
    const float threshold_linear = std::pow(10.0f, threshold_db / 20.0f);
    const float attack_coeff = std::exp(-1.0f / (attack_ms * kRate / 1000.0f));
    const float release_coeff = std::exp(-1.0f / (release_ms * kRate / 1000.0f));
    float envelope = 0.0f;
    for (long long i = 0; i < n; ++i) {
        const float input_abs = std::fabs(x[i]);
        if (input_abs > envelope) envelope = attack_coeff * envelope + (1.0f - attack_coeff) * input_abs;
        else envelope = release_coeff * envelope + (1.0f - release_coeff) * input_abs;
        float gain = 1.0f;
        if (envelope > threshold_linear) {
            const float over_db = 20.0f * std::log10(envelope / threshold_linear);
            const float compressed_db = over_db / ratio;
            const float reduction_db = over_db - compressed_db;
            gain = std::pow(10.0f, -reduction_db / 20.0f);
        }
        x[i] *= gain;
    }
    if (makeup_gain_db != 0.0f) apply_gain(x, n, makeup_gain_db);
}

void limit(float* x, long long n, float ceiling_db, float release_ms) {  // audio.rs:200-239
    if (n <= 0) return;
    const float ceiling_linear = std::pow(10.0f, ceiling_db / 20.0f);
    const float release_coeff = std::exp(-1.0f / (release_ms * kRate / 1000.0f));
    float gain_reduction = 1.0f;
    for (long long i = 0; i < n; ++i) {
        const float input_abs = std::fabs(x[i]);
        const float target_gain = input_abs > ceiling_linear ? ceiling_linear / input_abs : 1.0f;
        if (target_gain < gain_reduction) gain_reduction = target_gain;
        else gain_reduction = release_coeff * gain_reduction + (1.0f - release_coeff) * target_gain;
        x[i] *= gain_reduction;
    }
}

}  // namespace

ORC_API float orc_rms_db(const float* x, long long n) { return rms_db(x, n); }
ORC_API void orc_apply_gain(float* x, long long n, float gain_db) { apply_gain(x, n, gain_db); }
ORC_API void orc_normalize_rms(float* x, long long n, float target_db) { normalize_rms(x, n, target_db); }
ORC_API void orc_compress(float* x, long long n, float threshold_db, float ratio, float attack_ms, float release_ms, float makeup_gain_db) {
    compress(x, n, threshold_db, ratio, attack_ms, release_ms, makeup_gain_db);
}
ORC_API void orc_limit(float* x, long long n, float ceiling_db, float release_ms) { limit(x, n, ceiling_db, release_ms); }

// worker.rs:208-239: the three stages in the worker's order; flags / values in the layout of oh_preprocess_params
// (normalize, target_db, compress, threshold_db, ratio, attack_ms, release_ms, makeup_gain_db, limit, ceiling_db, limiter_release_ms);
// rms_db[2] = level before and after
ORC_API void orc_preprocess_audio(float* x, long long n, const int* flags /*[3]*/, const float* v /*[8]*/, float* rms_out /*[2] or null*/) {
    if (rms_out) rms_out[0] = rms_db(x, n);
    if (flags[0]) normalize_rms(x, n, v[0]);
    if (flags[1]) compress(x, n, v[1], v[2], v[3], v[4], v[5]);
    if (flags[2]) limit(x, n, v[6], v[7]);
    if (rms_out) rms_out[1] = rms_db(x, n);
}
