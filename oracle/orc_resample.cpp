// ORACLE (test infrastructure) — pre-engine DSP of the file route: sample conversion, mono mix and the sinc resampler to
// 16 kHz, i.e. what /root/reference/src/input/audio.rs does between the WAV reader and the engine:
//   :368-381  samples to f32 (float as is; integer s / 2^(bits-1))
//   :384-391  mono = sum of the frame's channels (f32, in order) / channels
//   :394-407  resample(.., ResamplingQuality::High) -> resample_sinc, :1007-1095
//   :415-424  zero padding to 1.1 s (WHISPER_MIN_DURATION_SECS, :31-34)
//
// resample_sinc drives the third-party crate rubato 0.16.2 (Cargo.toml:37, Cargo.lock:5571-5581; NOT in /root/reference):
// SincFixedIn::<f32>::new(ratio, 2.0, {sinc_len 256, f_cutoff 0.95, Linear, oversampling 256, BlackmanHarris2}, 1024, 1)
// fed chunk by chunk, the last chunk zero-padded and its output cut to ceil(len * ratio) (audio.rs:1046-1075).  What is
// restated below is rubato's published algorithm for exactly that configuration (asynchronous sinc resampler, fixed
// input size, linear interpolation between the two nearest of 256 oversampled sinc tables).  PARITY UNPINNED for the
// crate's internals (its f32 summation order in particular; outputs are compared at 1e-5); the restatement is anchored
// on the reference's own call site and on its own tests for this function (audio.rs:1166-1187: same rate is the
// identity, 4800 samples 48 kHz -> 16 kHz give 1600 +- 100 samples).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int kSincLen = 256, kOversampling = 256, kChunk = 1024;
constexpr float kCutoff = 0.95f;

float sincf_pi(float v) {
    if (v == 0.0f) return 1.0f;
    const float x = v * (float)M_PI;
    return std::sin(x) / x;
}

// rubato windows.rs: Blackman-Harris over npoints (periodic form), squared for "BlackmanHarris2"
std::vector<float> blackman_harris2(int npoints) {
    std::vector<float> w(npoints);
    const float pi2 = 2.0f * (float)M_PI, pi4 = 4.0f * (float)M_PI, pi6 = 6.0f * (float)M_PI, np_f = (float)npoints;
    const float a0 = 0.35875f, a1 = 0.48829f, a2 = 0.14128f, a3 = 0.01168f;
    for (int x = 0; x < npoints; ++x) {
        const float xf = (float)x;
        const float v = a0 - a1 * std::cos(pi2 * xf / np_f) + a2 * std::cos(pi4 * xf / np_f) - a3 * std::cos(pi6 * xf / np_f);
        w[x] = v * v;
    }
    return w;
}

// rubato sinc.rs make_sincs: `factor` tables of `npoints` taps, table (factor - n - 1) holds the taps at sub-sample offset n
std::vector<float> make_sincs(int npoints, int factor, float f_cutoff) {
    const int tot = npoints * factor;
    const std::vector<float> window = blackman_harris2(tot);
    std::vector<float> y(tot);
    float sum = 0.0f;
    for (int x = 0; x < tot; ++x) {
        const float val = window[x] * sincf_pi(((float)x - (float)(tot / 2)) * f_cutoff / (float)factor);
        sum += val;
        y[x] = val;
    }
    sum /= (float)factor;
    std::vector<float> sincs((size_t)factor * npoints);
    for (int p = 0; p < npoints; ++p)
        for (int n = 0; n < factor; ++n) sincs[(size_t)(factor - n - 1) * npoints + p] = y[(size_t)factor * p + n] / sum;
    return sincs;
}

// rubato's scalar inner product: eight running partial sums, added up at the end
float dot8(const float* wave, const float* sinc, int len) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < len; i += 8)
        for (int j = 0; j < 8; ++j) acc[j] += wave[i + j] * sinc[i + j];
    float s = 0.0f;
    for (int j = 0; j < 8; ++j) s += acc[j];
    return s;
}

struct SincFixedIn {
    double ratio;
    std::vector<float> sincs;
    std::vector<float> buffer;  // chunk + 2 * sinc_len
    double last_index;
    explicit SincFixedIn(double resample_ratio) : ratio(resample_ratio) {
        const float cutoff = resample_ratio >= 1.0 ? kCutoff : kCutoff * (float)resample_ratio;
        sincs = make_sincs(kSincLen, kOversampling, cutoff);
        buffer.assign(kChunk + 2 * kSincLen, 0.0f);
        last_index = -(double)(kSincLen / 2);
    }
    // one chunk of exactly kChunk frames in, a variable number of frames out
    void process(const float* in, std::vector<float>& out) {
        const double t_ratio = 1.0 / ratio;
        const long end_idx = (long)kChunk - ((long)kSincLen + 1) - (long)std::ceil(t_ratio);
        std::memmove(buffer.data(), buffer.data() + kChunk, sizeof(float) * 2 * kSincLen);
        std::memcpy(buffer.data() + 2 * kSincLen, in, sizeof(float) * kChunk);
        double idx = last_index;
        while (idx < (double)end_idx) {
            idx += t_ratio;
            long index = (long)std::floor(idx);
            long sub = (long)std::floor((idx - std::floor(idx)) * (double)kOversampling);
            const long i0 = index, s0 = sub;
            sub += 1;
            if (sub >= kOversampling) {
                sub -= kOversampling;
                index += 1;
            }
            const double fo = idx * (double)kOversampling;
            const float frac = (float)(fo - std::floor(fo));
            const float p0 = dot8(buffer.data() + (i0 + 2 * kSincLen), sincs.data() + (size_t)s0 * kSincLen, kSincLen);
            const float p1 = dot8(buffer.data() + (index + 2 * kSincLen), sincs.data() + (size_t)sub * kSincLen, kSincLen);
            out.push_back(p0 + frac * (p1 - p0));
        }
        last_index = idx - (double)kChunk;
    }
};

}  // namespace

// the 256 x 256 interpolation table for a ratio (tests compare the product's table with it)
ORC_API void orc_resample_sinc_table(double ratio, float* out /* [256][256] */) {
    const float cutoff = ratio >= 1.0 ? kCutoff : kCutoff * (float)ratio;
    const std::vector<float> t = make_sincs(kSincLen, kOversampling, cutoff);
    std::memcpy(out, t.data(), t.size() * sizeof(float));
}

// audio.rs:368-391: interleaved frames -> mono f32.  format 0: f32 samples; 1: i16; 2: i32 carrying `bits`-bit integers
ORC_API void orc_mono_mix(const void* pcm, int format, int bits, int channels, long long n_frames, float* out) {
    const float max_value = (float)(1 << ((format == 1 ? 16 : bits) - 1));
    for (long long f = 0; f < n_frames; ++f) {
        float sum = 0.0f;
        for (int c = 0; c < channels; ++c) {
            const long long i = f * channels + c;
            float v;
            if (format == 0) v = ((const float*)pcm)[i];
            else if (format == 1) v = (float)((const int16_t*)pcm)[i] / max_value;
            else v = (float)((const int32_t*)pcm)[i] / max_value;
            sum += v;
        }
        out[f] = channels > 1 ? sum / (float)channels : sum;
    }
}

// audio.rs:1007-1095.  Returns the number of output samples written (<= cap), or the number needed if out == nullptr.
ORC_API long long orc_resample_sinc(const float* samples, long long n, int from_rate, int to_rate, float* out, long long cap) {
    if (n <= 0) return 0;
    std::vector<float> res;
    if (from_rate == to_rate) {
        res.assign(samples, samples + n);
    } else {
        const double ratio = (double)to_rate / (double)from_rate;
        SincFixedIn rs(ratio);
        std::vector<float> in(kChunk), got;
        for (long long pos = 0; pos < n;) {
            const long long end = std::min<long long>(pos + kChunk, n), len = end - pos;
            std::memcpy(in.data(), samples + pos, sizeof(float) * len);
            if (len < kChunk) std::fill(in.begin() + len, in.end(), 0.0f);
            got.clear();
            rs.process(in.data(), got);
            if (len < kChunk) {  // "for the last chunk, only take the proportional amount"
                const size_t expected = (size_t)std::ceil((double)len * ratio);
                got.resize(std::min(expected, got.size()));
            }
            res.insert(res.end(), got.begin(), got.end());
            pos = end;
        }
    }
    if (out)
        for (long long i = 0; i < (long long)res.size() && i < cap; ++i) out[i] = res[i];
    return (long long)res.size();
}

// audio.rs:415-424: samples the loader appends so that the clip lasts 1.1 s
ORC_API long long orc_min_duration_padding(long long n_16k) {
    const double dur = (double)n_16k / 16000.0;
    if ((float)dur < 1.1f) return (long long)(16000.0f * 1.1f) - n_16k;
    return 0;
}
