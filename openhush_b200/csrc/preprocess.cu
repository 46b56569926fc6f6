// Optional level processing of the 16 kHz clip on the device ("optional normalise / limit"): the
// worker's hook between capture and WhisperEngine::transcribe, /root/reference/src/queue/worker.rs:196-240
// (preprocess_audio), i.e. in this order
//   1. AudioBuffer::normalize_rms   /root/reference/src/input/audio.rs:89-122   gain = 10^((target_db - rms_db) / 20)
//   2. AudioBuffer::compress        audio.rs:139-198   attack / release envelope follower, gain above the threshold, make-up gain
//   3. AudioBuffer::limit           audio.rs:200-239   instant attack, smoothed release towards 1
// Two of the three carry a recurrence over the samples whose coefficient depends on the data (|x| > envelope picks attack or
// release; target < gain picks the instant attack), so they do not reduce to a scan: they run one thread per clip, 32 clips
// per warp, every lane walking its own clip (a lane's 32-byte sectors are re-used by its next seven samples out of L1).
// Everything that is not on a recurrence — the compressor's log10 / pow gain curve, the gain products — is a plain
// element-wise kernel in between; the three passes alternate between the clip and one scratch array, so each reads one buffer
// (read-only path) and writes the other.  Arithmetic follows the reference operation by operation in f32 (separate multiplies and
// adds, no contraction; the RMS sum is the same sequential f32 sum), so what differs from the reference's own numbers is
// the last bits of log10f / powf / expf.
#include <cmath>

#include "engine.h"

namespace oh {

namespace {

struct LevelCoeffs {
    int normalize, compress, limit;
    float target_db;
    float threshold_linear, ratio, attack_coeff, release_coeff, makeup_linear;
    int has_makeup;
    float ceiling_linear, limiter_release_coeff;
};

// audio.rs:89-104: f32 sum of squares in sample order, sqrt of the mean, 20 log10
__device__ __forceinline__ float rms_db_of(float sum_squares, int n) {
    if (n <= 0) return -INFINITY;
    const float rms = __fsqrt_rn(__fdiv_rn(sum_squares, (float)n));
    return rms > 0.0f ? __fmul_rn(20.0f, log10f(rms)) : -INFINITY;
}

// pass A, one thread per clip: RMS before, the normalisation gain, the compressor's envelope (written to `env`)
__global__ void __launch_bounds__(32) level_envelope_kernel(const float* __restrict__ pcm, long long stride, const int* __restrict__ n_samples, int batch,
                                                            LevelCoeffs c, float* __restrict__ env, float* __restrict__ gain_out,
                                                            float* __restrict__ rms_out) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= batch) return;
    const int n = n_samples[b];
    const float* x = pcm + (size_t)b * stride;
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) sum = __fadd_rn(sum, __fmul_rn(x[i], x[i]));
    const float before = rms_db_of(sum, n);
    float gain = 1.0f;
    if (c.normalize && isfinite(before)) gain = powf(10.0f, __fdiv_rn(__fsub_rn(c.target_db, before), 20.0f));  // audio.rs:112-113, 125
    gain_out[b] = gain;
    if (rms_out) rms_out[2 * b] = before;
    if (!c.compress) return;
    float* e = env + (size_t)b * stride;
    float envelope = 0.0f;
    const float a1 = __fsub_rn(1.0f, c.attack_coeff), r1 = __fsub_rn(1.0f, c.release_coeff);
    for (int i = 0; i < n; ++i) {
        const float in = fabsf(__fmul_rn(x[i], gain));  // the compressor sees the normalised sample
        const float up = __fadd_rn(__fmul_rn(c.attack_coeff, envelope), __fmul_rn(a1, in));
        const float down = __fadd_rn(__fmul_rn(c.release_coeff, envelope), __fmul_rn(r1, in));
        envelope = in > envelope ? up : down;  // audio.rs:166-170
        e[i] = envelope;
    }
}

// pass B, element-wise: normalisation gain, compressor gain curve (audio.rs:173-182), make-up gain (audio.rs:186-188).
// Reads the clip and its envelope, writes the result over the envelope (`work`), so that pass C reads one buffer and writes
// the other and both can go through the read-only path.
__global__ void __launch_bounds__(256) level_gain_kernel(const float* __restrict__ pcm, long long stride, const int* __restrict__ n_samples, LevelCoeffs c,
                                                         float* __restrict__ work, const float* __restrict__ gain_in) {
    const int b = blockIdx.y;
    const int n = n_samples[b];
    const float g_norm = gain_in[b];
    const float* x = pcm + (size_t)b * stride;
    float* e = work + (size_t)b * stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = x[i];
        if (c.normalize) v = __fmul_rn(v, g_norm);
        if (c.compress) {
            const float envelope = e[i];
            float gain = 1.0f;
            if (envelope > c.threshold_linear) {
                const float over_db = __fmul_rn(20.0f, log10f(__fdiv_rn(envelope, c.threshold_linear)));
                const float compressed_db = __fdiv_rn(over_db, c.ratio);
                const float reduction_db = __fsub_rn(over_db, compressed_db);
                gain = powf(10.0f, __fdiv_rn(-reduction_db, 20.0f));
            }
            v = __fmul_rn(v, gain);
            if (c.has_makeup) v = __fmul_rn(v, c.makeup_linear);
        }
        e[i] = v;
    }
}

// pass C, one thread per clip: the limiter's gain recurrence (audio.rs:211-229) and the RMS after; `work` -> clip.
// Eight samples are in registers before the recurrence touches the first, so the loads do not sit on its chain.
__global__ void __launch_bounds__(32) level_limit_kernel(const float* __restrict__ work, float* __restrict__ pcm, long long stride,
                                                         const int* __restrict__ n_samples, int batch, LevelCoeffs c, float* __restrict__ rms_out) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    if (b >= batch) return;
    const int n = n_samples[b];
    const float* src = work + (size_t)b * stride;
    float* x = pcm + (size_t)b * stride;
    float gain_reduction = 1.0f, sum = 0.0f;
    const float r1 = __fsub_rn(1.0f, c.limiter_release_coeff);
    for (int i0 = 0; i0 < n; i0 += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = i0 + k < n ? src[i0 + k] : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i0 + k >= n) break;
            if (c.limit) {
                const float in = fabsf(v[k]);
                const float target = in > c.ceiling_linear ? __fdiv_rn(c.ceiling_linear, in) : 1.0f;
                const float rel = __fadd_rn(__fmul_rn(c.limiter_release_coeff, gain_reduction), __fmul_rn(r1, target));
                gain_reduction = target < gain_reduction ? target : rel;
                v[k] = __fmul_rn(v[k], gain_reduction);
            }
            x[i0 + k] = v[k];
            sum = __fadd_rn(sum, __fmul_rn(v[k], v[k]));
        }
    }
    if (rms_out) rms_out[2 * b + 1] = rms_db_of(sum, n);
}

}  // namespace

// d_pcm: [batch][stride] f32, processed in place; d_n: samples per clip; d_work: [batch][stride] scratch (envelope, then the
// compressed clip); d_gain: [batch]; d_rms: [batch][2] (before, after) or nullptr
int preprocess_levels_device(float* d_pcm, long long stride, const int* d_n, int batch, int max_n, const oh_preprocess_params& p, float* d_work, float* d_gain,
                             float* d_rms, cudaStream_t stream) {
    if (batch <= 0) return OH_OK;
    LevelCoeffs c{};
    const float sr = (float)kSampleRate;
    c.normalize = p.normalize != 0;
    c.target_db = p.target_db;
    c.compress = p.compress != 0 && p.ratio > 1.0f;  // audio.rs:147: a ratio <= 1 leaves the clip alone (no make-up gain either)
    c.threshold_linear = std::pow(10.0f, p.threshold_db / 20.0f);
    c.ratio = p.ratio;
    c.attack_coeff = std::exp(-1.0f / (p.attack_ms * sr / 1000.0f));
    c.release_coeff = std::exp(-1.0f / (p.release_ms * sr / 1000.0f));
    c.has_makeup = p.makeup_gain_db != 0.0f;
    c.makeup_linear = std::pow(10.0f, p.makeup_gain_db / 20.0f);
    c.limit = p.limit != 0;
    c.ceiling_linear = std::pow(10.0f, p.ceiling_db / 20.0f);
    c.limiter_release_coeff = std::exp(-1.0f / (p.limiter_release_ms * sr / 1000.0f));
    ProfScope ps(PC_RESAMPLE, 0.0, stream);
    level_envelope_kernel<<<(batch + 31) / 32, 32, 0, stream>>>(d_pcm, stride, d_n, batch, c, d_work, d_gain, d_rms);
    OH_COUNT_LAUNCH();
    const int gx = std::max(1, std::min((max_n + 255) / 256, 1184 / std::max(batch, 1) + 1));  // about eight CTAs per SM over the batch
    level_gain_kernel<<<dim3(gx, batch), 256, 0, stream>>>(d_pcm, stride, d_n, c, d_work, d_gain);
    OH_COUNT_LAUNCH();
    level_limit_kernel<<<(batch + 31) / 32, 32, 0, stream>>>(d_work, d_pcm, stride, d_n, batch, c, d_rms);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace oh
