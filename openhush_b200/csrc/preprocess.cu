// Optional level processing of the 16 kHz clip on the device ("optional normalise / limit"): the
// worker's hook between capture and WhisperEngine::transcribe, /root/reference/src/queue/worker.rs:196-240
// (preprocess_audio), i.e. in this order
//   1. AudioBuffer::normalize_rms   /root/reference/src/input/audio.rs:89-122   gain = 10^((target_db - rms_db) / 20)
//   2. AudioBuffer::compress        audio.rs:139-198   attack / release envelope follower, gain above the threshold, make-up gain
//   3. AudioBuffer::limit           audio.rs:200-239   instant attack, smoothed release towards 1
// Two of the three carry a recurrence over the samples whose coefficient depends on the data (|x| > envelope picks attack or
// release; target < gain picks the instant attack), so they do not reduce to a scan: a clip is walked sample by sample by ONE
// lane, and the cost of a pass is the length of that dependent chain (about a dozen cycles per sample).  A warp owns a clip:
// all 32 lanes move the data (the next 1024-sample tile is in flight in registers while the current one is worked on,
// finished tiles are stored coalesced) and compute the per-sample terms that do not depend on the recurrence's state (the
// limiter's division among them); lane 0 then runs the recurrence over the tile out of shared memory — so little besides
// the chain itself is serial, and clips run side by side on as many warps as there are clips.
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

constexpr int LV_TILE = 1024, LV_WARPS = 4;

// One warp sweeps one clip.  Per sample, `prep(v)` computes what does not depend on the recurrence's state — all 32 lanes
// do that for a whole tile at once (the limiter's division lives here) — and `step(v, aux)` advances the state and returns
// the value to store (kStore); lane 0 calls it in sample order, so a tile's only serial part is the chain itself.  The
// tile buffer is updated in place.
template <bool kStore, class P, class F>
__device__ __forceinline__ void warp_sweep(const float* __restrict__ src, float* __restrict__ dst, int n, float* s_buf, float* s_aux, int lane, P&& prep,
                                           F&& step) {
    float nxt[LV_TILE / 32];
    auto fetch = [&](int t0) {
#pragma unroll
        for (int k = 0; k < LV_TILE / 32; ++k) {
            const int idx = t0 + k * 32 + lane;
            nxt[k] = idx < n ? src[idx] : 0.0f;
        }
    };
    fetch(0);
    for (int t0 = 0; t0 < n; t0 += LV_TILE) {
#pragma unroll
        for (int k = 0; k < LV_TILE / 32; ++k) {
            s_buf[k * 32 + lane] = nxt[k];
            s_aux[k * 32 + lane] = prep(nxt[k]);
        }
        __syncwarp();
        if (t0 + LV_TILE < n) fetch(t0 + LV_TILE);  // lands while lane 0 walks the current tile
        if (lane == 0) {
            const int m = min(LV_TILE, n - t0);
            int i0 = 0;
            for (; i0 + 8 <= m; i0 += 8) {  // full groups: straight-line code
                float v[8], a[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[k] = s_buf[i0 + k];
                    a[k] = s_aux[i0 + k];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = step(v[k], a[k]);
                if (kStore) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) s_buf[i0 + k] = v[k];
                }
            }
            for (; i0 < m; ++i0) {  // the clip's last few samples
                const float o = step(s_buf[i0], s_aux[i0]);
                if (kStore) s_buf[i0] = o;
            }
        }
        __syncwarp();
        if (kStore) {
#pragma unroll
            for (int k = 0; k < LV_TILE / 32; ++k) {
                const int idx = t0 + k * 32 + lane;
                if (idx < n) dst[idx] = s_buf[k * 32 + lane];
            }
        }
        __syncwarp();
    }
}

// pass A, one warp per clip: RMS before, the normalisation gain, the compressor's envelope (written to `env`)
__global__ void __launch_bounds__(32 * LV_WARPS) level_envelope_kernel(const float* __restrict__ pcm, long long stride, const int* __restrict__ n_samples,
                                                                       int batch, LevelCoeffs c, float* __restrict__ env, float* __restrict__ gain_out,
                                                                       float* __restrict__ rms_out) {
    __shared__ float s_tile[LV_WARPS][LV_TILE], s_aux[LV_WARPS][LV_TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * LV_WARPS + warp;
    if (b >= batch) return;
    const int n = n_samples[b];
    const float* x = pcm + (size_t)b * stride;
    float sum = 0.0f;
    warp_sweep<false>(
        x, nullptr, n, s_tile[warp], s_aux[warp], lane, [&](float v) { return __fmul_rn(v, v); },
        [&](float, float sq) {
            sum = __fadd_rn(sum, sq);
            return 0.0f;
        });
    float gain = 1.0f;
    if (lane == 0) {  // lane 0's sum is the clip's
        const float before = rms_db_of(sum, n);
        if (c.normalize && isfinite(before)) gain = powf(10.0f, __fdiv_rn(__fsub_rn(c.target_db, before), 20.0f));  // audio.rs:112-113, 125
        gain_out[b] = gain;
        if (rms_out) rms_out[2 * b] = before;
    }
    if (!c.compress) return;
    gain = __shfl_sync(0xffffffffu, gain, 0);
    float envelope = 0.0f;
    const float a1 = __fsub_rn(1.0f, c.attack_coeff), r1 = __fsub_rn(1.0f, c.release_coeff);
    warp_sweep<true>(
        x, env + (size_t)b * stride, n, s_tile[warp], s_aux[warp], lane,
        [&](float v) { return fabsf(__fmul_rn(v, gain)); },  // the compressor sees the normalised sample
        [&](float, float in) {
            const float up = __fadd_rn(__fmul_rn(c.attack_coeff, envelope), __fmul_rn(a1, in));
            const float down = __fadd_rn(__fmul_rn(c.release_coeff, envelope), __fmul_rn(r1, in));
            envelope = in > envelope ? up : down;  // audio.rs:166-170
            return envelope;
        });
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

// pass C, one warp per clip: the limiter's gain recurrence (audio.rs:211-229) and the RMS after; `work` -> clip
__global__ void __launch_bounds__(32 * LV_WARPS) level_limit_kernel(const float* __restrict__ work, float* __restrict__ pcm, long long stride,
                                                                    const int* __restrict__ n_samples, int batch, LevelCoeffs c,
                                                                    float* __restrict__ rms_out) {
    __shared__ float s_tile[LV_WARPS][LV_TILE], s_aux[LV_WARPS][LV_TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * LV_WARPS + warp;
    if (b >= batch) return;
    const int n = n_samples[b];
    float gain_reduction = 1.0f, sum = 0.0f;
    const float r1 = __fsub_rn(1.0f, c.limiter_release_coeff);
    warp_sweep<true>(
        work + (size_t)b * stride, pcm + (size_t)b * stride, n, s_tile[warp], s_aux[warp], lane,
        [&](float v) {  // the gain this sample asks for
            const float in = fabsf(v);
            return in > c.ceiling_linear ? __fdiv_rn(c.ceiling_linear, in) : 1.0f;
        },
        [&](float v, float target) {
            if (c.limit) {
                const float rel = __fadd_rn(__fmul_rn(c.limiter_release_coeff, gain_reduction), __fmul_rn(r1, target));
                gain_reduction = target < gain_reduction ? target : rel;
                v = __fmul_rn(v, gain_reduction);
            }
            sum = __fadd_rn(sum, __fmul_rn(v, v));
            return v;
        });
    if (lane == 0 && rms_out) rms_out[2 * b + 1] = rms_db_of(sum, n);
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
    level_envelope_kernel<<<(batch + LV_WARPS - 1) / LV_WARPS, 32 * LV_WARPS, 0, stream>>>(d_pcm, stride, d_n, batch, c, d_work, d_gain, d_rms);
    OH_COUNT_LAUNCH();
    const int gx = std::max(1, std::min((max_n + 255) / 256, 1184 / std::max(batch, 1) + 1));  // about eight CTAs per SM over the batch
    level_gain_kernel<<<dim3(gx, batch), 256, 0, stream>>>(d_pcm, stride, d_n, c, d_work, d_gain);
    OH_COUNT_LAUNCH();
    level_limit_kernel<<<(batch + LV_WARPS - 1) / LV_WARPS, 32 * LV_WARPS, 0, stream>>>(d_work, d_pcm, stride, d_n, batch, c, d_rms);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace oh
