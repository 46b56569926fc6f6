// Pre-engine DSP on the device: the step immediately before the hot path on the file route.
// Replaces, for PCM that is already in memory, what /root/reference/src/input/audio.rs does between the WAV reader and
// WhisperEngine::transcribe:
//   :368-381  samples to f32 (integer formats: s / 2^(bits-1))
//   :384-391  mono = sum of a frame's channels / channels
//   :394-407  resample to 16 kHz with resample_sinc (:1007-1095): rubato 0.16 SincFixedIn, sinc_len 256, f_cutoff 0.95,
//             256x oversampled tables with linear interpolation between the two nearest, BlackmanHarris2 window, fed in
//             chunks of 1024 frames, the last chunk zero-padded and its output cut to ceil(len * ratio)
//   :415-424  zero padding to 1.1 s
// One kernel does all of it: a CTA per 1024-frame chunk stages the chunk (plus the 512 frames before it) as mono f32 in
// shared memory straight from the interleaved input, and every thread evaluates output samples as two 256-tap inner
// products against the oversampled table (L2-resident, 256 KB) blended by the sub-sample fraction.  The chunk
// bookkeeping (how many output samples each chunk yields) is the resampler's own recurrence, run on the host: it is
// O(outputs) additions and fixes the output length exactly.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "engine.h"

namespace oh {

namespace {

constexpr int RS_SINC = 256, RS_OVER = 256, RS_CHUNK = 1024, RS_THREADS = 256;

struct ChunkPlan {
    double last_index;  // resampler position (in input frames, relative to the chunk start) before the chunk's first output
    int out_off, n_out;
};

float sincf_pi(float v) {
    if (v == 0.0f) return 1.0f;
    const float x = v * (float)M_PI;
    return std::sin(x) / x;
}

// the [256 sub-sample offsets][256 taps] table for a ratio: Blackman-Harris squared window over all 65536 points, sinc at
// cutoff 0.95 (times the ratio when downsampling), normalised to unit DC gain per table
std::vector<float> make_sinc_table(double ratio) {
    const float cutoff = ratio >= 1.0 ? 0.95f : 0.95f * (float)ratio;
    const int tot = RS_SINC * RS_OVER;
    std::vector<float> y(tot);
    const float pi2 = 2.0f * (float)M_PI, pi4 = 4.0f * (float)M_PI, pi6 = 6.0f * (float)M_PI, np_f = (float)tot;
    float sum = 0.0f;
    for (int x = 0; x < tot; ++x) {
        const float xf = (float)x;
        float w = 0.35875f - 0.48829f * std::cos(pi2 * xf / np_f) + 0.14128f * std::cos(pi4 * xf / np_f) - 0.01168f * std::cos(pi6 * xf / np_f);
        w = w * w;
        const float val = w * sincf_pi((xf - (float)(tot / 2)) * cutoff / (float)RS_OVER);
        sum += val;
        y[x] = val;
    }
    sum /= (float)RS_OVER;
    std::vector<float> t((size_t)tot);
    for (int p = 0; p < RS_SINC; ++p)
        for (int n = 0; n < RS_OVER; ++n) t[(size_t)(RS_OVER - n - 1) * RS_SINC + p] = y[(size_t)RS_OVER * p + n] / sum;
    return t;
}

struct TableCache {  // device tables per (device, input rate)
    std::mutex mu;
    std::map<std::pair<int, int>, float*> tables;
} g_tables;

const float* device_table(int device, int from_rate) {
    std::lock_guard<std::mutex> lk(g_tables.mu);
    auto key = std::make_pair(device, from_rate);
    auto it = g_tables.tables.find(key);
    if (it != g_tables.tables.end()) return it->second;
    const std::vector<float> t = make_sinc_table((double)kSampleRate / (double)from_rate);
    float* d = nullptr;
    if (cudaMalloc(&d, t.size() * 4) != cudaSuccess || cudaMemcpy(d, t.data(), t.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    g_tables.tables[key] = d;
    return d;
}

// host side of the resampler's chunk loop (audio.rs:1046-1075 + rubato's index recurrence): output samples per chunk
std::vector<ChunkPlan> plan_chunks(long long n_frames, double ratio, long long* total_out) {
    std::vector<ChunkPlan> plan;
    const double t_ratio = 1.0 / ratio;
    const long end_idx = (long)RS_CHUNK - ((long)RS_SINC + 1) - (long)std::ceil(t_ratio);
    double last_index = -(double)(RS_SINC / 2);
    long long out = 0;
    for (long long pos = 0; pos < n_frames; pos += RS_CHUNK) {
        const long long len = std::min<long long>(RS_CHUNK, n_frames - pos);
        ChunkPlan c;
        c.last_index = last_index;
        c.out_off = (int)out;
        double idx = last_index;
        int n = 0;
        while (idx < (double)end_idx) {
            idx += t_ratio;
            ++n;
        }
        last_index = idx - (double)RS_CHUNK;
        if (len < RS_CHUNK) n = (int)std::min<long long>((long long)std::ceil((double)len * ratio), n);
        c.n_out = n;
        out += n;
        plan.push_back(c);
    }
    *total_out = out;
    return plan;
}

template <int FORMAT>  // 0 f32, 1 i16, 2 i32 holding `bits`-bit integers
__device__ __forceinline__ float load_mono(const void* pcm, long long frame, int channels, float inv_max) {
    float sum = 0.0f;
    for (int c = 0; c < channels; ++c) {
        const long long i = frame * channels + c;
        float v;
        if (FORMAT == 0) v = __ldg((const float*)pcm + i);
        else if (FORMAT == 1) v = (float)__ldg((const short*)pcm + i) * inv_max;
        else v = (float)__ldg((const int*)pcm + i) * inv_max;
        sum += v;
    }
    return channels > 1 ? sum / (float)channels : sum;
}

template <int FORMAT>
__global__ void __launch_bounds__(RS_THREADS) resample_sinc_kernel(const void* __restrict__ pcm, int channels, float inv_max, long long n_frames,
                                                                   const ChunkPlan* __restrict__ plan, const float* __restrict__ table, double t_ratio,
                                                                   float* __restrict__ out) {
    __shared__ float buf[RS_CHUNK + 2 * RS_SINC];  // [0, 512): the frames before the chunk (zeros before the clip), then the chunk
    const long long chunk0 = (long long)blockIdx.x * RS_CHUNK;
    for (int b = threadIdx.x; b < RS_CHUNK + 2 * RS_SINC; b += RS_THREADS) {
        const long long g = chunk0 + b - 2 * RS_SINC;
        buf[b] = (g >= 0 && g < n_frames) ? load_mono<FORMAT>(pcm, g, channels, inv_max) : 0.0f;
    }
    __syncthreads();
    const ChunkPlan cp = plan[blockIdx.x];
    for (int k = threadIdx.x; k < cp.n_out; k += RS_THREADS) {
        const double idx = cp.last_index + (double)(k + 1) * t_ratio;
        const double fl = floor(idx);
        long index = (long)fl;
        long sub = (long)floor((idx - fl) * (double)RS_OVER);
        const long i0 = index, s0 = sub;
        sub += 1;
        if (sub >= RS_OVER) {
            sub -= RS_OVER;
            index += 1;
        }
        const double fo = idx * (double)RS_OVER;
        const float frac = (float)(fo - floor(fo));
        const float* w0 = buf + (i0 + 2 * RS_SINC);
        const float* w1 = buf + (index + 2 * RS_SINC);
        const float4* t0 = reinterpret_cast<const float4*>(table + (size_t)s0 * RS_SINC);
        const float4* t1 = reinterpret_cast<const float4*>(table + (size_t)sub * RS_SINC);
        // eight running partial sums per inner product, added up at the end (the published scalar kernel's order)
        float a0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < RS_SINC; i += 8) {
            const float4 x0 = __ldg(t0 + i / 4), x1 = __ldg(t0 + i / 4 + 1), y0 = __ldg(t1 + i / 4), y1 = __ldg(t1 + i / 4 + 1);
            a0[0] = fmaf(w0[i + 0], x0.x, a0[0]); a0[1] = fmaf(w0[i + 1], x0.y, a0[1]); a0[2] = fmaf(w0[i + 2], x0.z, a0[2]); a0[3] = fmaf(w0[i + 3], x0.w, a0[3]);
            a0[4] = fmaf(w0[i + 4], x1.x, a0[4]); a0[5] = fmaf(w0[i + 5], x1.y, a0[5]); a0[6] = fmaf(w0[i + 6], x1.z, a0[6]); a0[7] = fmaf(w0[i + 7], x1.w, a0[7]);
            a1[0] = fmaf(w1[i + 0], y0.x, a1[0]); a1[1] = fmaf(w1[i + 1], y0.y, a1[1]); a1[2] = fmaf(w1[i + 2], y0.z, a1[2]); a1[3] = fmaf(w1[i + 3], y0.w, a1[3]);
            a1[4] = fmaf(w1[i + 4], y1.x, a1[4]); a1[5] = fmaf(w1[i + 5], y1.y, a1[5]); a1[6] = fmaf(w1[i + 6], y1.z, a1[6]); a1[7] = fmaf(w1[i + 7], y1.w, a1[7]);
        }
        float p0 = 0.0f, p1 = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p0 += a0[j];
            p1 += a1[j];
        }
        out[cp.out_off + k] = p0 + frac * (p1 - p0);
    }
}

// RMS level (dB full scale) of every `chunk`-sample block of a 16 kHz clip: one warp per block (the stand-in VAD engine of
// SURVEY.md section 8f N3 — the reference's Silero model is not in its tree; see openhush_b200/vad.py)
__global__ void __launch_bounds__(256) chunk_rms_db_kernel(const float* __restrict__ pcm, int chunk, int n_chunks, float* __restrict__ out) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= n_chunks) return;
    const float* x = pcm + (size_t)w * chunk;
    float acc = 0.0f;
    for (int i = lane; i < chunk; i += 32) acc = fmaf(x[i], x[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[w] = 20.0f * log10f(fmaxf(sqrtf(acc / (float)chunk), 1e-10f));
}

// same-rate input: conversion and mono mix only
template <int FORMAT>
__global__ void mono_mix_kernel(const void* __restrict__ pcm, int channels, float inv_max, long long n_frames, float* __restrict__ out) {
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < n_frames; f += (long long)gridDim.x * blockDim.x)
        out[f] = load_mono<FORMAT>(pcm, f, channels, inv_max);
}

}  // namespace

long long resample_out_len(int sample_rate, long long n_frames, bool pad_min_duration) {
    if (n_frames <= 0 || sample_rate <= 0) return 0;
    long long n = n_frames;
    if (sample_rate != kSampleRate) plan_chunks(n_frames, (double)kSampleRate / (double)sample_rate, &n);
    if (pad_min_duration && (float)n / (float)kSampleRate < 1.1f) n = (long long)((float)kSampleRate * 1.1f);
    return n;
}

// d_pcm: interleaved frames in device memory; d_out: device buffer of at least resample_out_len(.., true) floats
int resample_to_16k_device(oh_state* st, const void* d_pcm, int format, int bits, int channels, int sample_rate, long long n_frames, float* d_out,
                           long long* out_n, cudaStream_t stream) {
    if (!d_pcm || !d_out || n_frames <= 0 || channels < 1 || channels > 64 || sample_rate < 4000 || sample_rate > 768000 || format < 0 || format > 2 ||
        (format == 2 && (bits < 2 || bits > 32))) {
        set_error("resample: bad arguments (format 0 f32 / 1 i16 / 2 i32, 1..64 channels, 4 kHz..768 kHz)");
        return OH_ERR_INVALID;
    }
    if ((double)kSampleRate / (double)sample_rate > 2.0) {  // SincFixedIn is created with max_resample_ratio_relative 2.0 and sized for it
        set_error("resample: input rates below 8 kHz are not supported");
        return OH_ERR_UNSUPPORTED;
    }
    const float inv_max = 1.0f / (float)(1ll << ((format == 1 ? 16 : bits) - 1));
    long long n16 = n_frames;
    if (sample_rate == kSampleRate) {
        const int blocks = (int)std::min<long long>((n_frames + 255) / 256, 148 * 8);
        if (format == 0) mono_mix_kernel<0><<<blocks, 256, 0, stream>>>(d_pcm, channels, inv_max, n_frames, d_out);
        else if (format == 1) mono_mix_kernel<1><<<blocks, 256, 0, stream>>>(d_pcm, channels, inv_max, n_frames, d_out);
        else mono_mix_kernel<2><<<blocks, 256, 0, stream>>>(d_pcm, channels, inv_max, n_frames, d_out);
        OH_COUNT_LAUNCH();
    } else {
        const double ratio = (double)kSampleRate / (double)sample_rate;
        const float* table = device_table(st->ctx->params.gpu_device, sample_rate);
        if (!table) {
            set_error("resample: cannot allocate the interpolation table");
            return OH_ERR_CUDA;
        }
        const std::vector<ChunkPlan> plan = plan_chunks(n_frames, ratio, &n16);
        int rc;
        if ((rc = ensure(st->resample_plan, plan.size() * sizeof(ChunkPlan)))) return rc;
        OH_CUDA_CHECK(cudaMemcpyAsync(st->resample_plan.p, plan.data(), plan.size() * sizeof(ChunkPlan), cudaMemcpyHostToDevice, stream));  // pageable: staged before return
        const int grid = (int)plan.size();
        const ChunkPlan* dp = (const ChunkPlan*)st->resample_plan.p;
        ProfScope ps(PC_RESAMPLE, (double)n_frames * channels * (format == 1 ? 2 : 4) + (double)n16 * 4, stream);
        if (format == 0) resample_sinc_kernel<0><<<grid, RS_THREADS, 0, stream>>>(d_pcm, channels, inv_max, n_frames, dp, table, 1.0 / ratio, d_out);
        else if (format == 1) resample_sinc_kernel<1><<<grid, RS_THREADS, 0, stream>>>(d_pcm, channels, inv_max, n_frames, dp, table, 1.0 / ratio, d_out);
        else resample_sinc_kernel<2><<<grid, RS_THREADS, 0, stream>>>(d_pcm, channels, inv_max, n_frames, dp, table, 1.0 / ratio, d_out);
        OH_COUNT_LAUNCH();
    }
    OH_CUDA_CHECK(cudaGetLastError());
    // audio.rs:415-424: silence up to 1.1 s, below which the engine (like whisper.cpp) returns nothing
    if ((float)n16 / (float)kSampleRate < 1.1f) {
        const long long need = (long long)((float)kSampleRate * 1.1f);
        OH_CUDA_CHECK(cudaMemsetAsync(d_out + n16, 0, (size_t)(need - n16) * 4, stream));
        n16 = need;
    }
    if (out_n) *out_n = n16;
    return OH_OK;
}

int chunk_rms_db_device(const float* d_pcm, int chunk, int n_chunks, float* d_out, cudaStream_t stream) {
    if (n_chunks <= 0) return OH_OK;
    chunk_rms_db_kernel<<<(n_chunks + 7) / 8, 256, 0, stream>>>(d_pcm, chunk, n_chunks, d_out);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

const std::vector<float>& resample_table_host(int from_rate) {
    static std::mutex mu;
    static std::map<int, std::vector<float>> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(from_rate);
    if (it == cache.end()) it = cache.emplace(from_rate, make_sinc_table((double)kSampleRate / (double)from_rate)).first;
    return it->second;
}

}  // namespace oh
