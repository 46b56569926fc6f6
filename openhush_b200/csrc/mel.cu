// K1 / K2 — batched log-mel front end (replaces whisper.cpp's CPU-only `log_mel_spectrogram`, which the
// reference reaches through state.full() at /root/reference/src/engine/whisper.rs:266; algorithm per
// SURVEY.md Appendix A.3).
//
// K1 (mel_frames_kernel): one CTA = 32 consecutive frames of one clip.  The 5360 PCM samples those frames
//   touch are staged in shared memory once with coalesced float4 loads (reflect-pad on the left edge and
//   zero-pad past the clip are applied while staging), each warp then takes frames in turn: Hann window +
//   even/odd packing, 200-point complex Stockham FFT (fft200.cuh) ping-ponging between two per-warp shared
//   buffers, real-input split + |X|^2 for 201 bins, sparse mel filterbank, log10(max(.,1e-10)).  Results are
//   collected in a [n_mel][32] shared tile and written out as 128-byte runs.  The clip-wide maximum needed
//   by the normalisation is folded in with one atomicMax per CTA on an order-preserving int key.
// K2 is never a separate pass on the product path: clamp-to-(max-8) and (x+4)/4 are applied by the kernel
//   that consumes the mel (window_gather_kernel below, which writes the encoder's conv1 input).
//   mel_normalize_kernel exists for the stage API (oh_pcm_to_mel_*), which must return normalised mels.
//
// HBM traffic per 30 s clip: 1.92 MB PCM read + 4*n_mel*3001 B raw mel written.
#include "../../include/oh_whisper.h"
#include "fft200.cuh"
#include "kernels.h"

namespace oh {

namespace {

constexpr int FB = 32;                       // frames per CTA
constexpr int MEL_THREADS = 256;             // 8 warps
constexpr int MEL_WARPS = MEL_THREADS / 32;
constexpr int SEG = kHop * (FB - 1) + kNFFT;  // 5360 staged samples
constexpr int MAX_MEL = 128;
constexpr int MAX_FILT_W = kMaxMelFilterWeights;  // load_model rejects denser banks

__device__ __forceinline__ int float_to_ordered(float f) {
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

struct MelSmem {
    float pcm[SEG];
    float2 tw[fft::NTW];
    float hann[kNFFT];
    float2 bufA[MEL_WARPS][fft::N2];
    float2 bufB[MEL_WARPS][fft::N2];
    float power[MEL_WARPS][208];
    float tile[MAX_MEL][FB + 1];
    float filt_w[MAX_FILT_W];
    int filt_start[MAX_MEL], filt_count[MAX_MEL], filt_off[MAX_MEL];
    float red[MEL_WARPS];
};

__global__ void __launch_bounds__(MEL_THREADS) mel_frames_kernel(MelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
    const int clip = blockIdx.y;
    const int n = a.n_samples[clip];
    const int n_active = a.n_active[clip];
    const int f0 = blockIdx.x * FB;
    if (f0 >= n_active) return;
    const float* __restrict__ pcm = a.pcm[clip];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage constants and the PCM segment --------------------------------------------------------
    for (int i = tid; i < fft::NTW; i += MEL_THREADS) {
        s.tw[i] = a.twiddle[i];
        s.hann[i] = a.hann[i];
    }
    const int n_mel = a.n_mel;
    for (int i = tid; i < n_mel; i += MEL_THREADS) {
        s.filt_start[i] = a.filt_start[i];
        s.filt_count[i] = a.filt_count[i];
        s.filt_off[i] = a.filt_off[i];
    }
    for (int i = tid; i < a.n_filt_w; i += MEL_THREADS) s.filt_w[i] = a.filt_w[i];

    // padded index p = 160*f0 + i  <->  pcm index p - 200 (reflect for p < 200: pcm[200 - p]; zero past n)
    const long long src0 = (long long)kHop * f0 - kNFFT / 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0);
    for (int i4 = tid; i4 < SEG / 4; i4 += MEL_THREADS) {
        const long long src = src0 + 4 * i4;
        float4 v;
        if (aligned && src >= 0 && src + 3 < n) {
            v = __ldg(reinterpret_cast<const float4*>(pcm + src));
        } else {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                long long q = src + j;
                if (q < 0) q = -q;  // reflect: padded p -> pcm[200 - p] = pcm[-(p - 200)]
                t[j] = (q < n) ? pcm[q] : 0.0f;
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
        }
        *reinterpret_cast<float4*>(&s.pcm[4 * i4]) = v;
    }
    __syncthreads();

    float local_max = -10.0f;
    const int nf = min(FB, n_active - f0);
    for (int fl = warp; fl < nf; fl += MEL_WARPS) {
        const float* x = &s.pcm[kHop * fl];
        float2* A = s.bufA[warp];
        float2* B = s.bufB[warp];
        // Hann window, pack even/odd samples as one complex sequence
        for (int i = lane; i < fft::N2; i += 32) A[i] = make_float2(x[2 * i] * s.hann[2 * i], x[2 * i + 1] * s.hann[2 * i + 1]);
        __syncwarp();
        fft::stockham_pass<5>(A, B, s.tw, 1, lane, 32);
        __syncwarp();
        fft::stockham_pass<5>(B, A, s.tw, 5, lane, 32);
        __syncwarp();
        fft::stockham_pass<4>(A, B, s.tw, 25, lane, 32);
        __syncwarp();
        fft::stockham_pass<2>(B, A, s.tw, 100, lane, 32);
        __syncwarp();
        float* P = s.power[warp];
        for (int k = lane; k < kNBins; k += 32) P[k] = fft::real_power_bin(A, s.tw, k);
        __syncwarp();
        for (int m = lane; m < n_mel; m += 32) {
            const int st = s.filt_start[m], cnt = s.filt_count[m];
            const float* w = &s.filt_w[s.filt_off[m]];
            float acc = 0.0f;
            for (int k = 0; k < cnt; ++k) acc = fmaf(P[st + k], w[k], acc);
            const float v = log10f(fmaxf(acc, 1e-10f));
            s.tile[m][fl] = v;
            local_max = fmaxf(local_max, v);
        }
        __syncwarp();
    }
    local_max = warp_max(local_max);
    if (lane == 0) s.red[warp] = local_max;
    __syncthreads();
    if (tid == 0) {
        float mx = s.red[0];
        for (int w = 1; w < MEL_WARPS; ++w) mx = fmaxf(mx, s.red[w]);
        atomicMax(&a.max_key[clip], float_to_ordered(mx));
    }
    // ---- write the tile: runs of up to 32 consecutive frames per mel row -----------------------------
    float* out = a.raw + (size_t)clip * n_mel * a.ld;
    for (int idx = tid; idx < n_mel * FB; idx += MEL_THREADS) {
        const int m = idx / FB, fl = idx % FB;
        if (fl < nf) out[(size_t)m * a.ld + f0 + fl] = s.tile[m][fl];
    }
}

__global__ void mel_init_max_kernel(int* max_key, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) max_key[i] = float_to_ordered(-10.0f);  // zero-padding frames are log10(1e-10) = -10 and count (A.3 step 5)
}

// stage API only: normalised mel [n_mel][n_len] of one clip, padding frames included
__global__ void mel_normalize_kernel(const float* __restrict__ raw, int ld, int n_mel, int n_active, int n_len, const int* max_key,
                                     float* __restrict__ out) {
    const float floor_v = ordered_to_float(*max_key) - 8.0f;
    const long long total = (long long)n_mel * n_len;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / n_len), f = (int)(i % n_len);
        float v = f < n_active ? raw[(size_t)m * ld + f] : -10.0f;
        v = fmaxf(v, floor_v);
        out[i] = (v + 4.0f) * 0.25f;
    }
}

// Encoder input of one window per batch slot: normalised mel frames [seek, seek+3000) of the slot's clip,
// channel-last and zero-framed: X[slot][3002][n_mel] with rows 0 and 3001 zero, so that conv1 (k3 s1 p1) is a
// plain mat-mul over overlapping 3-row views (see encoder.cu).  Frames in [n_active, n_len) are the padding
// value -10 (then clamped), frames >= n_len are 0 (A.3 step 6).
template <typename T>
__global__ void __launch_bounds__(256) window_gather_kernel(const WindowDesc* __restrict__ wins, int n_mel_, T* __restrict__ X) {
    __shared__ float tile[32][MAX_MEL + 1];
    const int slot = blockIdx.y;
    const WindowDesc w = wins[slot];
    struct { int n_mel, ld; } a{n_mel_, w.ld};
    const int seek = w.seek;
    const int n_active = w.n_active, n_len = w.n_len;
    const float floor_v = ordered_to_float(*w.max_key) - 8.0f;
    const float* __restrict__ raw = w.raw;
    const int j0 = blockIdx.x * 32;  // window column
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int m = ty; m < a.n_mel; m += 8) {
        const int j = j0 + tx, f = seek + j;
        float v = 0.0f;
        if (j < kChunkFrames && f < n_len) {
            const float r = f < n_active ? raw[(size_t)m * a.ld + f] : -10.0f;
            v = (fmaxf(r, floor_v) + 4.0f) * 0.25f;
        }
        tile[tx][m] = v;
    }
    __syncthreads();
    T* out = X + (size_t)slot * (kChunkFrames + 2) * a.n_mel;
    for (int idx = threadIdx.x; idx < 32 * a.n_mel; idx += 256) {
        const int jl = idx / a.n_mel, m = idx % a.n_mel;
        const int j = j0 + jl;
        if (j < kChunkFrames) out[(size_t)(1 + j) * a.n_mel + m] = Cvt<T>::from_f32(tile[jl][m]);
    }
    if (blockIdx.x == 0)
        for (int m = threadIdx.x; m < a.n_mel; m += 256) {
            out[m] = Cvt<T>::from_f32(0.0f);
            out[(size_t)(kChunkFrames + 1) * a.n_mel + m] = Cvt<T>::from_f32(0.0f);
        }
}

// stage API (oh_encode_*): mel windows given directly as [batch][n_mel][3000] f32, already normalised
template <typename T>
__global__ void __launch_bounds__(256) window_from_mel_kernel(const float* __restrict__ mel, int n_mel, T* __restrict__ X) {
    __shared__ float tile[32][MAX_MEL + 1];
    const int slot = blockIdx.y;
    const float* __restrict__ src = mel + (size_t)slot * n_mel * kChunkFrames;
    const int j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int m = ty; m < n_mel; m += 8) {
        const int j = j0 + tx;
        tile[tx][m] = j < kChunkFrames ? src[(size_t)m * kChunkFrames + j] : 0.0f;
    }
    __syncthreads();
    T* out = X + (size_t)slot * (kChunkFrames + 2) * n_mel;
    for (int idx = threadIdx.x; idx < 32 * n_mel; idx += 256) {
        const int jl = idx / n_mel, m = idx % n_mel;
        const int j = j0 + jl;
        if (j < kChunkFrames) out[(size_t)(1 + j) * n_mel + m] = Cvt<T>::from_f32(tile[jl][m]);
    }
    if (blockIdx.x == 0)
        for (int m = threadIdx.x; m < n_mel; m += 256) {
            out[m] = Cvt<T>::from_f32(0.0f);
            out[(size_t)(kChunkFrames + 1) * n_mel + m] = Cvt<T>::from_f32(0.0f);
        }
}

}  // namespace

size_t mel_smem_bytes() { return sizeof(MelSmem); }

int launch_mel_frames(const MelArgs& a, int batch, int max_active_frames, cudaStream_t st) {
    static PerDeviceInit attr;
    if (!attr.ensure([] { return cudaFuncSetAttribute(mel_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem)) == cudaSuccess; })) {
        set_error("mel_frames: shared-memory opt-in failed");
        return OH_ERR_CUDA;
    }
    mel_init_max_kernel<<<(batch + 127) / 128, 128, 0, st>>>(a.max_key, batch);
    OH_COUNT_LAUNCH();
    dim3 grid((max_active_frames + FB - 1) / FB, batch);
    {
        // algorithmic bytes: PCM read once + raw log-mel written once (work is filled in by the caller's clip sizes:
        // here the upper bound batch * (4 * 160 * frames + 4 * n_mel * frames))
        ProfScope ps(PC_MEL, (double)batch * max_active_frames * (4.0 * kHop + 4.0 * a.n_mel), st);
        mel_frames_kernel<<<grid, MEL_THREADS, sizeof(MelSmem), st>>>(a);
    }
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

int launch_mel_normalize(const float* raw, int ld, int n_mel, int n_active, int n_len, const int* max_key, float* out, cudaStream_t st) {
    const long long total = (long long)n_mel * n_len;
    int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 8);
    mel_normalize_kernel<<<blocks, 256, 0, st>>>(raw, ld, n_mel, n_active, n_len, max_key, out);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

int launch_window_gather(const WindowDesc* d_windows, int n_mel, int n_slots, void* X, int act_type, cudaStream_t st) {
    dim3 grid((kChunkFrames + 31) / 32, n_slots);
    ProfScope ps(PC_GATHER, (double)n_slots * kChunkFrames * n_mel * (4.0 + (act_type == OH_COMPUTE_F32 ? 4 : 2)), st);
    if (act_type == OH_COMPUTE_F32) window_gather_kernel<float><<<grid, 256, 0, st>>>(d_windows, n_mel, (float*)X);
    else if (act_type == OH_COMPUTE_F16) window_gather_kernel<__half><<<grid, 256, 0, st>>>(d_windows, n_mel, (__half*)X);
    else window_gather_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(d_windows, n_mel, (__nv_bfloat16*)X);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

int launch_window_from_mel(const float* mel, int n_mel, int n_slots, void* X, int act_type, cudaStream_t st) {
    dim3 grid((kChunkFrames + 31) / 32, n_slots);
    if (act_type == OH_COMPUTE_F32) window_from_mel_kernel<float><<<grid, 256, 0, st>>>(mel, n_mel, (float*)X);
    else if (act_type == OH_COMPUTE_F16) window_from_mel_kernel<__half><<<grid, 256, 0, st>>>(mel, n_mel, (__half*)X);
    else window_from_mel_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(mel, n_mel, (__nv_bfloat16*)X);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace oh
