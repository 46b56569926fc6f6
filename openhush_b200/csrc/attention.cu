// K6 — encoder self-attention (non-causal, T = 1500, head dim 64), SURVEY.md A.4.
//
// encoder_attention_simple: the parity-mode kernel.  One CTA = 16 query rows of one (clip, head); the full
//   score rows live in shared memory, so the softmax is the reference's two-pass form (exact row max, then
//   exp and sum, then scale by 1/sum) rather than an online rescaling, and P.V accumulates sequentially over
//   keys in f32.  In the 16-bit modes P is rounded to the activation type before P.V, as ggml does.
// encoder_attention_mma: the throughput kernel for the 16-bit modes (flash-style, tensor cores), below.
#include <mma.h>

#include "../../include/oh_whisper.h"
#include "kernels.h"

namespace oh {

namespace {

constexpr int SQ = 16;    // queries per CTA
constexpr int KT = 64;    // keys per tile
constexpr int S_LD = 1504;

template <typename T>
__global__ void __launch_bounds__(256) encoder_attention_simple_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int d) {
    extern __shared__ __align__(16) float sm[];
    float* S = sm;                       // [SQ][S_LD]
    float* Qs = S + SQ * S_LD;           // [SQ][64]
    float* KVs = Qs + SQ * kHeadDim;     // [KT][65]
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * SQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = 3 * d;
    const T* base = qkv + (size_t)b * Tn * ld;
    const int nq = min(SQ, Tn - q0);

    for (int idx = tid; idx < SQ * kHeadDim; idx += 256) {
        const int q = idx / kHeadDim, c = idx % kHeadDim;
        Qs[idx] = q < nq ? Cvt<T>::to_f32(base[(size_t)(q0 + q) * ld + h * kHeadDim + c]) : 0.0f;
    }
    const float scale = 0.125f;  // 1/sqrt(64)

    // ---- pass 1: S = (Q K^T) * scale
    const int j = tid % KT, qg = tid / KT;
    for (int j0 = 0; j0 < Tn; j0 += KT) {
        __syncthreads();
        for (int idx = tid; idx < KT * kHeadDim; idx += 256) {
            const int key = idx / kHeadDim, c = idx % kHeadDim;
            KVs[key * 65 + c] = (j0 + key < Tn) ? Cvt<T>::to_f32(base[(size_t)(j0 + key) * ld + d + h * kHeadDim + c]) : 0.0f;
        }
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int c = 0; c < kHeadDim; ++c) {
            const float kv = KVs[j * 65 + c];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(Qs[(qg * 4 + q) * kHeadDim + c], kv, acc[q]);
        }
        if (j0 + j < Tn) {
#pragma unroll
            for (int q = 0; q < 4; ++q) S[(qg * 4 + q) * S_LD + j0 + j] = acc[q] * scale;
        }
    }
    __syncthreads();

    // ---- pass 2: row softmax (two rows per warp)
    for (int r = warp * 2; r < warp * 2 + 2; ++r) {
        float* row = S + r * S_LD;
        float mx = -INFINITY;
        for (int i = lane; i < Tn; i += 32) mx = fmaxf(mx, row[i]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int i = lane; i < Tn; i += 32) {
            const float e = expf(row[i] - mx);
            row[i] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int i = lane; i < Tn; i += 32) row[i] = Cvt<T>::to_f32(Cvt<T>::from_f32(row[i] * inv));
    }

    // ---- pass 3: O = P V
    const int c = tid % kHeadDim;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < Tn; j0 += KT) {
        __syncthreads();
        for (int idx = tid; idx < KT * kHeadDim; idx += 256) {
            const int key = idx / kHeadDim, cc = idx % kHeadDim;
            KVs[key * 65 + cc] = (j0 + key < Tn) ? Cvt<T>::to_f32(base[(size_t)(j0 + key) * ld + 2 * d + h * kHeadDim + cc]) : 0.0f;
        }
        __syncthreads();
        const int nk = min(KT, Tn - j0);
        for (int k = 0; k < nk; ++k) {
            const float v = KVs[k * 65 + c];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = fmaf(S[(qg * 4 + q) * S_LD + j0 + k], v, o[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int qq = qg * 4 + q;
        if (qq < nq) out[((size_t)b * Tn + q0 + qq) * d + h * kHeadDim + c] = Cvt<T>::from_f32(o[q]);
    }
}

template <typename T>
int launch_simple_t(const void* qkv, void* out, int B, int Tn, int d, cudaStream_t st) {
    const size_t smem = (size_t)(SQ * S_LD + SQ * kHeadDim + KT * 65) * sizeof(float);
    static PerDeviceInit attr;  // per instantiation and per device
    if (!attr.ensure([&] { return cudaFuncSetAttribute(encoder_attention_simple_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess; })) {
        set_error("encoder_attention_simple: shared-memory opt-in failed");
        return OH_ERR_CUDA;
    }
    dim3 grid((Tn + SQ - 1) / SQ, d / kHeadDim, B);
    ProfScope ps(PC_ATTN_ENC, 4.0 * B * (double)Tn * Tn * d, st);
    encoder_attention_simple_kernel<T><<<grid, 256, smem, st>>>((const T*)qkv, (T*)out, Tn, d);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace

int launch_encoder_attention_simple(const void* qkv, void* out, int type, int B, int Tn, int d, cudaStream_t st) {
    if (Tn > S_LD) {
        set_error("encoder_attention: T too large");
        return OH_ERR_INVALID;
    }
    if (type == DT_F32) return launch_simple_t<float>(qkv, out, B, Tn, d, st);
    if (type == DT_F16) return launch_simple_t<__half>(qkv, out, B, Tn, d, st);
    return launch_simple_t<__nv_bfloat16>(qkv, out, B, Tn, d, st);
}

}  // namespace oh

