// K4 (parity mode) — f32-accumulating SIMT mat-mul on CUDA cores with the same epilogue contract as the
// tensor-core kernel (kernels.h GemmArgs).  This is the arithmetic of OH_COMPUTE_F32: f16-stored weights are
// widened exactly, activations stay f32, every output is one sequential-k f32 dot product per thread — the
// closest device analogue of the reference CPU path's f32 accumulation, used for token-exact
// parity.  It also takes 16-bit activations so that the tensor-core pipeline can be cross-checked against it.
#include "../../include/oh_whisper.h"
#include "kernels.h"

namespace oh {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, THREADS = 256;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

template <typename TA, typename TW, typename TC>
__global__ void __launch_bounds__(THREADS) gemm_simt_kernel(GemmArgs g) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    pdl_wait();  // programmatic dependent launch (see common.cuh)
    pdl_launch_dependents();
    const int M = g.batch * g.rows;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const TA* __restrict__ A = reinterpret_cast<const TA*>(g.A);
    const TW* __restrict__ W = reinterpret_cast<const TW*>(g.W);

    // loader mapping: 128 rows x 2 chunks of 8 k
    const int lrow = tid >> 1, lk = (tid & 1) * 8;
    const int am = m0 + lrow;
    const bool a_ok = am < M;
    const TA* a_ptr = A;
    if (a_ok) {
        const int b = am / g.rows, i = am - b * g.rows;
        a_ptr = A + (long long)b * g.a_bstride + (long long)i * g.lda;
    }
    const int wn = n0 + lrow;
    const bool w_ok = wn < g.N;
    const TW* w_ptr = W + (long long)(w_ok ? wn : 0) * g.K;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < g.K; k0 += BK) {
        float av[8], wv[8];
        const int kk = k0 + lk;
        if (a_ok && kk + 8 <= g.K) load8<TA>(a_ptr + kk, av);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) av[j] = (a_ok && kk + j < g.K) ? Cvt<TA>::to_f32(a_ptr[kk + j]) : 0.0f;
        }
        if (w_ok && kk + 8 <= g.K) load8<TW>(w_ptr + kk, wv);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = (w_ok && kk + j < g.K) ? Cvt<TW>::to_f32(w_ptr[kk + j]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            As[lk + j][lrow] = av[j];
            Ws[lk + j][lrow] = wv[j];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]), a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[k][tx * TN]), b1 = *reinterpret_cast<const float4*>(&Ws[k][tx * TN + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

    TC* __restrict__ C = reinterpret_cast<TC*>(g.C);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= M) continue;
        const int b = m / g.rows, r = m - b * g.rows;
        TC* crow = C + (long long)(g.c_bidx ? g.c_bidx[b] : b) * g.c_bstride + (long long)r * g.ldc;
        const float* rrow = g.R ? g.R + (long long)b * g.r_bstride + (long long)r * g.ldr : nullptr;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += g.bias[n];
            if (g.gelu == GELU_LUT) v = gelu_lut(v, g.lut);
            else if (g.gelu == GELU_TANH) v = gelu_tanh(v);
            else if (g.gelu == GELU_FAST) {
                if (g.round_type == DT_F16) v = __half2float(__float2half_rn(gelu_fast(__half2float(__float2half_rn(v)))));
                else if (g.round_type == DT_BF16) v = __bfloat162float(__float2bfloat16_rn(gelu_fast(v)));
                else v = gelu_fast(v);
            }
            if (rrow) v += rrow[n];
            if (g.c_hsplit > 0) crow[(long long)(n / g.c_hsplit) * g.c_hstride + (n % g.c_hsplit)] = Cvt<TC>::from_f32(v);
            else crow[n] = Cvt<TC>::from_f32(v);
        }
    }
}

template <typename TA, typename TW>
int dispatch_c(const GemmArgs& g, dim3 grid, cudaStream_t st) {
    switch (g.c_type) {
        case DT_F32: OH_CUDA_CHECK(launch_k(gemm_simt_kernel<TA, TW, float>, grid, THREADS, 0, st, g)); break;
        case DT_F16: OH_CUDA_CHECK(launch_k(gemm_simt_kernel<TA, TW, __half>, grid, THREADS, 0, st, g)); break;
        default: OH_CUDA_CHECK(launch_k(gemm_simt_kernel<TA, TW, __nv_bfloat16>, grid, THREADS, 0, st, g)); break;
    }
    return OH_OK;
}
template <typename TA>
int dispatch_w(const GemmArgs& g, dim3 grid, cudaStream_t st) {
    return g.w_type == DT_BF16 ? dispatch_c<TA, __nv_bfloat16>(g, grid, st) : dispatch_c<TA, __half>(g, grid, st);
}

}  // namespace

int launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
    const int M = g.batch * g.rows;
    if (M <= 0 || g.N <= 0 || g.K <= 0) return OH_OK;
    // vector loads need 8-element (16/32-byte) aligned rows
    const bool aligned = (g.lda % 8 == 0) && (g.a_bstride % 8 == 0) && (g.K % 8 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 31) == 0) &&
                         ((reinterpret_cast<uintptr_t>(g.W) & 15) == 0);
    if (!aligned) {
        set_error("gemm_simt: unaligned operands");
        return OH_ERR_INVALID;
    }
    dim3 grid((g.N + BN - 1) / BN, (M + BM - 1) / BM);
    ProfScope ps(M <= 512 ? PC_GEMM_DEC : PC_GEMM_ENC, 2.0 * M * (double)g.N * g.K, st);
    switch (g.a_type) {
        case DT_F32: dispatch_w<float>(g, grid, st); break;
        case DT_F16: dispatch_w<__half>(g, grid, st); break;
        default: dispatch_w<__nv_bfloat16>(g, grid, st); break;
    }
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace oh
