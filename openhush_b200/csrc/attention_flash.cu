// K6 (throughput kernel) — fused encoder self-attention for the 16-bit modes: softmax(Q K^T / 8) V per
// (clip, head) without materialising the 1500 x 1500 score matrix (scores and P.V use
// 16-bit operands with f32 accumulation in the reference arithmetic too).
//
// One CTA = 64 query rows of one (clip, head), 4 warps x 16 rows.  K and V tiles of 64 keys stream through a
// two-stage cp.async ring in XOR-swizzled shared memory; Q.K^T and P.V are m16n8k16 tensor-core MMAs fed by
// ldmatrix; the softmax is the online (running max / running sum) form in f32 registers with exp2; P is rounded
// to the activation type as the MMA operand.  Keys >= T are zero-filled and masked; the output tile is staged
// through shared memory so that global stores are 16-byte and coalesced.
// Grid: (ceil(T/64), heads, clips) = 24 x 20 x B for large-v3 — far more CTAs than the 148 SMs can hold at once.
#include "../../include/oh_whisper.h"
#include "kernels.h"

namespace oh {

namespace {

constexpr int BQ = 64, BKV = 64, HD = 64, FA_THREADS = 128;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (row, col) of a [rows][64] 16-bit tile: 16-byte chunk index is XORed with (row & 7)
__device__ __forceinline__ int swz(int row, int col) { return row * HD + ((((col >> 3) ^ (row & 7)) << 3) | (col & 7)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T>
__device__ __forceinline__ uint32_t pack2(float x, float y);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float x, float y) {
    __half2 h = __floats2half2_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float x, float y) {
    __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T>
__global__ void __launch_bounds__(FA_THREADS) encoder_attention_flash_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int d) {
    __shared__ __align__(128) T Qs[BQ * HD];
    __shared__ __align__(128) T Ks[2][BKV * HD];
    __shared__ __align__(128) T Vs[2][BKV * HD];
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const size_t ld = 3 * (size_t)d;
    const T* base = qkv + (size_t)b * Tn * ld + h * HD;

    auto load_tile = [&](T* dst, int row0, int col_off) {
#pragma unroll
        for (int i = 0; i < (BKV * HD / 8) / FA_THREADS; ++i) {
            const int idx = tid + i * FA_THREADS;
            const int r = idx >> 3, ch = idx & 7;
            const bool ok = row0 + r < Tn;
            const T* src = base + (size_t)(ok ? row0 + r : 0) * ld + col_off + ch * 8;
            cp_async16(smem_addr(dst + swz(r, ch * 8)), src, ok);
        }
    };

    const int nkt = (Tn + BKV - 1) / BKV;
    load_tile(Qs, q0, 0);
    load_tile(Ks[0], 0, d);
    load_tile(Vs[0], 0, 2 * d);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = 0.0f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.0f, 0.0f};
    const float sc = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)

    for (int kt = 0; kt < nkt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nkt) {
            load_tile(Ks[buf ^ 1], (kt + 1) * BKV, d);
            load_tile(Vs[buf ^ 1], (kt + 1) * BKV, 2 * d);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 16 + (lane >> 4) * 8;
                ldmatrix_x4(smem_addr(Qs + swz(r, c)), qf[kk]);
            }
        }
        // ---- S = Q K^T (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[j][e] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t kb[4];
                const int r = jp * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(smem_addr(Ks[buf] + swz(r, c)), kb);
                mma16816<T>(s[2 * jp], qf[kk], kb[0], kb[1]);
                mma16816<T>(s[2 * jp + 1], qf[kk], kb[2], kb[3]);
            }
        }
        // ---- online softmax (rows g and g+8 of this warp's 16)
        const int key0 = kt * BKV;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = key0 + j * 8 + 2 * t + (e & 1);
                const float v = key < Tn ? s[j][e] * sc : -INFINITY;
                s[j][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        }
        float corr[2], m_new[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            m_new[r] = fmaxf(m_run[r], mx[r]);
            corr[r] = exp2f(m_run[r] - m_new[r]);
            m_run[r] = m_new[r];
            l_run[r] *= corr[r];
        }
        uint32_t pf[4][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f(s[j][0] - m_new[0]), p1 = exp2f(s[j][1] - m_new[0]);
            const float p2 = exp2f(s[j][2] - m_new[1]), p3 = exp2f(s[j][3] - m_new[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[j >> 1][(j & 1) * 2 + 0] = pack2<T>(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack2<T>(p2, p3);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= corr[0];
            o[j][1] *= corr[0];
            o[j][2] *= corr[1];
            o[j][3] *= corr[1];
        }
        // ---- O += P V
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                uint32_t vb[4];
                const int r = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = dp * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(smem_addr(Vs[buf] + swz(r, c)), vb);
                mma16816<T>(o[2 * dp], pf[ks], vb[0], vb[1]);
                mma16816<T>(o[2 * dp + 1], pf[ks], vb[2], vb[3]);
            }
        }
        __syncthreads();  // everyone is done with buffer `buf` before the next prefetch overwrites it
    }

    // ---- normalise, stage through shared memory (Qs is free now), coalesced store
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
        l_run[r] = 1.0f / l_run[r];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = j * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(Qs + swz(warp * 16 + g, c)) = pack2<T>(o[j][0] * l_run[0], o[j][1] * l_run[0]);
        *reinterpret_cast<uint32_t*>(Qs + swz(warp * 16 + g + 8, c)) = pack2<T>(o[j][2] * l_run[1], o[j][3] * l_run[1]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < (BQ * HD / 8) / FA_THREADS; ++i) {
        const int idx = tid + i * FA_THREADS;
        const int r = idx >> 3, ch = idx & 7;
        if (q0 + r < Tn) {
            const uint4 v = *reinterpret_cast<const uint4*>(Qs + swz(r, ch * 8));
            *reinterpret_cast<uint4*>(out + ((size_t)b * Tn + q0 + r) * d + h * HD + ch * 8) = v;
        }
    }
}

}  // namespace

int launch_encoder_attention_mma(const void* qkv, void* out, int type, int B, int Tn, int d, cudaStream_t st) {
    if (type == DT_F32) return launch_encoder_attention_simple(qkv, out, type, B, Tn, d, st);
    if (d % HD != 0 || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_error("encoder_attention_mma: unsupported shape or alignment");
        return OH_ERR_INVALID;
    }
    dim3 grid((Tn + BQ - 1) / BQ, d / HD, B);
    ProfScope ps(PC_ATTN_ENC, 4.0 * B * (double)Tn * Tn * d, st);
    if (type == DT_F16) encoder_attention_flash_kernel<__half><<<grid, FA_THREADS, 0, st>>>((const __half*)qkv, (__half*)out, Tn, d);
    else encoder_attention_flash_kernel<__nv_bfloat16><<<grid, FA_THREADS, 0, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, Tn, d);
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace oh
