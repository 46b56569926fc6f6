// K6 on the 5th-generation tensor cores — fused encoder self-attention softmax(Q K^T / 8) V per (window, head) for the
// 16-bit modes (reference arithmetic: whisper.cpp encoder attention, 16-bit operands with f32
// accumulation, reached from /root/reference/src/engine/whisper.rs:266).  The 1500 x 1500 score matrix never leaves
// the SM: S lives in TMEM, P goes through shared memory as the next MMA's operand, O partials come back through TMEM.
//
// One CTA = 128 query rows of one (window, head); 24 key blocks of 64; two CTAs per SM (100 KB of shared memory and
// 256 TMEM columns each) so that one CTA's softmax hides the other's barrier and TMEM latencies.  320 threads:
//   warp 0      TMA producer: Q tile once, then K_j and V_j boxes ([64 keys][64], 128-byte swizzle) through two rings,
//               all straight out of the packed qkv activation [B*T][3d] (3-D tensor maps, rows past T zero-filled)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer:
//                 S_j = Q K_j^T      M128 N64 K64   (A, B K-major)            -> TMEM columns [64*(j%2), +64)
//                 O_j = P_j V_j      M128 N64 K64   (A = P from tensor memory (or K-major shared memory), B = V MN-major:
//                                                    V is used as stored, no transpose)
//                                                                              -> TMEM columns [128 + 64*(j%2), +64)
//   warps 2..9  softmax, TWO threads per query row (TMEM lane = row; warps w and w+4 share a lane group, each takes 32 of
//               the block's 64 keys and 32 of the 64 output channels): half an S_j row into registers, row max (the two
//               halves meet through shared memory and a 64-thread named barrier), p = exp2(s*c - m) with the running
//               max / running sum of the online form, P_j rounded to the activation type and written to shared memory in
//               the swizzled K-major operand layout.  The output accumulator comes in two forms:
//                 fold (bf16): O_j partials are folded into f32 registers one block later
//                   (o = (o + O_{j-1}) * exp2(m_{j-1} - m_j)); the accumulator in TMEM is never rescaled;
//                 lazy (f16, the default there): O stays in ONE TMEM accumulator over all key blocks; P_j is taken against
//                   the maximum that accumulator is scaled to, and only when a row's maximum has grown by more than 8
//                   (p would pass 2^8) does the warp pull its 32 channels, rescale and store them back (tcgen05.ld/st) —
//                   after the first few blocks that almost never happens, so the per-block TMEM read-back, 32 FMAs per
//                   thread and one barrier round trip disappear (30.9 -> 27.4 ms per 32 windows).
//               Either way the MMAs of block j+1 overlap the softmax of block j.  Eight softmax warps per CTA (sixteen per
//               SM) instead of four: the kernel is bound by the latency of the per-row exp / TMEM chains, not by their
//               throughput.
// Every mbarrier wait is bounded (trap after ~4 s) so a protocol bug cannot hang the GPU.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/oh_whisper.h"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace oh {

namespace {

using namespace tcx;

constexpr int AT_BQ = 128, AT_BK = 64, AT_HD = 64;
constexpr int AT_THREADS = 320;                        // TMA warp, MMA warp, 8 softmax warps
constexpr int AT_SOFTMAX_THREADS = 256;
constexpr int AT_SK = 3, AT_SV = 3;                    // K / V ring depths
constexpr int AT_QBYTES = AT_BQ * AT_HD * 2;           // 16 KB
constexpr int AT_TILE = AT_BK * AT_HD * 2;             // 8 KB: K_j box, V_j box
constexpr int AT_PBYTES = AT_BQ * AT_BK * 2;           // 16 KB: P_j as one [128][64] K-major block
constexpr int AT_OFF_Q = 0, AT_OFF_K = AT_QBYTES, AT_OFF_V = AT_OFF_K + AT_SK * AT_TILE, AT_OFF_P = AT_OFF_V + AT_SV * AT_TILE;
constexpr int AT_OFF_MAX = AT_OFF_P + 2 * AT_PBYTES;   // row maxima of the two key halves, double-buffered: [2][2][128] f32
constexpr int AT_OFF_SUM = AT_OFF_MAX + 2 * 2 * AT_BQ * 4;  // row sums of the two halves at the end: [2][128] f32
constexpr int AT_OFF_BAR = AT_OFF_SUM + 2 * AT_BQ * 4;
constexpr int AT_SMEM_BYTES = AT_OFF_BAR + 256 + 1024 /*align slack*/;
constexpr int AT_TMEM_COLS = 256;
static_assert(AT_BK == 64 && AT_HD == 64, "operand descriptors below assume one 128-byte swizzle atom per row");

struct AtParams {
    int T, d, n_kb;
    int p_tmem;  // 1: P_j goes to the MMA through tensor memory (over the S_j columns it was computed from) instead of shared memory
    int o_lazy;  // 1: O accumulates in tensor memory over all key blocks; a row's accumulator is rescaled only when its maximum grows by > 8 (p <= 2^8)
    uint32_t idesc_qk, idesc_pv;
    void* out;
};

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {  // arguments are <= 0 here; flush-to-zero for very negative ones is what we want
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <typename T, bool kLazy>
__global__ void __launch_bounds__(AT_THREADS, 2)
encoder_attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AtParams p) {
    extern __shared__ uint8_t at_smem_raw[];
    const uint32_t smem_base = (smem_u32(at_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = at_smem_raw + (smem_base - smem_u32(at_smem_raw));
    const uint32_t bar_base = smem_base + AT_OFF_BAR;
    // barriers (8 B each)
    const uint32_t q_full = bar_base;
    auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto k_empty = [&](int s) { return bar_base + 8u * (1 + AT_SK + s); };
    auto v_full = [&](int s) { return bar_base + 8u * (1 + 2 * AT_SK + s); };
    auto v_empty = [&](int s) { return bar_base + 8u * (1 + 2 * AT_SK + AT_SV + s); };
    constexpr int B0 = 1 + 2 * AT_SK + 2 * AT_SV;
    auto s_full = [&](int i) { return bar_base + 8u * (B0 + i); };
    auto p_ready = [&](int i) { return bar_base + 8u * (B0 + 2 + i); };
    auto o_full = [&](int i) { return bar_base + 8u * (B0 + 4 + i); };
    auto o_empty = [&](int i) { return bar_base + 8u * (B0 + 6 + i); };
    const uint32_t tmem_slot = bar_base + 8u * (B0 + 8);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + AT_OFF_BAR + 8 * (B0 + 8));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int n_kb = p.n_kb;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_kv)) : "memory");
        mbar_init(q_full, 1);
        for (int s = 0; s < AT_SK; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(k_empty(s), 1);
        }
        for (int s = 0; s < AT_SV; ++s) {
            mbar_init(v_full(s), 1);
            mbar_init(v_empty(s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s_full(i), 1);
            mbar_init(p_ready(i), AT_SOFTMAX_THREADS);
            mbar_init(o_full(i), 1);
            mbar_init(o_empty(i), AT_SOFTMAX_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)AT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int cq = h * AT_HD, ck = p.d + h * AT_HD, cv = 2 * p.d + h * AT_HD;
            mbar_arrive_expect_tx(q_full, AT_QBYTES);
            tma_load_3d(&map_q, smem_base + AT_OFF_Q, q_full, cq, q0, b);
            int sk = 0, sv = 0;
            uint32_t pk = 0, pv = 0;
            for (int j = 0; j < n_kb; ++j) {
                mbar_wait(k_empty(sk), pk ^ 1);
                mbar_arrive_expect_tx(k_full(sk), AT_TILE);
                tma_load_3d(&map_kv, smem_base + AT_OFF_K + sk * AT_TILE, k_full(sk), ck, j * AT_BK, b);
                if (++sk == AT_SK) {
                    sk = 0;
                    pk ^= 1;
                }
                mbar_wait(v_empty(sv), pv ^ 1);
                mbar_arrive_expect_tx(v_full(sv), AT_TILE);
                tma_load_3d(&map_kv, smem_base + AT_OFF_V + sv * AT_TILE, v_full(sv), cv, j * AT_BK, b);
                if (++sv == AT_SV) {
                    sv = 0;
                    pv ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint64_t dq = make_smem_desc(smem_base + AT_OFF_Q);
            int sk = 0, sv = 0;
            uint32_t pk = 0, pv = 0;
            auto issue_qk = [&](int j) {  // S_j = Q K_j^T into S buffer j%2
                mbar_wait(k_full(sk), pk);
                tc_fence_after();
                const uint64_t dk = make_smem_desc(smem_base + AT_OFF_K + sk * AT_TILE);
                const uint32_t tS = tmem_base + (uint32_t)((j & 1) * AT_BK);
#pragma unroll
                for (int k = 0; k < AT_HD / 16; ++k) tc_mma_f16(tS, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), p.idesc_qk, k != 0 ? 1u : 0u);
                tc_commit(k_empty(sk));
                tc_commit(s_full(j & 1));
                if (++sk == AT_SK) {
                    sk = 0;
                    pk ^= 1;
                }
            };
            mbar_wait(q_full, 0);
            tc_fence_after();
            issue_qk(0);
            if (n_kb > 1) issue_qk(1);
            for (int j = 0; j < n_kb; ++j) {
                const int i = j & 1;
                const uint32_t ph = (uint32_t)((j >> 1) & 1);
                mbar_wait(p_ready(i), ph);  // P_j is in shared memory, S buffer i has been read out
                mbar_wait(v_full(sv), pv);
                if (j >= 2 && !kLazy) mbar_wait(o_empty(i), ph ^ 1);  // O_{j-2} has been folded into registers
                tc_fence_after();
                const uint64_t dp = make_smem_desc(smem_base + AT_OFF_P + i * AT_PBYTES);
                const uint64_t dv = make_smem_desc(smem_base + AT_OFF_V + sv * AT_TILE);
                // lazy mode: one accumulator for the whole row of key blocks (block 0 overwrites, the others add)
                const uint32_t tO = tmem_base + (uint32_t)(2 * AT_BK + (kLazy ? 0 : i * AT_HD));
                const uint32_t acc0 = (kLazy && j > 0) ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < AT_BK / 16; ++ks) {
                    // A: P ([128][64] K-major): in shared memory 32 B per k-step inside the swizzle atom; in tensor memory the 16
                    //    keys of a k-step are 8 packed columns of the S buffer P_j replaced
                    // B: V is MN-major ([key][64 channels], one 128-byte swizzle row per key): 16 keys = 2048 B per k-step
                    if (p.p_tmem) tc_mma_f16_ts(tO, tmem_base + (uint32_t)(i * AT_BK + ks * 8), dv + (uint64_t)(ks * (2048 >> 4)), p.idesc_pv, ks != 0 ? 1u : acc0);
                    else tc_mma_f16(tO, dp + (uint64_t)(ks * 2), dv + (uint64_t)(ks * (2048 >> 4)), p.idesc_pv, ks != 0 ? 1u : acc0);
                }
                tc_commit(v_empty(sv));
                tc_commit(o_full(i));
                if (++sv == AT_SV) {
                    sv = 0;
                    pv ^= 1;
                }
                if (j + 2 < n_kb) issue_qk(j + 2);
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax =====================
        const int lane_grp = warp & 3;            // TMEM lanes [32 * lane_grp, +32) are the ones this warp may read
        const int half = (warp - 2) >> 2;         // which 32 keys of a block / which 32 output channels this thread owns
        const int row = lane_grp * 32 + lane;     // query row inside the tile = TMEM lane
        const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
        const uint32_t pair_bar = 1u + (uint32_t)lane_grp;  // named barrier of the two warps that share these rows
        const float c = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e): scores live in the exp2 domain
        float m_run = -1e30f, l_run = 0.0f;
        float o[AT_HD / 2];
#pragma unroll
        for (int i = 0; i < AT_HD / 2; ++i) o[i] = 0.0f;
        uint8_t* Pgen = smem_gen + AT_OFF_P;
        float* sm_max = reinterpret_cast<float*>(smem_gen + AT_OFF_MAX);
        float* sm_sum = reinterpret_cast<float*>(smem_gen + AT_OFF_SUM);
        const int c0 = half * 32;                 // first key (and first channel) of this thread's half
        for (int j = 0; j < n_kb; ++j) {
            const int i = j & 1;
            const uint32_t ph = (uint32_t)((j >> 1) & 1);
            const int valid = min(AT_BK, p.T - j * AT_BK);  // keys of this block that exist
            mbar_wait(s_full(i), ph);
            tc_fence_after();
            const uint32_t tS = tmem_base + lane_addr + (uint32_t)(i * AT_BK + c0);
            uint32_t v[32];
            tc_ld_32x32(tS, v);
            tc_wait_ld();
            float mx = -1e30f;
            if (valid == AT_BK) {
#pragma unroll
                for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(v[k]));
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (c0 + k < valid) mx = fmaxf(mx, __uint_as_float(v[k]));
            }
            // the two halves of the row agree on the block's maximum (double-buffered by block parity: a thread may be one
            // block ahead of its partner, never two)
            sm_max[(i * 2 + half) * AT_BQ + row] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            mx = fmaxf(mx, sm_max[(i * 2 + (half ^ 1)) * AT_BQ + row]);
            float m_new = fmaxf(m_run, mx * c);
            float corr = ex2_approx(m_run - m_new);
            if (kLazy) {
                // The accumulator stays in tensor memory.  P_j is computed against the maximum the row's accumulator is scaled
                // to (m_run); only when the row's true maximum has grown by more than 8 (p would exceed 2^8) is the accumulator
                // pulled, rescaled and put back — rows are independent TMEM lanes, P_{j-1} V_{j-1} has completed (o_full) and
                // P_j V_j is not issued before this thread arrives on p_ready, so nobody else touches these columns meanwhile.
                // tcgen05.ld/st are warp-wide, so the warp rescales if any of its rows has to (the others multiply by one).  Its
                // partner warp sees the same maxima and takes the same decision for its 32 channels.
                bool grow = j > 0 && m_new > m_run + 8.0f;
                if (j > 0 && !grow) {
                    m_new = m_run;  // keep the scale
                    corr = 1.0f;
                }
                if (j == 0) corr = 1.0f;  // nothing accumulated yet
                if (__any_sync(0xffffffffu, grow)) {
                    const int ip = (j - 1) & 1;
                    mbar_wait(o_full(ip), (uint32_t)(((j - 1) >> 1) & 1));
                    tc_fence_after();
                    const uint32_t tOl = tmem_base + lane_addr + (uint32_t)(2 * AT_BK + c0);
                    uint32_t w[32];
                    tc_ld_32x32(tOl, w);
                    tc_wait_ld();
#pragma unroll
                    for (int k = 0; k < 32; ++k) w[k] = __float_as_uint(__uint_as_float(w[k]) * corr);
                    tc_st_32x32(tOl, w);
                    tc_wait_st();
                }
            }
            m_run = m_new;
            // p, row sum, this half of P_j (rounded to the activation type) into the swizzled K-major operand layout
            float lsum = 0.0f;
            uint8_t* Prow = Pgen + i * AT_PBYTES + row * 128;
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(v[k]), c, -m_new));
                float p1 = ex2_approx(fmaf(__uint_as_float(v[k + 1]), c, -m_new));
                if (valid != AT_BK) {
                    if (c0 + k >= valid) p0 = 0.0f;
                    if (c0 + k + 1 >= valid) p1 = 0.0f;
                }
                lsum += p0 + p1;
                pk[k >> 1] = pack2<T>(p0, p1);
            }
            if (p.p_tmem) {
                // both halves of the row hold their S values in registers (the pair barrier above came after the loads), so
                // the first 32 columns of the S buffer can take the row's 64 packed P values: this thread's 32 keys = 16 columns
                tc_st_32x16(tmem_base + lane_addr + (uint32_t)(i * AT_BK + half * 16), pk);
                tc_wait_st();
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ch = (c0 >> 3) + q;  // 16-byte chunk (8 keys) inside the 128-byte row
                    *reinterpret_cast<uint4*>(Prow + ((ch ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P_j (generic-proxy stores) -> visible to the MMA
            }
            l_run = l_run * corr + lsum;
            tc_fence_before();
            mbar_arrive(p_ready(i));
            // fold the previous block's partial output (same scale as o), then move both to the new maximum
            if (j >= 1 && !kLazy) {
                const int ip = (j - 1) & 1;
                mbar_wait(o_full(ip), (uint32_t)(((j - 1) >> 1) & 1));
                tc_fence_after();
                const uint32_t tO = tmem_base + lane_addr + (uint32_t)(2 * AT_BK + ip * AT_HD + c0);
                uint32_t w[32];
                tc_ld_32x32(tO, w);
                tc_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) o[k] = (o[k] + __uint_as_float(w[k])) * corr;
                tc_fence_before();
                mbar_arrive(o_empty(ip));
            }
        }
        {
            const int ip = (n_kb - 1) & 1;
            mbar_wait(o_full(ip), (uint32_t)(((n_kb - 1) >> 1) & 1));
            tc_fence_after();
            const uint32_t tO = tmem_base + lane_addr + (uint32_t)(2 * AT_BK + (kLazy ? 0 : ip * AT_HD) + c0);
            uint32_t w[32];
            tc_ld_32x32(tO, w);
            tc_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] += __uint_as_float(w[k]);
        }
        // the row's normaliser is the sum of the two halves' running sums (both carry the same running maximum)
        sm_sum[half * AT_BQ + row] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float l_tot = l_run + sm_sum[(half ^ 1) * AT_BQ + row];
        const int t = q0 + row;
        if (t < p.T) {
            const float inv = 1.0f / l_tot;
            T* orow = reinterpret_cast<T*>(p.out) + ((size_t)b * p.T + t) * p.d + h * AT_HD + c0;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
                uint4 u;
                u.x = pack2<T>(o[k] * inv, o[k + 1] * inv);
                u.y = pack2<T>(o[k + 2] * inv, o[k + 3] * inv);
                u.z = pack2<T>(o[k + 4] * inv, o[k + 5] * inv);
                u.w = pack2<T>(o[k + 6] * inv, o[k + 7] * inv);
                *reinterpret_cast<uint4*>(orow + k) = u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)AT_TMEM_COLS) : "memory");
    }
}

PerDeviceInit g_at_init;

}  // namespace

bool encoder_attention_tc_available() {
    return g_at_init.ensure([] {
        return gemm_tc_available() &&
               cudaFuncSetAttribute(encoder_attention_tc_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES) == cudaSuccess &&
               cudaFuncSetAttribute(encoder_attention_tc_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES) == cudaSuccess &&
               cudaFuncSetAttribute(encoder_attention_tc_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES) == cudaSuccess &&
               cudaFuncSetAttribute(encoder_attention_tc_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES) == cudaSuccess;
    });
}

int launch_encoder_attention_tc(const void* qkv, void* out, int type, int B, int Tn, int d, cudaStream_t st) {
    if (type == DT_F32 || !encoder_attention_tc_available()) return OH_ERR_UNSUPPORTED;
    if (d % AT_HD != 0 || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || Tn < 1) return OH_ERR_UNSUPPORTED;
    CUtensorMap map_q, map_kv;
    const uint64_t dims[3] = {(uint64_t)(3 * d), (uint64_t)Tn, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)(3 * d) * 2, (uint64_t)Tn * 3 * d * 2};
    const uint32_t box_q[3] = {AT_HD, AT_BQ, 1}, box_kv[3] = {AT_HD, AT_BK, 1};
    const CUtensorMapDataType dt = type == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (!tc_encode_tiled(&map_q, dt, 3, qkv, dims, strides, box_q) || !tc_encode_tiled(&map_kv, dt, 3, qkv, dims, strides, box_kv)) return OH_ERR_UNSUPPORTED;
    AtParams p;
    p.T = Tn; p.d = d; p.n_kb = (Tn + AT_BK - 1) / AT_BK;
    const uint32_t fmt = type == DT_BF16 ? 1u : 0u;
    const uint32_t base = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(AT_BQ >> 4) << 24);
    p.idesc_qk = base | ((uint32_t)(AT_BK >> 3) << 17);
    p.idesc_pv = base | ((uint32_t)(AT_HD >> 3) << 17) | (1u << 16);  // bit 16: B operand is MN-major
    p.out = out;
    // P_j reaches the second MMA through tensor memory (A operand in TMEM), written over the S_j columns it came from: 6x fewer
    // shared-memory wavefronts than the shared-memory route and 4 % faster on B200; OH_ATTN_P=smem selects the latter
    static const int p_tmem = [] {
        const char* e = getenv("OH_ATTN_P");
        return (e && strcmp(e, "smem") == 0) ? 0 : 1;
    }();
    p.p_tmem = p_tmem;
    // Accumulator: f16 keeps it in tensor memory (rescaled on demand); bf16 folds it into registers per block — with 8 bits
    // of mantissa P's rounding is the larger error term, and the fold path gives the row's dominant key exactly p = 1.
    // OH_ATTN_O=lazy / fold overrides the choice.
    static const int o_mode = [] {
        const char* e = getenv("OH_ATTN_O");
        return !e ? -1 : strcmp(e, "lazy") == 0 ? 1 : strcmp(e, "fold") == 0 ? 0 : -1;
    }();
    p.o_lazy = o_mode >= 0 ? o_mode : (type == DT_F16 ? 1 : 0);
    dim3 grid((Tn + AT_BQ - 1) / AT_BQ, d / AT_HD, B);
    ProfScope ps(PC_ATTN_ENC, 4.0 * B * (double)Tn * Tn * d, st);
    if (p.o_lazy) {
        if (type == DT_F16) OH_CUDA_CHECK(launch_k(encoder_attention_tc_kernel<__half, true>, grid, AT_THREADS, AT_SMEM_BYTES, st, map_q, map_kv, p));
        else OH_CUDA_CHECK(launch_k(encoder_attention_tc_kernel<__nv_bfloat16, true>, grid, AT_THREADS, AT_SMEM_BYTES, st, map_q, map_kv, p));
    } else {
        if (type == DT_F16) OH_CUDA_CHECK(launch_k(encoder_attention_tc_kernel<__half, false>, grid, AT_THREADS, AT_SMEM_BYTES, st, map_q, map_kv, p));
        else OH_CUDA_CHECK(launch_k(encoder_attention_tc_kernel<__nv_bfloat16, false>, grid, AT_THREADS, AT_SMEM_BYTES, st, map_q, map_kv, p));
    }
    OH_COUNT_LAUNCH();
    return OH_OK;
}

// start-up check against the mma.sync flash kernel: T with a ragged last block, two heads, two windows; the keys grow with
// time so that row maxima rise by far more than 2^8 from block to block (the on-demand rescale of the lazy accumulator runs)
int encoder_attention_tc_selftest(cudaStream_t st, float* max_rel_err) {
    if (!encoder_attention_tc_available()) return OH_ERR_UNSUPPORTED;
    const int B = 2, Tn = 300, d = 128;
    const size_t n_qkv = (size_t)B * Tn * 3 * d, n_out = (size_t)B * Tn * d;
    std::vector<__half> h(n_qkv);
    uint32_t s = 99u;
    for (size_t i = 0; i < n_qkv; ++i) {
        s = s * 1664525u + 1013904223u;
        const int t = (int)((i / (size_t)(3 * d)) % Tn), part = (int)((i / d) % 3);
        const float gain = part == 1 ? 0.25f + (float)t / 40.0f : 2.0f;
        h[i] = __float2half_rn((((s >> 8) & 0xFFFF) / 32768.0f - 1.0f) * gain);
    }
    void *dq = nullptr, *d0 = nullptr, *d1 = nullptr;
    OH_CUDA_CHECK(cudaMalloc(&dq, n_qkv * 2));
    OH_CUDA_CHECK(cudaMalloc(&d0, n_out * 2));
    OH_CUDA_CHECK(cudaMalloc(&d1, n_out * 2));
    OH_CUDA_CHECK(cudaMemcpy(dq, h.data(), n_qkv * 2, cudaMemcpyHostToDevice));
    OH_CUDA_CHECK(cudaMemset(d0, 0, n_out * 2));
    OH_CUDA_CHECK(cudaMemset(d1, 0, n_out * 2));
    int rc = launch_encoder_attention_mma(dq, d0, DT_F16, B, Tn, d, st);
    if (rc == OH_OK) rc = launch_encoder_attention_tc(dq, d1, DT_F16, B, Tn, d, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    float e = -1.0f;
    if (rc == OH_OK && ce == cudaSuccess) {
        std::vector<__half> r0(n_out), r1(n_out);
        cudaMemcpy(r0.data(), d0, n_out * 2, cudaMemcpyDeviceToHost);
        cudaMemcpy(r1.data(), d1, n_out * 2, cudaMemcpyDeviceToHost);
        double num = 0.0, den = 0.0;
        for (size_t i = 0; i < n_out; ++i) {
            const double a = __half2float(r0[i]), bb = __half2float(r1[i]);
            num += (a - bb) * (a - bb);
            den += a * a;
        }
        e = (float)std::sqrt(num / std::max(den, 1e-30));
        if (!(e < 5e-3f)) fprintf(stderr, "[ohwhisper] tcgen05 attention self-test: rel err %g\n", e);
    }
    cudaFree(dq); cudaFree(d0); cudaFree(d1);
    if (rc != OH_OK) return rc;
    if (ce != cudaSuccess) {
        set_error(std::string("encoder_attention_tc_selftest: ") + cudaGetErrorString(ce));
        return OH_ERR_CUDA;
    }
    if (max_rel_err) *max_rel_err = e;
    return (e >= 0.0f && e < 5e-3f) ? OH_OK : OH_ERR_CUDA;
}

}  // namespace oh
