// K4 (decode) on the 5th-generation tensor cores — mat-muls of the autoregressive step with M = resident decoder rows
// (<= 256).  The weight matrix is read exactly once; what bounds a plain tiled kernel at this shape is that EVERY CTA
// has to pull all of X[M][K] through its own L2 port (M x K x 2 bytes per SM, 1.3 MB for the MLP down-projection).
// Here a thread-block CLUSTER of CS CTAs shares one 128 x 64 output tile and splits K, so a CTA ingests only
// X[:, K/CS] and W[64][K/CS]; the CS partial tiles meet in distributed shared memory:
//
//   warp 0      TMA producer: k-blocks (64 wide) of this CTA's K-range, A box(es) [128][64] + W box [64][64], 4-stage ring
//               (M > 128: two A boxes per stage share the W box, 2-stage ring — 256 rows cost one pass over W)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M128 N64 K16, f32 accumulator in TMEM)
//   warps 2..9  tcgen05.ld the 128 x 64 partial tile(s) into shared memory (aliases the drained ring); two warps per TMEM
//               lane group, 32 columns each
//   all warps   request the residual values of the elements they will write (they do not depend on the partial sums)
//   cluster.sync
//   all warps   CTA r sums rows [r*128/CS, +128/CS) of the CS partial tiles in rank order (deterministic; no partial
//               sum touches HBM) into its own copy — eight 16-byte loads in flight per thread, remote shared memory
//               answers in ~215 cycles — arrives on the cluster barrier, then runs the epilogue from local shared memory
//               (bias / GELU / residual / type conversion, 16-byte coalesced stores) and only then waits on the barrier,
//               so the barrier latency hides under the stores
//
// The K-split depends on (N, K) only, never on M, and tensor-core rows are independent of each other, so a row's
// result does not depend on which other rows share the batch.
// (reference arithmetic: whisper.cpp decoder mat-muls, f16 operands with f32 accumulation, reached
// from /root/reference/src/engine/whisper.rs:266.)
#include <cooperative_groups.h>
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <vector>

#include "../../include/oh_whisper.h"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace cg = cooperative_groups;

namespace oh {

namespace {

using namespace tcx;

constexpr int SPK_M = 128, SPK_N = 64, SPK_K = 64, SPK_UMMA_K = 16;
constexpr int SPK_THREADS = 320;  // producer warp, MMA warp, 8 warps that drain TMEM; all ten run the reduction and the epilogue
constexpr int SPK_MAX_IT = 7;     // float4 items per thread in the epilogue for which the residual is prefetched (256 rows, CS >= 2)
constexpr int SPK_A_BYTES = SPK_M * SPK_K * 2, SPK_B_BYTES = SPK_N * SPK_K * 2;
constexpr int SPK_PS = SPK_N + 4;  // row stride (floats) of the partial tile: 16-byte aligned, spreads rows over banks
// MT = 128-row tiles of X per CTA (1 for M <= 128, 2 for M <= 256): both tiles share every W box, so 256 rows cost one
// pass over the weights and one cluster per output column block, not two
template <int MT>
struct SpkCfg {
    static constexpr int STAGES = MT == 1 ? 4 : 2;  // 96 KB / 80 KB of ring: two CTAs per SM either way
    static constexpr int STAGE_BYTES = MT * SPK_A_BYTES + SPK_B_BYTES;
    static constexpr int TMEM_COLS = 64 * MT;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;
    static constexpr int ROWS = MT * SPK_M;
    static_assert(ROWS * SPK_PS * 4 <= STAGES * STAGE_BYTES, "partial tile must fit in the drained ring");
};

struct SpkParams {
    int M, N, K, num_kb;
    uint32_t idesc;
    const float* bias;
    void* C;
    int c_type, ldc;
    const float* R;
    int ldr;
    int gelu, round_type;
    const uint16_t* lut;
};

// sums, for the rows this CTA owns, the CS partial tiles of the cluster in rank order; 8 / CS items per batch so that a
// thread has eight 16-byte (mostly remote) loads in flight before the first add
template <int CS>
__device__ __forceinline__ void spk_reduce_rows(cg::cluster_group& cluster, float* Ps, int crank, int rows_per, int n_items, int m_lim, int n_blk, int N) {
    constexpr int PER = 8 / CS;
    const float* remote[CS];
#pragma unroll
    for (int c = 0; c < CS; ++c) remote[c] = cluster.map_shared_rank(Ps, c);
    for (int i0 = threadIdx.x; i0 < n_items; i0 += SPK_THREADS * PER) {
        float4 part[PER][CS];
        bool ok[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int idx = i0 + j * SPK_THREADS;
            const int rl = crank * rows_per + idx / (SPK_N / 4), c4 = (idx % (SPK_N / 4)) * 4;
            ok[j] = idx < n_items && rl < m_lim && n_blk * SPK_N + c4 < N;
            if (ok[j]) {
#pragma unroll
                for (int c = 0; c < CS; ++c) part[j][c] = *reinterpret_cast<const float4*>(remote[c] + (size_t)rl * SPK_PS + c4);
            }
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (!ok[j]) continue;
            const int idx = i0 + j * SPK_THREADS;
            const int rl = crank * rows_per + idx / (SPK_N / 4), c4 = (idx % (SPK_N / 4)) * 4;
            float4 f = part[j][0];
#pragma unroll
            for (int c = 1; c < CS; ++c) {
                f.x += part[j][c].x; f.y += part[j][c].y; f.z += part[j][c].z; f.w += part[j][c].w;
            }
            *reinterpret_cast<float4*>(Ps + (size_t)rl * SPK_PS + c4) = f;  // the same thread reads it back in phase 2
        }
    }
}

template <int MT>
__global__ void __launch_bounds__(SPK_THREADS, 2)
gemm_splitk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const SpkParams p) {
    constexpr int SPK_STAGES = SpkCfg<MT>::STAGES, SPK_STAGE_BYTES = SpkCfg<MT>::STAGE_BYTES, SPK_TMEM_COLS = SpkCfg<MT>::TMEM_COLS,
                  SPK_ROWS = SpkCfg<MT>::ROWS;
    constexpr int SPK_W_OFF = MT * SPK_A_BYTES;  // stage layout: A tile 0 [, A tile 1], W
    extern __shared__ uint8_t spk_smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int n_blk = blockIdx.x / cs, m_blk = blockIdx.y;
    const uint32_t smem_base = (smem_u32(spk_smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem_gen = spk_smem_raw + (smem_base - smem_u32(spk_smem_raw));
    const uint32_t bar_base = smem_base + SPK_STAGES * SPK_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (SPK_STAGES + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * SPK_STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * SPK_STAGES + 1);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    float* Ps = reinterpret_cast<float*>(smem_gen);  // partial tile [128*MT][SPK_PS], valid once the MMAs have drained the ring

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // this CTA's k-blocks: an even split of the num_kb blocks in rank order (a function of K and CS only)
    const int kb_base = p.num_kb / cs, kb_rem = p.num_kb % cs;
    const int kb0 = crank * kb_base + min(crank, kb_rem);
    const int n_kb = kb_base + (crank < kb_rem ? 1 : 0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        for (int s = 0; s < SPK_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)SPK_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // programmatic dependent launch: the weights do not depend on the preceding kernel, so the first ring of W
            // boxes is requested before pdl_wait(); the activations follow once the predecessor has completed
            const int n_pre = min(n_kb, SPK_STAGES);
            for (int i = 0; i < n_pre; ++i) {
                mbar_arrive_expect_tx(full_bar(i), SPK_STAGE_BYTES);
                tma_load_2d(&map_w, smem_base + i * SPK_STAGE_BYTES + SPK_W_OFF, full_bar(i), (kb0 + i) * SPK_K, n_blk * SPK_N);
            }
            pdl_wait();
            pdl_launch_dependents();
            for (int i = 0; i < n_pre; ++i)
#pragma unroll
                for (int t = 0; t < MT; ++t)
                    tma_load_2d(&map_a, smem_base + i * SPK_STAGE_BYTES + t * SPK_A_BYTES, full_bar(i), (kb0 + i) * SPK_K, m_blk * SPK_ROWS + t * SPK_M);
            int stage = n_pre == SPK_STAGES ? 0 : n_pre;
            uint32_t phase = n_pre == SPK_STAGES ? 1 : 0;
            for (int i = n_pre; i < n_kb; ++i) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                const uint32_t sa = smem_base + stage * SPK_STAGE_BYTES, sb = sa + SPK_W_OFF;
                mbar_arrive_expect_tx(full_bar(stage), SPK_STAGE_BYTES);
                const int kc = (kb0 + i) * SPK_K;
                tma_load_2d(&map_w, sb, full_bar(stage), kc, n_blk * SPK_N);
#pragma unroll
                for (int t = 0; t < MT; ++t) tma_load_2d(&map_a, sa + t * SPK_A_BYTES, full_bar(stage), kc, m_blk * SPK_ROWS + t * SPK_M);
                if (++stage == SPK_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && n_kb > 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_kb; ++i) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * SPK_STAGE_BYTES, sb = sa + SPK_W_OFF;
                const uint64_t db = make_smem_desc(sb);
#pragma unroll
                for (int t = 0; t < MT; ++t) {
                    const uint64_t da = make_smem_desc(sa + t * SPK_A_BYTES);
#pragma unroll
                    for (int k = 0; k < SPK_K / SPK_UMMA_K; ++k)
                        tc_mma_f16(tmem_base + (uint32_t)(t * SPK_N), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (i | k) != 0 ? 1u : 0u);
                }
                tc_commit(empty_bar(stage));
                if (++stage == SPK_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            tc_commit(tfull_bar);  // arrives once every MMA above has completed (operands consumed, accumulator final)
        }
        __syncwarp();
    } else {
        // partial tile TMEM -> shared memory; thread = one accumulator row, one half (32 columns) of the tile
        const int lane_grp = warp & 3;  // TMEM lanes [32*lane_grp, +32) are the only ones this warp may read
        const int c0 = ((warp - 2) >> 2) * 32;
        const int row = lane_grp * 32 + lane;
        if (n_kb > 0) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
        }
#pragma unroll
        for (int t = 0; t < MT; ++t) {
            float* prow = Ps + (size_t)(t * SPK_M + row) * SPK_PS;
            if (n_kb > 0) {
                uint32_t v[32];
                tc_ld_32x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t * SPK_N + c0), v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(prow + c0 + j) =
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {  // K-range empty (more ranks than k-blocks): contributes zeros
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    pdl_wait();  // every thread orders itself after the preceding kernel before it reads the residual / writes C
    pdl_launch_dependents();

    // this CTA's share of the output tile: rows [crank * rows_per, +rows_per), as float4 items
    const int rows_per = SPK_ROWS / cs;
    const int n_items = rows_per * (SPK_N / 4);
    const int m_lim = min(p.M - m_blk * SPK_ROWS, SPK_ROWS);
    // The residual does not depend on the partial sums: its loads go out before the cluster barrier and land while the
    // partial tiles are being summed (each output element is read and written by the same thread, so an in-place
    // residual update is not disturbed).
    float4 rpre[SPK_MAX_IT];
    const bool pre_ok = p.R != nullptr && n_items <= SPK_MAX_IT * SPK_THREADS;
    if (pre_ok) {
#pragma unroll
        for (int it = 0; it < SPK_MAX_IT; ++it) {
            const int idx = (int)threadIdx.x + it * SPK_THREADS;
            const int rl = crank * rows_per + idx / (SPK_N / 4), n = n_blk * SPK_N + (idx % (SPK_N / 4)) * 4;
            if (idx < n_items && rl < m_lim && n + 4 <= p.N)
                rpre[it] = *reinterpret_cast<const float4*>(p.R + (size_t)(m_blk * SPK_ROWS + rl) * p.ldr + n);
        }
    }
    cluster.sync();  // every partial tile is complete and visible cluster-wide
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)SPK_TMEM_COLS) : "memory");
    }

    // ---- phase 1: sum the CS partials of this CTA's rows in rank order (deterministic) and park the sums in its own copy
    // of those rows — peers only ever read the rows THEY own, so the in-place update races with nobody.  Eight 16-byte
    // loads (8 / CS items x CS ranks) are in flight per thread before the first add: remote shared memory answers in
    // ~215 cycles, and one item at a time would pay that once per item.
    switch (cs) {
        case 8: spk_reduce_rows<8>(cluster, Ps, crank, rows_per, n_items, m_lim, n_blk, p.N); break;
        case 4: spk_reduce_rows<4>(cluster, Ps, crank, rows_per, n_items, m_lim, n_blk, p.N); break;
        case 2: spk_reduce_rows<2>(cluster, Ps, crank, rows_per, n_items, m_lim, n_blk, p.N); break;
        default: break;  // one rank: the partial tile is the sum
    }
    cluster.barrier_arrive();  // this CTA no longer reads remote shared memory; the matching wait is the last statement

    // ---- phase 2: epilogue from local shared memory (overlaps the cluster barrier)
    auto epilogue_item = [&](int idx, bool have_r, const float4& rr_pre) {
        const int rl = crank * rows_per + idx / (SPK_N / 4), c4 = (idx % (SPK_N / 4)) * 4;
        const int n = n_blk * SPK_N + c4;
        if (rl >= m_lim || n >= p.N) return;
        const float4 f4 = *reinterpret_cast<const float4*>(Ps + (size_t)rl * SPK_PS + c4);
        float f[4] = {f4.x, f4.y, f4.z, f4.w};
        const int r = m_blk * SPK_ROWS + rl;
        const bool full = n + 4 <= p.N;
        if (full) {
            if (p.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                f[0] += bb.x; f[1] += bb.y; f[2] += bb.z; f[3] += bb.w;
            }
            if (p.gelu != GELU_NONE) {
#pragma unroll
                for (int j = 0; j < 4; ++j) f[j] = epi_gelu(f[j], p.gelu, p.round_type, p.lut);
            }
            if (p.R) {
                const float4 rr = have_r ? rr_pre : *reinterpret_cast<const float4*>(p.R + (size_t)r * p.ldr + n);
                f[0] += rr.x; f[1] += rr.y; f[2] += rr.z; f[3] += rr.w;
            }
            if (p.c_type == DT_F32) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (size_t)r * p.ldc + n) = make_float4(f[0], f[1], f[2], f[3]);
            } else if (p.c_type == DT_F16) {
                __half2 h[2] = {__floats2half2_rn(f[0], f[1]), __floats2half2_rn(f[2], f[3])};
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.C) + (size_t)r * p.ldc + n) = *reinterpret_cast<uint2*>(h);
            } else {
                __nv_bfloat162 h[2] = {__floats2bfloat162_rn(f[0], f[1]), __floats2bfloat162_rn(f[2], f[3])};
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)r * p.ldc + n) = *reinterpret_cast<uint2*>(h);
            }
        } else {
            for (int j = 0; j < 4 && n + j < p.N; ++j) {
                float v = f[j];
                if (p.bias) v += p.bias[n + j];
                if (p.gelu != GELU_NONE) v = epi_gelu(v, p.gelu, p.round_type, p.lut);
                if (p.R) v += p.R[(size_t)r * p.ldr + n + j];
                if (p.c_type == DT_F32) reinterpret_cast<float*>(p.C)[(size_t)r * p.ldc + n + j] = v;
                else if (p.c_type == DT_F16) reinterpret_cast<__half*>(p.C)[(size_t)r * p.ldc + n + j] = __float2half_rn(v);
                else reinterpret_cast<__nv_bfloat16*>(p.C)[(size_t)r * p.ldc + n + j] = __float2bfloat16_rn(v);
            }
        }
    };
    if (pre_ok) {
#pragma unroll
        for (int it = 0; it < SPK_MAX_IT; ++it) {
            const int idx = (int)threadIdx.x + it * SPK_THREADS;
            if (idx < n_items) epilogue_item(idx, true, rpre[it]);
        }
    } else {
        const float4 none = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int idx = threadIdx.x; idx < n_items; idx += SPK_THREADS) epilogue_item(idx, false, none);
    }
    cluster.barrier_wait();  // nobody leaves while a peer may still read its partial tile
}

std::atomic<PFN_cuTensorMapEncodeTiled_v12000> g_spk_encode{nullptr};
PerDeviceInit g_spk_init;  // the shared-memory opt-in is per device

bool spk_init_device() {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return false;
    g_spk_encode.store(reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, PerDeviceInit::current_device()) != cudaSuccess || prop.major != 10) return false;
    return cudaFuncSetAttribute(gemm_splitk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SpkCfg<1>::SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(gemm_splitk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SpkCfg<2>::SMEM_BYTES) == cudaSuccess;
}

bool spk_encode_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_outer) {
    const uint64_t dims[2] = {inner, outer};
    const uint64_t strides[1] = {ld_elems * 2};
    const uint32_t box[2] = {SPK_K, box_outer};
    const uint32_t estr[2] = {1, 1};
    return g_spk_encode.load()(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// cluster size for an (N, K) problem: the largest power of two <= 8 that keeps every rank at least one k-block and the
// grid within one resident wave (2 CTAs per SM)
int gemm_splitk_cluster_size(int N, int K) {
    static const int cap = [] {
        const char* e = getenv("OH_SPLITK_CS");
        return e ? std::max(1, std::min(8, atoi(e))) : 8;
    }();
    const int n_tiles = (N + SPK_N - 1) / SPK_N, num_kb = (K + SPK_K - 1) / SPK_K;
    int cs = 1;
    while (cs * 2 <= cap && cs * 2 <= num_kb && n_tiles * cs * 2 <= 296) cs *= 2;  // 296 = 2 resident CTAs x 148 SMs (MT = 1)
    return cs;
}

int launch_gemm_splitk(const GemmArgs& g, cudaStream_t st) {
    if (!g_spk_init.ensure(spk_init_device)) return OH_ERR_UNSUPPORTED;
    const int M = g.batch * g.rows;
    if (g.a_type == DT_F32 || g.a_type != g.w_type || g.batch != 1 || M < 1 || M > 256 || g.n_taps != 1 || g.c_hsplit != 0) return OH_ERR_UNSUPPORTED;
    const int c_es = g.c_type == DT_F32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15) || (g.lda % 8) || (g.K % 8) || g.lda < g.K ||
        ((size_t)g.ldc * c_es % 16) || (reinterpret_cast<uintptr_t>(g.C) & 15) || (g.R && ((g.ldr % 4) || (reinterpret_cast<uintptr_t>(g.R) & 15))) ||
        (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15)))
        return OH_ERR_UNSUPPORTED;
    const int n_tiles = (g.N + SPK_N - 1) / SPK_N;
    if (n_tiles > 296) return OH_ERR_UNSUPPORTED;  // wide outputs (the logit projection) stay on the persistent kernel
    const CUtensorMapDataType dt = g.a_type == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap ma, mw;
    if (!spk_encode_map(&ma, dt, g.A, (uint64_t)g.K, (uint64_t)M, (uint64_t)g.lda, SPK_M)) return OH_ERR_UNSUPPORTED;
    if (!spk_encode_map(&mw, dt, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.K, SPK_N)) return OH_ERR_UNSUPPORTED;
    SpkParams p;
    p.M = M; p.N = g.N; p.K = g.K; p.num_kb = (g.K + SPK_K - 1) / SPK_K;
    const uint32_t fmt = g.a_type == DT_BF16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(SPK_N >> 3) << 17) | ((uint32_t)(SPK_M >> 4) << 24);
    p.bias = g.bias; p.C = g.C; p.c_type = g.c_type; p.ldc = g.ldc; p.R = g.R; p.ldr = g.ldr;
    p.gelu = g.gelu; p.round_type = g.round_type; p.lut = g.lut;
    const int cs = gemm_splitk_cluster_size(g.N, g.K);
    cudaLaunchConfig_t cfg{};
    const int mt = M > SPK_M ? 2 : 1;
    cfg.gridDim = dim3(n_tiles * cs, (M + mt * SPK_M - 1) / (mt * SPK_M));
    cfg.blockDim = dim3(SPK_THREADS);
    cfg.dynamicSmemBytes = mt == 2 ? SpkCfg<2>::SMEM_BYTES : SpkCfg<1>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl > 0 ? 2 : 1;
    OH_SKIP_IF(PC_GEMM_DEC);
    ProfScope ps(PC_GEMM_DEC, 2.0 * M * (double)g.N * g.K, st);
    if (mt == 2) OH_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_splitk_kernel<2>, ma, mw, p));
    else OH_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_splitk_kernel<1>, ma, mw, p));
    OH_COUNT_LAUNCH();
    return OH_OK;
}

// start-up check against the SIMT kernel: every cluster size, M / N / K tails, GELU, residual, both output types
int gemm_splitk_selftest(cudaStream_t st, float* max_rel_err) {
    struct Case {
        int M, N, K, c_type, gelu;
    };
    const Case cases[7] = {{64, 1280, 1280, DT_F32, 0}, {100, 3840, 1280, DT_F16, 0}, {37, 5120, 1280, DT_F16, 1}, {128, 1280, 5120, DT_F32, 0},
                           {200, 1300, 264, DT_F32, 0}, {3, 384, 384, DT_F16, 1}, {256, 1280, 5120, DT_F16, 1}};  // M > 128: two row tiles per CTA
    float worst = 0.0f;
    for (const Case& c : cases) {
        std::vector<__half> hA((size_t)c.M * c.K), hW((size_t)c.N * c.K);
        std::vector<float> hb(c.N), hR((size_t)c.M * c.N);
        uint32_t s = 4242u + (uint32_t)c.N + (uint32_t)c.K;
        auto rnd = [&]() {
            s = s * 1664525u + 1013904223u;
            return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
        };
        for (auto& v : hA) v = __float2half_rn(rnd());
        for (auto& v : hW) v = __float2half_rn(rnd() * 0.1f);
        for (auto& v : hb) v = rnd();
        for (auto& v : hR) v = rnd();
        const size_t c_es = c.c_type == DT_F32 ? 4 : 2;
        void *dA = nullptr, *dW = nullptr, *db = nullptr, *dR = nullptr, *dC0 = nullptr, *dC1 = nullptr;
        OH_CUDA_CHECK(cudaMalloc(&dA, hA.size() * 2));
        OH_CUDA_CHECK(cudaMalloc(&dW, hW.size() * 2));
        OH_CUDA_CHECK(cudaMalloc(&db, hb.size() * 4));
        OH_CUDA_CHECK(cudaMalloc(&dR, hR.size() * 4));
        OH_CUDA_CHECK(cudaMalloc(&dC0, (size_t)c.M * c.N * c_es));
        OH_CUDA_CHECK(cudaMalloc(&dC1, (size_t)c.M * c.N * c_es));
        OH_CUDA_CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemcpy(dR, hR.data(), hR.size() * 4, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemset(dC0, 0, (size_t)c.M * c.N * c_es));
        OH_CUDA_CHECK(cudaMemset(dC1, 0, (size_t)c.M * c.N * c_es));
        GemmArgs g;
        g.A = dA; g.a_type = DT_F16; g.lda = c.K; g.rows = c.M; g.W = dW; g.w_type = DT_F16; g.N = c.N; g.K = c.K; g.bias = (const float*)db;
        g.c_type = c.c_type; g.ldc = c.N; g.R = (const float*)dR; g.ldr = c.N; g.gelu = c.gelu ? GELU_FAST : GELU_NONE; g.round_type = DT_F16;
        g.C = dC0;
        int rc = launch_gemm_simt(g, st);
        g.C = dC1;
        if (rc == OH_OK) rc = launch_gemm_splitk(g, st);
        cudaError_t ce = cudaStreamSynchronize(st);
        std::vector<float> r0((size_t)c.M * c.N), r1((size_t)c.M * c.N);
        if (rc == OH_OK && ce == cudaSuccess) {
            if (c.c_type == DT_F32) {
                cudaMemcpy(r0.data(), dC0, r0.size() * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(r1.data(), dC1, r1.size() * 4, cudaMemcpyDeviceToHost);
            } else {
                std::vector<__half> t0(r0.size()), t1(r0.size());
                cudaMemcpy(t0.data(), dC0, t0.size() * 2, cudaMemcpyDeviceToHost);
                cudaMemcpy(t1.data(), dC1, t1.size() * 2, cudaMemcpyDeviceToHost);
                for (size_t i = 0; i < t0.size(); ++i) {
                    r0[i] = __half2float(t0[i]);
                    r1[i] = __half2float(t1[i]);
                }
            }
            double num = 0.0, den = 0.0;
            for (size_t i = 0; i < r0.size(); ++i) {
                num += (double)(r0[i] - r1[i]) * (r0[i] - r1[i]);
                den += (double)r0[i] * r0[i];
            }
            const float e = (float)std::sqrt(num / std::max(den, 1e-30));
            if (!(e < 2e-3f)) fprintf(stderr, "[ohwhisper] split-K tcgen05 self-test case M=%d N=%d K=%d: rel err %g\n", c.M, c.N, c.K, e);
            worst = std::max(worst, e);
        }
        cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dR); cudaFree(dC0); cudaFree(dC1);
        if (rc != OH_OK) return rc;
        if (ce != cudaSuccess) {
            set_error(std::string("gemm_splitk_selftest: ") + cudaGetErrorString(ce));
            return OH_ERR_CUDA;
        }
    }
    if (max_rel_err) *max_rel_err = worst;
    return worst < 2e-3f ? OH_OK : OH_ERR_CUDA;
}

}  // namespace oh
