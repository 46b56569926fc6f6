// K3/K4 — dense contractions on the 5th-generation tensor cores (sm_100a): TMA -> shared memory (128B swizzle)
// -> tcgen05.mma with the f32 accumulator in TMEM -> tcgen05.ld epilogue (bias / GELU / residual / type) -> HBM.
// Used for the conv stem (as 3-tap mat-muls), QKV / out / MLP projections, cross-K/V and the logit projection
// (reference arithmetic: whisper.cpp mat-muls with f16 operands and f32 accumulation,
// reached from /root/reference/src/engine/whisper.rs:266).
//
// One persistent CTA per SM, 320 threads, warp-specialised:
//   warp 0      TMA producer: per k-block one A box [128 rows][64 k] and one W box [BLOCK_N rows][64 k]
//   warp 1      TMEM allocator + single-thread MMA issuer (4 x tcgen05.mma M128 N=BLOCK_N K16 per k-block)
//   warps 2..9  epilogue: warp % 4 selects the 32 TMEM lanes (= 32 output rows) it may read; the two warps of a
//               lane group alternate over the 32-column chunks of the tile (one tcgen05.ld each)
// Pipelines: smem ring (full/empty mbarriers, freed by tcgen05.commit), and two TMEM accumulator buffers so the
// epilogue of tile i overlaps the MMAs of tile i+1.
// Tiles: 128 x BLOCK_N, ordered n-fastest so the CTAs in flight share A rows and the whole of W through L2.
// Every mbarrier wait is bounded (trap after ~4 s) so a protocol bug cannot hang the GPU.
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

namespace oh {

namespace {

using namespace tcx;

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int TC_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int SMEM_BUDGET = 200 * 1024;

struct TcParams {
    int batch, rows, N, K;
    int tiles_per_batch, n_blocks, num_tiles, num_kb, kb_per_tap, n_taps, tap_k;
    uint32_t idesc;
    const float* bias;
    void* C;
    int c_type;
    long long c_bstride;
    int ldc;
    const float* R;
    long long r_bstride;
    int ldr;
    int gelu, round_type;
    const uint16_t* lut;
    int c_hsplit;
    long long c_hstride;
    int stages;
    const int* c_bidx;  // optional: batch b writes the C block of index c_bidx[b] (scattered K/V slots)
};

template <int BLOCK_N>
struct TcConfig {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_w, const TcParams p) {
    using Cfg = TcConfig<BLOCK_N>;
    const int NS = p.stages;  // ring depth of this launch (<= Cfg::STAGES): a shallower ring leaves shared memory to co-resident kernels
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bar_base = smem_base + NS * Cfg::STAGE_BYTES;
    // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();  // programmatic dependent launch: barrier / TMEM set-up above overlaps the preceding kernel (see common.cuh)
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int n_blk = tile % p.n_blocks, m_blk = tile / p.n_blocks;
                const int b = m_blk / p.tiles_per_batch, i0 = (m_blk % p.tiles_per_batch) * BLOCK_M;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    const int tap = kb / p.kb_per_tap, kc = (kb % p.kb_per_tap) * BLOCK_K;
                    const CUtensorMap* ma = tap == 0 ? &map_a0 : (tap == 1 ? &map_a1 : &map_a2);
                    tma_load_3d(ma, sa, full_bar(stage), kc, i0, b);
                    tma_load_2d(&map_w, sb, full_bar(stage), tap * p.tap_k + kc, n_blk * BLOCK_N);
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
                        tc_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(tfull_bar(acc));  // accumulator complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        // 8 warps: TMEM lanes [32*lane_grp, +32) are the only ones a warp may read (lane_grp = warp % 4); the two warps
        // of a lane group take alternate 32-column chunks.  The residual of the next chunk is loaded before the
        // current one is finished so its latency overlaps, and the GELU variant is chosen once per chunk so that
        // the 32 elements form independent instruction chains.
        const int lane_grp = warp & 3, col_half = (warp - 2) >> 2;
        constexpr int N_CHUNKS = BLOCK_N >= 32 ? BLOCK_N / 32 : 1;
        const int gelu_mode = p.gelu == GELU_FAST ? (p.round_type == DT_F16 ? 4 : (p.round_type == DT_BF16 ? 5 : GELU_FAST)) : p.gelu;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n_blk = tile % p.n_blocks, m_blk = tile / p.n_blocks;
            const int b = m_blk / p.tiles_per_batch, i0 = (m_blk % p.tiles_per_batch) * BLOCK_M;
            const int row = i0 + lane_grp * 32 + lane;
            const bool row_ok = row < p.rows;
            const long long c_off = (long long)(p.c_bidx ? __ldg(p.c_bidx + b) : b) * p.c_bstride + (long long)row * p.ldc;
            const float* rrow = (p.R && row_ok) ? p.R + (long long)b * p.r_bstride + (long long)row * p.ldr : nullptr;
            const int n_lim = min(p.N, (n_blk + 1) * BLOCK_N);  // columns of this tile only
            float4 rnext[8];
            auto res_prefetch = [&](int chunk) {
                const int n0 = n_blk * BLOCK_N + chunk * 32;
                if (rrow && chunk < N_CHUNKS && n0 + 32 <= n_lim) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) rnext[j] = *reinterpret_cast<const float4*>(rrow + n0 + 4 * j);
                }
            };
            res_prefetch(col_half);  // does not depend on the accumulator
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int chunk = col_half; chunk < N_CHUNKS; chunk += 2) {
                const int c0 = chunk * 32;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
                if constexpr (BLOCK_N >= 32) {
                    tc_ld_32x32(taddr, v);
                } else {  // 16-column accumulator: never touch columns outside it
                    tc_ld_32x16(taddr, v);
#pragma unroll
                    for (int j = 16; j < 32; ++j) v[j] = 0u;
                }
                float4 rcur[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
                res_prefetch(chunk + 2);
                tc_wait_ld();
                const int n0 = n_blk * BLOCK_N + c0;
                // column offset inside the output row (head-split stores keep the 32-column chunk inside one head)
                const long long n_off = p.c_hsplit > 0 ? (long long)(n0 / p.c_hsplit) * p.c_hstride + (n0 % p.c_hsplit) : (long long)n0;
                if (row_ok && n0 < p.N) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    const bool full = n0 + 32 <= n_lim;
                    if (p.bias) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                                f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < n_lim) f[j] += p.bias[n0 + j];
                        }
                    }
                    switch (gelu_mode) {  // warp-uniform; each case is straight-line code over the 32 elements
                        case GELU_LUT:
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = gelu_lut(f[j], p.lut);
                            break;
                        case GELU_TANH:
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = gelu_tanh(f[j]);
                            break;
                        case GELU_FAST:
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
                            break;
                        case 4:
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __half2float(__float2half_rn(gelu_fast(__half2float(__float2half_rn(f[j])))));
                            break;
                        case 5:
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __bfloat162float(__float2bfloat16_rn(gelu_fast(f[j])));
                            break;
                        default:
                            break;
                    }
                    if (rrow) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                f[4 * j] += rcur[j].x; f[4 * j + 1] += rcur[j].y; f[4 * j + 2] += rcur[j].z; f[4 * j + 3] += rcur[j].w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < n_lim) f[j] += rrow[n0 + j];
                        }
                    }
                    if (p.c_type == DT_F32) {
                        float* crow = reinterpret_cast<float*>(p.C) + c_off + n_off;
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < n_lim) crow[j] = f[j];
                        }
                    } else if (p.c_type == DT_F16) {
                        __half* crow = reinterpret_cast<__half*>(p.C) + c_off + n_off;
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                __half2 h[4] = {__floats2half2_rn(f[j], f[j + 1]), __floats2half2_rn(f[j + 2], f[j + 3]),
                                                __floats2half2_rn(f[j + 4], f[j + 5]), __floats2half2_rn(f[j + 6], f[j + 7])};
                                *reinterpret_cast<uint4*>(crow + j) = *reinterpret_cast<uint4*>(h);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < n_lim) crow[j] = __float2half_rn(f[j]);
                        }
                    } else {
                        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + c_off + n_off;
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                __nv_bfloat162 h[4] = {__floats2bfloat162_rn(f[j], f[j + 1]), __floats2bfloat162_rn(f[j + 2], f[j + 3]),
                                                       __floats2bfloat162_rn(f[j + 4], f[j + 5]), __floats2bfloat162_rn(f[j + 6], f[j + 7])};
                                *reinterpret_cast<uint4*>(crow + j) = *reinterpret_cast<uint4*>(h);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < n_lim) crow[j] = __float2bfloat16_rn(f[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------------
std::atomic<PFN_cuTensorMapEncodeTiled_v12000> g_encode{nullptr};
int g_num_sms[PerDeviceInit::kMaxDevices] = {};
PerDeviceInit g_tc_init;  // shared-memory opt-ins and the SM count are per device

bool tc_init_device() {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return false;
    g_encode.store(reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn));
    const int dev = PerDeviceInit::current_device();
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) return false;
    g_num_sms[dev] = prop.multiProcessorCount;
    bool ok = cudaFuncSetAttribute(gemm_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcConfig<16>::SMEM_BYTES) == cudaSuccess &&
              cudaFuncSetAttribute(gemm_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcConfig<32>::SMEM_BYTES) == cudaSuccess &&
              cudaFuncSetAttribute(gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcConfig<64>::SMEM_BYTES) == cudaSuccess &&
              cudaFuncSetAttribute(gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcConfig<128>::SMEM_BYTES) == cudaSuccess &&
              cudaFuncSetAttribute(gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcConfig<256>::SMEM_BYTES) == cudaSuccess;
    return ok;
}

bool encode_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box) {
    uint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode.load()(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BLOCK_N>
int launch_t(const GemmArgs& g, cudaStream_t st) {
    using Cfg = TcConfig<BLOCK_N>;
    const CUtensorMapDataType dt = g.a_type == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const int n_taps = g.n_taps, tap_k = n_taps > 1 ? g.tap_k : g.K;
    CUtensorMap ma[3], mw;
    for (int t = 0; t < 3; ++t) {
        const int tt = t < n_taps ? t : 0;
        const uint64_t dims[3] = {(uint64_t)tap_k, (uint64_t)g.rows, (uint64_t)g.batch};
        const uint64_t strides[2] = {(uint64_t)g.lda * 2, (uint64_t)(g.batch > 1 ? g.a_bstride : (long long)g.rows * g.lda) * 2};
        const uint32_t box[3] = {BLOCK_K, BLOCK_M, 1};
        if (!encode_map(&ma[t], dt, 3, (const char*)g.A + (size_t)tt * tap_k * 2, dims, strides, box)) return OH_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
        const uint64_t strides[1] = {(uint64_t)g.K * 2};
        const uint32_t box[2] = {BLOCK_K, BLOCK_N};
        if (!encode_map(&mw, dt, 2, g.W, dims, strides, box)) return OH_ERR_UNSUPPORTED;
    }
    TcParams p;
    p.batch = g.batch; p.rows = g.rows; p.N = g.N; p.K = g.K;
    p.tiles_per_batch = (g.rows + BLOCK_M - 1) / BLOCK_M;
    p.n_blocks = (g.N + BLOCK_N - 1) / BLOCK_N;
    p.num_tiles = p.tiles_per_batch * g.batch * p.n_blocks;
    p.n_taps = n_taps;
    p.tap_k = tap_k;
    p.kb_per_tap = (tap_k + BLOCK_K - 1) / BLOCK_K;
    p.num_kb = p.kb_per_tap * n_taps;
    const uint32_t fmt = g.a_type == DT_BF16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
    p.bias = g.bias; p.C = g.C; p.c_type = g.c_type; p.c_bstride = g.c_bstride; p.ldc = g.ldc; p.c_bidx = g.c_bidx;
    p.R = g.R; p.r_bstride = g.r_bstride; p.ldr = g.ldr; p.gelu = g.gelu; p.round_type = g.round_type; p.lut = g.lut;
    p.c_hsplit = g.c_hsplit; p.c_hstride = g.c_hstride;
    // compact mode (encoder passes that run next to decode steps): at most 3 stages / 96 KB of ring, so that a decode CTA
    // (cross-attention: ~58 KB with a 3-deep ring, split-K mat-mul: 96 KB) fits on the same SM
    p.stages = g_gemm_compact > 0 ? std::min(Cfg::STAGES, std::max(2, g_gemm_compact_bytes / Cfg::STAGE_BYTES)) : Cfg::STAGES;
    const int smem_bytes = p.stages * Cfg::STAGE_BYTES + 1024 + 256;
    const int sm_cap = g_sm_limit > 0 ? std::min(g_sm_limit, g_num_sms[PerDeviceInit::current_device()]) : g_num_sms[PerDeviceInit::current_device()];
    const int grid = std::min(p.num_tiles, sm_cap);
    OH_SKIP_IF(g.N > 40000 ? PC_LOGITS : (g.batch * g.rows <= 512 ? PC_GEMM_DEC : PC_GEMM_ENC));  // timing experiments only
    ProfScope ps(g.batch * g.rows <= 512 ? PC_GEMM_DEC : PC_GEMM_ENC, 2.0 * g.batch * (double)g.rows * g.N * g.K, st);
    OH_CUDA_CHECK(launch_k(gemm_tc_kernel<BLOCK_N>, grid, TC_THREADS, smem_bytes, st, ma[0], ma[1], ma[2], mw, p));
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace

bool gemm_tc_available() { return g_tc_init.ensure(tc_init_device); }

bool tc_encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
    return gemm_tc_available() && encode_map(map, dt, rank, base, dims, strides_bytes, box);
}

int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
    if (!gemm_tc_available()) return OH_ERR_UNSUPPORTED;
    if (g.a_type == DT_F32 || g.a_type != g.w_type) return OH_ERR_UNSUPPORTED;
    const int M = g.batch * g.rows;
    if (M <= 0 || g.N <= 0 || g.K <= 0) return OH_OK;
    // TMA: 16-byte aligned bases and strides; vectorised epilogue: 16-byte aligned output rows
    const int c_es = g.c_type == DT_F32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15) || (g.lda % 8) || (g.K % 8) || (g.a_bstride % 8) ||
        ((size_t)g.ldc * c_es % 16) || ((size_t)g.c_bstride * c_es % 16) || (reinterpret_cast<uintptr_t>(g.C) & 15) || (g.R && ((g.ldr % 4) || (g.r_bstride % 4))) ||
        (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15)))
        return OH_ERR_UNSUPPORTED;
    if (g.n_taps > 1) {
        if (g.n_taps != 3 || g.tap_k * 3 != g.K || g.tap_k % 8) return OH_ERR_UNSUPPORTED;
    } else if (g.lda < g.K) {
        return OH_ERR_UNSUPPORTED;
    }
    if (g.c_hsplit > 0 && (g.c_hsplit % 32 != 0 || g.N % g.c_hsplit != 0)) return OH_ERR_UNSUPPORTED;
    if (M <= BLOCK_M) {
        // decode-time mat-muls (M = batch rows): the work is streaming W once, so cut N finely enough that most SMs pull
        const int want = 120;
        if ((g.N + 127) / 128 >= want) return launch_t<128>(g, st);
        if ((g.N + 63) / 64 >= want) return launch_t<64>(g, st);
        if ((g.N + 31) / 32 >= want) return launch_t<32>(g, st);
        return launch_t<16>(g, st);
    }
    if (g.N < 128) return launch_t<64>(g, st);
    if (g_gemm_compact > 0) return launch_t<128>(g, st);  // 128-wide tiles: 32 KB stages, 256 TMEM columns
    if (g.N % 256 == 0 || g.N >= 2048) return launch_t<256>(g, st);
    return launch_t<128>(g, st);
}

// Runs four small problems (tails in M, N and K; all three tile widths; the 3-tap conv view; both output
// types; bias + residual) through the tcgen05 kernel and through the SIMT kernel and compares.
int gemm_tc_selftest(cudaStream_t st, float* max_rel_err) {
    if (!gemm_tc_available()) return OH_ERR_UNSUPPORTED;
    struct Case {
        int batch, rows, N, K, lda, n_taps, tap_k, c_type;
        long long a_bstride;
    };
    const Case cases[7] = {
        {1, 300, 200, 136, 136, 1, 0, DT_F32, 0},        // BLOCK_N 128, tails everywhere
        {1, 70, 328, 256, 256, 1, 0, DT_F16, 0},         // BLOCK_N 64 (small M)
        {1, 260, 512, 192, 192, 1, 0, DT_F16, 0},        // BLOCK_N 256
        {2, 150, 256, 240, 80, 3, 80, DT_F32, 154 * 80}, // conv view: rows overlap, 3 taps of 80
        {1, 64, 1280, 320, 320, 1, 0, DT_F16, 0},        // decode shape -> BLOCK_N 16
        {1, 40, 3840, 128, 128, 1, 0, DT_F32, 0},        // decode shape -> BLOCK_N 32
        {2, 100, 256, 128, 128, 1, 0, DT_F16, 100 * 128}, // batched rows + head-split store (hsplit 64)
    };
    float worst = 0.0f;
    for (const Case& c : cases) {
        const size_t a_elems = c.batch > 1 ? (size_t)c.batch * c.a_bstride : (size_t)(c.rows + 2) * c.lda;
        const size_t M = (size_t)c.batch * c.rows;
        std::vector<__half> hA(a_elems), hW((size_t)c.N * c.K);
        std::vector<float> hb(c.N), hR(M * c.N);
        uint32_t s = 12345u + (uint32_t)c.N;
        auto rnd = [&]() {
            s = s * 1664525u + 1013904223u;
            return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
        };
        for (auto& v : hA) v = __float2half_rn(rnd());
        for (auto& v : hW) v = __float2half_rn(rnd() * 0.25f);
        for (auto& v : hb) v = rnd();
        for (auto& v : hR) v = rnd();
        const size_t c_es = c.c_type == DT_F32 ? 4 : 2;
        void *dA = nullptr, *dW = nullptr, *db = nullptr, *dR = nullptr, *dC0 = nullptr, *dC1 = nullptr;
        OH_CUDA_CHECK(cudaMalloc(&dA, hA.size() * 2));
        OH_CUDA_CHECK(cudaMalloc(&dW, hW.size() * 2));
        OH_CUDA_CHECK(cudaMalloc(&db, hb.size() * 4));
        OH_CUDA_CHECK(cudaMalloc(&dR, hR.size() * 4));
        OH_CUDA_CHECK(cudaMalloc(&dC0, M * c.N * c_es));
        OH_CUDA_CHECK(cudaMalloc(&dC1, M * c.N * c_es));
        OH_CUDA_CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemcpy(dR, hR.data(), hR.size() * 4, cudaMemcpyHostToDevice));
        OH_CUDA_CHECK(cudaMemset(dC0, 0, M * c.N * c_es));
        OH_CUDA_CHECK(cudaMemset(dC1, 0, M * c.N * c_es));
        GemmArgs g;
        g.A = dA; g.a_type = DT_F16; g.a_bstride = c.a_bstride; g.lda = c.lda; g.batch = c.batch; g.rows = c.rows;
        g.W = dW; g.w_type = DT_F16; g.N = c.N; g.K = c.K; g.bias = (const float*)db;
        g.c_type = c.c_type; g.c_bstride = (long long)c.rows * c.N; g.ldc = c.N;
        g.R = (const float*)dR; g.r_bstride = (long long)c.rows * c.N; g.ldr = c.N;
        g.gelu = GELU_FAST; g.round_type = DT_F16; g.n_taps = c.n_taps; g.tap_k = c.tap_k;
        if (c.batch == 2 && c.n_taps == 1) {  // the head-split case: [batch][N/64][rows][64]
            g.c_hsplit = 64;
            g.c_hstride = (long long)c.rows * 64;
            g.ldc = 64;
        }
        g.C = dC0;
        int rc = launch_gemm_simt(g, st);
        g.C = dC1;
        if (rc == OH_OK) rc = launch_gemm_tc(g, st);
        cudaError_t ce = cudaStreamSynchronize(st);
        std::vector<float> r0(M * c.N), r1(M * c.N);
        if (rc == OH_OK && ce == cudaSuccess) {
            if (c.c_type == DT_F32) {
                cudaMemcpy(r0.data(), dC0, r0.size() * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(r1.data(), dC1, r1.size() * 4, cudaMemcpyDeviceToHost);
            } else {
                std::vector<__half> t0(M * c.N), t1(M * c.N);
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
            if (!(e < 2e-3f)) fprintf(stderr, "[ohwhisper] tcgen05 self-test case M=%dx%d N=%d K=%d taps=%d: rel err %g\n", c.batch, c.rows, c.N, c.K, c.n_taps, e);
            worst = std::max(worst, e);
        }
        cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dR); cudaFree(dC0); cudaFree(dC1);
        if (rc != OH_OK) return rc;
        if (ce != cudaSuccess) {
            set_error(std::string("gemm_tc_selftest: ") + cudaGetErrorString(ce));
            return OH_ERR_CUDA;
        }
    }
    if (max_rel_err) *max_rel_err = worst;
    return OH_OK;
}

}  // namespace oh
