// K4 (decode) — mat-muls of the autoregressive step: M = resident decoder rows (<= 256), so the work is streaming
// the f16/bf16 weight matrix exactly once at HBM speed (SURVEY.md §8d "Decoder step: HBM"), not tensor throughput.
//
// Shape of the kernel:
//   * a CTA owns 64 weight rows (8 warps x 8 rows) and one K-range; a thread-block CLUSTER of CS CTAs shares the 64
//     weight rows and splits K, so that d x d projections still put >= 148 CTAs on the machine;
//   * weights go global -> registers directly (16-byte loads, 8 of them in flight per lane, every byte used once);
//     activations X[M][K-range] are staged in shared memory once per CTA (cp.async) and feed the A side of
//     m16n8k16 tensor-core MMAs through 16-byte LDS; the k order inside a 32-wide block is permuted identically for
//     both operands so that a lane's 16 contiguous weight bytes are exactly its B fragments of two MMAs;
//   * the CS partial tiles are summed in a fixed order through distributed shared memory (each CTA reduces and
//     stores its share of the rows), so the result is deterministic and no partial sum ever touches HBM;
//   * epilogue as everywhere: + bias, GELU, + residual, type conversion, 64-column coalesced row stores.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/oh_whisper.h"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace oh {

namespace {

constexpr int SK_THREADS = 256, SK_ROWS = 64, SK_MAXKB = 8;  // 8 warps x 8 weight rows; <= 8 k-blocks of 32 per sub-chunk

struct SkParams {
    const void* A;
    int lda, M;
    const void* W;
    int N, K;
    int kc;      // K-range per CTA (K / cluster size)
    int ks;      // sub-chunk of the K-range staged at a time (<= 256, multiple of 32, divides kc)
    const float* bias;
    void* C;
    int c_type, ldc;
    const float* R;
    int ldr;
    int gelu, round_type;
    const uint16_t* lut;
};

__device__ __forceinline__ void sk_cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
template <typename T>
__device__ __forceinline__ void sk_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void sk_mma<__half>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void sk_mma<__nv_bfloat16>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float sk_gelu(float v, int gelu, int round_type, const uint16_t* lut) {
    if (gelu == GELU_LUT) return gelu_lut(v, lut);
    if (gelu == GELU_TANH) return gelu_tanh(v);
    if (gelu == GELU_FAST) {
        if (round_type == DT_F16) return __half2float(__float2half_rn(gelu_fast(__half2float(__float2half_rn(v)))));
        if (round_type == DT_BF16) return __bfloat162float(__float2bfloat16_rn(gelu_fast(v)));
        return gelu_fast(v);
    }
    return v;
}

// MT = number of 16-row tiles of X (M_pad = 16 * MT)
template <typename T, int MT>
__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(const SkParams p) {
    extern __shared__ __align__(128) uint8_t sk_smem[];
    pdl_wait();  // programmatic dependent launch (see common.cuh)
    pdl_launch_dependents();
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int n_slice = blockIdx.x / cs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int MP = 16 * MT;
    const int m0 = blockIdx.y * MP;  // this CTA's rows of X / C (the weights are re-read from L2 by the other row blocks)
    const int xs = (p.ks % 64 == 0) ? p.ks + 32 : p.ks;  // row stride of the staged X in halves: 2*xs = 64 (mod 128) bytes -> conflict-free 16-byte LDS
    T* Xs = reinterpret_cast<T*>(sk_smem);
    const T* __restrict__ A = reinterpret_cast<const T*>(p.A);
    const T* __restrict__ W = reinterpret_cast<const T*>(p.W);

    const int wrow = n_slice * SK_ROWS + warp * 8 + g;  // this lane's weight row
    const bool wok = wrow < p.N;
    const T* wptr = W + (size_t)(wok ? wrow : 0) * p.K + crank * p.kc + 8 * t;

    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][e] = 0.0f;

    const int n_sub = p.kc / p.ks;
    const int nkb = p.ks / 32;
    for (int sc = 0; sc < n_sub; ++sc) {
        const int k0 = crank * p.kc + sc * p.ks;
        // weights of this sub-chunk: all loads issued before anything waits
        uint4 wreg[SK_MAXKB];
#pragma unroll
        for (int kb = 0; kb < SK_MAXKB; ++kb) {
            if (kb < nkb && wok) wreg[kb] = __ldg(reinterpret_cast<const uint4*>(wptr + sc * p.ks + kb * 32));
            else wreg[kb] = make_uint4(0u, 0u, 0u, 0u);
        }
        // stage X[:, k0 .. k0+ks) (rows >= M are zero)
        const int chunks_per_row = p.ks / 8;
        for (int idx = tid; idx < MP * chunks_per_row; idx += SK_THREADS) {
            const int r = idx / chunks_per_row, ch = idx - r * chunks_per_row;
            const bool ok = m0 + r < p.M;
            sk_cp_async16((uint32_t)__cvta_generic_to_shared(Xs + (size_t)r * xs + ch * 8), A + (size_t)(ok ? m0 + r : 0) * p.lda + k0 + ch * 8, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int kb = 0; kb < SK_MAXKB; ++kb) {
            if (kb < nkb) {
                const uint4 w = wreg[kb];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint4 xlo = *reinterpret_cast<const uint4*>(Xs + (size_t)(mt * 16 + g) * xs + kb * 32 + 8 * t);
                    const uint4 xhi = *reinterpret_cast<const uint4*>(Xs + (size_t)(mt * 16 + g + 8) * xs + kb * 32 + 8 * t);
                    sk_mma<T>(acc[mt], xlo.x, xhi.x, xlo.y, xhi.y, w.x, w.y);
                    sk_mma<T>(acc[mt], xlo.z, xhi.z, xlo.w, xhi.w, w.z, w.w);
                }
            }
        }
        __syncthreads();  // X buffer is reused by the next sub-chunk (and by the partial tile below)
    }

    // ---- partial tile [MP][64] f32 into this CTA's shared memory, then a fixed-order reduction across the cluster
    float* Ps = reinterpret_cast<float*>(sk_smem);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int col = warp * 8 + 2 * t;
        *reinterpret_cast<float2*>(Ps + (size_t)(mt * 16 + g) * SK_ROWS + col) = make_float2(acc[mt][0], acc[mt][1]);
        *reinterpret_cast<float2*>(Ps + (size_t)(mt * 16 + g + 8) * SK_ROWS + col) = make_float2(acc[mt][2], acc[mt][3]);
    }
    cluster.sync();
    // CTA `crank` finishes rows r = crank, crank + cs, ...; thread -> (row offset tid / 64, column tid % 64)
    const int col = tid & 63;
    const int n = n_slice * SK_ROWS + col;
    const float bias = (p.bias && n < p.N) ? p.bias[n] : 0.0f;
    const float* remote[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) remote[c] = cluster.map_shared_rank(Ps, c < cs ? c : 0);
    const int m_here = min(MP, p.M - m0);
    for (int rl = crank + cs * (tid >> 6); rl < m_here; rl += cs * (SK_THREADS / 64)) {
        const int r = m0 + rl;
        // all peers' partials are requested before the first add (remote shared memory is ~200 cycles away);
        // the summation order is fixed (rank 0..cs-1), so the result does not depend on timing
        float part[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) part[c] = c < cs ? remote[c][(size_t)rl * SK_ROWS + col] : 0.0f;
        float v = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) v += part[c];
        if (n < p.N) {
            v += bias;
            if (p.gelu != GELU_NONE) v = sk_gelu(v, p.gelu, p.round_type, p.lut);
            if (p.R) v += p.R[(size_t)r * p.ldr + n];
            if (p.c_type == DT_F32) reinterpret_cast<float*>(p.C)[(size_t)r * p.ldc + n] = v;
            else if (p.c_type == DT_F16) reinterpret_cast<__half*>(p.C)[(size_t)r * p.ldc + n] = __float2half_rn(v);
            else reinterpret_cast<__nv_bfloat16*>(p.C)[(size_t)r * p.ldc + n] = __float2bfloat16_rn(v);
        }
    }
    cluster.sync();  // nobody leaves while a peer may still read its partial tile
}

template <typename T, int MT>
int launch_mt(const SkParams& p, int cs, cudaStream_t st) {
    const int MP = 16 * MT;
    const size_t xs = (p.ks % 64 == 0) ? p.ks + 32 : p.ks;
    const size_t smem = std::max((size_t)MP * xs * 2, (size_t)MP * SK_ROWS * 4);
    {   // shared-memory opt-in: per device and per instantiation, grown on demand
        static std::mutex mu;
        static size_t configured[PerDeviceInit::kMaxDevices] = {};
        std::lock_guard<std::mutex> lk(mu);
        size_t& cur = configured[PerDeviceInit::current_device()];
        if (smem > cur) {
            OH_CUDA_CHECK(cudaFuncSetAttribute(gemm_skinny_kernel<T, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
    }
    const int n_slices = (p.N + SK_ROWS - 1) / SK_ROWS;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_slices * cs, (p.M + MP - 1) / MP);
    cfg.blockDim = dim3(SK_THREADS);
    cfg.dynamicSmemBytes = smem;
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
    ProfScope ps(PC_GEMM_DEC, 2.0 * p.M * (double)p.N * p.K, st);
    OH_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<T, MT>, p));
    OH_COUNT_LAUNCH();
    return OH_OK;
}

template <typename T>
int launch_t(const SkParams& p, int cs, cudaStream_t st) {
    // measured on B200 (scripts/bench_decode_step.py): one CTA column over all rows beats splitting the rows over
    // grid.y — the kernel is bound by per-CTA latency and by the L2 re-reads of W, not by MMA issue
    if (p.M <= 16) return launch_mt<T, 1>(p, cs, st);
    if (p.M <= 32) return launch_mt<T, 2>(p, cs, st);
    if (p.M <= 64) return launch_mt<T, 4>(p, cs, st);
    if (p.M <= 128) return launch_mt<T, 8>(p, cs, st);
    return launch_mt<T, 16>(p, cs, st);
}

}  // namespace

int launch_gemm_skinny(const GemmArgs& g, cudaStream_t st) {
    const int M = g.batch * g.rows;
    if (g.a_type == DT_F32 || g.a_type != g.w_type || g.batch != 1 || M < 1 || M > 256 || g.n_taps != 1 || g.c_hsplit != 0) return OH_ERR_UNSUPPORTED;
    if (g.K % 32 || g.lda % 8 || (reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15)) return OH_ERR_UNSUPPORTED;
    // cluster size: enough CTAs to cover the machine, K-range a multiple of 32
    const int n_slices = (g.N + SK_ROWS - 1) / SK_ROWS;
    int cs = 1;
    static const int cs_cap = [] {
        const char* e = getenv("OH_SKINNY_CS");
        return e ? std::max(1, atoi(e)) : 8;
    }();
    while (cs < cs_cap && n_slices * cs < 148 && (g.K / (cs * 2)) % 32 == 0 && g.K / (cs * 2) >= 64) cs *= 2;
    const int kc = g.K / cs;
    int ks = 0;
    for (int cand = std::min(kc, 32 * SK_MAXKB); cand >= 32; cand -= 32)
        if (kc % cand == 0) {
            ks = cand;
            break;
        }
    if (ks == 0) return OH_ERR_UNSUPPORTED;
    SkParams p;
    p.A = g.A; p.lda = g.lda; p.M = M; p.W = g.W; p.N = g.N; p.K = g.K; p.kc = kc; p.ks = ks;
    p.bias = g.bias; p.C = g.C; p.c_type = g.c_type; p.ldc = g.ldc; p.R = g.R; p.ldr = g.ldr;
    p.gelu = g.gelu; p.round_type = g.round_type; p.lut = g.lut;
    return g.a_type == DT_BF16 ? launch_t<__nv_bfloat16>(p, cs, st) : launch_t<__half>(p, cs, st);
}

// start-up check against the SIMT kernel (decode shapes: every cluster size, N tail, GELU, residual, both output types)
int gemm_skinny_selftest(cudaStream_t st, float* max_rel_err) {
    struct Case {
        int M, N, K, c_type, gelu;
    };
    const Case cases[5] = {{64, 1280, 1280, DT_F32, 0}, {100, 3840, 1280, DT_F16, 0}, {37, 5120, 1280, DT_F16, 1}, {64, 1280, 5120, DT_F32, 0}, {200, 1300, 256, DT_F32, 0}};
    float worst = 0.0f;
    for (const Case& c : cases) {
        std::vector<__half> hA((size_t)c.M * c.K), hW((size_t)c.N * c.K);
        std::vector<float> hb(c.N), hR((size_t)c.M * c.N);
        uint32_t s = 777u + (uint32_t)c.N + (uint32_t)c.K;
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
        if (rc == OH_OK) rc = launch_gemm_skinny(g, st);
        cudaError_t ce = cudaStreamSynchronize(st);
        if (rc == OH_OK && ce == cudaSuccess) {
            std::vector<float> r0((size_t)c.M * c.N), r1((size_t)c.M * c.N);
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
            if (!(e < 2e-3f)) fprintf(stderr, "[ohwhisper] skinny GEMM self-test case M=%d N=%d K=%d: rel err %g\n", c.M, c.N, c.K, e);
            worst = std::max(worst, e);
        }
        cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dR); cudaFree(dC0); cudaFree(dC1);
        if (rc != OH_OK) return rc;
        if (ce != cudaSuccess) {
            set_error(std::string("gemm_skinny_selftest: ") + cudaGetErrorString(ce));
            return OH_ERR_CUDA;
        }
    }
    if (max_rel_err) *max_rel_err = worst;
    return OH_OK;
}

}  // namespace oh
