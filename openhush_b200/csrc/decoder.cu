// K7 / K8 / K9 — the bandwidth-bound kernels of one autoregressive decoder step (SURVEY.md A.4 decoder,
// A.6 logit rules, A.7 greedy pick, A.5 per-token bookkeeping).  They replace whisper.cpp's
// `whisper_decode_internal` + `whisper_process_logits` + `whisper_sample_token`, reached from
// /root/reference/src/engine/whisper.rs:266.
//
// All per-sequence control state (position, current token, has_ts, seek_delta, result_len, failed/completed)
// lives in device memory (DecodeSeqState) and is advanced by the K9 kernel itself, so a greedy decode is a
// fixed sequence of launches per token with no host round trip — which is what lets the step be replayed as
// one CUDA graph.  Rows whose `active` flag is 0 exit immediately (no KV traffic).
#include <algorithm>

#include "../../include/oh_whisper.h"
#include "decode_rules.h"
#include "kernels.h"

namespace oh {

namespace {

// ------------------------------------------------------------------------------------------------- embed
template <typename TW>
__global__ void embed_kernel(const TW* __restrict__ emb, const float* __restrict__ pos_emb, DecodeSeqState s, int d, float* __restrict__ x) {
    pdl_wait();  // programmatic dependent launch: nothing above depends on the preceding kernel (see common.cuh)
    pdl_launch_dependents();
    const int b = blockIdx.x;
    if (!s.active[b]) return;
    const int tok = s.cur_token[b], pos = s.pos[b];
    const TW* e = emb + (size_t)tok * d;
    const float* p = pos_emb + (size_t)pos * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) x[(size_t)b * d + c] = Cvt<TW>::to_f32(e[c]) + p[c];
}

// ------------------------------------------------------------------------------------ K7+K8a: self attention
// grid (heads, rows), 128 threads.  Appends k,v of the current token at cache[slot][pos], then attends over [0, pos].
template <typename T>
__global__ void __launch_bounds__(128) self_attention_decode_kernel(const T* __restrict__ qkv, T* __restrict__ kcache, T* __restrict__ vcache,
                                                                    DecodeSeqState s, int d, int n_ctx, T* __restrict__ out) {
    pdl_wait();  // programmatic dependent launch: nothing above depends on the preceding kernel (see common.cuh)
    pdl_launch_dependents();
    const int h = blockIdx.x, b = blockIdx.y;
    if (!s.active[b]) return;
    __shared__ float q[kHeadDim];
    __shared__ float sc[448];
    __shared__ float red[4];
    __shared__ float part[2][kHeadDim];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pos = s.pos[b], slot = s.sslot[b];
    const int n_kv = pos + 1;
    const T* row = qkv + (size_t)b * 3 * d;
    // self cache of a cache line: [heads][n_ctx][64] — contiguous per (slot, head), like the cross cache
    T* K = kcache + ((size_t)slot * (d / kHeadDim) + h) * n_ctx * kHeadDim;
    T* V = vcache + ((size_t)slot * (d / kHeadDim) + h) * n_ctx * kHeadDim;
    if (tid < kHeadDim) {
        q[tid] = Cvt<T>::to_f32(row[h * kHeadDim + tid]);
        K[(size_t)pos * kHeadDim + tid] = row[d + h * kHeadDim + tid];
        V[(size_t)pos * kHeadDim + tid] = row[2 * d + h * kHeadDim + tid];
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = tid; j < n_kv; j += 128) {
        const T* kr = K + (size_t)j * kHeadDim;
        float acc = 0.0f;
#pragma unroll 16
        for (int c = 0; c < kHeadDim; ++c) acc = fmaf(q[c], Cvt<T>::to_f32(kr[c]), acc);
        acc *= 0.125f;
        sc[j] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.0f;
    for (int j = tid; j < n_kv; j += 128) {
        const float e = expf(sc[j] - mx);
        sc[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float inv = 1.0f / ((red[0] + red[1]) + (red[2] + red[3]));
    const int c = tid & 63, half_id = tid >> 6;
    float acc = 0.0f;
    for (int j = half_id; j < n_kv; j += 2) {
        const float p = Cvt<T>::to_f32(Cvt<T>::from_f32(sc[j] * inv));
        acc = fmaf(p, Cvt<T>::to_f32(V[(size_t)j * kHeadDim + c]), acc);
    }
    part[half_id][c] = acc;
    __syncthreads();
    if (tid < kHeadDim) out[(size_t)b * d + h * kHeadDim + tid] = Cvt<T>::from_f32(part[0][tid] + part[1][tid]);
}

// ----------------------------------------------------------------------------------- K8b: cross attention
// grid (heads, rows), 128 threads.  The cross cache of a slot is [T][2d] (k | v per encoder position).  Eight
// lanes share one key: each reads 8 consecutive elements (16 B in the 16-bit modes), so a warp instruction
// covers 4 keys x 128 B, and 4 independent keys are in flight per thread.
template <typename T>
struct Vec8 {
    float v[8];
    __device__ __forceinline__ void load(const T* p);     // read-only path (__ldg)
    __device__ __forceinline__ void load_cg(const T* p);  // coherent at L2 (data written earlier in the same kernel)
};
template <>
__device__ __forceinline__ void Vec8<float>::load_cg(const float* p) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void Vec8<__half>::load_cg(const __half* p) {
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void Vec8<__nv_bfloat16>::load_cg(const __nv_bfloat16* p) {
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void Vec8<float>::load(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void Vec8<__half>::load(const __half* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void Vec8<__nv_bfloat16>::load(const __nv_bfloat16* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

template <typename T>
__global__ void __launch_bounds__(128) cross_attention_decode_kernel(const T* __restrict__ q, const T* __restrict__ xkv, DecodeSeqState s, int d,
                                                                     int Tn, T* __restrict__ out) {
    pdl_wait();  // programmatic dependent launch: nothing above depends on the preceding kernel (see common.cuh)
    pdl_launch_dependents();
    const int h = blockIdx.x, b = blockIdx.y;
    if (!s.active[b]) return;
    __shared__ float sc[1504];
    __shared__ float red[4];
    __shared__ float part[4][kHeadDim];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane & 7, grp = lane >> 3;
    const int kslot = warp * 4 + grp;  // 0..15: this thread's key phase
    const int slot = s.slot[b];
    // cross cache of a slot: [2 (k,v)][heads][T][64] — every (slot, head) stream is contiguous
    const size_t ld = kHeadDim;
    const int H = d / kHeadDim;
    const T* Kb = xkv + (((size_t)slot * 2 + 0) * H + h) * Tn * kHeadDim + sub * 8;
    const T* Vb = xkv + (((size_t)slot * 2 + 1) * H + h) * Tn * kHeadDim + sub * 8;
    float qv[8];
    {
        const T* qr = q + (size_t)b * d + h * kHeadDim + sub * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) qv[i] = Cvt<T>::to_f32(qr[i]) ;
    }
    // ---- scores
    float mx = -INFINITY;
    for (int base = 0; base < Tn; base += 64) {  // warp-uniform trip count (full-mask shuffles inside)
        const int j0 = base + kslot;
        Vec8<T> kv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 16 * u;
            if (j < Tn) kv[u].load(Kb + (size_t)j * ld);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 16 * u;
            float acc = 0.0f;
            if (j < Tn) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(qv[i], kv[u].v[i], acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc *= 0.125f;
            if (j < Tn) {
                if (sub == 0) sc[j] = acc;
                mx = fmaxf(mx, acc);
            }
        }
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.0f;
    for (int j = tid; j < Tn; j += 128) {
        const float e = expf(sc[j] - mx);
        sc[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float inv = 1.0f / ((red[0] + red[1]) + (red[2] + red[3]));
    // ---- P.V
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.0f;
    for (int j0 = kslot; j0 < Tn; j0 += 64) {
        Vec8<T> vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 16 * u;
            if (j < Tn) vv[u].load(Vb + (size_t)j * ld);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 16 * u;
            if (j < Tn) {
                const float p = Cvt<T>::to_f32(Cvt<T>::from_f32(sc[j] * inv));
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(p, vv[u].v[i], o[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
        o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
    }
    if (grp == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) part[warp][sub * 8 + i] = o[i];
    }
    __syncthreads();
    if (tid < kHeadDim) out[(size_t)b * d + h * kHeadDim + tid] = Cvt<T>::from_f32((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]));
}

// eight 16-bit elements (one 16-byte load) -> f32
template <typename T>
__device__ __forceinline__ void raw8_to_f32(const uint4& u, float (&v)[8]);
template <>
__device__ __forceinline__ void raw8_to_f32<__half>(const uint4& u, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void raw8_to_f32<__nv_bfloat16>(const uint4& u, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

// ------------------------------------------------- K7+K8a (16-bit modes): self attention, one warp per (row, head)
// The self cache is short per (row, head) (<= 448 keys, typically ~100: 12.8 KB of K and of V, contiguous) but not
// L2-resident at 256 rows (131 MB per layer at position 100, re-read one step = 4 GB of other traffic later), so the
// kernel is a short HBM stream bound by latency chains: one warp per (row, head) with no block-level barrier, 4 groups
// of 8 lanes x 4 keys per sweep (8 independent 16-byte loads per lane, the next sweep requested before this one is
// consumed), per-group online softmax merged with shuffles at the end.  The step's k,v are appended first.
// Measured at 256 rows, position 100: 34 us per launch for 136 MB = 4.0 TB/s (0.6 of the measured HBM peak); what is
// left is the prologue's dependent loads (row state -> addresses -> first keys) and 2.2 waves of 16 warps per SM.
// Holding the register allocation to 5 or 6 CTAs per SM (96 / 80 registers instead of 118) was measured and does not pay:
// 1.02 ms per 256-row step at 4 CTAs, 1.11 at 5, 1.59 at 6 (spills).
template <typename T>
__global__ void __launch_bounds__(128) self_attention_decode_warp_kernel(const T* __restrict__ qkv, T* kcache, T* vcache, DecodeSeqState s, int d,
                                                                         int n_ctx, int rows, T* __restrict__ out) {
    pdl_wait();  // programmatic dependent launch: nothing above depends on the preceding kernel (see common.cuh)
    pdl_launch_dependents();
    const int H = d / kHeadDim;
    const int wid = blockIdx.x * 4 + (threadIdx.x >> 5);  // (row, head) index
    if (wid >= rows * H) return;
    const int b = wid / H, h = wid - b * H;
    if (!s.active[b]) return;
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const int pos = s.pos[b], slot = s.sslot[b];
    const int n_kv = pos + 1;
    const T* row = qkv + (size_t)b * 3 * d + h * kHeadDim;
    T* K = kcache + ((size_t)slot * H + h) * n_ctx * kHeadDim;
    T* V = vcache + ((size_t)slot * H + h) * n_ctx * kHeadDim;
    if (lane < 8) *reinterpret_cast<uint4*>(K + (size_t)pos * kHeadDim + lane * 8) = *reinterpret_cast<const uint4*>(row + d + lane * 8);
    else if (lane < 16) *reinterpret_cast<uint4*>(V + (size_t)pos * kHeadDim + (lane - 8) * 8) = *reinterpret_cast<const uint4*>(row + 2 * d + (lane - 8) * 8);
    __syncwarp();
    float qv[8];
    {
        const float sc = 0.125f * 1.44269504088896340736f;
        Vec8<T> qq;
        qq.load(row + sub * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) qv[i] = qq.v[i] * sc;
    }
    float m_run = -INFINITY, l_run = 0.0f, o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.0f;
    const T* Kb = K + sub * 8;
    const T* Vb = V + sub * 8;
    // 16 keys per sweep (4 per group); the next sweep's K/V rows are requested before the current one is consumed, so a
    // lane keeps up to 16 16-byte loads in flight (the kernel is bound by memory latency, not by arithmetic)
    uint4 kc[4], vc[4];
    auto load16 = [&](int base, uint4 (&kr)[4], uint4 (&vr)[4]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = base + grp + 4 * u;
            if (j < n_kv) {  // the cache is written by this kernel: coherent (L2) loads, not the read-only path
                kr[u] = __ldcg(reinterpret_cast<const uint4*>(Kb + (size_t)j * kHeadDim));
                vr[u] = __ldcg(reinterpret_cast<const uint4*>(Vb + (size_t)j * kHeadDim));
            }
        }
    };
    load16(0, kc, vc);
    for (int base = 0; base < n_kv; base += 16) {  // warp-uniform trip count
        uint4 kn[4], vn[4];
        if (base + 16 < n_kv) load16(base + 16, kn, vn);
        float sc4[4], mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = base + grp + 4 * u;
            float acc = 0.0f;
            if (j < n_kv) {
                float kf[8];
                raw8_to_f32<T>(kc[u], kf);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(qv[i], kf[i], acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            sc4[u] = j < n_kv ? acc : -INFINITY;
            mx = fmaxf(mx, sc4[u]);
        }
        if (mx > -INFINITY) {
            const float m_new = fmaxf(m_run, mx);
            const float corr = exp2f(m_run - m_new);
            m_run = m_new;
            l_run *= corr;
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] *= corr;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (base + grp + 4 * u < n_kv) {
                    const float p = exp2f(sc4[u] - m_new);
                    l_run += p;
                    const float pr = Cvt<T>::to_f32(Cvt<T>::from_f32(p));
                    float vf[8];
                    raw8_to_f32<T>(vc[u], vf);
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = fmaf(pr, vf[i], o[i]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            kc[u] = kn[u];
            vc[u] = vn[u];
        }
    }
    // merge the 4 groups (lanes sub, sub+8, sub+16, sub+24 hold the same channels)
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
        const float m_o = __shfl_xor_sync(0xffffffffu, m_run, off);
        const float l_o = __shfl_xor_sync(0xffffffffu, l_run, off);
        const float m_new = fmaxf(m_run, m_o);
        const float wa = m_run == -INFINITY ? 0.0f : exp2f(m_run - m_new);
        const float wb = m_o == -INFINITY ? 0.0f : exp2f(m_o - m_new);
        l_run = l_run * wa + l_o * wb;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float o_o = __shfl_xor_sync(0xffffffffu, o[i], off);
            o[i] = o[i] * wa + o_o * wb;
        }
        m_run = m_new;
    }
    if (grp == 0) {
        const float inv = 1.0f / l_run;
        T res[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) res[i] = Cvt<T>::from_f32(o[i] * inv);
        *reinterpret_cast<uint4*>(out + (size_t)b * d + h * kHeadDim + sub * 8) = *reinterpret_cast<const uint4*>(res);
    }
}

// ------------------------------------------------- K8b (16-bit modes): cross attention fed by bulk async copies
// The dominant HBM stream of a decode step: every row reads its slot's K and V (2 x 1500 x 64 per head, contiguous
// per (slot, head)).  One CTA per (head, row): a producer warp keeps XA_STAGES bulk copies (cp.async.bulk, i.e. the
// TMA engine in 1-D mode: global -> shared, completion on an mbarrier) in flight — 64 keys of K and of V per stage —
// so the memory pipeline depth is set by shared memory, not by registers; 8 consumer warps (32 groups of 8 lanes)
// read the staged rows with conflict-free 16-byte LDS, keep per-group online-softmax states and merge them at the end.
constexpr int XA_STAGES = 6, XA_KEYS = 64, XA_CONSUMERS = 256, XA_THREADS = XA_CONSUMERS + 32;

__device__ __forceinline__ uint32_t xa_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xa_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void xa_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xa_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void xa_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (clock64() - t0 > 8000000000LL) __trap();  // protocol bug: fail instead of hanging the GPU
    }
}
__device__ __forceinline__ void xa_bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
template <typename T>
__device__ __forceinline__ void xa_unpack8(const uint4& u, float (&v)[8]);
template <>
__device__ __forceinline__ void xa_unpack8<__half>(const uint4& u, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void xa_unpack8<__nv_bfloat16>(const uint4& u, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}

template <typename T>
__global__ void __launch_bounds__(XA_THREADS) cross_attention_decode_bulk_kernel(const T* __restrict__ q, const T* __restrict__ xkv, DecodeSeqState s,
                                                                                int d, int Tn, T* __restrict__ out, int n_stages) {
    // Programmatic dependent launch: the only input written by the preceding kernel (the q projection) is q.  The
    // row state was last written by the previous step's K9 and the cross K/V by the encoder, both complete before the
    // predecessor passed its own pdl_wait, so the producer warp starts streaming K/V at once and only the consumers
    // wait, right before they read q.
    const int h = blockIdx.x, b = blockIdx.y;
    if (!s.active[b]) {
        pdl_wait();
        return;
    }
    extern __shared__ __align__(128) uint8_t xa_smem[];
    // layout: [stage][K 64x64 | V 64x64] (16 KB per stage), then merge scratch, then the barriers
    constexpr int STAGE_BYTES = 2 * XA_KEYS * kHeadDim * 2;
    T* stage0 = reinterpret_cast<T*>(xa_smem);
    float* sm_o = reinterpret_cast<float*>(xa_smem + n_stages * STAGE_BYTES);  // [32][64]
    float* sm_m = sm_o + 32 * kHeadDim;                                           // [32]
    float* sm_l = sm_m + 32;                                                      // [32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_l + 32);                      // full[STAGES], empty[STAGES]
    const uint32_t bar0 = xa_smem_u32(bars);
    auto full_bar = [&](int st) { return bar0 + 8u * st; };
    auto empty_bar = [&](int st) { return bar0 + 8u * (XA_STAGES + st); };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = d / kHeadDim;
    const int slot = s.slot[b];
    const T* Kg = xkv + (((size_t)slot * 2 + 0) * H + h) * Tn * kHeadDim;
    const T* Vg = xkv + (((size_t)slot * 2 + 1) * H + h) * Tn * kHeadDim;
    const int n_it = (Tn + XA_KEYS - 1) / XA_KEYS;

    if (tid == 0) {
        for (int st = 0; st < n_stages; ++st) {
            xa_mbar_init(full_bar(st), 1);
            xa_mbar_init(empty_bar(st), XA_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == XA_CONSUMERS / 32) {
        // ===================== producer =====================
        if (lane == 0) {
            int st = 0;
            uint32_t phase = 0;
            for (int it = 0; it < n_it; ++it) {
                xa_mbar_wait(empty_bar(st), phase ^ 1);
                const int key0 = it * XA_KEYS;
                const uint32_t bytes = (uint32_t)min(XA_KEYS, Tn - key0) * kHeadDim * 2;
                const uint32_t dstK = xa_smem_u32(stage0) + st * STAGE_BYTES, dstV = dstK + STAGE_BYTES / 2;
                xa_mbar_expect_tx(full_bar(st), 2 * bytes);
                xa_bulk_copy(dstK, Kg + (size_t)key0 * kHeadDim, bytes, full_bar(st));
                xa_bulk_copy(dstV, Vg + (size_t)key0 * kHeadDim, bytes, full_bar(st));
                if (++st == n_stages) {
                    st = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ===================== consumers =====================
        const int sub = lane & 7, grp = lane >> 3;
        const int gidx = warp * 4 + grp;  // 0..31
        float qv[8];
        pdl_wait();
        pdl_launch_dependents();
        {
            const T* qr = q + (size_t)b * d + h * kHeadDim + sub * 8;
            const float sc = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[i] = Cvt<T>::to_f32(qr[i]) * sc;
        }
        float m_run = -INFINITY, l_run = 0.0f, o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.0f;
        int st = 0;
        uint32_t phase = 0;
        for (int it = 0; it < n_it; ++it) {
            xa_mbar_wait(full_bar(st), phase);
            const uint8_t* sK = xa_smem + st * STAGE_BYTES;
            const uint8_t* sV = sK + STAGE_BYTES / 2;
            const int valid = min(XA_KEYS, Tn - it * XA_KEYS);
            float sc2[2], kf[8];
            uint4 vraw[2];
            float mx = -INFINITY;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int key = gidx + 32 * u;
                const uint4 kraw = *reinterpret_cast<const uint4*>(sK + key * 128 + sub * 16);
                vraw[u] = *reinterpret_cast<const uint4*>(sV + key * 128 + sub * 16);
                xa_unpack8<T>(kraw, kf);
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc = fmaf(qv[i], kf[i], acc);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                sc2[u] = key < valid ? acc : -INFINITY;
                mx = fmaxf(mx, sc2[u]);
            }
            if (mx > -INFINITY) {
                const float m_new = fmaxf(m_run, mx);
                const float corr = exp2f(m_run - m_new);
                m_run = m_new;
                l_run *= corr;
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] *= corr;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (gidx + 32 * u < valid) {
                        const float p = exp2f(sc2[u] - m_new);
                        l_run += p;
                        const float pr = Cvt<T>::to_f32(Cvt<T>::from_f32(p));
                        float vf[8];
                        xa_unpack8<T>(vraw[u], vf);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = fmaf(pr, vf[i], o[i]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) xa_mbar_arrive(empty_bar(st));  // this warp is done with the stage
            if (++st == n_stages) {
                st = 0;
                phase ^= 1;
            }
        }
        if (sub == 0) {
            sm_m[gidx] = m_run;
            sm_l[gidx] = l_run;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) sm_o[gidx * kHeadDim + sub * 8 + i] = o[i];
    }
    __syncthreads();
    if (tid < kHeadDim) {
        float M = -INFINITY;
        for (int g2 = 0; g2 < 32; ++g2) M = fmaxf(M, sm_m[g2]);
        float L = 0.0f, O = 0.0f;
        for (int g2 = 0; g2 < 32; ++g2) {
            const float w = sm_m[g2] == -INFINITY ? 0.0f : exp2f(sm_m[g2] - M);
            L = fmaf(sm_l[g2], w, L);
            O = fmaf(sm_o[g2 * kHeadDim + tid], w, O);
        }
        out[(size_t)b * d + h * kHeadDim + tid] = Cvt<T>::from_f32(O / L);
    }
}
constexpr int xa_smem_bytes(int stages) { return stages * 2 * XA_KEYS * kHeadDim * 2 + 32 * kHeadDim * 4 + 64 * 4 + 2 * XA_STAGES * 8 + 64; }
constexpr int XA_SMEM_BYTES = xa_smem_bytes(XA_STAGES);

// ------------------------------------------------------------------------------------------------ K9
struct MaxIdx {
    float v;
    int i;
};
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) {  // larger value; on ties the lower index (A.7 strict '<' update)
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ MaxIdx warp_best(MaxIdx m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxIdx other;
        other.v = __shfl_xor_sync(0xffffffffu, m.v, o);
        other.i = __shfl_xor_sync(0xffffffffu, m.i, o);
        m = better(m, other);
    }
    return m;
}

struct LogitCtx {
    int is_initial, last_ts, penult_ts, has_ts, seek_delta;
};
__device__ __forceinline__ bool logit_masked(int i, const DecodeRules& r, const LogitCtx& c) {
    if (i == r.tok_not) return true;
    if (r.no_timestamps && i >= r.beg) return true;
    if (r.suppress_blank && c.is_initial && (i == r.eot || i == r.blank)) return true;
    if (i == r.sot || i == r.nosp || i == r.solm || i == r.translate || i == r.transcribe || i == r.prev) return true;
    if (i > r.sot && i <= r.sot + r.n_lang) return true;
    if (c.last_ts && (c.penult_ts ? i >= r.beg : i < r.eot)) return true;
    if (c.is_initial && r.tid0_initial >= 0 && i > r.beg + r.tid0_initial) return true;
    if (c.has_ts && i >= r.beg && i < r.beg + c.seek_delta / 2) return true;
    return false;
}

constexpr int K9_THREADS = 1024;

// sorted (best first) candidate list of the beam-mode top-k: larger logprob first, on ties the lower id
struct TopList {
    float v[kTopK];
    int i[kTopK];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < kTopK; ++j) {
            v[j] = -INFINITY;
            i[j] = 0x7fffffff;
        }
    }
    __device__ __forceinline__ void insert(float x, int id) {
        if (!(x > v[kTopK - 1] || (x == v[kTopK - 1] && id < i[kTopK - 1]))) return;
        v[kTopK - 1] = x;
        i[kTopK - 1] = id;
#pragma unroll
        for (int j = kTopK - 1; j > 0; --j) {
            const bool sw = v[j] > v[j - 1] || (v[j] == v[j - 1] && i[j] < i[j - 1]);
            if (sw) {
                const float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv;
                const int ti = i[j]; i[j] = i[j - 1]; i[j - 1] = ti;
            }
        }
    }
    __device__ __forceinline__ void merge_from_lane_xor(int off) {
        float ov[kTopK];
        int oi[kTopK];
#pragma unroll
        for (int j = 0; j < kTopK; ++j) {
            ov[j] = __shfl_xor_sync(0xffffffffu, v[j], off);
            oi[j] = __shfl_xor_sync(0xffffffffu, i[j], off);
        }
#pragma unroll
        for (int j = 0; j < kTopK; ++j) insert(ov[j], oi[j]);
    }
};

__global__ void __launch_bounds__(K9_THREADS) logits_step_kernel(const float* __restrict__ logits, int ld, DecodeSeqState s,
                                                                const DecodeRules* __restrict__ rp, float* __restrict__ dump_logprobs,
                                                                float* __restrict__ dump_probs, TopKOut topk) {
    pdl_wait();  // programmatic dependent launch: nothing above depends on the preceding kernel (see common.cuh)
    pdl_launch_dependents();
    const int b = blockIdx.x;
    if (!s.active[b]) return;
    const DecodeRules r = *rp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pos = s.pos[b];
    if (pos + 1 < s.prompt_len[b]) {  // still feeding the prompt: nothing to sample
        if (tid == 0) {
            s.pos[b] = pos + 1;
            s.cur_token[b] = s.prompt[(size_t)b * kMaxPrompt + pos + 1];
        }
        return;
    }
    __shared__ MaxIdx sm_txt[32], sm_ts[32];
    __shared__ float sm_a[32], sm_b[32];
    const int n_gen = s.n_gen[b];
    LogitCtx c;
    c.is_initial = n_gen == 0;
    const int* toks = s.tok_id + (size_t)b * kMaxGen;
    c.last_ts = n_gen > 0 && toks[n_gen - 1] >= r.beg;
    c.penult_ts = n_gen < 2 || toks[n_gen - 2] >= r.beg;
    c.has_ts = s.has_ts[b];
    c.seek_delta = s.seek_delta[b];
    const float* row = logits + (size_t)b * ld;
    const float temp = r.temperature;

    // pass 1: masked maxima (text ids < beg, timestamp ids >= beg), first index on ties
    MaxIdx mtxt{-INFINITY, 0x7fffffff}, mts{-INFINITY, 0x7fffffff};
    for (int i = tid; i < r.n_vocab; i += K9_THREADS) {
        if (logit_masked(i, r, c)) continue;
        float v = row[i];
        if (temp > 0.0f) v = v / temp;
        MaxIdx m{v, i};
        if (i < r.beg) mtxt = better(mtxt, m);
        else mts = better(mts, m);
    }
    mtxt = warp_best(mtxt);
    mts = warp_best(mts);
    if (lane == 0) {
        sm_txt[warp] = mtxt;
        sm_ts[warp] = mts;
    }
    __syncthreads();
    if (warp == 0) {
        mtxt = warp_best(sm_txt[lane]);
        mts = warp_best(sm_ts[lane]);
        if (lane == 0) {
            sm_txt[0] = mtxt;
            sm_ts[0] = mts;
        }
    }
    __syncthreads();
    mtxt = sm_txt[0];
    mts = sm_ts[0];
    const float mx = fmaxf(mtxt.v, mts.v);
    __syncthreads();

    // pass 2: sum exp(x - max) over text and over timestamps
    float st = 0.0f, sts = 0.0f;
    for (int i = tid; i < r.n_vocab; i += K9_THREADS) {
        if (logit_masked(i, r, c)) continue;
        float v = row[i];
        if (temp > 0.0f) v = v / temp;
        const float e = expf(v - mx);
        if (i < r.beg) st += e;
        else sts += e;
    }
    st = warp_sum(st);
    sts = warp_sum(sts);
    if (lane == 0) {
        sm_a[warp] = st;
        sm_b[warp] = sts;
    }
    __syncthreads();
    if (warp == 0) {
        st = warp_sum(sm_a[lane]);
        sts = warp_sum(sm_b[lane]);
        if (lane == 0) {
            sm_a[0] = st;
            sm_b[0] = sts;
        }
    }
    __syncthreads();
    st = sm_a[0];
    sts = sm_b[0];
    const float lse = logf(st + sts) + mx;
    const float ts_logprob = sts > 0.0f ? logf(sts) + mx - lse : -INFINITY;
    const float max_text_logprob = mtxt.v > -INFINITY ? mtxt.v - lse : -INFINITY;
    const bool force_ts = ts_logprob > max_text_logprob;

    if (topk.ids) {  // beam search: the row's best candidates of the filtered distribution; the host ranks across beams
        __shared__ float sm_tv[32][kTopK];
        __shared__ int sm_ti[32][kTopK];
        TopList tl;
        tl.init();
        for (int i = tid; i < r.n_vocab; i += K9_THREADS) {
            if (logit_masked(i, r, c) || (force_ts && i < r.beg)) continue;
            float v = row[i];
            if (temp > 0.0f) v = v / temp;
            tl.insert(v - lse, i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tl.merge_from_lane_xor(o);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < kTopK; ++j) {
                sm_tv[warp][j] = tl.v[j];
                sm_ti[warp][j] = tl.i[j];
            }
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < kTopK; ++j) {
                tl.v[j] = sm_tv[lane][j];
                tl.i[j] = sm_ti[lane][j];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tl.merge_from_lane_xor(o);
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < kTopK; ++j) {
                    topk.ids[(size_t)b * kTopK + j] = tl.v[j] > -INFINITY ? tl.i[j] : -1;
                    topk.logprobs[(size_t)b * kTopK + j] = tl.v[j];
                    topk.probs[(size_t)b * kTopK + j] = tl.v[j] > -INFINITY ? expf(tl.v[j]) : 0.0f;
                }
                int tid_tok = 0;
                float pt = 0.0f, ptsum = 0.0f;
                if (mts.v > -INFINITY) {
                    const float pmax = expf(mts.v - lse);
                    ptsum = expf(ts_logprob);
                    if (pmax > 0.0f) tid_tok = mts.i;
                    pt = pmax / (ptsum + 1e-10f);
                }
                topk.tid[b] = tid_tok;
                topk.pt[b] = pt;
                topk.ptsum[b] = ptsum;
            }
        }
        return;
    }

    if (dump_probs) {  // t > 0: the host samples (std::mt19937 + std::discrete_distribution, A.7)
        float* lp = dump_logprobs + (size_t)b * ld;
        float* pr = dump_probs + (size_t)b * ld;
        for (int i = tid; i < r.n_vocab; i += K9_THREADS) {
            float l = -INFINITY;
            if (!logit_masked(i, r, c) && !(force_ts && i < r.beg)) {
                float v = row[i];
                if (temp > 0.0f) v = v / temp;
                l = v - lse;
            }
            lp[i] = l;
            pr[i] = l == -INFINITY ? 0.0f : expf(l);
        }
        return;
    }

    if (tid == 0) {
        // greedy pick (A.7): first index with maximal probability
        int id = 0;
        float p = 0.0f, plog = 0.0f;
        MaxIdx pick{-INFINITY, 0x7fffffff};
        if (force_ts) pick = mts;
        else pick = better(mtxt, mts);
        if (pick.v > -INFINITY) {
            const float lpv = pick.v - lse;
            const float pv = expf(lpv);
            if (pv > 0.0f) {
                id = pick.i;
                p = pv;
                plog = lpv;
            }
        }
        int tid_tok = 0;
        float pt = 0.0f, ptsum = 0.0f;
        if (mts.v > -INFINITY) {
            const float pmax = expf(mts.v - lse);
            ptsum = expf(ts_logprob);
            if (pmax > 0.0f) tid_tok = mts.i;
            pt = pmax / (ptsum + 1e-10f);
        }
        if (id >= r.beg) {
            tid_tok = id;
            pt = p;
        }
        const int i = n_gen;
        const size_t o = (size_t)b * kMaxGen + i;
        s.tok_id[o] = id;
        s.tok_tid[o] = tid_tok;
        s.tok_p[o] = p;
        s.tok_plog[o] = plog;
        s.tok_pt[o] = pt;
        s.tok_ptsum[o] = ptsum;
        s.n_gen[b] = i + 1;
        SeqFlags f{c.has_ts, c.seek_delta, s.result_len[b], 0, 0};
        const bool live = apply_token_rules(f, id, i, s.seek[b], s.seek_end[b], r.beg, r.eot, r.max_tokens, r.no_timestamps, r.single_segment, r.n_max);
        s.has_ts[b] = f.has_ts;
        s.seek_delta[b] = f.seek_delta;
        s.result_len[b] = f.result_len;
        s.failed[b] = f.failed;
        s.completed[b] = f.completed;
        if (live) {
            s.cur_token[b] = id;
            s.pos[b] = pos + 1;
        } else {
            s.active[b] = 0;
            atomicSub(s.n_active_total, 1);
        }
    }
}

}  // namespace

int launch_embed(const void* tok_emb, int w_type, const float* pos_emb, const DecodeSeqState& s, int rows, int d, float* x, cudaStream_t st) {
    ProfScope ps(PC_EMBED, (double)rows * d * 10, st);
    if (w_type == DT_BF16) OH_CUDA_CHECK(launch_k(embed_kernel<__nv_bfloat16>, rows, 256, 0, st, (const __nv_bfloat16*)tok_emb, pos_emb, s, d, x));
    else OH_CUDA_CHECK(launch_k(embed_kernel<__half>, rows, 256, 0, st, (const __half*)tok_emb, pos_emb, s, d, x));
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

int launch_self_attention_decode(const void* qkv, void* kcache, void* vcache, int type, const DecodeSeqState& s, int rows, int d, int n_ctx,
                                 void* out, cudaStream_t st) {
    dim3 grid(d / kHeadDim, rows);
    if (n_ctx > 448) return OH_ERR_INVALID;
    OH_SKIP_IF(PC_SELF_ATTN);
    ProfScope ps(PC_SELF_ATTN, 0.0, st);
    if (type == DT_F32) OH_CUDA_CHECK(launch_k(self_attention_decode_kernel<float>, grid, 128, 0, st, (const float*)qkv, (float*)kcache, (float*)vcache, s, d, n_ctx, (float*)out));
    else if (type == DT_F16)
        OH_CUDA_CHECK(launch_k(self_attention_decode_warp_kernel<__half>, (rows * (d / kHeadDim) + 3) / 4, 128, 0, st, (const __half*)qkv, (__half*)kcache, (__half*)vcache, s, d, n_ctx, rows, (__half*)out));
    else
        OH_CUDA_CHECK(launch_k(self_attention_decode_warp_kernel<__nv_bfloat16>, (rows * (d / kHeadDim) + 3) / 4, 128, 0, st, (const __nv_bfloat16*)qkv, (__nv_bfloat16*)kcache, (__nv_bfloat16*)vcache, s, d, n_ctx, rows, (__nv_bfloat16*)out));
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

int launch_cross_attention_decode(const void* q, const void* xkv, int type, const DecodeSeqState& s, int rows, int d, int T, void* out,
                                  cudaStream_t st) {
    dim3 grid(d / kHeadDim, rows);
    if (T > 1504) return OH_ERR_INVALID;
    // algorithmic bytes: every row that holds a window streams its slot's K and V once (rows without one exit at once)
    OH_SKIP_IF(PC_CROSS_ATTN);
    const int rows_work = g_active_rows_hint >= 0 ? std::min(rows, g_active_rows_hint) : rows;
    ProfScope ps(PC_CROSS_ATTN, (double)rows_work * T * 2.0 * d * (type == DT_F32 ? 4 : 2), st);
    // OH_XATTN_OCC=1: one CTA per SM (shared memory padded past half an SM) so that other kernels always find room next
    // to it; the 6-stage ring of a single CTA per SM still covers the HBM latency
    // OH_XATTN_STAGES=n (2..6): depth of the K/V ring.  Fewer stages leave shared memory for a co-resident encoder CTA
    static const int xa_stages = [] {
        const char* o = getenv("OH_XATTN_STAGES");
        return o ? std::max(2, std::min(XA_STAGES, atoi(o))) : 4;  // measured on B200: 3..6 stages stream K/V equally fast (12.9 ms per 256-row step); 4 = 74 KB per CTA
    }();
    static const int xa_smem = [] {
        const char* o = getenv("OH_XATTN_OCC");
        return (o && atoi(o) == 1) ? 116 * 1024 : xa_smem_bytes(xa_stages);
    }();
    static const bool want_bulk = [] {
        const char* e = getenv("OH_XATTN");
        return !(e && strcmp(e, "online") == 0);
    }();
    static PerDeviceInit bulk_attr;  // the shared-memory opt-in is per device
    const bool use_bulk = want_bulk && type != DT_F32 && bulk_attr.ensure([] {
        return cudaFuncSetAttribute(cross_attention_decode_bulk_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, xa_smem) == cudaSuccess &&
               cudaFuncSetAttribute(cross_attention_decode_bulk_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, xa_smem) == cudaSuccess;
    });
    if (type == DT_F32) OH_CUDA_CHECK(launch_k(cross_attention_decode_kernel<float>, grid, 128, 0, st, (const float*)q, (const float*)xkv, s, d, T, (float*)out));
    else if (use_bulk && type == DT_F16)
        OH_CUDA_CHECK(launch_k(cross_attention_decode_bulk_kernel<__half>, grid, XA_THREADS, xa_smem, st, (const __half*)q, (const __half*)xkv, s, d, T, (__half*)out, xa_stages));
    else if (use_bulk)
        OH_CUDA_CHECK(launch_k(cross_attention_decode_bulk_kernel<__nv_bfloat16>, grid, XA_THREADS, xa_smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)xkv, s, d, T, (__nv_bfloat16*)out, xa_stages));
    else if (type == DT_F16)  // OH_XATTN=online, or the shared-memory opt-in failed: the two-pass kernel of the parity mode
        OH_CUDA_CHECK(launch_k(cross_attention_decode_kernel<__half>, grid, 128, 0, st, (const __half*)q, (const __half*)xkv, s, d, T, (__half*)out));
    else
        OH_CUDA_CHECK(launch_k(cross_attention_decode_kernel<__nv_bfloat16>, grid, 128, 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)xkv, s, d, T, (__nv_bfloat16*)out));
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

// self-attention cache lines follow the beams: grid (pair, layer * 2 + k/v); a line is [heads][n_ctx][64]
__global__ void __launch_bounds__(256) copy_self_lines_kernel(uint4* __restrict__ kcache, uint4* __restrict__ vcache, const int4* __restrict__ pairs,
                                                             int lines, int heads, int n_ctx, int vec_per_pos /* 64 * esz / 16 */) {
    pdl_wait();
    pdl_launch_dependents();
    const int4 pr = pairs[blockIdx.x];  // (src line, dst line, positions to copy, -)
    const int n_pos = pr.z;
    const int layer = blockIdx.y >> 1;
    uint4* base = (blockIdx.y & 1) ? vcache : kcache;
    const size_t line_vecs = (size_t)heads * n_ctx * vec_per_pos;
    const uint4* src = base + ((size_t)layer * lines + pr.x) * line_vecs;
    uint4* dst = base + ((size_t)layer * lines + pr.y) * line_vecs;
    const int per_head = n_pos * vec_per_pos;
    for (int idx = threadIdx.x; idx < heads * per_head; idx += blockDim.x) {
        const int h = idx / per_head, o = idx - h * per_head;
        dst[(size_t)h * n_ctx * vec_per_pos + o] = src[(size_t)h * n_ctx * vec_per_pos + o];
    }
}

int launch_copy_self_lines(void* kcache, void* vcache, int esz, const int4* pairs, int n_pairs, int n_layers, int lines, int heads, int n_ctx,
                           cudaStream_t st) {
    if (n_pairs <= 0) return OH_OK;
    OH_CUDA_CHECK(launch_k(copy_self_lines_kernel, dim3(n_pairs, n_layers * 2), 256, 0, st, (uint4*)kcache, (uint4*)vcache, pairs, lines, heads, n_ctx,
                           kHeadDim * esz / 16));
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

int launch_logits_step(const float* logits, int ld, const DecodeSeqState& s, const DecodeRules* r, int rows, float* dump_logprobs, float* dump_probs,
                       cudaStream_t st, const TopKOut& topk) {
    OH_SKIP_IF(PC_LOGITS);
    ProfScope ps(PC_LOGITS, (double)rows * ld * 4.0 * 2, st);
    OH_CUDA_CHECK(launch_k(logits_step_kernel, rows, K9_THREADS, 0, st, logits, ld, s, r, dump_logprobs, dump_probs, topk));
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace oh
