// Shared declarations of libohwhisper (product code; never includes anything from oracle/).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace oh {

void set_error(const std::string& msg);

#define OH_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            ::oh::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                            std::to_string(__LINE__));                                                   \
            return OH_ERR_CUDA;                                                                          \
        }                                                                                                \
    } while (0)

constexpr int kSampleRate = 16000;
constexpr int kNFFT = 400;
constexpr int kHop = 160;
constexpr int kNBins = 201;
constexpr int kChunkFrames = 3000;  // mel frames per 30 s window
constexpr int kHeadDim = 64;        // every Whisper size
constexpr int kMaxMelFilterWeights = 1024;  // sparse slaney bank staged in shared memory by the mel kernel (80 mel: 486, 128 mel: 522 weights)

// launch counter for bench.py's gpu_launches (host side, incremented by every kernel launch wrapper)
extern thread_local long long g_launch_count;
#define OH_COUNT_LAUNCH() (++::oh::g_launch_count)

// Optional per-kernel timing (oh_profile): every launch wrapper opens a ProfScope, which brackets the launch with
// two CUDA events on the launching stream when a profiler is attached to the calling thread (and costs one
// thread-local load otherwise).  `work` is the launch's algorithmic work: FLOPs for mat-muls, bytes otherwise.
enum ProfCat { PC_MEL = 0, PC_GATHER, PC_GEMM_ENC, PC_GEMM_DEC, PC_LN, PC_ATTN_ENC, PC_SELF_ATTN, PC_CROSS_ATTN, PC_EMBED, PC_LOGITS, PC_RESAMPLE, PC_COUNT };
struct Profiler {
    struct Rec {
        int cat;
        double work;
        cudaEvent_t e0, e1;
    };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> free_events;
};
extern thread_local Profiler* g_prof;
extern thread_local int g_active_rows_hint;  // rows that hold a window in the step being launched (-1: all); only sizes the profiler's `work`
// Timing experiments (scripts/bench_decode_step.py attributes a step by elimination): a DIAGNOSTIC build
// (`make DIAG=1`, -DOH_DIAG) can skip whole kernel categories through OH_SKIP_CATS; results are then meaningless.
// The release library has no such switch: the macro compiles to nothing.
#ifdef OH_DIAG
extern unsigned g_skip_mask;
#define OH_SKIP_IF(cat) do { if (::oh::g_skip_mask & (1u << (cat))) return OH_OK; } while (0)
#else
#define OH_SKIP_IF(cat) do { } while (0)
#endif

// One-time set-up that CUDA scopes per DEVICE (cudaFuncSetAttribute opt-ins, device properties, self-tests): `init`
// runs once per device ordinal, under a lock, on the calling thread's current device; later calls return what it
// returned.  (One worker thread per GPU inside one process is the multi-GPU host design: INTEGRATION.md section 3.)
class PerDeviceInit {
   public:
    static constexpr int kMaxDevices = 64;
    template <typename F>
    bool ensure(F&& init) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
        std::lock_guard<std::mutex> lk(mu_);
        if (!done_[dev]) {
            ok_[dev] = init();
            done_[dev] = true;
        }
        return ok_[dev];
    }
    static int current_device() {
        int dev = 0;
        cudaGetDevice(&dev);
        return dev < 0 || dev >= kMaxDevices ? 0 : dev;
    }

   private:
    std::mutex mu_;
    bool done_[kMaxDevices] = {};
    bool ok_[kMaxDevices] = {};
};
struct ProfScope {
    int idx = -1;
    cudaStream_t st;
    ProfScope(int cat, double work, cudaStream_t s) : st(s) {
        Profiler* p = g_prof;
        if (!p) return;
        Profiler::Rec r;
        r.cat = cat;
        r.work = work;
        for (cudaEvent_t* e : {&r.e0, &r.e1}) {
            if (!p->free_events.empty()) {
                *e = p->free_events.back();
                p->free_events.pop_back();
            } else {
                cudaEventCreate(e);
            }
        }
        cudaEventRecord(r.e0, st);
        idx = (int)p->recs.size();
        p->recs.push_back(r);
    }
    ~ProfScope() {
        if (idx >= 0 && g_prof) cudaEventRecord(g_prof->recs[idx].e1, st);
    }
};

// ---- programmatic dependent launch (decode step) ----------------------------------------------------
// The decoder step is a chain of ~360 short kernels.  Inside the step every kernel is launched with the
// programmatic-stream-serialization attribute: it may be scheduled while its predecessor is still running, does
// whatever does not depend on it (barrier / TMEM set-up, weight loads), and blocks in pdl_wait() until the predecessor
// has completed and flushed.  Rules every kernel of the chain follows: pdl_wait() before the first read of anything an
// earlier kernel wrote and before the first global write, on every path (early exits included) — that also keeps the
// ordering transitive.  Both instructions are no-ops in a kernel launched without the attribute.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
extern thread_local int g_sm_limit;  // > 0: persistent kernels launched by this thread size their grid for this many SMs (encoder on an SM partition)
extern thread_local int g_gemm_compact;        // > 0: tcgen05 mat-muls launched by this thread use 128-wide tiles and a ring of at most
extern thread_local int g_gemm_compact_bytes;  //      g_gemm_compact_bytes, leaving shared memory / TMEM to co-resident decode kernels
extern thread_local int g_pdl;  // > 0 while the calling thread is inside a decode step (set by decode_step_launch)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl > 0 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ---- 16-bit activation type helpers ---------------------------------------------------------------
template <typename T>
struct Cvt;
template <>
struct Cvt<float> {
    static __device__ __forceinline__ float to_f32(float v) { return v; }
    static __device__ __forceinline__ float from_f32(float v) { return v; }
};
template <>
struct Cvt<__half> {
    static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
};
template <>
struct Cvt<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
};

// GELU, tanh form.  Parity mode: ggml's f16 lookup table, table built on the host with
// tanhf exactly as ggml does.  Tensor-core modes: x * sigmoid(2u) in f32 on the 16-bit-rounded input.
__device__ __forceinline__ float gelu_lut(float x, const uint16_t* __restrict__ lut) {
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    unsigned short h = __half_as_ushort(__float2half_rn(x));
    return __half2float(__ushort_as_half(lut[h]));
}
// f32 parity mode: the same tanh formula evaluated in full f32 (no table, so no f16 rounding boundary that
// would amplify summation-order noise into token flips)
__device__ __forceinline__ float gelu_tanh(float x) {
    return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
}
__device__ __forceinline__ float gelu_fast(float x) {
    // 0.5 x (1 + tanh(u)) = x / (1 + exp(-2u)),  u = sqrt(2/pi) x (1 + 0.044715 x^2)
    const float u = 0.79788456080286535588f * x * (1.0f + 0.044715f * x * x);
    const float e = exp2f(-2.0f * 1.44269504088896340736f * u);
    return __fdividef(x, 1.0f + e);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace oh
