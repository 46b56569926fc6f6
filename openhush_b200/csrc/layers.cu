// K5 — LayerNorm (eps 1e-5) over rows of the f32 residual stream, output in the activation type of the
// following mat-mul (ggml rounds mat-mul inputs to f16, so emitting 16-bit here is lossless
// with respect to the reference arithmetic).  One warp per row, row kept in registers, two-pass statistics
// (mean, then variance of the centred values) like the reference's CPU kernel.
#include "../../include/oh_whisper.h"
#include "kernels.h"

namespace oh {

namespace {

constexpr int LN_MAX_PER_LANE = 40;  // d <= 1280

template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int rows, int d, const float* __restrict__ w,
                                                        const float* __restrict__ b, T* __restrict__ y) {
    pdl_wait();  // programmatic dependent launch (see common.cuh)
    pdl_launch_dependents();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + (size_t)row * d;
    float v[LN_MAX_PER_LANE];
    const int n4 = d >> 2;  // float4 chunks per row; lane handles chunks lane, lane+32, ...
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < LN_MAX_PER_LANE / 4; ++c) {
        const int idx = lane + 32 * c;
        if (idx < n4) {
            const float4 q = *reinterpret_cast<const float4*>(xr + 4 * idx);
            v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
            sum += (q.x + q.y) + (q.z + q.w);
        } else {
            v[4 * c] = v[4 * c + 1] = v[4 * c + 2] = v[4 * c + 3] = 0.0f;
        }
    }
    const float mean = warp_sum(sum) / (float)d;
    float sq = 0.0f;
#pragma unroll
    for (int c = 0; c < LN_MAX_PER_LANE / 4; ++c) {
        const int idx = lane + 32 * c;
        if (idx < n4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[4 * c + j] -= mean;
                sq = fmaf(v[4 * c + j], v[4 * c + j], sq);
            }
        }
    }
    const float var = warp_sum(sq) / (float)d;
    const float scale = 1.0f / sqrtf(var + 1e-5f);
    T* yr = y + (size_t)row * d;
#pragma unroll
    for (int c = 0; c < LN_MAX_PER_LANE / 4; ++c) {
        const int idx = lane + 32 * c;
        if (idx < n4) {
            const float4 ww = *reinterpret_cast<const float4*>(w + 4 * idx), bb = *reinterpret_cast<const float4*>(b + 4 * idx);
            const float o0 = v[4 * c] * scale * ww.x + bb.x, o1 = v[4 * c + 1] * scale * ww.y + bb.y;
            const float o2 = v[4 * c + 2] * scale * ww.z + bb.z, o3 = v[4 * c + 3] * scale * ww.w + bb.w;
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(yr + 4 * idx) = make_float4(o0, o1, o2, o3);
            } else {
                T t[4] = {Cvt<T>::from_f32(o0), Cvt<T>::from_f32(o1), Cvt<T>::from_f32(o2), Cvt<T>::from_f32(o3)};
                *reinterpret_cast<uint2*>(yr + 4 * idx) = *reinterpret_cast<uint2*>(t);
            }
        }
    }
}

}  // namespace

int launch_layernorm(const float* x, int rows, int d, const float* w, const float* b, void* y, int y_type, cudaStream_t st) {
    if (rows <= 0) return OH_OK;
    if (d % 4 != 0 || d > LN_MAX_PER_LANE * 32) {
        set_error("layernorm: unsupported width");
        return OH_ERR_INVALID;
    }
    const int warps = 8;
    OH_SKIP_IF(PC_LN);
    ProfScope ps(PC_LN, (double)rows * d * (4.0 + (y_type == DT_F32 ? 4 : 2)), st);
    dim3 grid((rows + warps - 1) / warps);
    if (y_type == DT_F32) OH_CUDA_CHECK(launch_k(layernorm_kernel<float>, grid, warps * 32, 0, st, x, rows, d, w, b, (float*)y));
    else if (y_type == DT_F16) OH_CUDA_CHECK(launch_k(layernorm_kernel<__half>, grid, warps * 32, 0, st, x, rows, d, w, b, (__half*)y));
    else OH_CUDA_CHECK(launch_k(layernorm_kernel<__nv_bfloat16>, grid, warps * 32, 0, st, x, rows, d, w, b, (__nv_bfloat16*)y));
    OH_COUNT_LAUNCH();
    OH_CUDA_CHECK(cudaGetLastError());
    return OH_OK;
}

}  // namespace oh
