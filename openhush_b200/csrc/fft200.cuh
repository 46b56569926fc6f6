// 400-point real DFT for the log-mel front end, as a 200-point complex Stockham
// autosort FFT (radices 5,5,4,2) over shared memory plus the real-input split step.
// Written as host/device code so the index arithmetic is unit-tested on the CPU (tests/test_fft_host.py
// builds this header into a tiny host library and compares with numpy.fft.rfft).
#pragma once
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define OH_HD __host__ __device__ __forceinline__
#else
#define OH_HD inline
#endif

namespace oh {
namespace fft {

constexpr int N2 = 200;    // complex FFT size
constexpr int NTW = 400;   // twiddle table: tw[i] = (cos(2 pi i/400), -sin(2 pi i/400)) = exp(-2 pi i * i/400)

OH_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
OH_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
OH_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// One Stockham pass of radix R over `in` -> `out` (both N2 complex), executed by `nthreads` cooperating
// threads of which this is `tid`.  Ns = product of the radices of the passes already done.
template <int R>
OH_HD void stockham_pass(const float2* in, float2* out, const float2* tw, int Ns, int tid, int nthreads) {
    constexpr int NB = N2 / R;  // butterflies in this pass
    for (int j = tid; j < NB; j += nthreads) {
        const int k = j % Ns;
        float2 v[R];
        // load + inter-pass twiddle exp(-2 pi i k t / (Ns R)); table step = 400 / (Ns R)
        const int tstep = NTW / (Ns * R);
#pragma unroll
        for (int t = 0; t < R; ++t) {
            float2 x = in[j + t * NB];
            if (t > 0 && k > 0) x = cmul(x, tw[(k * t * tstep) % NTW]);
            v[t] = x;
        }
        float2 o[R];
        if (R == 2) {
            o[0] = cadd(v[0], v[1]);
            o[1] = csub(v[0], v[1]);
        } else if (R == 4) {
            // exp(-2 pi i/4) = -i
            float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
            float2 mid = make_float2(d.y, -d.x);  // -i * d
            o[0] = cadd(a, c);
            o[1] = cadd(b, mid);
            o[2] = csub(a, c);
            o[3] = csub(b, mid);
        } else {
            // generic small DFT with table twiddles exp(-2 pi i r t / R): index r t (400/R)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 acc = v[0];
#pragma unroll
                for (int t = 1; t < R; ++t) acc = cadd(acc, cmul(v[t], tw[((r * t) % R) * (NTW / R)]));
                o[r] = acc;
            }
        }
        const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[j0 + r * Ns] = o[r];
    }
}

// |X[k]|^2, k in [0, 200], of the real 400-point sequence packed as z[n] = x[2n] + i x[2n+1], given Z = FFT200(z)
OH_HD float real_power_bin(const float2* Z, const float2* tw, int k) {
    const float2 a = Z[k == N2 ? 0 : k];
    const float2 bq = Z[(N2 - k) % N2];
    const float2 b = make_float2(bq.x, -bq.y);  // conj(Z[N2-k])
    const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
    const float2 dmb = csub(a, b);
    const float2 o = make_float2(0.5f * dmb.y, -0.5f * dmb.x);  // (a-b)/(2i)
    const float2 w = tw[k % NTW];                                // exp(-2 pi i k/400); k=200 -> -1
    const float2 x = cadd(e, cmul(w, o));
    return x.x * x.x + x.y * x.y;
}

}  // namespace fft
}  // namespace oh
