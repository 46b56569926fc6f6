// Host build of the device FFT header (openhush_b200/csrc/fft200.cuh) for CPU unit tests: emulates the
// `nthreads` cooperating lanes by looping over tid between passes (each pass is a barrier on the GPU).
#include <cmath>
#include <vector>

#include "../../openhush_b200/csrc/fft200.cuh"

using namespace oh::fft;

extern "C" __attribute__((visibility("default"))) void fft400_power(const float* x, float* p201, int nthreads) {
    std::vector<float2> tw(NTW), a(N2), b(N2);
    for (int i = 0; i < NTW; ++i) {
        double th = 2.0 * M_PI * i / NTW;
        tw[i] = make_float2((float)cos(th), (float)-sin(th));
    }
    for (int n = 0; n < N2; ++n) a[n] = make_float2(x[2 * n], x[2 * n + 1]);
    for (int t = 0; t < nthreads; ++t) stockham_pass<5>(a.data(), b.data(), tw.data(), 1, t, nthreads);
    for (int t = 0; t < nthreads; ++t) stockham_pass<5>(b.data(), a.data(), tw.data(), 5, t, nthreads);
    for (int t = 0; t < nthreads; ++t) stockham_pass<4>(a.data(), b.data(), tw.data(), 25, t, nthreads);
    for (int t = 0; t < nthreads; ++t) stockham_pass<2>(b.data(), a.data(), tw.data(), 100, t, nthreads);
    for (int k = 0; k <= N2; ++k) p201[k] = real_power_bin(a.data(), tw.data(), k);
}
