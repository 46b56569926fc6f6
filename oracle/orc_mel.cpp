// ORACLE (test infrastructure) — log-mel front end.  Restates SURVEY.md Appendix A.3, i.e. the
// `whisper_pcm_to_mel` step inside the call at /root/reference/src/engine/whisper.rs:266
// (whisper.cpp `log_mel_spectrogram`, out of tree; see orc_common.h header for provenance).
#include <algorithm>

#include "orc_common.h"

namespace orc {

namespace {

constexpr int SR = 16000, N_FFT = 400, HOP = 160, N_BINS = 201, CHUNK = 30;

struct Tables {
    float hann[N_FFT], sinv[N_FFT], cosv[N_FFT];
    Tables() {
        for (int i = 0; i < N_FFT; ++i) {
            hann[i] = 0.5f * (1.0f - cosf((2.0f * (float)M_PI * i) / N_FFT));  // periodic Hann
            double th = 2.0 * M_PI * i / N_FFT;
            sinv[i] = sinf((float)th);
            cosv[i] = cosf((float)th);
        }
    }
};
const Tables& tables() {
    static Tables t;
    return t;
}

// direct DFT for the odd-size leaf (N = 25), f32, twiddles from the 400-entry table
void dft_leaf(const float* in, int N, float* out) {
    const Tables& tb = tables();
    const int step = N_FFT / N;
    for (int k = 0; k < N; ++k) {
        float re = 0.f, im = 0.f;
        for (int n = 0; n < N; ++n) {
            int idx = (k * n * step) % N_FFT;
            re += in[n] * tb.cosv[idx];
            im -= in[n] * tb.sinv[idx];
        }
        out[2 * k] = re;
        out[2 * k + 1] = im;
    }
}

// radix-2 decimation in time down to an odd size, then the direct DFT (A.3 step 4)
void fft_rec(const float* in, int N, float* out, float* scratch) {
    if (N == 1) {
        out[0] = in[0];
        out[1] = 0.f;
        return;
    }
    const int half = N / 2;
    if (N - half * 2 == 1) {
        dft_leaf(in, N, out);
        return;
    }
    float* even = scratch;
    float* odd = scratch + half;
    float* even_f = scratch + 2 * half;          // 2*half floats
    float* odd_f = scratch + 4 * half;           // 2*half floats
    float* next = scratch + 6 * half;
    for (int i = 0; i < half; ++i) {
        even[i] = in[2 * i];
        odd[i] = in[2 * i + 1];
    }
    fft_rec(even, half, even_f, next);
    fft_rec(odd, half, odd_f, next);
    const Tables& tb = tables();
    const int step = N_FFT / N;
    for (int k = 0; k < half; ++k) {
        int idx = k * step;
        float re = tb.cosv[idx], im = -tb.sinv[idx];
        float ro = odd_f[2 * k], io = odd_f[2 * k + 1];
        out[2 * k] = even_f[2 * k] + re * ro - im * io;
        out[2 * k + 1] = even_f[2 * k + 1] + re * io + im * ro;
        out[2 * (k + half)] = even_f[2 * k] - re * ro + im * io;
        out[2 * (k + half) + 1] = even_f[2 * k + 1] - re * io - im * ro;
    }
}

}  // namespace

void log_mel_spectrogram(const Model& m, const float* pcm, int n, Mel& mel, int n_threads) {
    const Tables& tb = tables();
    const int n_mel = m.hp.n_mels;
    const int64_t pad1 = (int64_t)SR * CHUNK, pad2 = N_FFT / 2;
    std::vector<float> padded((size_t)(n + pad1 + 2 * pad2), 0.0f);
    std::copy(pcm, pcm + n, padded.begin() + pad2);
    for (int j = 0; j < pad2; ++j) padded[j] = pcm[pad2 - j];  // reflect: padded[0]=pcm[200] ... padded[199]=pcm[1]

    mel.n_mel = n_mel;
    mel.n_len = (int)((padded.size() - N_FFT) / HOP);
    mel.n_len_org = 1 + (int)((n + pad2 - N_FFT) / HOP);
    mel.data.assign((size_t)n_mel * mel.n_len, 0.0f);

    const int n_samples = n + (int)pad2;
    const int n_active = std::min(n_samples / HOP + 1, mel.n_len);
    const float* filt = m.mel_filters.data();
    (void)n_threads;
#pragma omp parallel
    {
        std::vector<float> fft_in(N_FFT), fft_out(2 * N_FFT), scratch(8 * N_FFT);
#pragma omp for schedule(static)
        for (int i = 0; i < mel.n_len; ++i) {
            if (i >= n_active) {
                for (int j = 0; j < n_mel; ++j) mel.data[(size_t)j * mel.n_len + i] = (float)log10(1e-10);
                continue;
            }
            const int off = i * HOP;
            const int valid = std::min(N_FFT, n_samples - off);
            for (int j = 0; j < valid; ++j) fft_in[j] = tb.hann[j] * padded[off + j];
            for (int j = std::max(valid, 0); j < N_FFT; ++j) fft_in[j] = 0.f;
            fft_rec(fft_in.data(), N_FFT, fft_out.data(), scratch.data());
            for (int k = 0; k < N_BINS; ++k) fft_out[k] = fft_out[2 * k] * fft_out[2 * k] + fft_out[2 * k + 1] * fft_out[2 * k + 1];
            for (int j = 0; j < n_mel; ++j) {
                double sum = 0.0;
                const float* fr = filt + (size_t)j * N_BINS;
                int k = 0;
                for (; k < N_BINS - 3; k += 4)
                    sum += fft_out[k] * fr[k] + fft_out[k + 1] * fr[k + 1] + fft_out[k + 2] * fr[k + 2] + fft_out[k + 3] * fr[k + 3];
                for (; k < N_BINS; ++k) sum += fft_out[k] * fr[k];
                sum = log10(std::max(sum, 1e-10));
                mel.data[(size_t)j * mel.n_len + i] = (float)sum;
            }
        }
    }
    // global max over ALL entries including the zero-padding frames, clamp to max-8, (x+4)/4  (A.3 step 5)
    double mmax = -1e20;
    for (float v : mel.data)
        if (v > mmax) mmax = v;
    mmax -= 8.0;
    for (float& v : mel.data) {
        if (v < mmax) v = (float)mmax;
        v = (float)((v + 4.0) / 4.0);
    }
}

}  // namespace orc
