// ORACLE (test infrastructure) — encoder / cross-KV / decoder arithmetic.  Restates SURVEY.md
// Appendix A.4 (whisper.cpp `whisper_encode_internal` / `whisper_decode_internal`, out of tree; reached
// from /root/reference/src/engine/whisper.rs:266).  ggml numerics are modelled by the Options knobs:
// f16-stored weights, activations rounded to f16 at mat-mul inputs, f32 accumulation, f16 K/V,
// tanh-GELU through an f16 table, LayerNorm/softmax reductions in double.
#include <immintrin.h>
#include <omp.h>

#include <algorithm>

#include "orc_common.h"

namespace orc {

namespace {

inline float hsum(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_hadd_ps(lo, lo);
    lo = _mm_hadd_ps(lo, lo);
    return _mm_cvtss_f32(lo);
}

// MB rows of A against two rows of W, all K-contiguous; f32 accumulation
template <int MB, bool F16>
inline void dot_kernel(const float* a, int lda, const uint8_t* w0, const uint8_t* w1, int K, float* c0, float* c1) {
    __m256 acc0[MB], acc1[MB];
    for (int m = 0; m < MB; ++m) acc0[m] = acc1[m] = _mm256_setzero_ps();
    int k = 0;
    for (; k + 8 <= K; k += 8) {
        __m256 b0, b1;
        if (F16) {
            b0 = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)(w0 + 2 * (size_t)k)));
            b1 = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)(w1 + 2 * (size_t)k)));
        } else {
            b0 = _mm256_loadu_ps((const float*)(w0 + 4 * (size_t)k));
            b1 = _mm256_loadu_ps((const float*)(w1 + 4 * (size_t)k));
        }
        for (int m = 0; m < MB; ++m) {
            __m256 av = _mm256_loadu_ps(a + (size_t)m * lda + k);
            acc0[m] = _mm256_fmadd_ps(av, b0, acc0[m]);
            acc1[m] = _mm256_fmadd_ps(av, b1, acc1[m]);
        }
    }
    for (int m = 0; m < MB; ++m) {
        float s0 = hsum(acc0[m]), s1 = hsum(acc1[m]);
        for (int kk = k; kk < K; ++kk) {
            float x = a[(size_t)m * lda + kk];
            float y0 = F16 ? f16_to_f32(((const uint16_t*)w0)[kk]) : ((const float*)w0)[kk];
            float y1 = F16 ? f16_to_f32(((const uint16_t*)w1)[kk]) : ((const float*)w1)[kk];
            s0 += x * y0;
            s1 += x * y1;
        }
        c0[m] = s0;
        c1[m] = s1;
    }
}

template <bool F16>
void gemm_nt_t(const float* A, int lda, const uint8_t* W, size_t ldw_bytes, int M, int N, int K, const float* bias, float* C, int ldc,
               bool parallel) {
    const int NB = 16;
    const int n_blocks = (N + NB - 1) / NB;
#pragma omp parallel for schedule(static) if (parallel)
    for (int nb = 0; nb < n_blocks; ++nb) {
        const int n0 = nb * NB, n1 = std::min(N, n0 + NB);
        for (int m0 = 0; m0 < M; m0 += 4) {
            const int mb = std::min(4, M - m0);
            for (int n = n0; n < n1; n += 2) {
                const int nn = (n + 1 < n1) ? n + 1 : n;
                const uint8_t* w0 = W + (size_t)n * ldw_bytes;
                const uint8_t* w1 = W + (size_t)nn * ldw_bytes;
                float c0[4], c1[4];
                const float* a = A + (size_t)m0 * lda;
                switch (mb) {
                    case 4: dot_kernel<4, F16>(a, lda, w0, w1, K, c0, c1); break;
                    case 3: dot_kernel<3, F16>(a, lda, w0, w1, K, c0, c1); break;
                    case 2: dot_kernel<2, F16>(a, lda, w0, w1, K, c0, c1); break;
                    default: dot_kernel<1, F16>(a, lda, w0, w1, K, c0, c1); break;
                }
                for (int m = 0; m < mb; ++m) {
                    C[(size_t)(m0 + m) * ldc + n] = c0[m] + (bias ? bias[n] : 0.f);
                    if (nn != n) C[(size_t)(m0 + m) * ldc + nn] = c1[m] + (bias ? bias[nn] : 0.f);
                }
            }
        }
    }
}

// C[M][N] = A[M][K] . W[N][K]^T + bias
void gemm_w(const float* A, int lda, const Mat& W, int M, const float* bias, float* C, int ldc, bool parallel = true) {
    if (W.is_f16) gemm_nt_t<true>(A, lda, (const uint8_t*)W.h.data(), (size_t)W.cols * 2, M, W.rows, W.cols, bias, C, ldc, parallel);
    else gemm_nt_t<false>(A, lda, (const uint8_t*)W.f.data(), (size_t)W.cols * 4, M, W.rows, W.cols, bias, C, ldc, parallel);
}
void gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, float* C, int ldc) {
    gemm_nt_t<false>(A, lda, (const uint8_t*)W, (size_t)ldw * 4, M, N, K, nullptr, C, ldc, false);
}

void round_inplace(float* x, size_t n, int mode) {
    if (mode == ACT_NONE) return;
#pragma omp parallel for schedule(static) if (n > 65536)
    for (size_t i = 0; i < n; ++i) x[i] = round_act(x[i], mode);
}

inline float gelu_one(const Model& m, int mode, float x) {
    if (mode == GELU_TANH_F16LUT) {
        if (x <= -10.0f) return 0.0f;
        if (x >= 10.0f) return x;
        return f16_to_f32(m.gelu_table[f32_to_f16(x)]);
    }
    if (mode == GELU_TANH) return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
void gelu_inplace(const Model& m, int mode, float* x, size_t n) {
#pragma omp parallel for schedule(static) if (n > 65536)
    for (size_t i = 0; i < n; ++i) x[i] = gelu_one(m, mode, x[i]);
}

// y = (x - mean) / sqrt(var + 1e-5) * w + b, reductions in double (ggml_float)
void layer_norm(const float* x, int rows, int d, const float* w, const float* b, float* y) {
#pragma omp parallel for schedule(static) if (rows > 16)
    for (int r = 0; r < rows; ++r) {
        const float* xr = x + (size_t)r * d;
        float* yr = y + (size_t)r * d;
        double sum = 0.0;
        for (int i = 0; i < d; ++i) sum += (double)xr[i];
        float mean = (float)(sum / d);
        double sum2 = 0.0;
        for (int i = 0; i < d; ++i) {
            float v = xr[i] - mean;
            yr[i] = v;
            sum2 += (double)(v * v);
        }
        float variance = (float)(sum2 / d);
        const float scale = 1.0f / sqrtf(variance + 1e-5f);
        for (int i = 0; i < d; ++i) yr[i] = yr[i] * scale * w[i] + b[i];
    }
}

// in-place softmax(scale * s) over n entries; exp sum in double
void softmax_row(float* s, int n, float scale) {
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) {
        s[i] *= scale;
        mx = std::max(mx, s[i]);
    }
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        float e = (s[i] == -INFINITY) ? 0.0f : expf(s[i] - mx);
        s[i] = e;
        sum += (double)e;
    }
    const float inv = (float)(1.0 / sum);
    for (int i = 0; i < n; ++i) s[i] *= inv;
}

}  // namespace

void encode(const Model& m, const Options& opt, const float* mel, std::vector<float>& enc_out, std::vector<float>* conv_out) {
    const int C = m.hp.n_mels, d = m.hp.n_audio_state, T = m.hp.n_audio_ctx, T2 = 2 * T, H = m.hp.n_audio_head, dh = d / H;
    if (opt.n_threads > 0) omp_set_num_threads(opt.n_threads);

    // conv1 (k3 s1 p1) as im2col + mat-mul, +bias, GELU
    std::vector<float> col1((size_t)T2 * 3 * C);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < T2; ++i)
        for (int t = 0; t < 3; ++t) {
            int p = i + t - 1;
            for (int c = 0; c < C; ++c) col1[(size_t)i * 3 * C + t * C + c] = (p >= 0 && p < T2) ? mel[(size_t)c * T2 + p] : 0.f;
        }
    round_inplace(col1.data(), col1.size(), opt.act_round);
    std::vector<float> x1((size_t)(T2 + 2) * d, 0.f);  // one zero row before and after for conv2's padding
    float* x1p = x1.data() + d;
    gemm_w(col1.data(), 3 * C, m.conv1, T2, m.conv1_b.data(), x1p, d);
    gelu_inplace(m, opt.gelu, x1p, (size_t)T2 * d);
    round_inplace(x1p, (size_t)T2 * d, opt.act_round);
    col1.clear();
    col1.shrink_to_fit();

    // conv2 (k3 s2 p1): im2col row i is rows 2i-1, 2i, 2i+1 of x1, already contiguous in time-major layout
    std::vector<float> x((size_t)T * d);
    gemm_w(x1.data(), 2 * d, m.conv2, T, m.conv2_b.data(), x.data(), d);
    gelu_inplace(m, opt.gelu, x.data(), x.size());
    for (size_t i = 0; i < x.size(); ++i) x[i] += m.enc_pos[i];
    if (conv_out) *conv_out = x;
    x1.clear();
    x1.shrink_to_fit();

    std::vector<float> h((size_t)T * d), q((size_t)T * d), k((size_t)T * d), v((size_t)T * d), att((size_t)T * d);
    std::vector<float> hid((size_t)T * 4 * d);
    const float scale = 1.0f / sqrtf((float)dh);
    for (const EncLayer& L : m.enc) {
        layer_norm(x.data(), T, d, L.ln1_w.data(), L.ln1_b.data(), h.data());
        round_inplace(h.data(), h.size(), opt.act_round);
        gemm_w(h.data(), d, L.attn.q, T, L.attn.qb.data(), q.data(), d);
        gemm_w(h.data(), d, L.attn.k, T, nullptr, k.data(), d);
        gemm_w(h.data(), d, L.attn.v, T, L.attn.vb.data(), v.data(), d);
        round_inplace(q.data(), q.size(), opt.act_round);
        round_inplace(k.data(), k.size(), opt.act_round);
        round_inplace(v.data(), v.size(), opt.act_round);
#pragma omp parallel
        {
            std::vector<float> S((size_t)T * T), Vt((size_t)dh * T), O((size_t)T * dh);
#pragma omp for schedule(dynamic)
            for (int hd = 0; hd < H; ++hd) {
                const float* qh = q.data() + hd * dh;
                const float* kh = k.data() + hd * dh;
                const float* vh = v.data() + hd * dh;
                gemm_f32(qh, d, kh, d, T, T, dh, S.data(), T);
                for (int i = 0; i < T; ++i) {
                    softmax_row(S.data() + (size_t)i * T, T, scale);
                    if (opt.act_round != ACT_NONE)
                        for (int j = 0; j < T; ++j) S[(size_t)i * T + j] = round_act(S[(size_t)i * T + j], opt.act_round);
                }
                for (int j = 0; j < T; ++j)
                    for (int c = 0; c < dh; ++c) Vt[(size_t)c * T + j] = vh[(size_t)j * d + c];
                gemm_f32(S.data(), T, Vt.data(), T, T, dh, T, O.data(), dh);
                for (int i = 0; i < T; ++i)
                    for (int c = 0; c < dh; ++c) att[(size_t)i * d + hd * dh + c] = O[(size_t)i * dh + c];
            }
        }
        round_inplace(att.data(), att.size(), opt.act_round);
        gemm_w(att.data(), d, L.attn.o, T, L.attn.ob.data(), h.data(), d);
        for (size_t i = 0; i < x.size(); ++i) x[i] += h[i];

        layer_norm(x.data(), T, d, L.ln2_w.data(), L.ln2_b.data(), h.data());
        round_inplace(h.data(), h.size(), opt.act_round);
        gemm_w(h.data(), d, L.mlp0, T, L.mlp0_b.data(), hid.data(), 4 * d);
        gelu_inplace(m, opt.gelu, hid.data(), hid.size());
        round_inplace(hid.data(), hid.size(), opt.act_round);
        gemm_w(hid.data(), 4 * d, L.mlp2, T, L.mlp2_b.data(), h.data(), d);
        for (size_t i = 0; i < x.size(); ++i) x[i] += h[i];
    }
    enc_out.resize((size_t)T * d);
    layer_norm(x.data(), T, d, m.ln_post_w.data(), m.ln_post_b.data(), enc_out.data());
}

void cross_kv(const Model& m, const Options& opt, const std::vector<float>& enc_out, std::vector<float>& ck, std::vector<float>& cv) {
    const int d = m.hp.n_audio_state, dt = m.hp.n_text_state, T = m.hp.n_audio_ctx, L = m.hp.n_text_layer;
    std::vector<float> e = enc_out;
    round_inplace(e.data(), e.size(), opt.act_round);
    ck.resize((size_t)L * T * dt);
    cv.resize((size_t)L * T * dt);
    for (int l = 0; l < L; ++l) {
        gemm_w(e.data(), d, m.dec[l].cross.k, T, nullptr, ck.data() + (size_t)l * T * dt, dt);
        gemm_w(e.data(), d, m.dec[l].cross.v, T, m.dec[l].cross.vb.data(), cv.data() + (size_t)l * T * dt, dt);
    }
    round_inplace(ck.data(), ck.size(), opt.act_round);  // the cross cache is f16 in ggml
    round_inplace(cv.data(), cv.size(), opt.act_round);
}

void decode_token(const Model& m, const Options& opt, const std::vector<float>& ck, const std::vector<float>& cv, KVCache& kv, int token,
                  int n_past, float* logits, float* hidden_out) {
    const int d = m.hp.n_text_state, H = m.hp.n_text_head, dh = d / H, T = m.hp.n_audio_ctx, NC = m.hp.n_text_ctx, L = m.hp.n_text_layer;
    if (kv.k.empty()) {
        kv.k.assign((size_t)L * NC * d, 0.f);
        kv.v.assign((size_t)L * NC * d, 0.f);
    }
    std::vector<float> x(d), h(d), q(d), kk(d), vv(d), att(d), o(d), hid(4 * d);
    for (int i = 0; i < d; ++i) {
        float e = m.tok_emb.is_f16 ? f16_to_f32(m.tok_emb.h[(size_t)token * d + i]) : m.tok_emb.f[(size_t)token * d + i];
        x[i] = e + m.dec_pos[(size_t)n_past * d + i];
    }
    const float scale = 1.0f / sqrtf((float)dh);
    const int n_kv = n_past + 1;
    std::vector<float> s(std::max(n_kv, T) * (size_t)H);
    for (int l = 0; l < L; ++l) {
        const DecLayer& Lw = m.dec[l];
        // --- masked self-attention with KV cache
        layer_norm(x.data(), 1, d, Lw.ln1_w.data(), Lw.ln1_b.data(), h.data());
        round_inplace(h.data(), d, opt.act_round);
        gemm_w(h.data(), d, Lw.attn.q, 1, Lw.attn.qb.data(), q.data(), d);
        gemm_w(h.data(), d, Lw.attn.k, 1, nullptr, kk.data(), d);
        gemm_w(h.data(), d, Lw.attn.v, 1, Lw.attn.vb.data(), vv.data(), d);
        float* Kc = kv.k.data() + (size_t)l * NC * d;
        float* Vc = kv.v.data() + (size_t)l * NC * d;
        for (int i = 0; i < d; ++i) {
            Kc[(size_t)n_past * d + i] = round_act(kk[i], opt.act_round);
            Vc[(size_t)n_past * d + i] = round_act(vv[i], opt.act_round);
            q[i] = round_act(q[i], opt.act_round);
        }
        for (int hd = 0; hd < H; ++hd) {
            float* sh = s.data() + (size_t)hd * n_kv;
            for (int j = 0; j < n_kv; ++j) {
                float acc = 0.f;
                for (int c = 0; c < dh; ++c) acc += q[hd * dh + c] * Kc[(size_t)j * d + hd * dh + c];
                sh[j] = acc;
            }
            softmax_row(sh, n_kv, scale);
            for (int c = 0; c < dh; ++c) {
                float acc = 0.f;
                for (int j = 0; j < n_kv; ++j) acc += round_act(sh[j], opt.act_round) * Vc[(size_t)j * d + hd * dh + c];
                att[hd * dh + c] = acc;
            }
        }
        round_inplace(att.data(), d, opt.act_round);
        gemm_w(att.data(), d, Lw.attn.o, 1, Lw.attn.ob.data(), o.data(), d);
        for (int i = 0; i < d; ++i) x[i] += o[i];

        // --- cross-attention over the encoder states
        layer_norm(x.data(), 1, d, Lw.lnx_w.data(), Lw.lnx_b.data(), h.data());
        round_inplace(h.data(), d, opt.act_round);
        gemm_w(h.data(), d, Lw.cross.q, 1, Lw.cross.qb.data(), q.data(), d);
        round_inplace(q.data(), d, opt.act_round);
        const float* Kx = ck.data() + (size_t)l * T * d;
        const float* Vx = cv.data() + (size_t)l * T * d;
#pragma omp parallel for schedule(static)
        for (int hd = 0; hd < H; ++hd) {
            float* sh = s.data() + (size_t)hd * T;
            for (int j = 0; j < T; ++j) {
                float acc = 0.f;
                for (int c = 0; c < dh; ++c) acc += q[hd * dh + c] * Kx[(size_t)j * d + hd * dh + c];
                sh[j] = acc;
            }
            softmax_row(sh, T, scale);
            if (opt.act_round != ACT_NONE)
                for (int j = 0; j < T; ++j) sh[j] = round_act(sh[j], opt.act_round);
            float acc[128];
            for (int c = 0; c < dh; ++c) acc[c] = 0.f;
            for (int j = 0; j < T; ++j)
                for (int c = 0; c < dh; ++c) acc[c] += sh[j] * Vx[(size_t)j * d + hd * dh + c];
            for (int c = 0; c < dh; ++c) att[hd * dh + c] = acc[c];
        }
        round_inplace(att.data(), d, opt.act_round);
        gemm_w(att.data(), d, Lw.cross.o, 1, Lw.cross.ob.data(), o.data(), d);
        for (int i = 0; i < d; ++i) x[i] += o[i];

        // --- MLP
        layer_norm(x.data(), 1, d, Lw.ln2_w.data(), Lw.ln2_b.data(), h.data());
        round_inplace(h.data(), d, opt.act_round);
        gemm_w(h.data(), d, Lw.mlp0, 1, Lw.mlp0_b.data(), hid.data(), 4 * d);
        gelu_inplace(m, opt.gelu, hid.data(), hid.size());
        round_inplace(hid.data(), hid.size(), opt.act_round);
        gemm_w(hid.data(), 4 * d, Lw.mlp2, 1, Lw.mlp2_b.data(), o.data(), d);
        for (int i = 0; i < d; ++i) x[i] += o[i];
    }
    if (!logits && !hidden_out) return;
    layer_norm(x.data(), 1, d, m.ln_w.data(), m.ln_b.data(), h.data());
    if (hidden_out) std::memcpy(hidden_out, h.data(), sizeof(float) * d);
    if (logits) {
        round_inplace(h.data(), d, opt.act_round);
        gemm_w(h.data(), d, m.tok_emb, 1, nullptr, logits, m.hp.n_vocab);
    }
}

}  // namespace orc
