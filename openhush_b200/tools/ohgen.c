/*
 * ohgen — deterministic synthetic inputs for the transcription hot path.
 *
 * There is no network and no checkpoint on the build or GPU boxes, so every model file and
 * every PCM clip used by the tests and by bench.py is generated here from a counter-based
 * integer RNG (SplitMix64 of seed+index), i.e. element i of a tensor never depends on how many
 * threads produced it.  Nothing in this file is on the product path or in the oracle: it only
 * writes inputs that both of them read.
 *
 * What it writes:
 *   - a whisper.cpp legacy-"ggml" model container (the format named by the
 *     reference's model catalogue, /root/reference/src/engine/whisper.rs:71-79 "ggml-<name>.bin"):
 *     hparams, slaney mel filterbank, vocabulary, tensors named by OpenAI state-dict keys, f16 for
 *     >=2-D weights (ftype 1), f32 for biases / LayerNorm / positional embeddings.
 *   - 16 kHz mono f32 clips: band-limited noise + three parabolic-sine chirps, |x| <= 0.5.
 *
 * The weights are random but *shaped* (SURVEY.md §7 hard part 2): the tied token embedding gets a
 * larger scale so greedy distributions are peaked, and the EOT / timestamp rows get a further gain
 * so decodes contain timestamp pairs and terminate — this is what makes whisper's timestamp rules,
 * seek logic and segmenting actually execute on synthetic weights.
 *
 * Build: see openhush_b200/csrc/Makefile (target libohgen.so).  C99 + OpenMP.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OHGEN_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- RNG */

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* uniform in [-1, 1) with a 24-bit mantissa: exact in f32, no libm */
static inline float unif_pm1(uint64_t seed, uint64_t idx) {
    uint32_t m = (uint32_t)(splitmix64(seed + idx) >> 40); /* 24 bits */
    return (float)m * (1.0f / 8388608.0f) - 1.0f;
}

static inline uint64_t tensor_seed(uint64_t seed, uint64_t tensor_index) {
    return splitmix64(seed * 0x100000001B3ull + tensor_index) & 0x0000FFFFFFFFFFFFull;
}

/* ---------------------------------------------------------------- audio */

/* parabolic sine on a 32-bit phase accumulator (Bhaskara-like, no libm): phase in [0,2^32) */
static inline float psin_u32(uint32_t ph) {
    /* map to x in [-1,1): x = ph/2^31 - 1, sin(pi*x) ~= 4x(1-|x|) */
    float x = (float)(int32_t)(ph ^ 0x80000000u) * (1.0f / 2147483648.0f);
    float ax = x < 0 ? -x : x;
    float y = 4.0f * x * (1.0f - ax);
    float ay = y < 0 ? -y : y;
    return 0.225f * (y * ay - y) + y; /* refinement, |err| < 1e-3 */
}

/*
 * One clip: 4-tap smoothed uniform noise (amp 0.12) + three linear chirps whose start frequency,
 * slope and amplitude depend on the seed.  Peak amplitude stays below 0.5.
 * Clips of different lengths with the same seed share their common prefix.
 */
OHGEN_API void ohgen_audio(uint64_t seed, int64_t n, float* out) {
    uint64_t s = splitmix64(seed ^ 0xA5D10A5D10ull) & 0x0000FFFFFFFFFFFFull;
    /* chirp parameters from the seed */
    double f0[3], f1[3];
    float amp[3];
    for (int c = 0; c < 3; ++c) {
        double u0 = (double)(splitmix64(s + 1000 + c) >> 11) / 9007199254740992.0;
        double u1 = (double)(splitmix64(s + 2000 + c) >> 11) / 9007199254740992.0;
        double u2 = (double)(splitmix64(s + 3000 + c) >> 11) / 9007199254740992.0;
        f0[c] = 120.0 + 2500.0 * u0;             /* Hz at t=0 */
        f1[c] = 200.0 + 5000.0 * u1;             /* Hz at t=30 s */
        amp[c] = 0.05f + 0.07f * (float)u2;      /* <= 0.12 */
    }
    const double SR = 16000.0, T = 480000.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float nz = 0.0f;
        for (int k = 0; k < 4; ++k) {
            int64_t j = i - k;
            if (j >= 0) nz += unif_pm1(s, (uint64_t)j);
        }
        float v = 0.03f * nz; /* |.| <= 0.12 */
        for (int c = 0; c < 3; ++c) {
            /* phase(t) = f0*t + 0.5*(f1-f0)/T * t^2 cycles; evaluated in double, wrapped to u32 */
            double t = (double)i;
            double cyc = (f0[c] * t + 0.5 * (f1[c] - f0[c]) * t * t / T) / SR;
            double fr = cyc - floor(cyc);
            uint32_t ph = (uint32_t)(fr * 4294967296.0);
            /* slow amplitude envelope so there are louder and quieter stretches */
            uint32_t eph = (uint32_t)((uint64_t)i * (uint64_t)(8947u + 1733u * (uint32_t)c));
            float env = 0.6f + 0.4f * psin_u32(eph);
            v += amp[c] * env * psin_u32(ph);
        }
        out[i] = v;
    }
}

/* ---------------------------------------------------------------- mel filterbank (slaney) */

static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / (200.0 / 3.0);
    const double logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = 1000.0 / (200.0 / 3.0);
    const double logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

/* filters[n_mel][201], the layout stored in the model file (Appendix A.1) */
OHGEN_API void ohgen_mel_filters(int n_mel, float* filters) {
    const int n_fft_bins = 201;
    double* pts = (double*)malloc(sizeof(double) * (n_mel + 2));
    double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0);
    for (int i = 0; i < n_mel + 2; ++i) pts[i] = mel_to_hz(m0 + (m1 - m0) * (double)i / (double)(n_mel + 1));
    for (int m = 0; m < n_mel; ++m) {
        double lo = pts[m], ce = pts[m + 1], hi = pts[m + 2];
        double enorm = 2.0 / (hi - lo);
        for (int k = 0; k < n_fft_bins; ++k) {
            double f = (double)k * 16000.0 / 400.0;
            double up = (f - lo) / (ce - lo), down = (hi - f) / (hi - ce);
            double w = up < down ? up : down;
            if (w < 0) w = 0;
            filters[m * n_fft_bins + k] = (float)(w * enorm);
        }
    }
    free(pts);
}

/* ---------------------------------------------------------------- model file */

typedef struct {
    int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
} ohgen_hparams;

typedef struct {
    double emb_logit_std;  /* target std of a text-token logit (sets the tied embedding scale) */
    double ts_gain;        /* extra gain on timestamp rows (ids >= beg) */
    double eot_gain;       /* extra gain on the EOT row */
    double w_gain;         /* gain on 1/sqrt(fan_in) for all linear / conv weights */
    double emb_norm;       /* L2 norm of a text-token embedding row (kept small so the tied projection does not
                              just echo the input token); decoder.ln.weight is scaled by emb_logit_std/emb_norm */
    double qk_gain;        /* extra gain on attention query and key weights: peaky, input-dependent softmaxes */
    double ts_slope;       /* timestamp row k gets ts_gain * (1 - ts_slope * k / 1500): early timestamps are likelier,
                              so (with the monotone rule) decoded timestamps walk forward as in speech */
} ohgen_shaping;

static void w32(FILE* f, int32_t v) { fwrite(&v, 4, 1, f); }

static void write_tensor_header(FILE* f, const char* name, int n_dims, const int64_t* torch_shape, int ttype) {
    w32(f, n_dims);
    w32(f, (int32_t)strlen(name));
    w32(f, ttype);
    for (int i = n_dims - 1; i >= 0; --i) w32(f, (int32_t)torch_shape[i]); /* ne[] is reversed */
    fwrite(name, 1, strlen(name), f);
}

/* value(i) = offset + scale * row_gain(row) * U(-1,1); stored f16 (ttype 1) or f32 (ttype 0) */
static void write_random_tensor(FILE* f, const char* name, int n_dims, const int64_t* shape, int ttype,
                                uint64_t seed, float scale, float offset,
                                const float* row_gain /* per leading-dim row, or NULL */) {
    write_tensor_header(f, name, n_dims, shape, ttype);
    int64_t n = 1;
    for (int i = 0; i < n_dims; ++i) n *= shape[i];
    int64_t row_len = n / shape[0];
    const int64_t CH = 1 << 22;
    void* buf = malloc((size_t)CH * (ttype ? 2 : 4));
    for (int64_t base = 0; base < n; base += CH) {
        int64_t m = n - base < CH ? n - base : CH;
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < m; ++j) {
            int64_t i = base + j;
            float g = row_gain ? row_gain[i / row_len] : 1.0f;
            float v = offset + scale * g * unif_pm1(seed, (uint64_t)i);
            if (ttype) ((_Float16*)buf)[j] = (_Float16)v;
            else ((float*)buf)[j] = v;
        }
        fwrite(buf, ttype ? 2 : 4, (size_t)m, f);
    }
    free(buf);
}

static void base26(int id, char* out) {
    /* deterministic pseudo-word for a token id: 1..4 lowercase letters */
    int len = 1 + (int)(splitmix64(0x70C0ull + (uint64_t)id) % 4);
    uint64_t h = splitmix64(0xB26ull + (uint64_t)id);
    for (int i = 0; i < len; ++i) { out[i] = (char)('a' + (h % 26)); h /= 26; }
    out[len] = 0;
}

/*
 * Writes ggml-<name>.bin.  Returns 0 on success.
 * Special ids follow Appendix A.1: eot = 50256 (+1 if multilingual); beg = n_vocab - 1501.
 */
OHGEN_API int ohgen_write_model(const char* path, const ohgen_hparams* hp, uint64_t seed, const ohgen_shaping* sh) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    static char iobuf[1 << 22];
    setvbuf(f, iobuf, _IOFBF, sizeof iobuf);

    w32(f, 0x67676d6c);
    fwrite(hp, sizeof(int32_t), 11, f);

    /* mel filters */
    w32(f, hp->n_mels);
    w32(f, 201);
    float* filt = (float*)malloc(sizeof(float) * hp->n_mels * 201);
    ohgen_mel_filters(hp->n_mels, filt);
    fwrite(filt, sizeof(float), (size_t)hp->n_mels * 201, f);
    free(filt);

    /* vocabulary: 50257 entries for multilingual files, 50256 for English-only ones */
    const int multilingual = hp->n_vocab >= 51865;
    const int n_in_file = multilingual ? 50257 : 50256;
    w32(f, n_in_file);
    for (int id = 0; id < n_in_file; ++id) {
        char word[16], tok[24];
        if (id == 220) { strcpy(tok, " "); }           /* suppress_blank needs the id of " " (A.6 step 3) */
        else if (id < 256 && id >= 33 && id < 127) { tok[0] = (char)id; tok[1] = 0; }
        else {
            base26(id, word);
            if (splitmix64(0x5Cull + (uint64_t)id) % 3 != 0) { tok[0] = ' '; strcpy(tok + 1, word); }
            else strcpy(tok, word);
        }
        uint32_t len = (uint32_t)strlen(tok);
        fwrite(&len, 4, 1, f);
        fwrite(tok, 1, len, f);
    }

    const int64_t d = hp->n_audio_state, dt = hp->n_text_state, V = hp->n_vocab;
    const int wt = hp->ftype ? 1 : 0; /* weight ttype */
    uint64_t ti = 0;
    char name[128];
    const float g = (float)sh->w_gain;
    const float qk = (float)sh->qk_gain;
#define SEED() tensor_seed(seed, ti++)
#define LIN(nm, out_, in_, withbias) LING(nm, out_, in_, withbias, 1.0f)
#define LING(nm, out_, in_, withbias, gg)                                                        \
    do {                                                                                        \
        int64_t shp[2] = {(out_), (in_)};                                                       \
        snprintf(name, sizeof name, "%s.weight", nm);                                           \
        write_random_tensor(f, name, 2, shp, wt, SEED(), (gg) * g * 1.7320508f / sqrtf((float)(in_)), 0.f, NULL); \
        if (withbias) {                                                                         \
            int64_t bs[1] = {(out_)};                                                           \
            snprintf(name, sizeof name, "%s.bias", nm);                                         \
            write_random_tensor(f, name, 1, bs, 0, SEED(), 0.1f, 0.f, NULL);                    \
        } else { (void)SEED(); }                                                                \
    } while (0)
#define LNORM(nm, dim) LNORMG(nm, dim, 1.0f)
#define LNORMG(nm, dim, gm)                                                                     \
    do {                                                                                        \
        int64_t bs[1] = {(dim)};                                                                \
        snprintf(name, sizeof name, "%s.weight", nm);                                           \
        write_random_tensor(f, name, 1, bs, 0, SEED(), 0.1f * (gm), (gm), NULL);                \
        snprintf(name, sizeof name, "%s.bias", nm);                                             \
        write_random_tensor(f, name, 1, bs, 0, SEED(), 0.1f, 0.f, NULL);                        \
    } while (0)

    /* ---- encoder ---- */
    { int64_t s2[2] = {hp->n_audio_ctx, d};
      write_random_tensor(f, "encoder.positional_embedding", 2, s2, 0, SEED(), 0.5f, 0.f, NULL); }
    { int64_t s3[3] = {d, hp->n_mels, 3}; int64_t sb[2] = {d, 1};
      write_random_tensor(f, "encoder.conv1.weight", 3, s3, wt, SEED(), g * 1.7320508f / sqrtf(3.f * hp->n_mels), 0.f, NULL);
      write_random_tensor(f, "encoder.conv1.bias", 2, sb, 0, SEED(), 0.1f, 0.f, NULL); }
    { int64_t s3[3] = {d, d, 3}; int64_t sb[2] = {d, 1};
      write_random_tensor(f, "encoder.conv2.weight", 3, s3, wt, SEED(), g * 1.7320508f / sqrtf(3.f * d), 0.f, NULL);
      write_random_tensor(f, "encoder.conv2.bias", 2, sb, 0, SEED(), 0.1f, 0.f, NULL); }
    for (int l = 0; l < hp->n_audio_layer; ++l) {
        char p[96];
        snprintf(p, sizeof p, "encoder.blocks.%d.attn_ln", l); LNORM(p, d);
        snprintf(p, sizeof p, "encoder.blocks.%d.attn.query", l); LING(p, d, d, 1, qk);
        snprintf(p, sizeof p, "encoder.blocks.%d.attn.key", l); LING(p, d, d, 0, qk);
        snprintf(p, sizeof p, "encoder.blocks.%d.attn.value", l); LIN(p, d, d, 1);
        snprintf(p, sizeof p, "encoder.blocks.%d.attn.out", l); LIN(p, d, d, 1);
        snprintf(p, sizeof p, "encoder.blocks.%d.mlp_ln", l); LNORM(p, d);
        snprintf(p, sizeof p, "encoder.blocks.%d.mlp.0", l); LIN(p, 4 * d, d, 1);
        snprintf(p, sizeof p, "encoder.blocks.%d.mlp.2", l); LIN(p, d, 4 * d, 1);
    }
    LNORM("encoder.ln_post", d);

    /* ---- decoder ---- */
    { int64_t s2[2] = {hp->n_text_ctx, dt};
      write_random_tensor(f, "decoder.positional_embedding", 2, s2, 0, SEED(), 1.5f * 1.7320508f / sqrtf((float)dt), 0.f, NULL); }
    {
        /* tied token embedding: rows of norm emb_norm (U(-a,a) has std a/sqrt(3)); the logit scale comes from
         * decoder.ln.weight below: logit_v = E_v . (gamma * xhat), |xhat|^2 ~ d => std = emb_norm * gamma. */
        float a = (float)(sh->emb_norm * 1.7320508 / sqrt((double)dt));
        float* rg = (float*)malloc(sizeof(float) * V);
        const int eot = multilingual ? 50257 : 50256;
        const int beg = (int)V - 1501;
        for (int64_t v = 0; v < V; ++v) rg[v] = 1.0f;
        for (int64_t v = beg; v < V; ++v) rg[v] = (float)(sh->ts_gain * (1.0 - sh->ts_slope * (double)(v - beg) / 1500.0));
        rg[eot] = (float)sh->eot_gain;
        int64_t s2[2] = {V, dt};
        write_random_tensor(f, "decoder.token_embedding.weight", 2, s2, wt, SEED(), a, 0.f, rg);
        free(rg);
    }
    for (int l = 0; l < hp->n_text_layer; ++l) {
        char p[96];
        snprintf(p, sizeof p, "decoder.blocks.%d.attn_ln", l); LNORM(p, dt);
        snprintf(p, sizeof p, "decoder.blocks.%d.attn.query", l); LING(p, dt, dt, 1, qk);
        snprintf(p, sizeof p, "decoder.blocks.%d.attn.key", l); LING(p, dt, dt, 0, qk);
        snprintf(p, sizeof p, "decoder.blocks.%d.attn.value", l); LIN(p, dt, dt, 1);
        snprintf(p, sizeof p, "decoder.blocks.%d.attn.out", l); LIN(p, dt, dt, 1);
        snprintf(p, sizeof p, "decoder.blocks.%d.cross_attn_ln", l); LNORM(p, dt);
        snprintf(p, sizeof p, "decoder.blocks.%d.cross_attn.query", l); LING(p, dt, dt, 1, qk);
        snprintf(p, sizeof p, "decoder.blocks.%d.cross_attn.key", l); LING(p, dt, d, 0, qk);
        snprintf(p, sizeof p, "decoder.blocks.%d.cross_attn.value", l); LIN(p, dt, d, 1);
        snprintf(p, sizeof p, "decoder.blocks.%d.cross_attn.out", l); LIN(p, dt, dt, 1);
        snprintf(p, sizeof p, "decoder.blocks.%d.mlp_ln", l); LNORM(p, dt);
        snprintf(p, sizeof p, "decoder.blocks.%d.mlp.0", l); LIN(p, 4 * dt, dt, 1);
        snprintf(p, sizeof p, "decoder.blocks.%d.mlp.2", l); LIN(p, dt, 4 * dt, 1);
    }
    LNORMG("decoder.ln", dt, (float)(sh->emb_logit_std / sh->emb_norm));

    int rc = ferror(f) ? -2 : 0;
    if (fclose(f) != 0) rc = -3;
    return rc;
}
