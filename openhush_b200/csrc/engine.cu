// Buffers, stage runners (mel / encoder / decoder step) and the batched full() driver.
//
// run_full() is the device-side replacement of whisper.cpp's `whisper_full_with_state` (SURVEY.md A.5, A.7,
// A.8), reached from /root/reference/src/engine/whisper.rs:266.  Each batch item is one independent call of
// the reference; items advance window by window in lock-step rounds:
//   gather mel window at item.seek -> encoder -> cross-KV -> greedy decode (device state machine, CUDA graph)
//   -> host scoring / temperature fallback -> segments -> seek += seek_delta.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "decode_rules.h"
#include "engine.h"

namespace oh {

thread_local long long g_launch_count = 0;
thread_local int g_pdl = 0;
thread_local Profiler* g_prof = nullptr;
unsigned g_skip_mask = 0;  // OH_SKIP_CATS: timing experiments only (kernel categories that are not launched)
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const std::string& get_error() { return g_error; }

int ensure(DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return OH_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    OH_CUDA_CHECK(cudaMalloc(&b.p, bytes));
    OH_CUDA_CHECK(cudaMemset(b.p, 0, bytes));
    b.bytes = bytes;
    return OH_OK;
}
static void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

static inline size_t rup(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- layout of the DecodeSeqState block (ints unless noted), rows = max_batch -------------------------
namespace seqlay {
enum { SLOT, ACTIVE, POS, CUR, PLEN, NGEN, HAS_TS, SDELTA, RLEN, FAILED, COMPLETED, SEEK, SEEKEND, N_SCALARS };
}
struct SeqLayout {
    size_t rows, prompt_off, tok_off[6], counter_off, total_words;
    explicit SeqLayout(size_t r) : rows(r) {
        size_t o = seqlay::N_SCALARS * rows;
        prompt_off = o;
        o += rows * kMaxPrompt;
        for (int i = 0; i < 6; ++i) {
            tok_off[i] = o;
            o += rows * kMaxGen;
        }
        counter_off = o;
        total_words = o + 4;
    }
};

static DecodeSeqState seq_view(int32_t* base, const SeqLayout& L) {
    DecodeSeqState s;
    using namespace seqlay;
    s.slot = base + SLOT * L.rows;
    s.active = base + ACTIVE * L.rows;
    s.pos = base + POS * L.rows;
    s.cur_token = base + CUR * L.rows;
    s.prompt_len = base + PLEN * L.rows;
    s.n_gen = base + NGEN * L.rows;
    s.has_ts = base + HAS_TS * L.rows;
    s.seek_delta = base + SDELTA * L.rows;
    s.result_len = base + RLEN * L.rows;
    s.failed = base + FAILED * L.rows;
    s.completed = base + COMPLETED * L.rows;
    s.seek = base + SEEK * L.rows;
    s.seek_end = base + SEEKEND * L.rows;
    s.prompt = base + L.prompt_off;
    s.tok_id = base + L.tok_off[0];
    s.tok_tid = base + L.tok_off[1];
    s.tok_p = reinterpret_cast<float*>(base + L.tok_off[2]);
    s.tok_plog = reinterpret_cast<float*>(base + L.tok_off[3]);
    s.tok_pt = reinterpret_cast<float*>(base + L.tok_off[4]);
    s.tok_ptsum = reinterpret_cast<float*>(base + L.tok_off[5]);
    s.n_active_total = base + L.counter_off;
    return s;
}

// the same state block seen from row `row0` on (row groups of the interleaved decode step)
static DecodeSeqState seq_offset(DecodeSeqState s, int row0) {
    s.slot += row0; s.active += row0; s.pos += row0; s.cur_token += row0; s.prompt_len += row0; s.n_gen += row0; s.has_ts += row0;
    s.seek_delta += row0; s.result_len += row0; s.failed += row0; s.completed += row0; s.seek += row0; s.seek_end += row0;
    s.prompt += (size_t)row0 * kMaxPrompt;
    s.tok_id += (size_t)row0 * kMaxGen; s.tok_tid += (size_t)row0 * kMaxGen; s.tok_p += (size_t)row0 * kMaxGen;
    s.tok_plog += (size_t)row0 * kMaxGen; s.tok_pt += (size_t)row0 * kMaxGen; s.tok_ptsum += (size_t)row0 * kMaxGen;
    return s;
}

int state_alloc(oh_state* st) {
    const Model& m = *st->ctx->model;
    const HParams& hp = m.hp;
    const int d = hp.n_audio_state, C = hp.n_mels, T = hp.n_audio_ctx, L = hp.n_text_layer, V = hp.n_vocab;
    st->esz = st->ctx->act_type == DT_F32 ? 4 : 2;
    const size_t e = st->esz;
    st->max_batch = std::max(1, st->ctx->params.max_batch);
    int chunk = std::min(st->max_batch, 64);
    if (const char* s = getenv("OH_ENC_CHUNK")) chunk = std::max(1, std::min(st->max_batch, atoi(s)));
    st->enc_chunk = chunk;
    if (const char* s = getenv("OH_GEMM")) st->gemm_mode = strcmp(s, "simt") == 0 ? 1 : 0;
    if (const char* s = getenv("OH_ATTN")) st->attn_mode = strcmp(s, "simple") == 0 ? 1 : (strcmp(s, "flash") == 0 ? 2 : 0);  // flash: the mma.sync kernel
    if (const char* s = getenv("OH_GRAPHS")) st->use_graphs = atoi(s) != 0;
    if (const char* s = getenv("OH_PDL")) st->use_pdl = atoi(s) != 0;
    if (const char* s = getenv("OH_DUAL")) st->dual_min_rows = atoi(s) > 0 ? atoi(s) : (1 << 30);  // OH_DUAL=0: never interleave row groups
    if (const char* s = getenv("OH_GELU")) st->gelu_lut_f32 = strcmp(s, "lut") == 0;
    if (const char* s = getenv("OH_SKIP_CATS")) g_skip_mask = (unsigned)strtoul(s, nullptr, 0);
    if (const char* s = getenv("OH_SKINNY")) st->skinny_mode = atoi(s) == 0 ? 1 : 0;  // OH_SKINNY=0: decode mat-muls through the persistent tcgen05 kernel
    if (const char* s = getenv("OH_DECODE_GEMM")) st->skinny_mode = strcmp(s, "skinny") == 0 ? 0 : (strcmp(s, "tc") == 0 ? 1 : 2);

    OH_CUDA_CHECK(cudaStreamCreateWithFlags(&st->own_stream, cudaStreamNonBlocking));
    st->stream = st->own_stream;
    OH_CUDA_CHECK(cudaStreamCreateWithFlags(&st->stream2, cudaStreamNonBlocking));
    for (auto& ev : st->ev) OH_CUDA_CHECK(cudaEventCreate(&ev));
    for (auto& ev : st->ev_dual) OH_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

    const size_t W = chunk;
    int rc;
    if ((rc = ensure(st->xin, W * (kChunkFrames + 2) * C * e))) return rc;
    if ((rc = ensure(st->x1, W * (kChunkFrames + 2) * d * e))) return rc;
    if ((rc = ensure(st->x, W * T * d * 4))) return rc;
    if ((rc = ensure(st->h, W * T * d * e))) return rc;
    if ((rc = ensure(st->qkv, W * T * 3 * d * e))) return rc;
    if ((rc = ensure(st->hid, W * T * 4 * d * e))) return rc;
    const size_t S = st->max_batch;
    if ((rc = ensure(st->xkv, (size_t)L * S * T * 2 * d * e))) return rc;
    if ((rc = ensure(st->kself, (size_t)L * S * hp.n_text_ctx * d * e))) return rc;
    if ((rc = ensure(st->vself, (size_t)L * S * hp.n_text_ctx * d * e))) return rc;
    if ((rc = ensure(st->dx, S * d * 4))) return rc;
    if ((rc = ensure(st->dh, S * d * e))) return rc;
    if ((rc = ensure(st->dqkv, S * 3 * d * e))) return rc;
    if ((rc = ensure(st->datt, S * d * e))) return rc;
    if ((rc = ensure(st->dq, S * d * e))) return rc;
    if ((rc = ensure(st->dhid, S * 4 * d * e))) return rc;
    st->ld_logits = (int)rup(V, 8);
    if ((rc = ensure(st->logits, S * st->ld_logits * 4))) return rc;
    SeqLayout lay(S);
    if ((rc = ensure(st->seq_dev, lay.total_words * 4))) return rc;
    st->seq = seq_view((int32_t*)st->seq_dev.p, lay);
    if ((rc = ensure(st->rules_dev, sizeof(DecodeRules)))) return rc;
    st->pinned_bytes = lay.total_words * 4 + 4096;
    OH_CUDA_CHECK(cudaMallocHost(&st->pinned, st->pinned_bytes));
    memset(st->pinned, 0, st->pinned_bytes);
    st->rngs.assign(S, std::mt19937(0));
    st->items.clear();
    return OH_OK;
}

void state_release(oh_state* st) {
    for (auto& kv : st->step_graphs) cudaGraphExecDestroy(kv.second);
    st->step_graphs.clear();
    DevBuf* all[] = {&st->pcm_stage, &st->mel_raw, &st->clip_meta, &st->xin, &st->x1, &st->x, &st->h, &st->qkv, &st->hid, &st->enc_f32,
                     &st->xkv, &st->kself, &st->vself, &st->dx, &st->dh, &st->dqkv, &st->datt, &st->dq, &st->dhid, &st->logits,
                     &st->dump_lp, &st->dump_p, &st->seq_dev, &st->rules_dev};
    for (DevBuf* b : all) release(*b);
    if (st->pinned) cudaFreeHost(st->pinned);
    st->pinned = nullptr;
    if (st->prof) {
        if (g_prof == st->prof) g_prof = nullptr;
        for (auto& r : st->prof->recs) {
            cudaEventDestroy(r.e0);
            cudaEventDestroy(r.e1);
        }
        for (auto& e : st->prof->free_events) cudaEventDestroy(e);
        delete st->prof;
        st->prof = nullptr;
    }
    for (auto& ev : st->ev)
        if (ev) cudaEventDestroy(ev);
    if (st->own_stream) cudaStreamDestroy(st->own_stream);
    st->own_stream = nullptr;
    if (st->stream2) cudaStreamDestroy(st->stream2);
    st->stream2 = nullptr;
    for (auto& ev : st->ev_dual) {
        if (ev) cudaEventDestroy(ev);
        ev = nullptr;
    }
}

// ------------------------------------------------------------------------------------------------ mel
// clip_meta layout (device): [pcm ptrs (8 B each)] [n_samples] [n_active] [n_len] [max_key] [clip_of_slot] [seek_of_slot]
struct ClipMeta {
    const float** pcm;
    int *n_samples, *n_active, *n_len, *max_key, *clip_of_slot, *seek_of_slot;
};
static ClipMeta clip_meta_view(void* base, size_t cap) {
    ClipMeta c;
    char* p = (char*)base;
    c.pcm = (const float**)p;
    p += cap * 8;
    c.n_samples = (int*)p;
    p += cap * 4;
    c.n_active = (int*)p;
    p += cap * 4;
    c.n_len = (int*)p;
    p += cap * 4;
    c.max_key = (int*)p;
    p += cap * 4;
    c.clip_of_slot = (int*)p;
    p += cap * 4;
    c.seek_of_slot = (int*)p;
    return c;
}
static size_t clip_meta_bytes(size_t cap) { return cap * (8 + 6 * 4); }

int run_mel(oh_state* st, const float* const* pcm, const int* n, int batch, bool pcm_on_device) {
    const Model& m = *st->ctx->model;
    const int n_mel = m.hp.n_mels;
    st->items.assign(batch, Item());
    size_t stage_floats = 0;
    int max_active = 0;
    std::vector<size_t> off(batch);
    for (int i = 0; i < batch; ++i) {
        Item& it = st->items[i];
        it.n_samples = n[i];
        if (n[i] <= kNFFT / 2) {  // cannot reflect-pad: whisper.cpp reports "failed to compute log mel spectrogram"
            it.rc = OH_ERR_MEL;
            it.done = true;
            it.n_len = it.n_len_org = it.n_active = 0;
        } else {
            it.n_len = (n[i] + kSampleRate * 30) / kHop;
            it.n_len_org = 1 + (n[i] + kNFFT / 2 - kNFFT) / kHop;
            it.n_active = std::min((n[i] + kNFFT / 2) / kHop + 1, it.n_len);
        }
        off[i] = stage_floats;
        stage_floats += rup((size_t)std::max(n[i], 0), 4);
        max_active = std::max(max_active, it.n_active);
    }
    st->mel_clips = batch;
    st->mel_ld = (int)rup(std::max(max_active, 1), 32);
    int rc;
    const size_t cap = std::max<size_t>(batch, st->max_batch);
    if ((rc = ensure(st->clip_meta, clip_meta_bytes(cap)))) return rc;
    if ((rc = ensure(st->mel_raw, (size_t)batch * n_mel * st->mel_ld * 4))) return rc;
    ClipMeta cm = clip_meta_view(st->clip_meta.p, cap);

    std::vector<const float*> ptrs(batch);
    std::vector<int> ns(batch), na(batch), nl(batch);
    if (!pcm_on_device) {
        if ((rc = ensure(st->pcm_stage, std::max<size_t>(stage_floats, 4) * 4))) return rc;
        for (int i = 0; i < batch; ++i) {
            ptrs[i] = (const float*)st->pcm_stage.p + off[i];
            if (n[i] > 0)
            {
                OH_CUDA_CHECK(cudaMemcpyAsync((float*)st->pcm_stage.p + off[i], pcm[i], (size_t)n[i] * 4, cudaMemcpyHostToDevice, st->stream));
                st->timings.h2d_bytes += (long long)n[i] * 4;
            }
        }
    } else {
        for (int i = 0; i < batch; ++i) ptrs[i] = pcm[i];
    }
    for (int i = 0; i < batch; ++i) {
        ns[i] = st->items[i].n_samples;
        na[i] = st->items[i].n_active;
        nl[i] = st->items[i].n_len;
    }
    OH_CUDA_CHECK(cudaMemcpyAsync(cm.pcm, ptrs.data(), batch * 8, cudaMemcpyHostToDevice, st->stream));
    OH_CUDA_CHECK(cudaMemcpyAsync(cm.n_samples, ns.data(), batch * 4, cudaMemcpyHostToDevice, st->stream));
    OH_CUDA_CHECK(cudaMemcpyAsync(cm.n_active, na.data(), batch * 4, cudaMemcpyHostToDevice, st->stream));
    OH_CUDA_CHECK(cudaMemcpyAsync(cm.n_len, nl.data(), batch * 4, cudaMemcpyHostToDevice, st->stream));
    // the vectors above must outlive the async copies from pageable memory: those are staged synchronously by the runtime
    MelArgs a;
    a.pcm = cm.pcm;
    a.n_samples = cm.n_samples;
    a.n_active = cm.n_active;
    a.raw = (float*)st->mel_raw.p;
    a.ld = st->mel_ld;
    a.max_key = cm.max_key;
    a.n_mel = n_mel;
    a.twiddle = st->ctx->d_twiddle;
    a.hann = st->ctx->d_hann;
    a.filt_start = m.mel.start;
    a.filt_count = m.mel.count;
    a.filt_off = m.mel.offset;
    a.filt_w = m.mel.weights;
    a.n_filt_w = m.mel.n_weights;
    if (max_active > 0) {
        if ((rc = launch_mel_frames(a, batch, max_active, st->stream))) return rc;
    }
    return OH_OK;
}

// ------------------------------------------------------------------------------------------------ GEMM dispatch
static int gemm(oh_state* st, GemmArgs g, cudaStream_t stream = nullptr) {
    if (!stream) stream = st->stream;
    const Model& m = *st->ctx->model;
    const int act = st->ctx->act_type;
    g.w_type = m.compute_type == OH_COMPUTE_BF16 ? DT_BF16 : DT_F16;
    g.lut = m.gelu_lut;
    g.round_type = act;
    if (g.gelu != GELU_NONE) g.gelu = act == DT_F32 ? (st->gelu_lut_f32 ? GELU_LUT : GELU_TANH) : GELU_FAST;
    if (act != DT_F32 && st->gemm_mode == 0 && gemm_tc_available()) {
        if (g.batch * g.rows <= 256 && st->skinny_mode != 1) {
            int rc = st->skinny_mode == 2 ? launch_gemm_splitk(g, stream) : launch_gemm_skinny(g, stream);
            if (rc != OH_ERR_UNSUPPORTED) return rc;
        }
        int rc = launch_gemm_tc(g, stream);
        if (rc != OH_ERR_UNSUPPORTED) return rc;
    }
    return launch_gemm_simt(g, stream);
}

// ------------------------------------------------------------------------------------------------ encoder
int run_encoder(oh_state* st, int nw, int slot0, float* d_enc_f32_out, const int* slot_list) {
    const Model& m = *st->ctx->model;
    const HParams& hp = m.hp;
    const int d = hp.n_audio_state, C = hp.n_mels, T = hp.n_audio_ctx, act = st->ctx->act_type;
    const size_t e = st->esz;
    cudaStream_t s = st->stream;
    int rc;
    char* x1 = (char*)st->x1.p;
    float* x = (float*)st->x.p;
    {  // conv1: k3 s1 p1 as a mat-mul over 3-row views of the zero-framed channel-last window; + bias, GELU
        GemmArgs g;
        g.A = st->xin.p;
        g.a_type = act;
        g.a_bstride = (long long)(kChunkFrames + 2) * C;
        g.lda = C;
        g.batch = nw;
        g.rows = kChunkFrames;
        g.W = m.conv1.w;
        g.N = d;
        g.K = 3 * C;
        g.n_taps = 3;  // TMA path: one tensor map per tap over the overlapping 3-row view
        g.tap_k = C;
        g.bias = m.conv1.b;
        g.C = x1 + (size_t)d * e;  // row 1 of each window (row 0 and row 3001 stay zero)
        g.c_type = act;
        g.c_bstride = (long long)(kChunkFrames + 2) * d;
        g.ldc = d;
        g.gelu = 1;
        if ((rc = gemm(st, g))) return rc;
    }
    {  // conv2: k3 s2 p1: row i of the output reads rows 2i-1..2i+1 = 3d contiguous elements starting at padded row 2i
        GemmArgs g;
        g.A = x1;
        g.a_type = act;
        g.a_bstride = (long long)(kChunkFrames + 2) * d;
        g.lda = 2 * d;
        g.batch = nw;
        g.rows = T;
        g.W = m.conv2.w;
        g.N = d;
        g.K = 3 * d;
        g.n_taps = 3;
        g.tap_k = d;
        g.bias = m.conv2.b;
        g.C = x;
        g.c_type = DT_F32;
        g.c_bstride = (long long)T * d;
        g.ldc = d;
        g.gelu = 1;
        g.R = m.enc_pos;  // + positional embedding, broadcast over the batch
        g.r_bstride = 0;
        g.ldr = d;
        if ((rc = gemm(st, g))) return rc;
    }
    const int M = nw * T;
    for (const EncLayer& L : m.enc) {
        if ((rc = launch_layernorm(x, M, d, L.ln1.w, L.ln1.b, st->h.p, act, s))) return rc;
        GemmArgs g;
        g.A = st->h.p; g.a_type = act; g.lda = d; g.rows = M;
        g.W = L.qkv.w; g.N = 3 * d; g.K = d; g.bias = L.qkv.b;
        g.C = st->qkv.p; g.c_type = act; g.ldc = 3 * d;
        if ((rc = gemm(st, g))) return rc;
        if (act != DT_F32 && st->attn_mode != 1) {
            rc = st->attn_mode == 0 ? launch_encoder_attention_tc(st->qkv.p, st->h.p, act, nw, T, d, s) : OH_ERR_UNSUPPORTED;
            if (rc == OH_ERR_UNSUPPORTED) rc = launch_encoder_attention_mma(st->qkv.p, st->h.p, act, nw, T, d, s);
        } else {
            rc = launch_encoder_attention_simple(st->qkv.p, st->h.p, act, nw, T, d, s);
        }
        if (rc) return rc;
        GemmArgs o;
        o.A = st->h.p; o.a_type = act; o.lda = d; o.rows = M;
        o.W = L.out.w; o.N = d; o.K = d; o.bias = L.out.b;
        o.C = x; o.c_type = DT_F32; o.ldc = d; o.R = x; o.ldr = d;
        if ((rc = gemm(st, o))) return rc;
        if ((rc = launch_layernorm(x, M, d, L.ln2.w, L.ln2.b, st->h.p, act, s))) return rc;
        GemmArgs f0;
        f0.A = st->h.p; f0.a_type = act; f0.lda = d; f0.rows = M;
        f0.W = L.mlp0.w; f0.N = 4 * d; f0.K = d; f0.bias = L.mlp0.b;
        f0.C = st->hid.p; f0.c_type = act; f0.ldc = 4 * d; f0.gelu = 1;
        if ((rc = gemm(st, f0))) return rc;
        GemmArgs f2;
        f2.A = st->hid.p; f2.a_type = act; f2.lda = 4 * d; f2.rows = M;
        f2.W = L.mlp2.w; f2.N = d; f2.K = 4 * d; f2.bias = L.mlp2.b;
        f2.C = x; f2.c_type = DT_F32; f2.ldc = d; f2.R = x; f2.ldr = d;
        if ((rc = gemm(st, f2))) return rc;
    }
    if ((rc = launch_layernorm(x, M, d, m.ln_post.w, m.ln_post.b, st->h.p, act, s))) return rc;
    if (d_enc_f32_out) {
        if ((rc = launch_layernorm(x, M, d, m.ln_post.w, m.ln_post.b, d_enc_f32_out, DT_F32, s))) return rc;
    }
    // cross-attention K/V of every decoder layer, straight into the slots' caches: [L][slots][T][2d]
    // slot_list (optional): window w goes to slot slot_list[w]; runs of consecutive slots share one mat-mul
    for (int w0 = 0; w0 < nw;) {
        int w1 = w0 + 1;
        if (slot_list)
            while (w1 < nw && slot_list[w1] == slot_list[w1 - 1] + 1) ++w1;
        else
            w1 = nw;
        const int first_slot = slot_list ? slot_list[w0] : slot0;
        for (int l = 0; l < hp.n_text_layer; ++l) {
            const DecLayer& D = m.dec[l];
            GemmArgs g;  // one "batch" per window so that the slot index separates from the row index
            g.A = (char*)st->h.p + (size_t)w0 * T * d * e; g.a_type = act; g.lda = d; g.batch = w1 - w0; g.rows = T; g.a_bstride = (long long)T * d;
            g.W = D.xkv.w; g.N = 2 * d; g.K = d; g.bias = D.xkv.b;
            g.C = (char*)st->xkv.p + ((size_t)l * st->max_batch + first_slot) * T * 2 * d * e;
            g.c_type = act;
            g.c_bstride = (long long)T * 2 * d;       // one slot: [2 (k,v)][heads][T][64]
            g.c_hsplit = kHeadDim;                    // column n = (kv * heads + head) * 64 + c
            g.c_hstride = (long long)T * kHeadDim;
            g.ldc = kHeadDim;
            if ((rc = gemm(st, g))) return rc;
        }
        w0 = w1;
    }
    return OH_OK;
}

// ------------------------------------------------------------------------------------------------ decoder step
// One decoder step for rows [row0, row0 + rows) on stream `s`.  `after_q0` (optional) is recorded right after the first
// layer's cross-attention query projection: the interleaved step starts the second row group there.
static int decode_step_launch(oh_state* st, int row0, int rows, bool with_k9, bool dump, cudaStream_t s, cudaEvent_t after_q0 = nullptr) {
    const Model& m = *st->ctx->model;
    const HParams& hp = m.hp;
    const int d = hp.n_text_state, T = hp.n_audio_ctx, act = st->ctx->act_type, NC = hp.n_text_ctx;
    const size_t e = st->esz;
    const size_t r0 = (size_t)row0;
    float* x = (float*)st->dx.p + r0 * d;
    char* dh = (char*)st->dh.p + r0 * d * e;
    char* dqkv = (char*)st->dqkv.p + r0 * 3 * d * e;
    char* datt = (char*)st->datt.p + r0 * d * e;
    char* dq = (char*)st->dq.p + r0 * d * e;
    char* dhid = (char*)st->dhid.p + r0 * 4 * d * e;
    float* logits = (float*)st->logits.p + r0 * st->ld_logits;
    const DecodeSeqState seq = seq_offset(st->seq, row0);
    int rc;
    struct PdlScope {  // kernels launched below may start while their predecessor drains (common.cuh)
        const bool on;
        explicit PdlScope(bool enable) : on(enable) { g_pdl += on ? 1 : 0; }
        ~PdlScope() { g_pdl -= on ? 1 : 0; }
    } pdl_scope(st->use_pdl && g_prof == nullptr);
    if ((rc = launch_embed(m.tok_emb.w, m.compute_type == OH_COMPUTE_BF16 ? DT_BF16 : DT_F16, m.dec_pos, seq, rows, d, x, s))) return rc;
    for (int l = 0; l < hp.n_text_layer; ++l) {
        const DecLayer& D = m.dec[l];
        if ((rc = launch_layernorm(x, rows, d, D.ln1.w, D.ln1.b, dh, act, s))) return rc;
        GemmArgs g;
        g.A = dh; g.a_type = act; g.lda = d; g.rows = rows;
        g.W = D.qkv.w; g.N = 3 * d; g.K = d; g.bias = D.qkv.b;
        g.C = dqkv; g.c_type = act; g.ldc = 3 * d;
        if ((rc = gemm(st, g, s))) return rc;
        char* kc = (char*)st->kself.p + (size_t)l * st->max_batch * NC * d * e;
        char* vc = (char*)st->vself.p + (size_t)l * st->max_batch * NC * d * e;
        if ((rc = launch_self_attention_decode(dqkv, kc, vc, act, seq, rows, d, NC, datt, s))) return rc;
        GemmArgs o;
        o.A = datt; o.a_type = act; o.lda = d; o.rows = rows;
        o.W = D.out.w; o.N = d; o.K = d; o.bias = D.out.b;
        o.C = x; o.c_type = DT_F32; o.ldc = d; o.R = x; o.ldr = d;
        if ((rc = gemm(st, o, s))) return rc;

        if ((rc = launch_layernorm(x, rows, d, D.lnx.w, D.lnx.b, dh, act, s))) return rc;
        GemmArgs q;
        q.A = dh; q.a_type = act; q.lda = d; q.rows = rows;
        q.W = D.xq.w; q.N = d; q.K = d; q.bias = D.xq.b;
        q.C = dq; q.c_type = act; q.ldc = d;
        if ((rc = gemm(st, q, s))) return rc;
        if (l == 0 && after_q0) OH_CUDA_CHECK(cudaEventRecord(after_q0, s));
        const char* xkv = (const char*)st->xkv.p + (size_t)l * st->max_batch * T * 2 * d * e;
        if ((rc = launch_cross_attention_decode(dq, xkv, act, seq, rows, d, T, datt, s))) return rc;
        GemmArgs xo;
        xo.A = datt; xo.a_type = act; xo.lda = d; xo.rows = rows;
        xo.W = D.xout.w; xo.N = d; xo.K = d; xo.bias = D.xout.b;
        xo.C = x; xo.c_type = DT_F32; xo.ldc = d; xo.R = x; xo.ldr = d;
        if ((rc = gemm(st, xo, s))) return rc;

        if ((rc = launch_layernorm(x, rows, d, D.ln2.w, D.ln2.b, dh, act, s))) return rc;
        GemmArgs f0;
        f0.A = dh; f0.a_type = act; f0.lda = d; f0.rows = rows;
        f0.W = D.mlp0.w; f0.N = 4 * d; f0.K = d; f0.bias = D.mlp0.b;
        f0.C = dhid; f0.c_type = act; f0.ldc = 4 * d; f0.gelu = 1;
        if ((rc = gemm(st, f0, s))) return rc;
        GemmArgs f2;
        f2.A = dhid; f2.a_type = act; f2.lda = 4 * d; f2.rows = rows;
        f2.W = D.mlp2.w; f2.N = d; f2.K = 4 * d; f2.bias = D.mlp2.b;
        f2.C = x; f2.c_type = DT_F32; f2.ldc = d; f2.R = x; f2.ldr = d;
        if ((rc = gemm(st, f2, s))) return rc;
    }
    if ((rc = launch_layernorm(x, rows, d, m.ln.w, m.ln.b, dh, act, s))) return rc;
    GemmArgs lg;
    lg.A = dh; lg.a_type = act; lg.lda = d; lg.rows = rows;
    lg.W = m.tok_emb.w; lg.N = hp.n_vocab; lg.K = d;
    lg.C = logits; lg.c_type = DT_F32; lg.ldc = st->ld_logits;
    if ((rc = gemm(st, lg, s))) return rc;
    if (with_k9) {
        float* lp = nullptr;
        float* pr = nullptr;
        if (dump) {
            if ((rc = ensure(st->dump_lp, (size_t)st->max_batch * st->ld_logits * 4))) return rc;
            if ((rc = ensure(st->dump_p, (size_t)st->max_batch * st->ld_logits * 4))) return rc;
            lp = (float*)st->dump_lp.p + r0 * st->ld_logits;
            pr = (float*)st->dump_p.p + r0 * st->ld_logits;
        }
        if ((rc = launch_logits_step(logits, st->ld_logits, seq, (const DecodeRules*)st->rules_dev.p, rows, lp, pr, s))) return rc;
    }
    return OH_OK;
}

// The whole step.  With enough rows the batch is cut into two row groups that run on two streams, the second one
// starting when the first reaches its first cross-attention: from then on one group streams cross-attention K/V (the
// HBM-bound part of a step) while the other walks through its latency-bound chain of projections and LayerNorms, so
// the two kinds of work overlap instead of alternating.  Rows are independent, so the results do not change.
// Measured on B200 (large-v3, 128 rows): 9.3 ms against 8.8 ms for the single chain — the projections of one group
// queue behind the other group's resident cross-attention CTAs — so the interleave is opt-in (OH_DUAL=<min rows>).
static int decode_step_all(oh_state* st, int rows, bool with_k9, bool dump) {
    cudaStream_t s = st->stream;
    if (rows < st->dual_min_rows || dump || g_prof != nullptr) return decode_step_launch(st, 0, rows, with_k9, dump, s);
    const int ra = (rows + 1) / 2;
    int rc;
    OH_CUDA_CHECK(cudaEventRecord(st->ev_dual[0], s));
    OH_CUDA_CHECK(cudaStreamWaitEvent(st->stream2, st->ev_dual[0], 0));
    if ((rc = decode_step_launch(st, 0, ra, with_k9, false, s, st->ev_dual[1]))) return rc;
    OH_CUDA_CHECK(cudaStreamWaitEvent(st->stream2, st->ev_dual[1], 0));
    if ((rc = decode_step_launch(st, ra, rows - ra, with_k9, false, st->stream2))) return rc;
    OH_CUDA_CHECK(cudaEventRecord(st->ev_dual[2], st->stream2));
    OH_CUDA_CHECK(cudaStreamWaitEvent(s, st->ev_dual[2], 0));
    return OH_OK;
}

int run_decode_step(oh_state* st, int rows, bool dump_probs) { return decode_step_all(st, rows, true, dump_probs); }

static int run_decode_step_graph(oh_state* st, int rows, bool with_k9 = true) {
    if (!st->use_graphs || st->profiling) return decode_step_all(st, rows, with_k9, false);
    const int key = rows * 2 + (with_k9 ? 1 : 0);
    auto it = st->step_graphs.find(key);
    if (it == st->step_graphs.end()) {
        // first use of this row count: one direct step (sets kernel attributes, faults in lazily allocated
        // buffers), capture on the second, replay from then on
        if (!st->warmed_rows.count(key)) {
            st->warmed_rows.insert(key);
            return decode_step_all(st, rows, with_k9, false);
        }
        cudaGraph_t graph = nullptr;
        const long long launches_before = g_launch_count;
        OH_CUDA_CHECK(cudaStreamBeginCapture(st->stream, cudaStreamCaptureModeThreadLocal));
        int rc = decode_step_all(st, rows, with_k9, false);
        cudaError_t ce = cudaStreamEndCapture(st->stream, &graph);
        g_launch_count = launches_before;  // nothing ran during capture
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        OH_CUDA_CHECK(ce);
        cudaGraphExec_t exec = nullptr;
        OH_CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        it = st->step_graphs.emplace(key, exec).first;
    }
    OH_CUDA_CHECK(cudaGraphLaunch(it->second, st->stream));
    g_launch_count += 11LL * st->ctx->model->hp.n_text_layer + (with_k9 ? 4 : 3);  // kernels inside one replayed step
    return OH_OK;
}

const int* clip_max_key(oh_state* st, int batch, int clip) {
    ClipMeta cm = clip_meta_view(st->clip_meta.p, std::max<size_t>(batch, st->max_batch));
    return cm.max_key + clip;
}

// stage API: feed tokens[batch][n_tokens] at positions n_past.. and return the logits of the last position
int decode_tokens_stage(oh_state* st, const int* tokens, int n_tokens, int n_past, int batch, float* out_logits) {
    SeqLayout lay(st->max_batch);
    int32_t* hb = (int32_t*)st->pinned;
    DecodeSeqState hs = seq_view(hb, lay);
    memset(hb, 0, seqlay::N_SCALARS * lay.rows * 4);
    cudaStream_t s = st->stream;
    int rc;
    for (int i = 0; i < n_tokens; ++i) {
        for (int r = 0; r < batch; ++r) {
            hs.slot[r] = r;
            hs.active[r] = 1;
            hs.pos[r] = n_past + i;
            hs.cur_token[r] = tokens[(size_t)r * n_tokens + i];
        }
        OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, seqlay::N_SCALARS * lay.rows * 4, cudaMemcpyHostToDevice, s));
        if ((rc = run_decode_step_graph(st, batch, false))) return rc;
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    if (out_logits) {
        const int V = st->ctx->model->hp.n_vocab;
        OH_CUDA_CHECK(cudaMemcpy2DAsync(out_logits, (size_t)V * 4, st->logits.p, (size_t)st->ld_logits * 4, (size_t)V * 4, batch,
                                        cudaMemcpyDeviceToHost, s));
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    return OH_OK;
}

// ------------------------------------------------------------------------------------------------ full()
namespace {

struct PassRow {
    int item, slot;
    std::vector<int> prompt;
};
struct PassResult {
    std::vector<TokenData> tokens;
    int result_len = 0, seek_delta = 0;
    bool has_ts = false, failed = false, completed = false;
    double sum_logprobs = 0, avg_logprobs = 0, entropy = 0, score = 0;
};

void sequence_score(const oh_full_params& p, PassResult& r) {  // A.8
    const int n = r.result_len;
    double sum = 0.0;
    for (int i = 0; i < n && i < (int)r.tokens.size(); ++i) sum += r.tokens[i].plog;
    r.sum_logprobs = sum;
    r.avg_logprobs = n > 0 ? sum / n : -INFINITY;
    double penalty = n;
    if (p.length_penalty > 0.0f) penalty = std::pow((5.0 + n) / 6.0, (double)p.length_penalty);
    r.score = n > 0 ? sum / penalty : -INFINITY;
    std::map<int, int> counts;
    int cnt = 0;
    for (int i = std::max(0, n - 32); i < n; ++i) {
        counts[r.tokens[i].id]++;
        cnt++;
    }
    double ent = 0.0;
    for (auto& kv : counts) {
        const double q = (double)kv.second / cnt;
        ent -= q * std::log(q);
    }
    r.entropy = ent;
}

void emit_segments(const Model& m, const oh_full_params& p, Item& it, const PassResult& best, const std::vector<int>& prompt, int n_prompt_init) {
    const Vocab& vc = m.vocab;
    const auto& tk = best.tokens;
    std::vector<int> new_past;
    if (!prompt.empty() && prompt.front() == vc.token_prev) new_past.insert(new_past.end(), prompt.begin() + 1, prompt.end() - n_prompt_init);
    for (int i = 0; i < best.result_len && i < (int)tk.size(); ++i) new_past.push_back(tk[i].id);
    it.prompt_past.swap(new_past);
    const int seek = it.seek;
    if (!tk.empty()) {
        int i0 = 0;
        int64_t t0 = seek + 2 * (tk.front().tid - vc.token_beg);
        std::string text;
        for (int i = 0; i < (int)tk.size(); ++i) {
            if (p.print_special || tk[i].id < vc.token_eot) text += vc.id_to_token[tk[i].id];
            if (tk[i].id > vc.token_beg && !p.single_segment) {
                const int64_t t1 = seek + 2 * (tk[i].tid - vc.token_beg);
                if (!text.empty()) {
                    Segment s;
                    s.t0 = t0;
                    s.t1 = t1;
                    s.text = text;
                    for (int j = i0; j <= i; ++j) s.tokens.push_back(tk[j]);
                    it.segments.push_back(std::move(s));
                }
                text.clear();
                while (i < (int)tk.size() && tk[i].id > vc.token_beg) i++;
                i--;
                t0 = t1;
                i0 = i + 1;
            }
        }
        if (!text.empty()) {
            Segment s;
            s.t0 = t0;
            s.t1 = seek + best.seek_delta;
            s.text = text;
            for (int j = i0; j < (int)tk.size(); ++j) s.tokens.push_back(tk[j]);
            it.segments.push_back(std::move(s));
        }
    }
    it.seek += best.seek_delta;
}

// host side of one sampled token (A.7, t > 0): std::discrete_distribution over the row's probabilities
TokenData sample_on_host(const Vocab& vc, const float* probs, const float* logprobs, std::mt19937& rng) {
    TokenData r;
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = vc.token_beg; i < vc.n_vocab; ++i) {
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                r.tid = i;
            }
        }
        r.pt = (float)(max_ts / (sum_ts + 1e-10));
        r.ptsum = (float)sum_ts;
    }
    std::discrete_distribution<> dist(probs, probs + vc.n_vocab);
    r.id = dist(rng);
    r.p = probs[r.id];
    r.plog = logprobs[r.id];
    if (r.id >= vc.token_beg) {
        r.tid = r.id;
        r.pt = r.p;
    }
    return r;
}

}  // namespace

static int decode_pass(oh_state* st, const oh_full_params& p, bool no_timestamps, const std::vector<PassRow>& rows_in, float temperature,
                       std::vector<PassResult>& out) {
    const Model& m = *st->ctx->model;
    const Vocab& vc = m.vocab;
    const int rows = (int)rows_in.size();
    const int n_max = m.hp.n_text_ctx / 2 - 4;
    SeqLayout lay(st->max_batch);
    int32_t* hb = (int32_t*)st->pinned;
    DecodeSeqState hs = seq_view(hb, lay);
    memset(hb, 0, lay.total_words * 4);
    int max_prompt = 0;
    for (int r = 0; r < rows; ++r) {
        const PassRow& pr = rows_in[r];
        const Item& it = st->items[pr.item];
        const int pl = (int)pr.prompt.size();
        if (pl > kMaxPrompt || pl < 1) {
            set_error("prompt too long");
            return OH_ERR_DECODE;
        }
        max_prompt = std::max(max_prompt, pl);
        hs.slot[r] = pr.slot;
        hs.active[r] = 1;
        hs.pos[r] = 0;
        hs.cur_token[r] = pr.prompt[0];
        hs.prompt_len[r] = pl;
        for (int i = 0; i < pl; ++i) hs.prompt[(size_t)r * kMaxPrompt + i] = pr.prompt[i];
        hs.seek_delta[r] = 100 * 30;
        hs.seek[r] = it.seek;
        hs.seek_end[r] = it.seek_end;
    }
    *hs.n_active_total = rows;
    DecodeRules R;
    R.n_vocab = vc.n_vocab; R.eot = vc.token_eot; R.sot = vc.token_sot; R.translate = vc.token_translate; R.transcribe = vc.token_transcribe;
    R.solm = vc.token_solm; R.prev = vc.token_prev; R.nosp = vc.token_nosp; R.tok_not = vc.token_not; R.beg = vc.token_beg;
    R.blank = vc.token_blank; R.n_lang = vc.num_languages();
    R.no_timestamps = no_timestamps; R.suppress_blank = p.suppress_blank; R.single_segment = p.single_segment; R.max_tokens = p.max_tokens;
    R.n_max = n_max;
    R.tid0_initial = p.max_initial_ts > 0.0f ? (int)std::round(p.max_initial_ts / (30.0f / (float)m.hp.n_audio_ctx)) : -1;
    R.temperature = temperature;
    cudaStream_t s = st->stream;
    OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
    OH_CUDA_CHECK(cudaMemcpyAsync(st->rules_dev.p, &R, sizeof R, cudaMemcpyHostToDevice, s));
    st->timings.h2d_bytes += (long long)lay.total_words * 4 + (long long)sizeof R;
    OH_CUDA_CHECK(cudaStreamSynchronize(s));  // R and hb are reused below

    const int max_steps = max_prompt + n_max - 1;
    int rc;
    if (temperature < 1e-6f) {
        int* counter = (int*)((char*)st->pinned + lay.total_words * 4 + 64);
        for (int step = 0; step < max_steps; ++step) {
            if ((rc = run_decode_step_graph(st, rows))) return rc;
            st->timings.n_decode_steps++;
            if ((step & 3) == 3 || step == max_steps - 1) {
                OH_CUDA_CHECK(cudaMemcpyAsync(counter, st->seq.n_active_total, 4, cudaMemcpyDeviceToHost, s));
                st->timings.d2h_bytes += 4;
                OH_CUDA_CHECK(cudaStreamSynchronize(s));
                if (*counter <= 0) break;
            }
        }
        OH_CUDA_CHECK(cudaMemcpyAsync(hb, st->seq_dev.p, lay.total_words * 4, cudaMemcpyDeviceToHost, s));
        st->timings.d2h_bytes += (long long)lay.total_words * 4;
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
    } else {
        // t > 0: probabilities come back to the host and are sampled with the same libstdc++ calls as upstream
        std::vector<float> probs((size_t)rows * st->ld_logits), logprobs((size_t)rows * st->ld_logits);
        int n_live = rows;
        for (int step = 0; step < max_steps && n_live > 0; ++step) {
            if ((rc = decode_step_all(st, rows, true, true))) return rc;
            st->timings.n_decode_steps++;
            OH_CUDA_CHECK(cudaMemcpyAsync(probs.data(), st->dump_p.p, probs.size() * 4, cudaMemcpyDeviceToHost, s));
            OH_CUDA_CHECK(cudaMemcpyAsync(logprobs.data(), st->dump_lp.p, logprobs.size() * 4, cudaMemcpyDeviceToHost, s));
            st->timings.d2h_bytes += (long long)probs.size() * 8;
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
            for (int r = 0; r < rows; ++r) {
                if (!hs.active[r]) continue;
                const int pos = hs.pos[r];
                if (pos + 1 < hs.prompt_len[r]) {  // mirrors the kernel's prompt feeding
                    hs.pos[r] = pos + 1;
                    hs.cur_token[r] = hs.prompt[(size_t)r * kMaxPrompt + pos + 1];
                    continue;
                }
                TokenData t = sample_on_host(vc, probs.data() + (size_t)r * st->ld_logits, logprobs.data() + (size_t)r * st->ld_logits,
                                             st->rngs[hs.slot[r]]);
                const int i = hs.n_gen[r];
                const size_t o = (size_t)r * kMaxGen + i;
                hs.tok_id[o] = t.id; hs.tok_tid[o] = t.tid; hs.tok_p[o] = t.p; hs.tok_plog[o] = t.plog; hs.tok_pt[o] = t.pt; hs.tok_ptsum[o] = t.ptsum;
                hs.n_gen[r] = i + 1;
                SeqFlags f{hs.has_ts[r], hs.seek_delta[r], hs.result_len[r], 0, 0};
                const bool live = apply_token_rules(f, t.id, i, hs.seek[r], hs.seek_end[r], vc.token_beg, vc.token_eot, p.max_tokens, no_timestamps,
                                                    p.single_segment, n_max);
                hs.has_ts[r] = f.has_ts; hs.seek_delta[r] = f.seek_delta; hs.result_len[r] = f.result_len;
                hs.failed[r] = f.failed; hs.completed[r] = f.completed;
                if (live) {
                    hs.cur_token[r] = t.id;
                    hs.pos[r] = pos + 1;
                } else {
                    hs.active[r] = 0;
                    n_live--;
                }
            }
            // scalars + token ids (the kernel reads the last two sampled ids)
            OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, (lay.tok_off[1]) * 4, cudaMemcpyHostToDevice, s));
            st->timings.h2d_bytes += (long long)lay.tok_off[1] * 4;
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
        }
    }
    out.assign(rows, PassResult());
    for (int r = 0; r < rows; ++r) {
        PassResult& pr = out[r];
        const int ng = hs.n_gen[r];
        pr.tokens.resize(ng);
        for (int i = 0; i < ng; ++i) {
            const size_t o = (size_t)r * kMaxGen + i;
            pr.tokens[i] = TokenData{hs.tok_id[o], hs.tok_tid[o], hs.tok_p[o], hs.tok_plog[o], hs.tok_pt[o], hs.tok_ptsum[o]};
        }
        pr.result_len = hs.result_len[r];
        pr.seek_delta = hs.seek_delta[r];
        pr.has_ts = hs.has_ts[r];
        pr.failed = hs.failed[r];
        pr.completed = hs.completed[r];
        st->timings.n_tokens += ng;
    }
    return OH_OK;
}

// ------------------------------------------------------------------------------------------------ beam search
// BASELINE config 4 / A.7 "Beam".  The reference never selects it (/root/reference/src/engine/whisper.rs:243 hard-codes
// Greedy) and the pinned upstream's candidate sampler cannot be verified offline, so this follows SURVEY.md A.5/A.7
// as written — per live decoder the top beam_size ids of the filtered distribution, candidates ranked by
// sum_logprobs_all, assigned to the decoders in order skipping repeats, KV cache inherited from the proposing decoder
// — decision for decision the same procedure the test oracle restates.  It is a host-driven lock-step loop over
// beam_size decoder rows (the K9 kernel runs in "dump" mode: rules + log-softmax on the device, the distributions come
// back to the host); items and windows are processed one after the other.  Correctness path, not a throughput path.
namespace {

struct BeamDec {
    std::vector<TokenData> tokens;
    int seek_delta = 100 * 30, result_len = 0, slot = 0;
    bool has_ts = false, failed = false, completed = false;
    double sum_logprobs_all = 0.0;
    std::vector<float> probs, logprobs;
};
struct BeamCand {
    int decoder_idx, seek_delta, result_len;
    bool has_ts;
    std::vector<TokenData> tokens;
    double sum_logprobs_all;
};

// the k most probable ids, best first (ties: lower id); tid / pt / ptsum as for every sampled token (A.7)
std::vector<TokenData> topk_on_host(const Vocab& vc, const float* probs, const float* logprobs, int k) {
    TokenData base;
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = vc.token_beg; i < vc.n_vocab; ++i) {
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                base.tid = i;
            }
        }
        base.pt = (float)(max_ts / (sum_ts + 1e-10));
        base.ptsum = (float)sum_ts;
    }
    std::vector<int> ids(vc.n_vocab);
    for (int i = 0; i < vc.n_vocab; ++i) ids[i] = i;
    k = std::min(k, vc.n_vocab);
    std::partial_sort(ids.begin(), ids.begin() + k, ids.end(), [&](int a, int b) {
        if (logprobs[a] != logprobs[b]) return logprobs[a] > logprobs[b];
        return a < b;
    });
    std::vector<TokenData> out;
    for (int c = 0; c < k; ++c) {
        TokenData r = base;
        r.id = ids[c];
        r.p = probs[r.id];
        r.plog = logprobs[r.id];
        if (r.id >= vc.token_beg) {
            r.tid = r.id;
            r.pt = r.p;
        }
        out.push_back(r);
    }
    return out;
}

int copy_self_kv(oh_state* st, int src_slot, int dst_slot) {
    const HParams& hp = st->ctx->model->hp;
    const size_t per_slot = (size_t)hp.n_text_ctx * hp.n_text_state * st->esz;
    for (int l = 0; l < hp.n_text_layer; ++l) {
        const size_t base = (size_t)l * st->max_batch * per_slot;
        OH_CUDA_CHECK(cudaMemcpyAsync((char*)st->kself.p + base + dst_slot * per_slot, (char*)st->kself.p + base + src_slot * per_slot, per_slot,
                                      cudaMemcpyDeviceToDevice, st->stream));
        OH_CUDA_CHECK(cudaMemcpyAsync((char*)st->vself.p + base + dst_slot * per_slot, (char*)st->vself.p + base + src_slot * per_slot, per_slot,
                                      cudaMemcpyDeviceToDevice, st->stream));
    }
    return OH_OK;
}
int copy_cross_kv(oh_state* st, int src_slot, int dst_slot) {
    const HParams& hp = st->ctx->model->hp;
    const size_t per_slot = (size_t)hp.n_audio_ctx * 2 * hp.n_text_state * st->esz;
    for (int l = 0; l < hp.n_text_layer; ++l) {
        const size_t base = (size_t)l * st->max_batch * per_slot;
        OH_CUDA_CHECK(cudaMemcpyAsync((char*)st->xkv.p + base + dst_slot * per_slot, (char*)st->xkv.p + base + src_slot * per_slot, per_slot,
                                      cudaMemcpyDeviceToDevice, st->stream));
    }
    return OH_OK;
}

// one window, temperature 0, nb decoders on rows / slots 0..nb-1 (cross K/V of the window already in every slot)
int beam_pass(oh_state* st, const oh_full_params& p, bool no_timestamps, const Item& it, const std::vector<int>& prompt, int nb, PassResult& best) {
    const Model& m = *st->ctx->model;
    const Vocab& vc = m.vocab;
    const int n_max = m.hp.n_text_ctx / 2 - 4;
    const int pl = (int)prompt.size();
    if (pl > kMaxPrompt || pl < 1) {
        set_error("prompt too long");
        return OH_ERR_DECODE;
    }
    SeqLayout lay(st->max_batch);
    int32_t* hb = (int32_t*)st->pinned;
    DecodeSeqState hs = seq_view(hb, lay);
    memset(hb, 0, lay.total_words * 4);
    for (int r = 0; r < nb; ++r) {
        hs.slot[r] = r;
        hs.prompt_len[r] = pl;
        for (int i = 0; i < pl; ++i) hs.prompt[(size_t)r * kMaxPrompt + i] = prompt[i];
        hs.seek_delta[r] = 100 * 30;
        hs.seek[r] = it.seek;
        hs.seek_end[r] = it.seek_end;
    }
    hs.active[0] = 1;
    hs.cur_token[0] = prompt[0];
    DecodeRules R;
    R.n_vocab = vc.n_vocab; R.eot = vc.token_eot; R.sot = vc.token_sot; R.translate = vc.token_translate; R.transcribe = vc.token_transcribe;
    R.solm = vc.token_solm; R.prev = vc.token_prev; R.nosp = vc.token_nosp; R.tok_not = vc.token_not; R.beg = vc.token_beg;
    R.blank = vc.token_blank; R.n_lang = vc.num_languages();
    R.no_timestamps = no_timestamps; R.suppress_blank = p.suppress_blank; R.single_segment = p.single_segment; R.max_tokens = p.max_tokens;
    R.n_max = n_max;
    R.tid0_initial = p.max_initial_ts > 0.0f ? (int)std::round(p.max_initial_ts / (30.0f / (float)m.hp.n_audio_ctx)) : -1;
    R.temperature = 0.0f;
    cudaStream_t s = st->stream;
    OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
    OH_CUDA_CHECK(cudaMemcpyAsync(st->rules_dev.p, &R, sizeof R, cudaMemcpyHostToDevice, s));
    OH_CUDA_CHECK(cudaStreamSynchronize(s));
    const size_t ld = st->ld_logits;
    int rc;
    auto upload_rows = [&]() -> int {  // scalars, prompts and sampled-token ids (the kernel reads the last two ids of a row)
        OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.tok_off[1] * 4, cudaMemcpyHostToDevice, s));
        st->timings.h2d_bytes += (long long)lay.tok_off[1] * 4;
        return OH_OK;
    };
    // the prompt through decoder 0 (A.5: "decode(prompt) -> process_logits(decoder0, is_initial)")
    // (the kernel advances pos / cur_token through the prompt on the device by itself; the pinned host mirror must not be
    // touched while an asynchronous upload of it could still be in flight, so nothing is uploaded in this phase)
    for (int i = 0; i < pl; ++i) {
        if ((rc = decode_step_all(st, nb, true, true))) return rc;
        st->timings.n_decode_steps++;
    }
    hs.pos[0] = pl - 1;
    hs.cur_token[0] = prompt[pl - 1];
    std::vector<BeamDec> dec(nb);
    dec[0].probs.resize(ld);
    dec[0].logprobs.resize(ld);
    OH_CUDA_CHECK(cudaMemcpyAsync(dec[0].probs.data(), st->dump_p.p, ld * 4, cudaMemcpyDeviceToHost, s));
    OH_CUDA_CHECK(cudaMemcpyAsync(dec[0].logprobs.data(), st->dump_lp.p, ld * 4, cudaMemcpyDeviceToHost, s));
    st->timings.d2h_bytes += (long long)ld * 8;
    for (int j = 1; j < nb; ++j)  // "copy logits/probs/kv to the other decoders"
        if ((rc = copy_self_kv(st, 0, j))) return rc;
    OH_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int j = 0; j < nb; ++j) {
        dec[j].slot = j;
        if (j > 0) {
            dec[j].probs = dec[0].probs;
            dec[j].logprobs = dec[0].logprobs;
        }
    }

    for (int i = 0; i < n_max; ++i) {
        // ---- candidates of every live decoder, ranked
        std::vector<BeamCand> cands;
        for (int j = 0; j < nb; ++j) {
            BeamDec& d = dec[j];
            if (d.completed || d.failed) continue;
            for (const TokenData& tok : topk_on_host(vc, d.probs.data(), d.logprobs.data(), nb)) {
                BeamCand c{j, d.seek_delta, d.result_len, d.has_ts, d.tokens, d.sum_logprobs_all + tok.plog};
                c.tokens.push_back(tok);
                cands.push_back(std::move(c));
            }
        }
        std::stable_sort(cands.begin(), cands.end(), [](const BeamCand& a, const BeamCand& b) {
            if (a.sum_logprobs_all != b.sum_logprobs_all) return a.sum_logprobs_all > b.sum_logprobs_all;
            return a.decoder_idx < b.decoder_idx;
        });
        auto same_tokens = [](const BeamCand& a, const BeamCand& b) {
            if (a.tokens.size() != b.tokens.size()) return false;
            for (size_t q = 0; q < a.tokens.size(); ++q)
                if (a.tokens[q].id != b.tokens[q].id) return false;
            return true;
        };
        std::vector<int> take(nb, -1);
        size_t cur_c = 0;
        for (int j = 0; j < nb; ++j) {
            if (dec[j].completed || dec[j].failed) continue;
            if (cur_c >= cands.size()) cur_c = 0;
            const size_t pick = cur_c++;
            while (cur_c < cands.size() && i > 0 && same_tokens(cands[cur_c], cands[pick])) ++cur_c;
            take[j] = (int)pick;
        }
        // ---- KV caches follow the candidates: a cache moves with its slot index; only a cache that two decoders inherit
        // (or that its owner keeps while another decoder inherits it) is copied, into the slot of a decoder that nobody follows
        std::vector<int> old_slot(nb), new_slot(nb, -1);
        std::vector<char> claimed(st->max_batch, 0);
        for (int j = 0; j < nb; ++j) old_slot[j] = dec[j].slot;
        for (int j = 0; j < nb; ++j)
            if (take[j] >= 0 && cands[take[j]].decoder_idx == j) {
                new_slot[j] = old_slot[j];
                claimed[old_slot[j]] = 1;
            }
        for (int j = 0; j < nb; ++j)
            if (take[j] >= 0 && new_slot[j] < 0) {
                const int src = old_slot[cands[take[j]].decoder_idx];
                if (!claimed[src]) {
                    new_slot[j] = src;
                    claimed[src] = 1;
                }
            }
        for (int j = 0; j < nb; ++j)
            if (take[j] >= 0 && new_slot[j] < 0) {
                int f = -1;
                for (int q = 0; q < nb; ++q)
                    if (!claimed[old_slot[q]]) {
                        f = old_slot[q];
                        break;
                    }
                if (f < 0) {
                    set_error("beam search: no free KV slot");
                    return OH_ERR_DECODE;
                }
                if ((rc = copy_self_kv(st, old_slot[cands[take[j]].decoder_idx], f))) return rc;
                new_slot[j] = f;
                claimed[f] = 1;
            }
        {   // decoders that are finished keep distinct (unused) slots
            int q = 0;
            for (int j = 0; j < nb; ++j)
                if (new_slot[j] < 0) {
                    while (claimed[old_slot[q]]) ++q;
                    new_slot[j] = old_slot[q];
                    claimed[old_slot[q]] = 1;
                }
        }
        for (int j = 0; j < nb; ++j) {
            dec[j].slot = new_slot[j];
            if (take[j] < 0) continue;
            const BeamCand& c = cands[take[j]];
            dec[j].tokens = c.tokens;
            dec[j].seek_delta = c.seek_delta;
            dec[j].has_ts = c.has_ts;
            dec[j].result_len = c.result_len;
            dec[j].sum_logprobs_all = c.sum_logprobs_all;
        }
        // ---- A.5 bookkeeping for the token each live decoder ended up with
        bool all_done = true;
        for (int j = 0; j < nb; ++j) {
            BeamDec& d = dec[j];
            if (d.completed || d.failed) continue;
            SeqFlags f{d.has_ts ? 1 : 0, d.seek_delta, d.result_len, 0, 0};
            apply_token_rules(f, d.tokens.back().id, i, it.seek, it.seek_end, vc.token_beg, vc.token_eot, p.max_tokens, no_timestamps, p.single_segment,
                              n_max);
            d.has_ts = f.has_ts != 0;
            d.seek_delta = f.seek_delta;
            d.result_len = f.result_len;
            d.failed = f.failed != 0;
            d.completed = f.completed != 0;
            if (!d.completed && !d.failed) all_done = false;
        }
        if (all_done || i + 1 == n_max) break;
        // ---- one token per live decoder
        for (int j = 0; j < nb; ++j) {
            const BeamDec& d = dec[j];
            const bool live = !d.completed && !d.failed;
            hs.active[j] = live ? 1 : 0;
            hs.slot[j] = d.slot;
            if (!live) continue;
            hs.pos[j] = pl + i;
            hs.cur_token[j] = d.tokens.back().id;
            hs.n_gen[j] = (int)d.tokens.size();
            hs.has_ts[j] = d.has_ts ? 1 : 0;
            hs.seek_delta[j] = d.seek_delta;
            hs.result_len[j] = d.result_len;
            for (size_t q = 0; q < d.tokens.size(); ++q) hs.tok_id[(size_t)j * kMaxGen + q] = d.tokens[q].id;
        }
        if ((rc = upload_rows())) return rc;
        if ((rc = decode_step_all(st, nb, true, true))) return rc;
        st->timings.n_decode_steps++;
        for (int j = 0; j < nb; ++j) {
            BeamDec& d = dec[j];
            if (d.completed || d.failed) continue;
            d.probs.resize(ld);
            d.logprobs.resize(ld);
            OH_CUDA_CHECK(cudaMemcpyAsync(d.probs.data(), (float*)st->dump_p.p + (size_t)j * ld, ld * 4, cudaMemcpyDeviceToHost, s));
            OH_CUDA_CHECK(cudaMemcpyAsync(d.logprobs.data(), (float*)st->dump_lp.p + (size_t)j * ld, ld * 4, cudaMemcpyDeviceToHost, s));
            st->timings.d2h_bytes += (long long)ld * 8;
        }
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    // ---- rank the decoders (A.5 / A.8): first best score among the ones that did not fail
    double best_score = -INFINITY;
    int best_id = 0;
    std::vector<PassResult> res(nb);
    for (int j = 0; j < nb; ++j) {
        PassResult& r = res[j];
        const BeamDec& d = dec[j];
        r.tokens = d.tokens;
        r.result_len = d.result_len;
        r.seek_delta = d.seek_delta;
        r.has_ts = d.has_ts;
        r.failed = d.failed;
        r.completed = d.completed;
        if (r.failed) continue;
        r.tokens.resize(std::min<size_t>(r.tokens.size(), (size_t)r.result_len));
        sequence_score(p, r);
        if (r.result_len > 32 && r.entropy < p.entropy_thold) {
            r.failed = true;
            continue;
        }
        if (best_score < r.score) {
            best_score = r.score;
            best_id = j;
        }
    }
    best = res[best_id];
    for (const BeamDec& d : dec) st->timings.n_tokens += (int)d.tokens.size();
    return OH_OK;
}

}  // namespace

// full() for strategy = BEAM_SEARCH: items one after the other, windows in order (A.5), beam_size decoder rows
static int run_full_beam(oh_state* st, const oh_full_params& p, int batch, const std::vector<float>& temperatures, const std::vector<int>& prompt_init,
                         bool no_timestamps, const ClipMeta& cm, std::chrono::steady_clock::time_point t_begin, long long launches0) {
    using clk = std::chrono::steady_clock;
    const Model& m = *st->ctx->model;
    const Vocab& vc = m.vocab;
    const HParams& hp = m.hp;
    const int nb = std::max(1, p.beam_size);
    if (nb > st->max_batch) {
        set_error("beam search needs oh_context_params.max_batch >= beam_size (one KV-cache slot per beam)");
        return OH_ERR_INVALID;
    }
    cudaStream_t s = st->stream;
    int rc;
    auto build_prompt = [&](const Item& it, float t_cur, std::vector<int>& prompt) {
        prompt.clear();
        if (!it.prompt_past.empty() && t_cur < 0.5f && p.n_max_text_ctx > 0) {
            const int n_take = std::min(std::min(p.n_max_text_ctx, hp.n_text_ctx / 2), (int)it.prompt_past.size());
            prompt.push_back(vc.token_prev);
            prompt.insert(prompt.end(), it.prompt_past.end() - n_take, it.prompt_past.end());
        }
        prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());
    };
    for (int i = 0; i < batch; ++i) {
        Item& it = st->items[i];
        while (!it.done) {
            if (it.seek + 100 >= it.seek_end) {
                it.done = true;
                break;
            }
            // ---- encode the window into slot 0, then give every beam its own copy of the cross K/V
            auto t0 = clk::now();
            const int slot0 = 0;
            OH_CUDA_CHECK(cudaMemcpyAsync(cm.clip_of_slot, &i, 4, cudaMemcpyHostToDevice, s));
            OH_CUDA_CHECK(cudaMemcpyAsync(cm.seek_of_slot, &it.seek, 4, cudaMemcpyHostToDevice, s));
            WindowGatherArgs wa;
            wa.raw = (const float*)st->mel_raw.p;
            wa.ld = st->mel_ld;
            wa.n_mel = hp.n_mels;
            wa.max_key = cm.max_key;
            wa.n_active = cm.n_active;
            wa.n_len = cm.n_len;
            wa.clip_of_slot = cm.clip_of_slot;
            wa.seek_of_slot = cm.seek_of_slot;
            if ((rc = launch_window_gather(wa, 1, st->xin.p, st->ctx->act_type, s))) return rc == OH_ERR_CUDA ? rc : OH_ERR_ENCODE;
            if ((rc = run_encoder(st, 1, 0, nullptr, &slot0))) return rc == OH_ERR_CUDA ? rc : OH_ERR_ENCODE;
            for (int j = 1; j < nb; ++j)
                if ((rc = copy_cross_kv(st, 0, j))) return rc;
            OH_CUDA_CHECK(cudaStreamSynchronize(s));  // &i / &it.seek were read by the copies above
            st->timings.n_windows++;
            st->timings.encode_ms += std::chrono::duration<float, std::milli>(clk::now() - t0).count();
            if (it.seek > 0 && it.seek + 500 >= it.seek_end) it.prompt_past.clear();  // short tail: drop the prompt (upstream behaviour)

            auto t1 = clk::now();
            PassResult best;
            std::vector<int> prompt;
            for (size_t ti = 0; ti < temperatures.size(); ++ti) {
                const float t_cur = temperatures[ti];
                build_prompt(it, t_cur, prompt);
                if (t_cur <= 0.0f) {
                    if ((rc = beam_pass(st, p, no_timestamps, it, prompt, nb, best))) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
                } else {
                    // A.5: BEAM at t > 0 runs best_of decoders with sampling; best_of defaults to -1 for this strategy -> one decoder
                    std::vector<PassRow> rows = {PassRow{i, 0, prompt}};
                    std::vector<PassResult> res;
                    if ((rc = decode_pass(st, p, no_timestamps, rows, t_cur, res))) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
                    best = res[0];
                    if (!best.failed) {
                        best.tokens.resize(std::min<size_t>(best.tokens.size(), (size_t)best.result_len));
                        sequence_score(p, best);
                        if (best.result_len > 32 && best.entropy < p.entropy_thold) best.failed = true;
                    }
                }
                if (ti + 1 != temperatures.size() && (best.failed || best.avg_logprobs < p.logprob_thold)) {
                    st->timings.n_fallbacks++;
                    continue;
                }
                break;
            }
            emit_segments(m, p, it, best, prompt, (int)prompt_init.size());
            st->timings.decode_ms += std::chrono::duration<float, std::milli>(clk::now() - t1).count();
        }
    }
    st->timings.total_ms = std::chrono::duration<float, std::milli>(clk::now() - t_begin).count();
    st->timings.n_kernel_launches = g_launch_count - launches0;
    return OH_OK;
}

int run_full(oh_state* st, const oh_full_params& p, const float* const* pcm, const int* n, int batch, bool pcm_on_device) {
    using clk = std::chrono::steady_clock;
    const Model& m = *st->ctx->model;
    const Vocab& vc = m.vocab;
    const HParams& hp = m.hp;
    if (batch < 1) {
        set_error("empty batch");
        return OH_ERR_INVALID;
    }
    if (p.strategy != OH_SAMPLING_GREEDY && p.strategy != OH_SAMPLING_BEAM_SEARCH) {
        set_error("unknown sampling strategy");
        return OH_ERR_UNSUPPORTED;
    }
    if (p.suppress_non_speech_tokens || p.offset_ms != 0 || p.duration_ms != 0) {
        set_error("unsupported oh_full_params (suppress_non_speech_tokens / offset_ms / duration_ms)");
        return OH_ERR_UNSUPPORTED;
    }
    st->timings = oh_timings{};
    const long long launches0 = g_launch_count;
    auto t_begin = clk::now();
    cudaStream_t s = st->stream;
    int rc;

    // the prompts of earlier calls survive only with no_context = false
    std::vector<std::vector<int>> carried(batch);
    if (!p.no_context)
        for (int i = 0; i < batch && i < (int)st->items.size(); ++i) carried[i] = st->items[i].prompt_past;

    if ((rc = run_mel(st, pcm, n, batch, pcm_on_device))) return rc;
    OH_CUDA_CHECK(cudaStreamSynchronize(s));
    auto t_mel = clk::now();
    st->timings.mel_ms = std::chrono::duration<float, std::milli>(t_mel - t_begin).count();

    std::vector<float> temperatures;
    if (p.temperature_inc > 0.0f) {
        for (float t = p.temperature; t < 1.0f + 1e-6f; t += p.temperature_inc) temperatures.push_back(t);
    } else {
        temperatures.push_back(p.temperature);
    }
    std::vector<int> prompt_init = {vc.token_sot};
    bool no_timestamps = p.no_timestamps != 0;
    int lang_id = 0;
    if (vc.is_multilingual()) {
        const char* lang = (p.language && strcmp(p.language, "auto") != 0 && p.language[0]) ? p.language : "en";
        lang_id = lang_id_from_code(lang);
        if (lang_id < 0 || lang_id >= vc.num_languages()) lang_id = 0;
        prompt_init.push_back(vc.token_sot + 1 + lang_id);
        prompt_init.push_back(p.translate ? vc.token_translate : vc.token_transcribe);
    }
    if (hp.n_text_layer == 2 && hp.n_vocab != 51866) no_timestamps = true;  // "distilled" heuristic (A.1)
    if (no_timestamps) prompt_init.push_back(vc.token_not);

    for (int i = 0; i < batch; ++i) {
        Item& it = st->items[i];
        it.lang_id = lang_id;
        it.prompt_past = carried[i];
        it.seek = 0;
        it.seek_end = it.n_len_org;
        if (!it.done && it.seek_end < 100) it.done = true;  // < 1 s: no segments, rc 0 (A.3 step 7)
    }

    ClipMeta cm = clip_meta_view(st->clip_meta.p, std::max<size_t>(batch, st->max_batch));
    if (p.strategy == OH_SAMPLING_BEAM_SEARCH) return run_full_beam(st, p, batch, temperatures, prompt_init, no_timestamps, cm, t_begin, launches0);

    // ------------------------------------------------------------------------------------------------------------
    // Scheduler.  Every KV-cache slot is one decoder row.  Windows are admitted to free slots (encoder pass over the
    // admitted windows, cross-KV straight into their slots), rows decode greedily under the device state machine, and
    // every CHECK steps the host looks at the `active` flags: a finished row is scored, its segments are emitted, its
    // item's seek advances, and the item's next window (A.5: windows of one call are sequentially dependent) is queued
    // for admission while the other rows keep decoding — rows never wait for the slowest row of a "round".
    // Rows that need a temperature fallback are parked and resolved together once the greedy rows have drained.
    // ------------------------------------------------------------------------------------------------------------
    struct SlotInfo {
        int item = -1;
        bool armed = false;
        std::vector<int> prompt;
    };
    const int nslots = std::min(batch, st->max_batch);  // more items than slots: items queue for a free slot
    const int n_max = hp.n_text_ctx / 2 - 4;
    const bool lockstep = getenv("OH_SCHED") && strcmp(getenv("OH_SCHED"), "lockstep") == 0;
    // host look-in period (steps) and the number of waiting windows that triggers an encoder pass while rows are decoding
    const int CHECK = getenv("OH_CHECK") ? std::max(1, atoi(getenv("OH_CHECK"))) : 8;
    const int admit_threshold = lockstep ? nslots : (getenv("OH_ADMIT") ? std::max(1, atoi(getenv("OH_ADMIT"))) : std::max(1, std::min(8, nslots / 8)));
    std::vector<SlotInfo> slots(nslots);
    std::vector<int> free_slots;
    for (int i = nslots - 1; i >= 0; --i) free_slots.push_back(i);  // pop_back yields ascending slot ids
    std::vector<int> need_window;                                     // items whose next window awaits admission
    for (int i = 0; i < batch; ++i) {
        Item& it = st->items[i];
        if (it.done) continue;
        if (it.seek + 100 >= it.seek_end) it.done = true;
        else need_window.push_back(i);
    }
    std::vector<PassRow> parked;  // rows waiting for a temperature fallback (slot stays reserved)
    SeqLayout lay(st->max_batch);
    int32_t* hb = (int32_t*)st->pinned;
    DecodeSeqState hs = seq_view(hb, lay);
    memset(hb, 0, lay.total_words * 4);
    OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
    st->timings.h2d_bytes += (long long)lay.total_words * 4;
    bool rules_dirty = true;
    int n_armed = 0;
    const size_t scalar_words = lay.prompt_off + (size_t)lay.rows * kMaxPrompt;  // scalars + prompts
    const size_t token_words = lay.counter_off - lay.tok_off[0];

    auto build_prompt = [&](const Item& it, float t_cur, std::vector<int>& prompt) {
        prompt.clear();
        if (!it.prompt_past.empty() && t_cur < 0.5f && p.n_max_text_ctx > 0) {
            const int n_take = std::min(std::min(p.n_max_text_ctx, hp.n_text_ctx / 2), (int)it.prompt_past.size());
            prompt.push_back(vc.token_prev);
            prompt.insert(prompt.end(), it.prompt_past.end() - n_take, it.prompt_past.end());
        }
        prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());
    };
    // accept a finished window: segments, prompt carry-over, seek advance, next window or done
    auto finish_window = [&](int slot, const PassResult& res) {
        SlotInfo& si = slots[slot];
        Item& it = st->items[si.item];
        emit_segments(m, p, it, res, si.prompt, (int)prompt_init.size());
        if (it.seek + 100 >= it.seek_end) it.done = true;
        else need_window.push_back(si.item);
        si.item = -1;
        si.armed = false;
        free_slots.push_back(slot);
    };
    // score one t=0 result; returns true if it is accepted at this temperature index
    auto evaluate = [&](PassResult& pr, size_t ti) {
        if (!pr.failed) {
            pr.tokens.resize(std::min<size_t>(pr.tokens.size(), (size_t)pr.result_len));
            sequence_score(p, pr);
            if (pr.result_len > 32 && pr.entropy < p.entropy_thold) pr.failed = true;
        }
        if (ti + 1 != temperatures.size() && (pr.failed || pr.avg_logprobs < p.logprob_thold)) return false;
        return true;
    };

    while (true) {
        // ---- admission: encode the next windows of waiting items into free slots
        if (!need_window.empty() && !free_slots.empty() && (n_armed == 0 || (!lockstep && (int)need_window.size() >= admit_threshold))) {
            auto t0 = clk::now();
            const int k = (int)std::min(need_window.size(), free_slots.size());
            std::sort(free_slots.begin(), free_slots.end(), std::greater<int>());
            std::vector<int> win_item(need_window.begin(), need_window.begin() + k), win_slot(k), win_seek(k);
            need_window.erase(need_window.begin(), need_window.begin() + k);
            for (int i = 0; i < k; ++i) {
                win_slot[i] = free_slots.back();
                free_slots.pop_back();
                win_seek[i] = st->items[win_item[i]].seek;
            }
            OH_CUDA_CHECK(cudaMemcpyAsync(cm.clip_of_slot, win_item.data(), k * 4, cudaMemcpyHostToDevice, s));
            OH_CUDA_CHECK(cudaMemcpyAsync(cm.seek_of_slot, win_seek.data(), k * 4, cudaMemcpyHostToDevice, s));
            for (int c0 = 0; c0 < k; c0 += st->enc_chunk) {
                const int nw = std::min(st->enc_chunk, k - c0);
                WindowGatherArgs wa;
                wa.raw = (const float*)st->mel_raw.p;
                wa.ld = st->mel_ld;
                wa.n_mel = hp.n_mels;
                wa.max_key = cm.max_key;
                wa.n_active = cm.n_active;
                wa.n_len = cm.n_len;
                wa.clip_of_slot = cm.clip_of_slot + c0;
                wa.seek_of_slot = cm.seek_of_slot + c0;
                if ((rc = launch_window_gather(wa, nw, st->xin.p, st->ctx->act_type, s))) return rc == OH_ERR_CUDA ? rc : OH_ERR_ENCODE;
                if ((rc = run_encoder(st, nw, 0, nullptr, win_slot.data() + c0))) return rc == OH_ERR_CUDA ? rc : OH_ERR_ENCODE;
                st->timings.n_windows += nw;
            }
            // arm the rows (the host mirror of the scalars is current: nothing ran since the last download)
            for (int i = 0; i < k; ++i) {
                const int r = win_slot[i];
                SlotInfo& si = slots[r];
                Item& it = st->items[win_item[i]];
                if (it.seek > 0 && it.seek + 500 >= it.seek_end) it.prompt_past.clear();  // short tail: drop the prompt (upstream behaviour)
                si.item = win_item[i];
                si.armed = true;
                build_prompt(it, temperatures[0], si.prompt);
                const int pl = (int)si.prompt.size();
                if (pl > kMaxPrompt) {
                    set_error("prompt too long");
                    return OH_ERR_DECODE;
                }
                hs.slot[r] = r;
                hs.active[r] = 1;
                hs.pos[r] = 0;
                hs.cur_token[r] = si.prompt[0];
                hs.prompt_len[r] = pl;
                for (int j = 0; j < pl; ++j) hs.prompt[(size_t)r * kMaxPrompt + j] = si.prompt[j];
                hs.n_gen[r] = 0;
                hs.has_ts[r] = 0;
                hs.seek_delta[r] = 100 * 30;
                hs.result_len[r] = 0;
                hs.failed[r] = 0;
                hs.completed[r] = 0;
                hs.seek[r] = it.seek;
                hs.seek_end[r] = it.seek_end;
                st->timings.n_prompt_steps += pl - 1;
            }
            OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, scalar_words * 4, cudaMemcpyHostToDevice, s));
            st->timings.h2d_bytes += (long long)scalar_words * 4;
            n_armed += k;
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
            st->timings.encode_ms += std::chrono::duration<float, std::milli>(clk::now() - t0).count();
        }
        if (n_armed == 0) {
            if (!parked.empty()) {
                // ---- temperature fallback for the parked rows, lock-step per temperature (A.5 ladder)
                auto t0 = clk::now();
                std::vector<PassRow> pending;
                pending.swap(parked);
                for (size_t ti = 1; ti < temperatures.size() && !pending.empty(); ++ti) {
                    for (PassRow& pr : pending) {
                        build_prompt(st->items[pr.item], temperatures[ti], pr.prompt);
                        slots[pr.slot].prompt = pr.prompt;
                    }
                    std::vector<PassResult> res;
                    if ((rc = decode_pass(st, p, no_timestamps, pending, temperatures[ti], res))) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
                    std::vector<PassRow> next;
                    for (size_t r = 0; r < pending.size(); ++r) {
                        if (evaluate(res[r], ti)) finish_window(pending[r].slot, res[r]);
                        else {
                            st->timings.n_fallbacks++;
                            next.push_back(pending[r]);
                        }
                    }
                    pending.swap(next);
                }
                rules_dirty = true;
                memset(hb, 0, lay.total_words * 4);  // decode_pass reused the block: start the greedy rows from a clean slate
                OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
                st->timings.decode_ms += std::chrono::duration<float, std::milli>(clk::now() - t0).count();
                continue;
            }
            if (need_window.empty()) break;
            continue;  // admission fires because nothing is armed
        }
        // ---- CHECK greedy steps for every slot (inactive rows exit at once)
        auto t1 = clk::now();
        if (rules_dirty) {
            DecodeRules R;
            R.n_vocab = vc.n_vocab; R.eot = vc.token_eot; R.sot = vc.token_sot; R.translate = vc.token_translate; R.transcribe = vc.token_transcribe;
            R.solm = vc.token_solm; R.prev = vc.token_prev; R.nosp = vc.token_nosp; R.tok_not = vc.token_not; R.beg = vc.token_beg;
            R.blank = vc.token_blank; R.n_lang = vc.num_languages();
            R.no_timestamps = no_timestamps; R.suppress_blank = p.suppress_blank; R.single_segment = p.single_segment; R.max_tokens = p.max_tokens;
            R.n_max = n_max;
            R.tid0_initial = p.max_initial_ts > 0.0f ? (int)std::round(p.max_initial_ts / (30.0f / (float)hp.n_audio_ctx)) : -1;
            R.temperature = temperatures[0];
            OH_CUDA_CHECK(cudaMemcpyAsync(st->rules_dev.p, &R, sizeof R, cudaMemcpyHostToDevice, s));
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
            rules_dirty = false;
        }
        if (temperatures[0] >= 1e-6f) {
            // a non-zero base temperature samples on the host from the first pass on: use the lock-step path for everything
            for (int r = 0; r < nslots; ++r)
                if (slots[r].armed) parked.push_back(PassRow{slots[r].item, r, slots[r].prompt});
            memset(hb, 0, scalar_words * 4);
            n_armed = 0;
            // run them as "fallback index 0"
            std::vector<PassRow> pending;
            pending.swap(parked);
            for (size_t ti = 0; ti < temperatures.size() && !pending.empty(); ++ti) {
                for (PassRow& pr : pending) {
                    build_prompt(st->items[pr.item], temperatures[ti], pr.prompt);
                    slots[pr.slot].prompt = pr.prompt;
                }
                std::vector<PassResult> res;
                if ((rc = decode_pass(st, p, no_timestamps, pending, temperatures[ti], res))) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
                std::vector<PassRow> next;
                for (size_t r = 0; r < pending.size(); ++r) {
                    if (evaluate(res[r], ti)) finish_window(pending[r].slot, res[r]);
                    else {
                        st->timings.n_fallbacks++;
                        next.push_back(pending[r]);
                    }
                }
                pending.swap(next);
            }
            rules_dirty = true;
            memset(hb, 0, lay.total_words * 4);
            OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
            continue;
        }
        for (int step = 0; step < CHECK; ++step) {
            if ((rc = run_decode_step_graph(st, nslots))) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
            st->timings.n_decode_steps++;
            st->timings.n_row_steps += nslots;
        }
        OH_CUDA_CHECK(cudaMemcpyAsync(hb, st->seq_dev.p, scalar_words * 4, cudaMemcpyDeviceToHost, s));
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
        st->timings.d2h_bytes += (long long)scalar_words * 4;
        bool any_finished = false;
        for (int r = 0; r < nslots; ++r)
            if (slots[r].armed && !hs.active[r]) any_finished = true;
        if (any_finished) {
            OH_CUDA_CHECK(cudaMemcpyAsync(hb + lay.tok_off[0], (int32_t*)st->seq_dev.p + lay.tok_off[0], token_words * 4, cudaMemcpyDeviceToHost, s));
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
            st->timings.d2h_bytes += (long long)token_words * 4;
            for (int r = 0; r < nslots; ++r) {
                if (!slots[r].armed || hs.active[r]) continue;
                PassResult pr;
                const int ng = hs.n_gen[r];
                pr.tokens.resize(ng);
                for (int i = 0; i < ng; ++i) {
                    const size_t o = (size_t)r * kMaxGen + i;
                    pr.tokens[i] = TokenData{hs.tok_id[o], hs.tok_tid[o], hs.tok_p[o], hs.tok_plog[o], hs.tok_pt[o], hs.tok_ptsum[o]};
                }
                pr.result_len = hs.result_len[r];
                pr.seek_delta = hs.seek_delta[r];
                pr.has_ts = hs.has_ts[r];
                pr.failed = hs.failed[r];
                pr.completed = hs.completed[r];
                st->timings.n_tokens += ng;
                slots[r].armed = false;
                n_armed--;
                if (evaluate(pr, 0)) finish_window(r, pr);
                else {
                    st->timings.n_fallbacks++;
                    parked.push_back(PassRow{slots[r].item, r, slots[r].prompt});
                }
            }
        }
        st->timings.decode_ms += std::chrono::duration<float, std::milli>(clk::now() - t1).count();
    }
    st->timings.total_ms = std::chrono::duration<float, std::milli>(clk::now() - t_begin).count();
    st->timings.n_kernel_launches = g_launch_count - launches0;
    return OH_OK;
}

}  // namespace oh
