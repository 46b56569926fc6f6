// Buffers, stage runners (mel / encoder / decoder step) and the batched full() driver.
//
// run_full() is the device-side replacement of whisper.cpp's `whisper_full_with_state` (SURVEY.md A.5, A.7,
// A.8), reached from /root/reference/src/engine/whisper.rs:266.  Each batch item is one independent call of
// the reference; items advance window by window in lock-step rounds:
//   gather mel window at item.seek -> encoder -> cross-KV -> greedy decode (device state machine, CUDA graph)
//   -> host scoring / temperature fallback -> segments -> seek += seek_delta.
#include <cuda.h>
#include <cudaTypedefs.h>

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
thread_local int g_sm_limit = 0;
thread_local int g_gemm_compact = 0;
thread_local int g_gemm_compact_bytes = 96 * 1024;
thread_local Profiler* g_prof = nullptr;
thread_local int g_active_rows_hint = -1;
#ifdef OH_DIAG
unsigned g_skip_mask = 0;  // OH_SKIP_CATS (diagnostic build only): kernel categories that are not launched
#endif
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
void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

static inline size_t rup(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- layout of the DecodeSeqState block (ints unless noted; SeqLayout in engine.h) ------------------------------
DecodeSeqState seq_view(int32_t* base, const SeqLayout& L) {
    DecodeSeqState s;
    using namespace seqlay;
    s.slot = base + SLOT * L.rows;
    s.sslot = base + SSLOT * L.rows;
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
    s.slot += row0; s.sslot += row0; s.active += row0; s.pos += row0; s.cur_token += row0; s.prompt_len += row0; s.n_gen += row0; s.has_ts += row0;
    s.seek_delta += row0; s.result_len += row0; s.failed += row0; s.completed += row0; s.seek += row0; s.seek_end += row0;
    s.prompt += (size_t)row0 * kMaxPrompt;
    s.tok_id += (size_t)row0 * kMaxGen; s.tok_tid += (size_t)row0 * kMaxGen; s.tok_p += (size_t)row0 * kMaxGen;
    s.tok_plog += (size_t)row0 * kMaxGen; s.tok_pt += (size_t)row0 * kMaxGen; s.tok_ptsum += (size_t)row0 * kMaxGen;
    return s;
}

// ---- encoder stream on an SM partition (green context) --------------------------------------------------------------
// The encoder is tensor-bound, a decode step is HBM- and latency-bound; run concurrently they hide each other, but the
// persistent tcgen05 mat-mul holds every SM it lands on for milliseconds, and the step's short kernels would queue behind
// it.  OH_ENC_SMS=n confines the encoder stream to n SMs (CUDA green context; the driver rounds n to its partition
// granularity), so that the decode stream always finds the other SMs free.  Unset / 0: a plain second stream.
namespace {
template <typename F>
F driver_fn(const char* name) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<F>(fn);
}
}  // namespace

static bool create_partition_stream(oh_state* st, int want_sms) {
    auto get_dev = driver_fn<PFN_cuDeviceGet>("cuDeviceGet");
    auto get_res = driver_fn<PFN_cuDeviceGetDevResource>("cuDeviceGetDevResource");
    auto split = driver_fn<PFN_cuDevSmResourceSplitByCount>("cuDevSmResourceSplitByCount");
    auto gen = driver_fn<PFN_cuDevResourceGenerateDesc>("cuDevResourceGenerateDesc");
    auto gcreate = driver_fn<PFN_cuGreenCtxCreate>("cuGreenCtxCreate");
    auto gstream = driver_fn<PFN_cuGreenCtxStreamCreate>("cuGreenCtxStreamCreate");
    if (!get_dev || !get_res || !split || !gen || !gcreate || !gstream) return false;
    CUdevice dev;
    if (get_dev(&dev, st->ctx->params.gpu_device) != CUDA_SUCCESS) return false;
    CUdevResource all{}, part{}, rest{};
    if (get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned int n_groups = 1;
    if (split(&part, &n_groups, &all, &rest, 0, (unsigned)want_sms) != CUDA_SUCCESS || n_groups < 1) return false;
    CUdevResourceDesc desc = nullptr;
    if (gen(&desc, &part, 1) != CUDA_SUCCESS) return false;
    CUgreenCtx g = nullptr;
    if (gcreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream cs = nullptr;
    if (gstream(&cs, g, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
        if (auto gdestroy = driver_fn<PFN_cuGreenCtxDestroy>("cuGreenCtxDestroy")) gdestroy(g);
        return false;
    }
    st->green_ctx = g;
    st->enc_stream = (cudaStream_t)cs;
    st->enc_sms = (int)part.sm.smCount;
    return true;
}

int state_alloc(oh_state* st) {
    const Model& m = *st->ctx->model;
    const HParams& hp = m.hp;
    const int d = hp.n_audio_state, C = hp.n_mels, T = hp.n_audio_ctx, L = hp.n_text_layer, V = hp.n_vocab;
    st->esz = st->ctx->act_type == DT_F32 ? 4 : 2;
    const size_t e = st->esz;
    st->max_batch = std::max(1, st->ctx->params.max_batch);
    int chunk = std::min(st->max_batch, 32);
    if (const char* s = getenv("OH_ENC_CHUNK")) chunk = std::max(1, std::min(std::max(st->max_batch, 1), atoi(s)));
    st->enc_chunk = chunk;
    // look-ahead slots: windows are encoded (on the encoder stream) before a decoder row is free for them, so a row
    // that finishes is re-armed at the next look-in instead of waiting for an encoder pass
    int lookahead = st->max_batch >= 16 ? std::min(st->max_batch, 3 * chunk) : 0;
    if (const char* s = getenv("OH_LOOKAHEAD")) lookahead = std::max(0, atoi(s));
    st->n_slots = st->max_batch + lookahead;
    if (const char* s = getenv("OH_ASYNC_ENC")) st->async_encoder = atoi(s) != 0;
    if (const char* s = getenv("OH_CHECK")) st->check_steps = std::max(1, atoi(s));
    if (const char* s = getenv("OH_ENC_COMPACT")) st->enc_compact_bytes = atoi(s) > 0 ? std::max(64, atoi(s)) * 1024 : 0;  // KB of ring per encoder mat-mul CTA
    if (const char* s = getenv("OH_SCHED_TRACE")) {
        st->trace = fopen(s, "a");
        if (st->trace) fprintf(st->trace, "step,armed,ready,need_window,passes,free_slots,tickets\n");
    }
    st->starve_rows = std::max(1, std::min(8, st->max_batch / 16));
    if (const char* s = getenv("OH_STARVE")) st->starve_rows = std::max(1, atoi(s));
    if (const char* s = getenv("OH_GEMM")) st->gemm_mode = strcmp(s, "simt") == 0 ? 1 : 0;
    if (const char* s = getenv("OH_ATTN")) st->attn_mode = strcmp(s, "simple") == 0 ? 1 : (strcmp(s, "flash") == 0 ? 2 : 0);  // flash: the mma.sync kernel
    if (const char* s = getenv("OH_GRAPHS")) st->use_graphs = atoi(s) != 0;
    if (const char* s = getenv("OH_PDL")) st->use_pdl = atoi(s) != 0;
    if (const char* s = getenv("OH_DUAL")) st->dual_min_rows = atoi(s) > 0 ? atoi(s) : (1 << 30);  // OH_DUAL=0: never interleave row groups
    if (const char* s = getenv("OH_GELU")) st->gelu_lut_f32 = strcmp(s, "lut") == 0;
#ifdef OH_DIAG
    if (const char* s = getenv("OH_SKIP_CATS")) g_skip_mask = (unsigned)strtoul(s, nullptr, 0);
#endif
    if (const char* s = getenv("OH_DECODE_GEMM")) st->splitk_decode = strcmp(s, "tc") != 0;  // OH_DECODE_GEMM=tc: decode mat-muls through the persistent tcgen05 kernel

    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // lo = numerically largest = least urgent
    OH_CUDA_CHECK(cudaStreamCreateWithPriority(&st->own_stream, cudaStreamNonBlocking, prio_hi));
    st->stream = st->own_stream;
    OH_CUDA_CHECK(cudaStreamCreateWithFlags(&st->stream2, cudaStreamNonBlocking));
    OH_CUDA_CHECK(cudaStreamCreateWithFlags(&st->fe_stream, cudaStreamNonBlocking));
    int enc_sms = 0;
    if (const char* s = getenv("OH_ENC_SMS")) enc_sms = std::max(0, atoi(s));
    if (!(enc_sms > 0 && create_partition_stream(st, enc_sms))) {
        st->green_ctx = nullptr;
        st->enc_sms = 0;
        // without a partition the encoder stream outranks the decode stream: its CTAs are placed first whenever SMs free up,
        // so encoder passes finish at their stand-alone speed and the rows are never left without a window
        // (OH_ENC_PRIO=lo: the encoder only gets what the decode steps leave idle)
        const char* pe = getenv("OH_ENC_PRIO");
        const bool enc_hi = !(pe && strcmp(pe, "lo") == 0);
        if (enc_hi) {
            cudaStreamDestroy(st->own_stream);
            OH_CUDA_CHECK(cudaStreamCreateWithPriority(&st->own_stream, cudaStreamNonBlocking, prio_lo));
            st->stream = st->own_stream;
        }
        OH_CUDA_CHECK(cudaStreamCreateWithPriority(&st->enc_stream, cudaStreamNonBlocking, enc_hi ? prio_hi : prio_lo));
    }
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
    st->win_desc_cap = 4 * chunk;  // ring: up to two passes in flight plus the one being filled
    if ((rc = ensure(st->win_desc, (size_t)st->win_desc_cap * sizeof(WindowDesc)))) return rc;
    OH_CUDA_CHECK(cudaMallocHost((void**)&st->win_desc_host, (size_t)st->win_desc_cap * sizeof(WindowDesc)));
    if ((rc = ensure(st->slot_ring, (size_t)st->win_desc_cap * sizeof(int)))) return rc;
    OH_CUDA_CHECK(cudaMallocHost((void**)&st->slot_ring_host, (size_t)st->win_desc_cap * sizeof(int)));
    const size_t R = st->max_batch, S = st->n_slots;
    if ((rc = ensure(st->xkv, (size_t)L * S * T * 2 * d * e))) return rc;
    if ((rc = ensure(st->kself, (size_t)L * R * hp.n_text_ctx * d * e))) return rc;
    if ((rc = ensure(st->vself, (size_t)L * R * hp.n_text_ctx * d * e))) return rc;
    if ((rc = ensure(st->dx, R * d * 4))) return rc;
    if ((rc = ensure(st->dh, R * d * e))) return rc;
    if ((rc = ensure(st->dqkv, R * 3 * d * e))) return rc;
    if ((rc = ensure(st->datt, R * d * e))) return rc;
    if ((rc = ensure(st->dq, R * d * e))) return rc;
    if ((rc = ensure(st->dhid, R * 4 * d * e))) return rc;
    st->ld_logits = (int)rup(V, 8);
    if ((rc = ensure(st->logits, R * st->ld_logits * 4))) return rc;
    SeqLayout lay(R);
    if ((rc = ensure(st->seq_dev, lay.total_words * 4))) return rc;
    st->seq = seq_view((int32_t*)st->seq_dev.p, lay);
    if ((rc = ensure(st->rules_dev, sizeof(DecodeRules)))) return rc;
    st->pinned_bytes = lay.total_words * 4 + 4096;
    OH_CUDA_CHECK(cudaMallocHost(&st->pinned, st->pinned_bytes));
    memset(st->pinned, 0, st->pinned_bytes);
    st->items.clear();
    Sched& sc = st->sched;
    sc.rows.assign(R, RowInfo());
    sc.free_slots.clear();
    for (int i = (int)S - 1; i >= 0; --i) sc.free_slots.push_back(i);  // pop_back yields ascending slot ids
    return OH_OK;
}

void state_release(oh_state* st) {
    sched_release(st);
    if (st->trace) fclose(st->trace);
    st->trace = nullptr;
    for (auto& kv : st->step_graphs) cudaGraphExecDestroy(kv.second);
    st->step_graphs.clear();
    DevBuf* all[] = {&st->pcm_stage, &st->mel_raw, &st->clip_meta, &st->xin, &st->x1, &st->x, &st->h, &st->qkv, &st->hid, &st->enc_f32,
                     &st->win_desc, &st->slot_ring, &st->xkv, &st->kself, &st->vself, &st->dx, &st->dh, &st->dqkv, &st->datt, &st->dq, &st->dhid, &st->logits,
                     &st->dump_lp, &st->dump_p, &st->topk_dev, &st->beam_pairs, &st->resample_plan, &st->resample_in, &st->resample_out, &st->level_pcm, &st->level_env, &st->level_aux, &st->seq_dev, &st->rules_dev};
    for (DevBuf* b : all) release(*b);
    if (st->pinned) cudaFreeHost(st->pinned);
    st->pinned = nullptr;
    if (st->win_desc_host) cudaFreeHost(st->win_desc_host);
    st->win_desc_host = nullptr;
    if (st->slot_ring_host) cudaFreeHost(st->slot_ring_host);
    st->slot_ring_host = nullptr;
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
    if (st->fe_stream) cudaStreamDestroy(st->fe_stream);
    st->fe_stream = nullptr;
    if (st->enc_stream) cudaStreamDestroy(st->enc_stream);
    st->enc_stream = nullptr;
    if (st->green_ctx) {
        if (auto gdestroy = driver_fn<PFN_cuGreenCtxDestroy>("cuGreenCtxDestroy")) gdestroy((CUgreenCtx)st->green_ctx);
        st->green_ctx = nullptr;
    }
    for (auto& ev : st->ev_dual) {
        if (ev) cudaEventDestroy(ev);
        ev = nullptr;
    }
}

// ------------------------------------------------------------------------------------------------ mel
// Per-clip bookkeeping of A.3 (n_len, n_len_org, frames that see samples) and the staging layout of the clips' PCM.
MelPlan plan_mel(std::vector<Item>& items, const int* n, int batch) {
    MelPlan pl;
    items.assign(batch, Item());
    pl.pcm_off.resize(batch);
    for (int i = 0; i < batch; ++i) {
        Item& it = items[i];
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
        pl.pcm_off[i] = pl.stage_floats;
        pl.stage_floats += rup((size_t)std::max(n[i], 0), 4);
        pl.max_active = std::max(pl.max_active, it.n_active);
    }
    pl.ld = (int)rup(std::max(pl.max_active, 1), 32);
    return pl;
}

// clip_meta layout (device): [pcm ptrs (8 B each)] [n_samples] [n_active] [max_key]
struct ClipMeta {
    const float** pcm;
    int *n_samples, *n_active, *max_key;
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
    c.max_key = (int*)p;
    return c;
}
static size_t clip_meta_bytes(size_t cap) { return cap * (8 + 3 * 4); }

int launch_mel(oh_state* st, const MelPlan& pl, const std::vector<Item>& items, const float* const* pcm, bool pcm_on_device, DevBuf& pcm_stage,
               DevBuf& mel_raw, DevBuf& clip_meta, cudaStream_t stream, long long* h2d_bytes) {
    const Model& m = *st->ctx->model;
    const int n_mel = m.hp.n_mels, batch = (int)items.size();
    int rc;
    if ((rc = ensure(clip_meta, clip_meta_bytes(batch)))) return rc;
    if ((rc = ensure(mel_raw, (size_t)batch * n_mel * pl.ld * 4))) return rc;
    ClipMeta cm = clip_meta_view(clip_meta.p, batch);
    std::vector<const float*> ptrs(batch);
    std::vector<int> ns(batch), na(batch);
    if (!pcm_on_device) {
        if ((rc = ensure(pcm_stage, std::max<size_t>(pl.stage_floats, 4) * 4))) return rc;
        for (int i = 0; i < batch; ++i) {
            ptrs[i] = (const float*)pcm_stage.p + pl.pcm_off[i];
            if (items[i].n_samples > 0) {
                OH_CUDA_CHECK(cudaMemcpyAsync((float*)pcm_stage.p + pl.pcm_off[i], pcm[i], (size_t)items[i].n_samples * 4, cudaMemcpyHostToDevice, stream));
                if (h2d_bytes) *h2d_bytes += (long long)items[i].n_samples * 4;
            }
        }
    } else {
        for (int i = 0; i < batch; ++i) ptrs[i] = pcm[i];
    }
    for (int i = 0; i < batch; ++i) {
        ns[i] = items[i].n_samples;
        na[i] = items[i].n_active;
    }
    // the vectors are pageable: the runtime stages such copies before cudaMemcpyAsync returns
    OH_CUDA_CHECK(cudaMemcpyAsync(cm.pcm, ptrs.data(), (size_t)batch * 8, cudaMemcpyHostToDevice, stream));
    OH_CUDA_CHECK(cudaMemcpyAsync(cm.n_samples, ns.data(), (size_t)batch * 4, cudaMemcpyHostToDevice, stream));
    OH_CUDA_CHECK(cudaMemcpyAsync(cm.n_active, na.data(), (size_t)batch * 4, cudaMemcpyHostToDevice, stream));
    MelArgs a;
    a.pcm = cm.pcm;
    a.n_samples = cm.n_samples;
    a.n_active = cm.n_active;
    a.raw = (float*)mel_raw.p;
    a.ld = pl.ld;
    a.max_key = cm.max_key;
    a.n_mel = n_mel;
    a.twiddle = st->ctx->d_twiddle;
    a.hann = st->ctx->d_hann;
    a.filt_start = m.mel.start;
    a.filt_count = m.mel.count;
    a.filt_off = m.mel.offset;
    a.filt_w = m.mel.weights;
    a.n_filt_w = m.mel.n_weights;
    if (pl.max_active > 0) {
        if ((rc = launch_mel_frames(a, batch, pl.max_active, stream))) return rc;
    }
    return OH_OK;
}

const int* clip_max_key(const DevBuf& clip_meta, int batch, int clip) { return clip_meta_view(clip_meta.p, batch).max_key + clip; }

int run_mel(oh_state* st, const float* const* pcm, const int* n, int batch, bool pcm_on_device) {
    MelPlan pl = plan_mel(st->items, n, batch);
    st->mel_clips = batch;
    st->mel_ld = pl.ld;
    return launch_mel(st, pl, st->items, pcm, pcm_on_device, st->pcm_stage, st->mel_raw, st->clip_meta, st->stream, &st->timings.h2d_bytes);
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
        if (g.batch * g.rows <= 256 && st->splitk_decode) {
            int rc = launch_gemm_splitk(g, stream);
            if (rc != OH_ERR_UNSUPPORTED) return rc;
        }
        int rc = launch_gemm_tc(g, stream);
        if (rc != OH_ERR_UNSUPPORTED) return rc;
    }
    return launch_gemm_simt(g, stream);
}

// ------------------------------------------------------------------------------------------------ encoder
int run_encoder(oh_state* st, int nw, int slot0, float* d_enc_f32_out, const int* slot_list, cudaStream_t stream) {
    const Model& m = *st->ctx->model;
    const HParams& hp = m.hp;
    const int d = hp.n_audio_state, C = hp.n_mels, T = hp.n_audio_ctx, act = st->ctx->act_type;
    const size_t e = st->esz;
    cudaStream_t s = stream ? stream : st->stream;
    // on an SM partition the persistent mat-mul sizes its grid for the partition, not for the whole device
    struct SmLimit {
        int saved;
        explicit SmLimit(int v) : saved(g_sm_limit) { g_sm_limit = v; }
        ~SmLimit() { g_sm_limit = saved; }
    } sm_limit((stream && stream == st->enc_stream && st->enc_sms > 0) ? st->enc_sms : 0);
    struct Compact {  // OH_ENC_COMPACT: the pass shares the SMs with decode steps (see gemm_tc.cu)
        int saved;
        explicit Compact(int v) : saved(g_gemm_compact) { g_gemm_compact = v; }
        ~Compact() { g_gemm_compact = saved; }
    } compact((st->enc_compact_bytes > 0 && ((stream && stream == st->enc_stream) || getenv("OH_ENC_COMPACT_ALL"))) ? 1 : 0);
    if (st->enc_compact_bytes > 0) g_gemm_compact_bytes = st->enc_compact_bytes;
    auto gemm = [&](oh_state* st_, GemmArgs g_) { return oh::gemm(st_, g_, s); };
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
    // cross-attention K/V of every decoder layer, straight into the slots' caches: [L][slots][2][heads][T][64].
    // slot_list (optional): window w goes to slot slot_list[w]; the mat-mul scatters its batches through a device copy of
    // the list, so one launch per layer serves the pass however fragmented the free slots are
    const int* d_slots = nullptr;
    if (slot_list) {
        const int area = st->slot_cursor++ % 4;
        int* hs = st->slot_ring_host + (size_t)area * st->enc_chunk;
        int* ds = (int*)st->slot_ring.p + (size_t)area * st->enc_chunk;
        memcpy(hs, slot_list, (size_t)nw * sizeof(int));
        OH_CUDA_CHECK(cudaMemcpyAsync(ds, hs, (size_t)nw * sizeof(int), cudaMemcpyHostToDevice, s));
        d_slots = ds;
    }
    for (int l = 0; l < hp.n_text_layer; ++l) {
        const DecLayer& D = m.dec[l];
        GemmArgs g;  // one "batch" per window so that the slot index separates from the row index
        g.A = st->h.p; g.a_type = act; g.lda = d; g.batch = nw; g.rows = T; g.a_bstride = (long long)T * d;
        g.W = D.xkv.w; g.N = 2 * d; g.K = d; g.bias = D.xkv.b;
        g.C = (char*)st->xkv.p + ((size_t)l * st->n_slots + (slot_list ? 0 : slot0)) * T * 2 * d * e;
        g.c_type = act;
        g.c_bstride = (long long)T * 2 * d;       // one slot: [2 (k,v)][heads][T][64]
        g.c_bidx = d_slots;
        g.c_hsplit = kHeadDim;                    // column n = (kv * heads + head) * 64 + c
        g.c_hstride = (long long)T * kHeadDim;
        g.ldc = kHeadDim;
        if ((rc = gemm(st, g))) return rc;
    }
    return OH_OK;
}

// ------------------------------------------------------------------------------------------------ decoder step
// One decoder step for rows [row0, row0 + rows) on stream `s`.  `after_q0` (optional) is recorded right after the first
// layer's cross-attention query projection: the interleaved step starts the second row group there.
static int decode_step_launch(oh_state* st, int row0, int rows, int k9_mode, cudaStream_t s, cudaEvent_t after_q0 = nullptr) {
    const bool with_k9 = k9_mode != K9_NONE, dump = k9_mode == K9_DUMP;
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
        const char* xkv = (const char*)st->xkv.p + (size_t)l * st->n_slots * T * 2 * d * e;
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
        TopKOut tk;
        if (k9_mode == K9_TOPK) {  // beam search: [ids | logprobs | probs] x kTopK per row, then tid, pt, ptsum per row
            const size_t R = st->max_batch;
            if ((rc = ensure(st->topk_dev, R * (3 * kTopK + 3) * 4))) return rc;
            int32_t* base = (int32_t*)st->topk_dev.p;
            tk.ids = base + r0 * kTopK;
            tk.logprobs = (float*)(base + R * kTopK) + r0 * kTopK;
            tk.probs = (float*)(base + 2 * R * kTopK) + r0 * kTopK;
            tk.tid = base + 3 * R * kTopK + r0;
            tk.pt = (float*)(base + 3 * R * kTopK + R) + r0;
            tk.ptsum = (float*)(base + 3 * R * kTopK + 2 * R) + r0;
        }
        if ((rc = launch_logits_step(logits, st->ld_logits, seq, (const DecodeRules*)st->rules_dev.p, rows, lp, pr, s, tk))) return rc;
    }
    return OH_OK;
}

// The whole step.  With enough rows the batch is cut into two row groups that run on two streams, the second one
// starting when the first reaches its first cross-attention: from then on one group streams cross-attention K/V (the
// HBM-bound part of a step) while the other walks through its latency-bound chain of projections and LayerNorms, so
// the two kinds of work overlap instead of alternating.  Rows are independent, so the results do not change.
// Measured on B200 (large-v3, 128 rows): 9.3 ms against 8.8 ms for the single chain — the projections of one group
// queue behind the other group's resident cross-attention CTAs — so the interleave is opt-in (OH_DUAL=<min rows>).
static int decode_step_all(oh_state* st, int rows, int k9_mode) {
    cudaStream_t s = st->stream;
    if (rows < st->dual_min_rows || k9_mode == K9_DUMP || k9_mode == K9_TOPK || g_prof != nullptr) return decode_step_launch(st, 0, rows, k9_mode, s);
    const int ra = (rows + 1) / 2;
    int rc;
    OH_CUDA_CHECK(cudaEventRecord(st->ev_dual[0], s));
    OH_CUDA_CHECK(cudaStreamWaitEvent(st->stream2, st->ev_dual[0], 0));
    if ((rc = decode_step_launch(st, 0, ra, k9_mode, s, st->ev_dual[1]))) return rc;
    OH_CUDA_CHECK(cudaStreamWaitEvent(st->stream2, st->ev_dual[1], 0));
    if ((rc = decode_step_launch(st, ra, rows - ra, k9_mode, st->stream2))) return rc;
    OH_CUDA_CHECK(cudaEventRecord(st->ev_dual[2], st->stream2));
    OH_CUDA_CHECK(cudaStreamWaitEvent(s, st->ev_dual[2], 0));
    return OH_OK;
}

int run_decode_step(oh_state* st, int rows, bool dump_probs) { return decode_step_all(st, rows, dump_probs ? K9_DUMP : K9_GREEDY); }

int run_decode_step_graph(oh_state* st, int rows, int k9_mode) {
    if (!st->use_graphs || st->profiling || k9_mode == K9_DUMP) return decode_step_all(st, rows, k9_mode);
    if (k9_mode == K9_TOPK) {  // its buffer must exist before the step is captured
        int rc = ensure(st->topk_dev, (size_t)st->max_batch * (3 * kTopK + 3) * 4);
        if (rc) return rc;
    }
    const bool with_k9 = k9_mode != K9_NONE;
    const int key = rows * 4 + k9_mode;
    auto it = st->step_graphs.find(key);
    if (it == st->step_graphs.end()) {
        // first use of this row count: one direct step (sets kernel attributes, faults in lazily allocated
        // buffers), capture on the second, replay from then on
        if (!st->warmed_rows.count(key)) {
            st->warmed_rows.insert(key);
            return decode_step_all(st, rows, k9_mode);
        }
        cudaGraph_t graph = nullptr;
        const long long launches_before = g_launch_count;
        OH_CUDA_CHECK(cudaStreamBeginCapture(st->stream, cudaStreamCaptureModeThreadLocal));
        int rc = decode_step_all(st, rows, k9_mode);
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

const int* clip_max_key(oh_state* st, int batch, int clip) { return clip_max_key(st->clip_meta, batch, clip); }

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
            hs.sslot[r] = r;
            hs.active[r] = 1;
            hs.pos[r] = n_past + i;
            hs.cur_token[r] = tokens[(size_t)r * n_tokens + i];
        }
        OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, seqlay::N_SCALARS * lay.rows * 4, cudaMemcpyHostToDevice, s));
        if ((rc = run_decode_step_graph(st, batch, K9_NONE))) return rc;
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

}  // namespace oh
