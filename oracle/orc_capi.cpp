// ORACLE (test infrastructure) — extern "C" surface for ctypes (oracle/oracle.py) and bench.py's CPU legs.
// Shapes follow include/oh_whisper.h so a parity test reads the same on both sides.
#include <omp.h>

#include "../include/oh_whisper.h"
#include "orc_common.h"

using namespace orc;

#define ORC_API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_err;

ORC_API const char* orc_last_error() { return g_err.c_str(); }

ORC_API void* orc_model_load(const char* path) {
    std::string err;
    Model* m = load_model(path, &err);
    if (!m) g_err = err;
    return m;
}
ORC_API void orc_model_free(void* m) { delete (Model*)m; }

ORC_API void orc_model_hparams(void* mp, int32_t* out11) { std::memcpy(out11, &((Model*)mp)->hp, sizeof(int32_t) * 11); }

// eot, sot, translate, transcribe, solm, prev, nosp, not, beg, id(" ")
ORC_API void orc_model_special_tokens(void* mp, int32_t* out10) {
    const Vocab& v = ((Model*)mp)->vocab;
    int32_t t[10] = {v.token_eot, v.token_sot, v.token_translate, v.token_transcribe, v.token_solm,
                     v.token_prev, v.token_nosp, v.token_not, v.token_beg, -1};
    auto it = v.token_to_id.find(" ");
    if (it != v.token_to_id.end()) t[9] = it->second;
    std::memcpy(out10, t, sizeof t);
}
ORC_API const char* orc_token_to_str(void* mp, int id) {
    const Vocab& v = ((Model*)mp)->vocab;
    return (id >= 0 && id < v.n_vocab) ? v.id_to_token[id].c_str() : "";
}
ORC_API const float* orc_model_mel_filters(void* mp) { return ((Model*)mp)->mel_filters.data(); }
ORC_API const uint16_t* orc_model_gelu_table(void* mp) { return ((Model*)mp)->gelu_table.data(); }

ORC_API void* orc_state_new(void* mp) {
    State* s = new State();
    s->model = (Model*)mp;
    s->decoders.resize(1);
    return s;
}
ORC_API void orc_state_free(void* s) { delete (State*)s; }
ORC_API void orc_state_set_options(void* sp, int gelu, int act_round, int n_threads) {
    State* s = (State*)sp;
    s->opt.gelu = gelu;
    s->opt.act_round = act_round;
    s->opt.n_threads = n_threads;
    if (n_threads > 0) omp_set_num_threads(n_threads);
}

// ---- stages ---------------------------------------------------------------------------------------
ORC_API int orc_pcm_to_mel(void* sp, const float* pcm, int n) {
    State* s = (State*)sp;
    if (n <= 200) return -1;
    log_mel_spectrogram(*s->model, pcm, n, s->mel, s->opt.n_threads);
    return 0;
}
ORC_API void orc_mel_dims(void* sp, int* n_mel, int* n_len, int* n_len_org) {
    State* s = (State*)sp;
    *n_mel = s->mel.n_mel;
    *n_len = s->mel.n_len;
    *n_len_org = s->mel.n_len_org;
}
ORC_API const float* orc_mel_data(void* sp) { return ((State*)sp)->mel.data.data(); }

// mel_window: [n_mel][3000]; out_enc: [1500][d]; out_conv (optional): [1500][d] after conv stem + pos-emb
ORC_API int orc_encode(void* sp, const float* mel_window, float* out_enc, float* out_conv) {
    State* s = (State*)sp;
    std::vector<float> conv;
    encode(*s->model, s->opt, mel_window, s->enc_out, out_conv ? &conv : nullptr);
    cross_kv(*s->model, s->opt, s->enc_out, s->ck, s->cv);
    if (out_enc) std::memcpy(out_enc, s->enc_out.data(), s->enc_out.size() * 4);
    if (out_conv) std::memcpy(out_conv, conv.data(), conv.size() * 4);
    return 0;
}
ORC_API const float* orc_cross_k(void* sp) { return ((State*)sp)->ck.data(); }
ORC_API const float* orc_cross_v(void* sp) { return ((State*)sp)->cv.data(); }

// tokens at positions n_past .. n_past+n-1 of decoder 0; logits of the last one -> out_logits[V]
ORC_API int orc_decode(void* sp, const int* tokens, int n, int n_past, float* out_logits) {
    State* s = (State*)sp;
    const Model& m = *s->model;
    if (n_past + n > m.hp.n_text_ctx) return -7;
    Decoder& d = s->decoders[0];
    for (int i = 0; i < n; ++i)
        decode_token(m, s->opt, s->ck, s->cv, d.kv, tokens[i], n_past + i, i + 1 == n ? out_logits : nullptr, nullptr);
    return 0;
}

// known-answer access to A.6/A.7: logits[V] in; logits/logprobs/probs [V] out; returns the greedy TokenData
ORC_API int orc_process_logits(void* mp, const oh_full_params* cp, const int* tokens_cur, int n_tokens_cur, int has_ts, int seek_delta,
                               float temperature, float* logits_inout, float* logprobs_out, float* probs_out, int* out_id, int* out_tid,
                               float* out_p, float* out_plog, float* out_pt, float* out_ptsum) {
    const Model& m = *(Model*)mp;
    FullParams p;
    p.no_timestamps = cp->no_timestamps;
    p.suppress_blank = cp->suppress_blank;
    p.max_initial_ts = cp->max_initial_ts;
    Decoder d;
    for (int i = 0; i < n_tokens_cur; ++i) {
        TokenData t;
        t.id = tokens_cur[i];
        d.tokens.push_back(t);
    }
    d.has_ts = has_ts != 0;
    d.seek_delta = seek_delta;
    d.logits.assign(logits_inout, logits_inout + m.hp.n_vocab);
    process_logits(m, p, d, temperature, n_tokens_cur == 0);
    TokenData r = sample_token(m, d, true);
    std::memcpy(logits_inout, d.logits.data(), d.logits.size() * 4);
    if (logprobs_out) std::memcpy(logprobs_out, d.logprobs.data(), d.logprobs.size() * 4);
    if (probs_out) std::memcpy(probs_out, d.probs.data(), d.probs.size() * 4);
    *out_id = r.id;
    *out_tid = r.tid;
    *out_p = r.p;
    *out_plog = r.plog;
    *out_pt = r.pt;
    *out_ptsum = r.ptsum;
    return 0;
}

// ---- full() ---------------------------------------------------------------------------------------
ORC_API int orc_full(void* sp, const oh_full_params* cp, const float* pcm, int n) {
    State* s = (State*)sp;
    FullParams p;
    p.strategy = cp->strategy;
    p.n_max_text_ctx = cp->n_max_text_ctx;
    p.translate = cp->translate;
    p.no_context = cp->no_context;
    p.no_timestamps = cp->no_timestamps;
    p.single_segment = cp->single_segment;
    p.max_tokens = cp->max_tokens;
    p.suppress_blank = cp->suppress_blank;
    p.temperature = cp->temperature;
    p.max_initial_ts = cp->max_initial_ts;
    p.length_penalty = cp->length_penalty;
    p.temperature_inc = cp->temperature_inc;
    p.entropy_thold = cp->entropy_thold;
    p.logprob_thold = cp->logprob_thold;
    p.no_speech_thold = cp->no_speech_thold;
    p.best_of = cp->greedy_best_of;
    p.beam_size = cp->beam_size;
    p.language = (cp->language && std::strcmp(cp->language, "auto") != 0) ? cp->language : "en";
    return full(*s, p, pcm, n);
}
ORC_API int orc_full_n_segments(void* sp) { return (int)((State*)sp)->result_all.size(); }
ORC_API const char* orc_full_get_segment_text(void* sp, int i) { return ((State*)sp)->result_all[i].text.c_str(); }
ORC_API int64_t orc_full_get_segment_t0(void* sp, int i) { return ((State*)sp)->result_all[i].t0; }
ORC_API int64_t orc_full_get_segment_t1(void* sp, int i) { return ((State*)sp)->result_all[i].t1; }
ORC_API int orc_full_n_tokens(void* sp, int i) { return (int)((State*)sp)->result_all[i].tokens.size(); }
ORC_API int orc_full_get_token_id(void* sp, int i, int j) { return ((State*)sp)->result_all[i].tokens[j].id; }
ORC_API int orc_full_get_token_tid(void* sp, int i, int j) { return ((State*)sp)->result_all[i].tokens[j].tid; }
ORC_API float orc_full_get_token_p(void* sp, int i, int j) { return ((State*)sp)->result_all[i].tokens[j].p; }
ORC_API float orc_full_get_token_plog(void* sp, int i, int j) { return ((State*)sp)->result_all[i].tokens[j].plog; }
ORC_API int orc_full_lang_id(void* sp) { return ((State*)sp)->lang_id; }
// n_encode, n_decode_tokens, n_fallbacks, n_windows
ORC_API void orc_full_stats(void* sp, int* out4) {
    State* s = (State*)sp;
    out4[0] = s->n_encode;
    out4[1] = s->n_decode_tokens;
    out4[2] = s->n_fallbacks;
    out4[3] = (int)s->window_seeks.size();
}
ORC_API int orc_full_window_seek(void* sp, int w) { return ((State*)sp)->window_seeks[w]; }
ORC_API float orc_full_window_temperature(void* sp, int w) { return ((State*)sp)->window_temps[w]; }
ORC_API int orc_full_window_n_tokens(void* sp, int w) { return (int)((State*)sp)->window_tokens[w].size(); }
ORC_API int orc_full_window_token(void* sp, int w, int i) { return ((State*)sp)->window_tokens[w][i]; }

// models the bf16 compute mode of the product, which re-rounds the f16-stored weights to bf16 at load time
// (a bf16-rounded f16 value is still exactly representable in f16, so the storage type does not change)
static void round_mat_bf16(Mat& m) {
    if (m.is_f16) for (auto& h : m.h) h = f32_to_f16(round_bf16(f16_to_f32(h)));
    else for (auto& f : m.f) f = round_bf16(f);
}
ORC_API void orc_model_round_weights_bf16(void* mp) {
    Model* m = (Model*)mp;
    round_mat_bf16(m->conv1);
    round_mat_bf16(m->conv2);
    round_mat_bf16(m->tok_emb);
    auto attn = [](AttnW& a) { round_mat_bf16(a.q); round_mat_bf16(a.k); round_mat_bf16(a.v); round_mat_bf16(a.o); };
    for (auto& L : m->enc) { attn(L.attn); round_mat_bf16(L.mlp0); round_mat_bf16(L.mlp2); }
    for (auto& L : m->dec) { attn(L.attn); attn(L.cross); round_mat_bf16(L.mlp0); round_mat_bf16(L.mlp2); }
}

ORC_API int orc_lang_id(const char* code) { return lang_id_from_code(code); }
