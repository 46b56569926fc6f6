/*
 * ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the transcription hot path that OpenHush reaches through
 * /root/reference/src/engine/whisper.rs:266 (`state.full(params, &samples)`).  The arithmetic of that
 * call lives in the crates.io dependency whisper-rs 0.13.2 -> whisper-rs-sys 0.11.1
 * (/root/reference/Cargo.toml:25, /root/reference/Cargo.lock:7693-7711), which vendors whisper.cpp/ggml
 * inside the crate tarball.  That source is NOT under /root/reference and not on this machine, and
 * there is no Rust toolchain, so this file restates the published whisper.cpp algorithm as specified
 * in SURVEY.md Appendix A (A.1 model file, A.3 log-mel, A.4 network, A.5 full() control flow,
 * A.6 logit filters, A.7 sampling, A.8 scoring/segments).
 *
 * PARITY UNPINNED: the reference's tests hold no golden vector, known-answer test or fixture for
 * this path (/root/reference/src/engine/whisper.rs:958-981 is #[ignore]d and
 * model-less).  The restatement is instead cross-checked against HuggingFace transformers' Whisper
 * (tests/golden/make_hf_golden.py) and against known-answer cases derived from the spec.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything in this directory.  The product library (openhush_b200/csrc) never links or loads it.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

namespace orc {

// ---- numerics knobs ------------------------------------------------------
enum GeluMode { GELU_TANH_F16LUT = 0, GELU_TANH = 1, GELU_ERF = 2 };
enum ActRound { ACT_NONE = 0, ACT_F16 = 1, ACT_BF16 = 2 };

struct Options {
    int gelu = GELU_TANH_F16LUT;  // ggml evaluates GELU through a 65536-entry f16 table (A.4)
    int act_round = ACT_F16;      // ggml rounds activations to f16 at every mat-mul with f16 weights (A.4)
    int n_threads = 0;            // 0 = OpenMP default
};

// ---- f16 helpers --------------------------------------------------------------------------------
static inline float f16_to_f32(uint16_t h) {
    _Float16 v;
    std::memcpy(&v, &h, 2);
    return (float)v;
}
static inline uint16_t f32_to_f16(float f) {
    _Float16 v = (_Float16)f;  // round-to-nearest-even
    uint16_t h;
    std::memcpy(&h, &v, 2);
    return h;
}
static inline float round_f16(float f) { return (float)(_Float16)f; }
static inline float round_bf16(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return f;
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}
static inline float round_act(float f, int mode) {
    return mode == ACT_F16 ? round_f16(f) : mode == ACT_BF16 ? round_bf16(f) : f;
}

// ---- model --------------------------------------------------------------------------------------
struct HParams {
    int32_t n_vocab = 0, n_audio_ctx = 0, n_audio_state = 0, n_audio_head = 0, n_audio_layer = 0;
    int32_t n_text_ctx = 0, n_text_state = 0, n_text_head = 0, n_text_layer = 0, n_mels = 0, ftype = 0;
};

// A weight matrix [rows][cols] kept in the file's storage type (f16 or f32), K-contiguous.
struct Mat {
    int rows = 0, cols = 0;
    std::vector<uint16_t> h;  // used when is_f16
    std::vector<float> f;     // used otherwise
    bool is_f16 = false;
};

struct Vocab {
    int n_vocab = 0;
    std::vector<std::string> id_to_token;
    std::map<std::string, int> token_to_id;
    int token_eot = 50256, token_sot = 50257;
    int token_translate = 50357, token_transcribe = 50358, token_solm = 50359, token_prev = 50360;
    int token_nosp = 50361, token_not = 50362, token_beg = 50363;
    bool is_multilingual() const { return n_vocab >= 51865; }
    int num_languages() const { return n_vocab - 51765 - (is_multilingual() ? 1 : 0); }
};

struct AttnW {
    Mat q, k, v, o;
    std::vector<float> qb, vb, ob;
};
struct EncLayer {
    std::vector<float> ln1_w, ln1_b, ln2_w, ln2_b;
    AttnW attn;
    Mat mlp0, mlp2;
    std::vector<float> mlp0_b, mlp2_b;
};
struct DecLayer {
    std::vector<float> ln1_w, ln1_b, lnx_w, lnx_b, ln2_w, ln2_b;
    AttnW attn, cross;
    Mat mlp0, mlp2;
    std::vector<float> mlp0_b, mlp2_b;
};

struct Model {
    HParams hp;
    Vocab vocab;
    std::vector<float> mel_filters;  // [n_mel][201]
    // encoder
    std::vector<float> enc_pos;  // [1500][d]
    Mat conv1, conv2;            // [d][3*n_mel] and [d][3*d], re-laid as [out][tap*C + c] at load
    std::vector<float> conv1_b, conv2_b;
    std::vector<EncLayer> enc;
    std::vector<float> ln_post_w, ln_post_b;
    // decoder
    std::vector<float> dec_pos;  // [448][d]
    Mat tok_emb;                 // [V][d]
    std::vector<DecLayer> dec;
    std::vector<float> ln_w, ln_b;
    std::vector<uint16_t> gelu_table;  // f16 -> f16, built at load time with this host's tanhf (ggml does the same)
};

Model* load_model(const char* path, std::string* err);

// ---- mel (A.3) ----------------------------------------------------------------------------------
struct Mel {
    int n_mel = 0, n_len = 0, n_len_org = 0;
    std::vector<float> data;  // [n_mel][n_len]
};
void log_mel_spectrogram(const Model& m, const float* pcm, int n, Mel& out, int n_threads);

// ---- network (A.4) ------------------------------------------------------------------------------
struct KVCache {  // one decoder's self-attention cache
    std::vector<float> k, v;  // [L][n_text_ctx][d]
};
struct State;

void encode(const Model& m, const Options& opt, const float* mel_window /*[n_mel][3000]*/, std::vector<float>& enc_out /*[1500][d]*/,
            std::vector<float>* conv_out /*optional [1500][d] after pos-emb*/);
void cross_kv(const Model& m, const Options& opt, const std::vector<float>& enc_out, std::vector<float>& ck, std::vector<float>& cv /*[L][1500][d]*/);
// one token at position n_past; appends to kv; writes logits[V] if logits != nullptr
void decode_token(const Model& m, const Options& opt, const std::vector<float>& ck, const std::vector<float>& cv, KVCache& kv, int token,
                  int n_past, float* logits, float* hidden_out /*optional [d] after final ln*/);

// ---- decoding (A.5-A.8) -------------------------------------------------------------------------
struct TokenData {
    int id = 0, tid = 0;
    float p = 0, plog = 0, pt = 0, ptsum = 0;
};
struct Segment {
    int64_t t0 = 0, t1 = 0;
    std::string text;
    std::vector<TokenData> tokens;
};

struct FullParams {
    int strategy = 0;  // 0 greedy, 1 beam
    int n_max_text_ctx = 16384;
    int translate = 0, no_context = 1, no_timestamps = 0, single_segment = 0;
    int max_tokens = 0;
    int suppress_blank = 1;
    float temperature = 0.0f, max_initial_ts = 1.0f, length_penalty = -1.0f;
    float temperature_inc = 0.2f, entropy_thold = 2.4f, logprob_thold = -1.0f, no_speech_thold = 0.6f;
    int best_of = 1, beam_size = 5;
    std::string language = "en";
};

struct Decoder {
    KVCache kv;
    std::vector<TokenData> tokens;  // tokens_cur
    int result_len = 0, seek_delta = 0;
    bool has_ts = false, failed = false, completed = false;
    double sum_logprobs_all = 0, sum_logprobs = 0, avg_logprobs = 0, entropy = 0, score = 0;
    std::vector<float> logits, logprobs, probs;
    std::mt19937 rng{0};  // seeded once, never reseeded (A.7)
};

struct State {
    const Model* model = nullptr;
    Options opt;
    Mel mel;
    std::vector<float> enc_out, ck, cv;
    std::vector<Decoder> decoders;
    std::vector<Segment> result_all;
    std::vector<int> prompt_past;
    int lang_id = 0;
    // statistics for benches / tests
    int n_encode = 0, n_decode_tokens = 0, n_fallbacks = 0;
    std::vector<int> window_seeks;       // seek of every encoded window
    std::vector<float> window_temps;     // temperature finally accepted per window
    std::vector<std::vector<int>> window_tokens;  // accepted decoder's raw token ids per window (segments drop the
                                                  // second timestamp of a pair, so they cannot be teacher-forced)
};

void process_logits(const Model& m, const FullParams& p, Decoder& dec, float temperature, bool is_initial);
TokenData sample_token(const Model& m, Decoder& dec, bool best);
std::vector<TokenData> sample_token_topk(const Model& m, Decoder& dec, int k);
int full(State& st, const FullParams& p, const float* pcm, int n);
int lang_id_from_code(const char* code);

}  // namespace orc
