// Context / state objects behind the C ABI, the stage runners and the continuous-batching full() scheduler.
#pragma once
#include <chrono>
#include <deque>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "../../include/oh_whisper.h"
#include "kernels.h"
#include "model.h"

struct oh_context {
    oh::Model* model = nullptr;
    oh_context_params params{};
    float2* d_twiddle = nullptr;  // [400]
    float* d_hann = nullptr;      // [400]
    int act_type = oh::DT_F16;    // element type of activations / KV caches (DT_F32 in the parity mode)
    int n_states = 0;
    bool freed = false;
};

namespace oh {

struct TokenData {
    int id = 0, tid = 0;
    float p = 0, plog = 0, pt = 0, ptsum = 0;
};
struct Segment {
    int64_t t0 = 0, t1 = 0;
    std::string text;
    std::vector<TokenData> tokens;
};

// one independent transcription job inside a batched call (== one oh_full call)
struct Item {
    int n_samples = 0;
    int n_len = 0, n_len_org = 0, n_active = 0;
    int seek = 0, seek_end = 0;
    bool done = false;
    int rc = 0;
    int lang_id = 0;
    std::vector<int> prompt_past;
    std::vector<Segment> segments;
};

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

// ---- the scheduler's objects (full.cu) ----------------------------------------------------------------------------
// A submission is one oh_full / oh_full_batch / oh_stream_submit call: its items, their log-mel, and the decode policy
// derived from its params.  Submissions stay live until they are waited for; windows of different live submissions share
// the decoder rows, which is what keeps the rows full across calls.
struct Submission {
    int ticket = 0;
    oh_full_params p{};
    std::string language;  // owns p.language
    std::vector<float> temperatures;
    std::vector<int> prompt_init;
    bool no_timestamps = false;
    std::vector<Item> items;
    std::vector<std::mt19937>* rngs = nullptr;  // per item (A.7); the blocking calls use the state's persistent generators
    std::vector<std::mt19937> own_rngs;
    int n_done = 0;
    DevBuf pcm_stage, mel_raw, clip_meta;  // clip_meta: [max_key int per clip]
    int mel_ld = 0;
    cudaEvent_t ev_mel = nullptr;  // recorded on the front-end stream after this submission's log-mel
    std::chrono::steady_clock::time_point t_begin;
    bool finished() const { return n_done >= (int)items.size(); }
};

struct WinRef {  // a window of an item: waiting for the encoder, in an encoder pass, or encoded and waiting for a row
    Submission* sub = nullptr;
    int item = -1;
    int slot = -1;
};
struct EncoderPass {
    cudaEvent_t done = nullptr;
    std::vector<WinRef> wins;
};
struct RowInfo {
    Submission* sub = nullptr;
    int item = -1, slot = -1;
    bool armed = false;
    std::vector<int> prompt;
};
struct ParkedRow {  // a window whose t = 0 result failed the A.5 checks: waits (with its cross K/V slot) for the temperature ladder
    Submission* sub = nullptr;
    int item = -1, slot = -1;
    std::vector<int> prompt;
};

// beam search (full.cu): the beam_size decoders of one window and the row group that hosts them
struct BeamDec {
    std::vector<TokenData> tokens;
    int seek_delta = 100 * 30, result_len = 0, line = 0;
    bool has_ts = false, failed = false, completed = false;
    double sum_logprobs_all = 0.0;
    std::vector<TokenData> cand;  // this decoder's top candidates from the last step, best first
};
struct BeamCand {
    int decoder_idx, seek_delta, result_len;
    bool has_ts;
    std::vector<TokenData> tokens;
    double sum_logprobs_all;
};
struct BeamGroup {
    bool busy = false, in_prompt = true;
    Submission* sub = nullptr;
    int item = -1, xslot = -1;
    std::vector<int> prompt;
    int fed = 0;  // prompt tokens that have gone through the decoder
    int i = 0;    // index of the token being chosen (A.5's loop variable)
    std::vector<BeamDec> dec;
};


struct Sched {
    std::deque<std::unique_ptr<Submission>> subs;  // live submissions in ticket order
    int next_ticket = 1;
    std::vector<int> free_slots;
    std::vector<RowInfo> rows;
    int n_armed = 0;
    std::deque<WinRef> need_window;  // windows waiting for an encoder pass (continuations first)
    std::deque<EncoderPass> passes;  // in flight on the encoder stream, oldest first
    std::deque<WinRef> ready;        // encoded, waiting for a free row
    std::vector<ParkedRow> parked;
    std::vector<cudaEvent_t> event_pool;
    std::vector<DevBuf> buf_pool;    // device buffers of finished submissions, reused by the next ones
    bool rules_valid = false;
    DecodeRules rules{};
    bool mirror_dirty = false;       // the pinned host mirror of the row scalars has changes that are not on the device yet
    int desc_cursor = 0;             // ring position in the pinned window-descriptor area
    std::vector<BeamGroup> beam_groups;   // beam search: one group of beam_size rows per window in flight
    std::vector<int4> beam_kv_copies;     // cache-line copies queued for the next step
    int beam_nb = 0;
};

}  // namespace oh

struct oh_state {
    oh_context* ctx = nullptr;
    cudaStream_t own_stream = nullptr, stream = nullptr;  // decode stream (the caller may substitute its own)
    cudaStream_t enc_stream = nullptr;                     // encoder passes run here, concurrently with decode steps
    cudaStream_t fe_stream = nullptr;                      // front end: PCM upload + log-mel of a new submission
    void* green_ctx = nullptr;                             // CUgreenCtx of the encoder's SM partition (OH_ENC_SMS), or null
    int enc_sms = 0;                                       // SMs of that partition (0: the encoder stream may use every SM)
    cudaStream_t stream2 = nullptr;  // second row group of the interleaved decode step
    cudaEvent_t ev_dual[3] = {nullptr, nullptr, nullptr};
    int dual_min_rows = 1 << 30;     // OH_DUAL=n: batches of >= n rows decode as two interleaved row groups (measured slower on B200: off)
    int max_batch = 1;   // decoder rows resident on the GPU (oh_context_params.max_batch)
    int n_slots = 1;     // cross-attention K/V slots: rows + look-ahead (windows encoded before a row is free for them)
    int enc_chunk = 1;   // windows per encoder pass
    int esz = 2;         // bytes per activation element
    bool async_encoder = true;  // OH_ASYNC_ENC=0: encoder passes are issued on the decode stream (serialised, round-1 behaviour)
    int check_steps = 8;        // decode steps between two host look-ins
    int enc_compact_bytes = 0;  // OH_ENC_COMPACT=<KB>: encoder mat-muls on the encoder stream use 128-wide tiles and this much ring (co-residency experiment)
    int starve_rows = 8;        // empty rows at which the decode steps pause for the encoder pass in flight (OH_STARVE)

    // --- stage API mel (oh_pcm_to_mel_*): its own small submission-like block
    oh::DevBuf pcm_stage, mel_raw, clip_meta;
    int mel_ld = 0, mel_clips = 0;
    // --- encoder activations (enc_chunk windows)
    oh::DevBuf xin, x1, x, h, qkv, hid, enc_f32;
    oh::DevBuf win_desc;           // device WindowDesc ring
    oh::WindowDesc* win_desc_host = nullptr;  // pinned mirror of the ring
    int win_desc_cap = 0;
    oh::DevBuf slot_ring;          // device copies of the passes' slot lists (same ring discipline)
    int* slot_ring_host = nullptr;
    int slot_cursor = 0;
    // --- caches
    oh::DevBuf xkv;                // [L][n_slots][2][heads][1500][64]
    oh::DevBuf kself, vself;       // [L][rows][heads][448][64]
    // --- decoder activations (max_batch rows)
    oh::DevBuf dx, dh, dqkv, datt, dq, dhid, logits, dump_lp, dump_p, topk_dev;
    int ld_logits = 0;
    oh::DevBuf resample_plan, resample_in, resample_out;  // pre-engine DSP (resample.cu): chunk plan, staging of the host entry point
    oh::DevBuf level_pcm, level_env, level_aux;           // level processing (preprocess.cu): host-entry staging, envelope scratch, n / gain / rms
    oh::DevBuf beam_pairs;         // beam search: (src line, dst line, positions) of the step's cache-line copies
    oh::DevBuf seq_dev;            // DecodeSeqState arrays
    oh::DevBuf rules_dev;
    oh::DecodeSeqState seq{};
    void* pinned = nullptr;        // pinned host mirror of the DecodeSeqState block + small scratch
    size_t pinned_bytes = 0;
    std::map<int, cudaGraphExec_t> step_graphs;  // keyed by rows
    std::set<int> warmed_rows;
    bool use_graphs = true;
    int gemm_mode = 0;             // 0 auto (tcgen05 when 16-bit), 1 force SIMT
    int attn_mode = 0;             // 0 auto, 1 force the simple kernel
    bool splitk_decode = true;     // decode mat-muls (M <= 256): cluster split-K tcgen05 kernel; false (OH_DECODE_GEMM=tc): the persistent tcgen05 kernel
    bool use_pdl = true;           // OH_PDL=0: decode-step kernels are launched fully serialised
    bool gelu_lut_f32 = false;     // OH_GELU=lut: f32 mode evaluates GELU through ggml's f16 table instead of f32 tanh
    oh::Profiler* prof = nullptr;
    bool profiling = false;

    oh::Sched sched;
    FILE* trace = nullptr;           // OH_SCHED_TRACE=<file>: scheduler occupancy log (diagnostics)
    long long trace_steps = 0;

    // --- results of the last waited submission
    std::vector<oh::Item> items;
    std::vector<std::mt19937> rngs;  // one per item index of the blocking calls, seeded 0 on first use, never reseeded (A.7)
    oh_timings timings{};            // accumulates since the previous wait returned
    oh_timings timings_out{};        // what oh_get_timings reports: the activity up to the last wait
    long long launches_mark = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string scratch_text;
};

namespace oh {

int state_alloc(oh_state* st);
void state_release(oh_state* st);

// stages
// stage-API log-mel into the state's own block (oh_pcm_to_mel_*); items of the call land in st->items
int run_mel(oh_state* st, const float* const* pcm, const int* n, int batch, bool pcm_on_device);
// encoder pass over the windows staged in st->xin; window w's cross K/V goes to slot slot_list[w] (default slot0 + w)
int run_encoder(oh_state* st, int n_windows, int slot0, float* d_enc_f32_out /* nullable, [n_windows][1500][d] */,
                const int* slot_list = nullptr, cudaStream_t stream = nullptr /* default: the decode stream */);
int run_decode_step(oh_state* st, int rows, bool dump_probs);
enum { K9_NONE = 0, K9_GREEDY = 1, K9_DUMP = 2, K9_TOPK = 3 };  // what follows the logit projection of a step
int run_decode_step_graph(oh_state* st, int rows, int k9_mode = K9_GREEDY);

// the full() driver (full.cu)
// persistent_rng: sampled (t > 0) tokens draw from the state's generators, which continue across calls (oh_full: upstream's
// behaviour); otherwise every item draws from a generator seeded 0 for this call (batched / streamed jobs: each equals oh_full on a fresh state)
int run_full(oh_state* st, const oh_full_params& p, const float* const* pcm, const int* n, int batch, bool pcm_on_device, bool persistent_rng);
int sched_submit(oh_state* st, const oh_full_params& p, const float* const* pcm, const int* n, int batch, bool pcm_on_device, bool persistent_rngs,
                 int* ticket, bool carry_context = false);
int sched_wait(oh_state* st, int ticket);
int sched_pending(const oh_state* st);
void sched_abort(oh_state* st);    // drop every live submission, reset rows and slots
void sched_release(oh_state* st);

// pre-engine DSP (resample.cu): interleaved PCM -> mono 16 kHz f32, the reference's load_wav_file steps on the device
long long resample_out_len(int sample_rate, long long n_frames, bool pad_min_duration);
int resample_to_16k_device(oh_state* st, const void* d_pcm, int format, int bits, int channels, int sample_rate, long long n_frames, float* d_out,
                           long long* out_n, cudaStream_t stream);

int chunk_rms_db_device(const float* d_pcm, int chunk, int n_chunks, float* d_out, cudaStream_t stream);
int preprocess_levels_device(float* d_pcm, long long stride, const int* d_n, int batch, int max_n, const oh_preprocess_params& p, float* d_work, float* d_gain,
                             float* d_rms, cudaStream_t stream);

int ensure(DevBuf& b, size_t bytes);
void release(DevBuf& b);
const int* clip_max_key(oh_state* st, int batch, int clip);  // stage-API block
const int* clip_max_key(const DevBuf& clip_meta, int batch, int clip);
int decode_tokens_stage(oh_state* st, const int* tokens, int n_tokens, int n_past, int batch, float* out_logits);

// mel of `batch` clips into caller-provided buffers on `stream` (shared by the stage API and the scheduler)
struct MelPlan {
    std::vector<size_t> pcm_off;  // float offset of each clip in the staging buffer
    size_t stage_floats = 0;
    int max_active = 0, ld = 0;
};
MelPlan plan_mel(std::vector<Item>& items, const int* n, int batch);
int launch_mel(oh_state* st, const MelPlan& plan, const std::vector<Item>& items, const float* const* pcm, bool pcm_on_device, DevBuf& pcm_stage,
               DevBuf& mel_raw, DevBuf& clip_meta, cudaStream_t stream, long long* h2d_bytes);

// DecodeSeqState block layout helpers (engine.cu)
namespace seqlay {
enum { SLOT, SSLOT, ACTIVE, POS, CUR, PLEN, NGEN, HAS_TS, SDELTA, RLEN, FAILED, COMPLETED, SEEK, SEEKEND, N_SCALARS };
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
DecodeSeqState seq_view(int32_t* base, const SeqLayout& L);
DecodeRules make_rules(const Model& m, const oh_full_params& p, bool no_timestamps, float temperature);

}  // namespace oh
