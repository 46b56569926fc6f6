// Context / state objects behind the C ABI and the batched full() driver.
#pragma once
#include <map>
#include <set>
#include <memory>
#include <random>
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

}  // namespace oh

struct oh_state {
    oh_context* ctx = nullptr;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t stream2 = nullptr;  // second row group of the interleaved decode step
    cudaEvent_t ev_dual[3] = {nullptr, nullptr, nullptr};
    int dual_min_rows = 1 << 30;     // OH_DUAL=n: batches of >= n rows decode as two interleaved row groups (measured slower on B200: off)
    int max_batch = 1;   // KV-cache slots
    int enc_chunk = 1;   // windows per encoder pass
    int esz = 2;         // bytes per activation element

    // --- mel
    oh::DevBuf pcm_stage;          // device copy of host PCM (concatenated)
    oh::DevBuf mel_raw;            // [clips][n_mel][mel_ld]
    int mel_ld = 0, mel_clips = 0;
    oh::DevBuf clip_meta;          // device: pcm ptrs | n_samples | n_active | n_len | max_key | clip_of_slot | seek_of_slot
    // --- encoder activations (enc_chunk windows)
    oh::DevBuf xin, x1, x, h, qkv, hid, enc_f32;
    // --- caches (max_batch slots)
    oh::DevBuf xkv;                // [L][slots][1500][2d]
    oh::DevBuf kself, vself;       // [L][slots][448][d]
    // --- decoder activations (max_batch rows)
    oh::DevBuf dx, dh, dqkv, datt, dq, dhid, logits, dump_lp, dump_p;
    int ld_logits = 0;
    oh::DevBuf seq_dev;            // DecodeSeqState arrays
    oh::DevBuf rules_dev;
    oh::DecodeSeqState seq{};
    std::vector<int> seq_host_i;   // staging
    void* pinned = nullptr;        // small pinned scratch
    size_t pinned_bytes = 0;
    std::map<int, cudaGraphExec_t> step_graphs;  // keyed by rows
    std::set<int> warmed_rows;
    bool use_graphs = true;
    int gemm_mode = 0;             // 0 auto (tcgen05 when 16-bit), 1 force SIMT
    int attn_mode = 0;             // 0 auto, 1 force the simple kernel
    int skinny_mode = 2;           // decode mat-muls: 2 cluster split-K tcgen05 kernel (default), 0 register-streaming mma.sync kernel, 1 persistent tcgen05 kernel
    bool use_pdl = true;           // OH_PDL=0: decode-step kernels are launched fully serialised
    bool gelu_lut_f32 = false;     // OH_GELU=lut: f32 mode evaluates GELU through ggml's f16 table instead of f32 tanh
    oh::Profiler* prof = nullptr;
    bool profiling = false;

    // --- results
    std::vector<oh::Item> items;
    std::vector<std::mt19937> rngs;  // one per slot, seeded 0 at state creation, never reseeded (A.7)
    oh_timings timings{};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string scratch_text;
};

namespace oh {

int state_alloc(oh_state* st);
void state_release(oh_state* st);

// stages
int run_mel(oh_state* st, const float* const* pcm, const int* n, int batch, bool pcm_on_device);
int run_encoder(oh_state* st, int n_windows, int slot0, float* d_enc_f32_out /* nullable, [n_windows][1500][d] */,
                const int* slot_list = nullptr /* per-window KV slot; default slot0 + w */);
int run_decode_step(oh_state* st, int rows, bool dump_probs);
int run_full(oh_state* st, const oh_full_params& p, const float* const* pcm, const int* n, int batch, bool pcm_on_device);

int ensure(DevBuf& b, size_t bytes);
const int* clip_max_key(oh_state* st, int batch, int clip);
int decode_tokens_stage(oh_state* st, const int* tokens, int n_tokens, int n_past, int batch, float* out_logits);

}  // namespace oh
