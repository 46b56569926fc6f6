// Launch wrappers of every CUDA kernel in libohwhisper.  All return OH_OK or an OH_ERR_* code.
#pragma once
#include "common.cuh"

namespace oh {

// ---- K1/K2: log-mel (mel.cu) ------------------------------------------------------------------------
struct MelArgs {
    const float* const* pcm;  // [batch] device pointers to 16 kHz mono f32 PCM
    const int* n_samples;     // [batch]
    const int* n_active;      // [batch] frames that see real samples: min((n+200)/160 + 1, n_len)
    float* raw;               // [batch][n_mel][ld] un-normalised log10 mel
    int ld;
    int* max_key;             // [batch] order-preserving int key of the clip-wide max
    int n_mel;
    const float2* twiddle;    // [400] exp(-2 pi i k/400)
    const float* hann;        // [400] periodic Hann
    const int *filt_start, *filt_count, *filt_off;
    const float* filt_w;
    int n_filt_w;
};
size_t mel_smem_bytes();
int launch_mel_frames(const MelArgs& a, int batch, int max_active_frames, cudaStream_t st);
int launch_mel_normalize(const float* raw, int ld, int n_mel, int n_active, int n_len, const int* max_key, float* out, cudaStream_t st);

// one encoder-input window: frames [seek, seek + 3000) of a clip's un-normalised log-mel (device array, one per window of a pass)
struct WindowDesc {
    const float* raw;    // the clip's raw log10 mel, [n_mel][ld]
    const int* max_key;  // the clip-wide max (order-preserving int key), written by the mel kernel
    int ld, n_active, n_len, seek;
};
// act_type: OH_COMPUTE_F32 / F16 / BF16 selects the element type of X ([windows][3002][n_mel])
int launch_window_gather(const WindowDesc* d_windows, int n_mel, int n_windows, void* X, int act_type, cudaStream_t st);
int launch_window_from_mel(const float* mel, int n_mel, int n_slots, void* X, int act_type, cudaStream_t st);

// ---- K3/K4: dense contractions (gemm_simt.cu, gemm_tc.cu) --------------------------------------------
// C(b,i,n) = [gelu]( sum_k A(b,i,k) W(n,k) + bias[n] ) + R(b,i,n)
//   A(b,i,k) = A[b*a_bstride + i*lda + k]   (overlapping rows allowed: conv as mat-mul)
//   C(b,i,n) = C[b*c_bstride + i*ldc + n],  R likewise with r_bstride / ldr (r_bstride = 0 broadcasts, e.g. pos-emb)
enum { GELU_NONE = 0, GELU_LUT = 1, GELU_FAST = 2, GELU_TANH = 3 };
enum { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };
struct GemmArgs {
    const void* A = nullptr;
    int a_type = DT_F32;
    long long a_bstride = 0;
    int lda = 0;
    int batch = 1, rows = 0;  // M = batch * rows
    const void* W = nullptr;  // [N][K], 16-bit
    int w_type = DT_F16;
    int N = 0, K = 0;
    const float* bias = nullptr;
    void* C = nullptr;
    int c_type = DT_F32;
    long long c_bstride = 0;
    int ldc = 0;
    const float* R = nullptr;
    long long r_bstride = 0;
    int ldr = 0;
    int gelu = GELU_NONE;
    const uint16_t* lut = nullptr;
    int round_type = DT_F32;  // 16-bit pipelines: GELU input and output are rounded to this type (ggml's f16 table semantics)
    // conv-as-mat-mul hint for the TMA path: K = n_taps * tap_k and lda < K (rows overlap).  The linear formula
    // above is unchanged; the tensor-core kernel loads tap t through its own tensor map based at A + t*tap_k.
    int n_taps = 1;
    int tap_k = 0;
    // head-split store (cross-attention K/V cache, contiguous per head): when c_hsplit > 0 column n goes to
    //   C[b*c_bstride + (n / c_hsplit) * c_hstride + i*ldc + (n % c_hsplit)]      (c_hsplit must be a multiple of 32)
    int c_hsplit = 0;
    long long c_hstride = 0;
    // scattered batch store (cross K/V of an encoder pass into non-adjacent slots): batch b writes block c_bidx[b] (device array)
    // instead of block b; nullptr = identity.  Honoured by the tcgen05 and SIMT kernels (the decode-time kernels take batch = 1).
    const int* c_bidx = nullptr;
};
// one-off numerical self-check of the tcgen05 path against the SIMT kernel (called at context creation)
int gemm_tc_selftest(cudaStream_t st, float* max_rel_err);
// decode-time mat-muls on tcgen05 with cluster split-K and a DSMEM reduction (gemm_splitk.cu); the default for M <= 256
int launch_gemm_splitk(const GemmArgs& g, cudaStream_t st);
int gemm_splitk_selftest(cudaStream_t st, float* max_rel_err);
int launch_gemm_simt(const GemmArgs& g, cudaStream_t st);
// tcgen05/TMEM/TMA path; requires 16-bit A of the same type as W.  Returns OH_ERR_UNSUPPORTED for shapes it does not take.
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st);
bool gemm_tc_available();

// ---- K5: LayerNorm (layers.cu) ------------------------------------------------------------------------
int launch_layernorm(const float* x, int rows, int d, const float* w, const float* b, void* y, int y_type, cudaStream_t st);

// ---- K6: encoder attention (attention.cu) --------------------------------------------------------------
// qkv: [B*T][3d] (q|k|v per token) of type t; out: [B*T][d] of type t.  Non-causal, softmax(q k^T / 8) v per head.
int launch_encoder_attention_simple(const void* qkv, void* out, int type, int B, int T, int d, cudaStream_t st);
int launch_encoder_attention_mma(const void* qkv, void* out, int type, int B, int T, int d, cudaStream_t st);
// tcgen05 / TMEM / TMA flash attention (attention_tc.cu); 16-bit modes, returns OH_ERR_UNSUPPORTED otherwise
int launch_encoder_attention_tc(const void* qkv, void* out, int type, int B, int T, int d, cudaStream_t st);
bool encoder_attention_tc_available();
int encoder_attention_tc_selftest(cudaStream_t st, float* max_rel_err);

// ---- K7/K8/K9: decoder step (decoder.cu) ----------------------------------------------------------------
constexpr int kMaxPrompt = 232;   // [prev] + 224 past tokens + sot/lang/task/notimestamps, rounded up
constexpr int kMaxGen = 224;      // n_max = 448/2 - 4 = 220 sampled tokens per window, rounded up

struct DecodeSeqState {  // struct of arrays, one entry per batch row; lives in device memory
    int* slot;        // cross-attention K/V slot of this row (the encoded window it decodes; beams of a window share one)
    int* sslot;       // self-attention K/V cache line of this row (a beam inherits its parent's by index)
    int* active;      // 1 while the row still decodes
    int* pos;         // tokens already in the self-attention cache (= n_past of the next step)
    int* cur_token;   // input token of the next step
    int* prompt_len;
    int* prompt;      // [rows][kMaxPrompt]
    int* n_gen;       // sampled tokens so far (i of A.5)
    int* has_ts;
    int* seek_delta;
    int* result_len;
    int* failed;
    int* completed;
    int* seek;
    int* seek_end;
    int* tok_id;      // [rows][kMaxGen]
    int* tok_tid;
    float* tok_p;
    float* tok_plog;
    float* tok_pt;
    float* tok_ptsum;
    int* n_active_total;  // single counter
};

struct DecodeRules {  // A.5/A.6 constants for one decode pass
    int n_vocab, eot, sot, translate, transcribe, solm, prev, nosp, tok_not, beg, blank, n_lang;
    int no_timestamps, suppress_blank, single_segment, max_tokens, n_max;
    int tid0_initial;  // round(max_initial_ts / 0.02), or -1 when the rule is off
    float temperature;
};

int launch_embed(const void* tok_emb, int w_type, const float* pos_emb, const DecodeSeqState& s, int rows, int d, float* x, cudaStream_t st);
// appends this step's k,v (from qkv [rows][3d]) to the self cache and attends over positions [0, pos]
int launch_self_attention_decode(const void* qkv, void* kcache, void* vcache, int type, const DecodeSeqState& s, int rows, int d, int n_ctx,
                                 void* out, cudaStream_t st);
// q: [rows][d]; cross cache of a slot: [T][2d] (k | v per encoder position); out [rows][d]
int launch_cross_attention_decode(const void* q, const void* xkv, int type, const DecodeSeqState& s, int rows, int d, int T, void* out,
                                  cudaStream_t st);
// K9 in beam mode: instead of picking, a row reports its kTopK best ids of the filtered distribution (best first; ties: lower
// id) with logprob and prob, and the timestamp statistics recorded with every token (A.7): tid, pt, ptsum.
constexpr int kTopK = 8;  // beam_size <= kTopK
struct TopKOut {
    int* ids = nullptr;      // [rows][kTopK]; nullptr: not in top-k mode
    float* logprobs = nullptr;
    float* probs = nullptr;
    int* tid = nullptr;      // [rows]
    float* pt = nullptr;
    float* ptsum = nullptr;
};
// K9: logit filters + log-softmax + timestamp-mass rule + greedy pick + A.5 state update, one CTA per row.
// With dump_probs != nullptr the row's logprobs/probs are written ([rows][ld]) and nothing is sampled (t > 0 path).
// `r` points to DEVICE memory so that a captured graph of the step survives a change of rules (temperature ladder).
// With topk.ids != nullptr the row's top-k candidates are written and nothing is sampled (beam search; the host ranks).
int launch_logits_step(const float* logits, int ld, const DecodeSeqState& s, const DecodeRules* r, int rows, float* dump_logprobs,
                       float* dump_probs, cudaStream_t st, const TopKOut& topk = TopKOut());
// self-attention cache lines of beam search: copy positions [0, n_pos) of line src -> dst for every pair, all layers, K and V,
// in one launch.  pairs: device array of n_pairs (src, dst, n_pos, -) int4; caches are [layer][lines][heads][n_ctx][64].
int launch_copy_self_lines(void* kcache, void* vcache, int esz, const int4* pairs, int n_pairs, int n_layers, int lines, int heads, int n_ctx,
                           cudaStream_t st);

}  // namespace oh
