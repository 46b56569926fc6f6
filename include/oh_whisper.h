/*
 * oh_whisper.h — C ABI of the B200-native transcription hot path (libohwhisper.so).
 *
 * This is the drop-in boundary of SURVEY.md §8(b).  OpenHush has no plugin ABI; its only FFI seam on
 * the hot path is whisper-rs -> whisper.h.  Every entry point below replaces exactly one whisper.h
 * symbol that `src/engine/whisper.rs` reaches through whisper-rs 0.13.2 (Cargo.toml:25), keeps that
 * symbol's argument meaning, ownership and error codes, and is what a Rust `extern "C"` block (see
 * INTEGRATION.md) binds instead of whisper-rs-sys.  Plain pointers and sizes only; no exceptions or
 * panics cross this boundary (int codes / NULL).
 *
 * Ownership: the caller owns PCM buffers and params (borrowed for the duration of a call only); the
 * library owns contexts, states and every pointer it returns (valid until the next oh_full* call on the
 * same state, or until the state is freed).
 * Threading: a state is used by one thread at a time but may be created on one thread and used on
 * another (the daemon moves the engine into its worker thread: src/daemon.rs:1803-1805,
 * src/queue/worker.rs:100-103); every entry point re-selects its CUDA device.  Different states /
 * contexts may run concurrently (one per GPU).
 * There is no CPU fallback: if no sm_100 device is usable, oh_init_from_file returns NULL.
 */
#ifndef OH_WHISPER_H
#define OH_WHISPER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OH_API __attribute__((visibility("default")))

typedef struct oh_context oh_context; /* model weights on one GPU     (whisper_context) */
typedef struct oh_state oh_state;     /* buffers + results of a call  (whisper_state)   */

/* return codes of oh_full*: same values whisper-rs maps: 0 ok, -1 mel failed,
 * -6/7 encode failed, -7/8 decode failed; other non-zero = generic error */
enum {
    OH_OK = 0,
    OH_ERR_MEL = -1,
    OH_ERR_INVALID = -2,
    OH_ERR_ENCODE = -6,
    OH_ERR_DECODE = -7,
    OH_ERR_CUDA = -100,
    OH_ERR_UNSUPPORTED = -101
};

/* arithmetic of the dense contractions */
enum oh_compute_type {
    OH_COMPUTE_F32 = 0, /* f32 activations, f32 accumulation on CUDA cores: the parity mode          */
    OH_COMPUTE_F16 = 1, /* f16 operands on tcgen05 tensor cores, f32 accumulation (ggml's own choice: */
                        /*   f16 weights, activations rounded to f16 at every mat-mul)                */
    OH_COMPUTE_BF16 = 2 /* bf16 operands on tcgen05 tensor cores, f32 accumulation                    */
};

/* replaces whisper_context_params / WhisperContextParameters (src/engine/whisper.rs:156-157) */
typedef struct oh_context_params {
    int use_gpu;      /* must be non-zero: there is no CPU path (use_gpu=false -> NULL)               */
    int flash_attn;   /* accepted for ABI shape; attention is always fused                            */
    int gpu_device;   /* CUDA ordinal; the multi-GPU workers set this (config.rs:921-929 gpu.devices) */
    int compute_type; /* enum oh_compute_type; default OH_COMPUTE_F16                                 */
    int max_batch;    /* decoder rows / KV-cache slots resident on the GPU (default 1); oh_full_batch may be   */
                      /* given more items than this: they queue for a free slot                                */
} oh_context_params;

/* BEAM_SEARCH (never selected by the reference, whisper.rs:243): beam_size decoders per window; needs
 * oh_context_params.max_batch >= beam_size and beam_size <= 8.  The resident rows form max_batch / beam_size groups: that
 * many windows decode side by side, the beams of a window sharing its cross-attention K/V */
enum oh_sampling_strategy { OH_SAMPLING_GREEDY = 0, OH_SAMPLING_BEAM_SEARCH = 1 };

/* replaces whisper_full_params / FullParams (src/engine/whisper.rs:243-263); defaults in
 * oh_full_default_params are SURVEY.md Appendix A.2 */
typedef struct oh_full_params {
    int strategy;
    int n_threads;      /* accepted, unused */
    int n_max_text_ctx; /* 16384 */
    int offset_ms;
    int duration_ms;
    int translate;
    int no_context;
    int no_timestamps;
    int single_segment;
    int print_special;
    int print_progress;
    int print_realtime;
    int print_timestamps;
    int max_tokens;
    int suppress_blank;
    int suppress_non_speech_tokens; /* must be 0 */
    float temperature;
    float max_initial_ts;
    float length_penalty;
    float temperature_inc;
    float entropy_thold;
    float logprob_thold;
    float no_speech_thold;
    int greedy_best_of;
    int beam_size;
    float beam_patience;
    const char* language; /* borrowed; NULL or "auto" behave as "en" (A.2: detect_language is never set) */
} oh_full_params;

/* ---- lifecycle -------------------------------------------------------------------------------- */

/* whisper_context_default_params()  <- WhisperContextParameters::default(), whisper.rs:156 */
OH_API oh_context_params oh_context_default_params(void);

/* whisper_init_from_file_with_params()  <- WhisperContext::new_with_params, whisper.rs:159.
 * Reads a legacy-ggml `ggml-<name>.bin` (Appendix A.1), uploads the weights.  NULL on any failure
 * (the engine maps NULL to WhisperError::LoadFailed, whisper.rs:160). */
OH_API oh_context* oh_init_from_file(const char* path_model, oh_context_params params);

/* whisper_init_state()  <- ctx.create_state(), whisper.rs:167.  Allocates every device buffer once. */
OH_API oh_state* oh_init_state(oh_context* ctx);

/* whisper_free_state() / whisper_free()  <- Drop.  Either order is legal; NULL is a no-op. */
OH_API void oh_free_state(oh_state* state);
OH_API void oh_free(oh_context* ctx);

/* last error message of the calling thread ("" if none); for LoadFailed(format!(..)) texts */
OH_API const char* oh_last_error(void);

/* ---- the hot call ----------------------------------------------------------------------------- */

/* whisper_full_default_params(strategy)  <- FullParams::new(Greedy{best_of:1}), whisper.rs:243 */
OH_API oh_full_params oh_full_default_params(int strategy);

/* whisper_full_with_state()  <- state.full(params, &samples), whisper.rs:266-268.
 * 16 kHz mono f32 PCM in HOST memory, borrowed for the call.  Blocking.  Results are stored in `state`. */
OH_API int oh_full(oh_context* ctx, oh_state* state, oh_full_params params, const float* samples, int n_samples);

/* additive: `batch` independent oh_full calls run as one batched job (BASELINE configs 2-5).
 * Item i is exactly oh_full(samples[i], n_samples[i]) on a freshly created state; results are read with the *_item getters.
 * (Sampled tokens of the temperature fallback, A.7: oh_full draws from the state's generator, which is seeded 0 at state
 * creation and continues across calls like upstream's; every item of a batched or streamed call draws from a generator
 * seeded 0 for that call, so its result does not depend on what the state or its neighbours decoded before.) */
OH_API int oh_full_batch(oh_context* ctx, oh_state* state, oh_full_params params, const float* const* samples,
                         const int* n_samples, int batch);

/* as oh_full_batch, but the PCM already lives in DEVICE memory (bench.py's kernel-only `value`) */
OH_API int oh_full_batch_device(oh_context* ctx, oh_state* state, oh_full_params params, const float* const* d_samples,
                                const int* n_samples, int batch);

/* additive: continuous batching ACROSS calls (the daemon's jobs are independent, queue/mod.rs:18-27).
 * oh_stream_submit registers `batch` jobs (PCM in host memory, or device memory with on_device != 0; borrowed until the
 * call returns: it uploads the PCM and computes the log-mel) and returns a ticket; the jobs start to share the resident
 * decoder rows with the jobs of earlier tickets.  oh_stream_wait(ticket) drives the scheduler until every job of that
 * ticket is done; its results are then read with the *_item getters (until the next wait / oh_full* on the state).  Jobs of
 * later tickets keep decoding meanwhile, so the rows never drain between calls.  oh_full_batch == submit + wait.
 * All live tickets must share one set of oh_full_params (a different set waits for the live tickets to finish); sampled
 * (t > 0) tokens of a streamed job draw from a generator seeded 0 per job.  One thread at a time per state. */
OH_API int oh_stream_submit(oh_context* ctx, oh_state* state, oh_full_params params, const float* const* samples, const int* n_samples,
                            int batch, int on_device, int* out_ticket);
OH_API int oh_stream_wait(oh_context* ctx, oh_state* state, int ticket);
OH_API int oh_stream_pending(oh_state* state); /* tickets submitted and not yet waited for */

/* ---- results (item 0 unless *_item) ----------------------------------------------------------- */

/* whisper_full_n_segments_from_state()  <- state.full_n_segments(), whisper.rs:271 */
OH_API int oh_full_n_segments(oh_state* state);
/* whisper_full_get_segment_text_from_state()  <- state.full_get_segment_text(i), whisper.rs:277.
 * NUL-terminated bytes, possibly invalid UTF-8 (the engine skips such segments). */
OH_API const char* oh_full_get_segment_text(oh_state* state, int i_segment);
/* whisper_full_get_segment_t0/t1_from_state(): 10 ms units */
OH_API int64_t oh_full_get_segment_t0(oh_state* state, int i_segment);
OH_API int64_t oh_full_get_segment_t1(oh_state* state, int i_segment);
/* whisper_full_n_tokens_from_state() / whisper_full_get_token_id_from_state() / ..._get_token_p... */
OH_API int oh_full_n_tokens(oh_state* state, int i_segment);
OH_API int oh_full_get_token_id(oh_state* state, int i_segment, int i_token);
OH_API float oh_full_get_token_p(oh_state* state, int i_segment, int i_token);
/* whisper_full_lang_id_from_state()  <- state.full_lang_id_from_state(), whisper.rs:290 */
OH_API int oh_full_lang_id(oh_state* state);

/* per-item return code of a batched call: 0, or OH_ERR_MEL for a clip too short to reflect-pad (its item has no segments) */
OH_API int oh_full_item_rc(oh_state* state, int item);
OH_API int oh_full_n_segments_item(oh_state* state, int item);
OH_API const char* oh_full_get_segment_text_item(oh_state* state, int item, int i_segment);
OH_API int64_t oh_full_get_segment_t0_item(oh_state* state, int item, int i_segment);
OH_API int64_t oh_full_get_segment_t1_item(oh_state* state, int item, int i_segment);
OH_API int oh_full_n_tokens_item(oh_state* state, int item, int i_segment);
OH_API int oh_full_get_token_id_item(oh_state* state, int item, int i_segment, int i_token);
OH_API float oh_full_get_token_plog_item(oh_state* state, int item, int i_segment, int i_token);
OH_API int oh_full_get_token_tid_item(oh_state* state, int item, int i_segment, int i_token);

/* fixed-size segment records for the multi-GPU gather: writes up to `cap` int64
 * words {item, n_seg, then per segment: t0, t1, n_tok, token ids...}; returns words needed */
OH_API int64_t oh_full_pack_records(oh_state* state, int64_t* out, int64_t cap);

/* ---- model introspection (whisper_n_vocab, whisper_token_eot, ... in whisper.h) ---------------- */
OH_API int oh_n_vocab(oh_context* ctx);
OH_API int oh_n_audio_ctx(oh_context* ctx);
OH_API int oh_n_audio_state(oh_context* ctx);
OH_API int oh_n_audio_layer(oh_context* ctx);
OH_API int oh_n_text_ctx(oh_context* ctx);
OH_API int oh_n_text_state(oh_context* ctx);
OH_API int oh_n_text_layer(oh_context* ctx);
OH_API int oh_n_mels(oh_context* ctx);
OH_API int oh_is_multilingual(oh_context* ctx);
OH_API int oh_token_eot(oh_context* ctx);
OH_API int oh_token_sot(oh_context* ctx);
OH_API int oh_token_beg(oh_context* ctx);
OH_API int oh_token_not(oh_context* ctx);
OH_API const char* oh_token_to_str(oh_context* ctx, int token);
OH_API int oh_lang_id(const char* lang); /* whisper_lang_id(): -1 if unknown */
/* model-file handling: how long oh_init_from_file spent parsing + uploading the file (the engine's
 * load / idle-unload cycle, src/daemon.rs:2241-2283), and how many tensors of the file were block-quantised (ggml q4_0 / q4_1 /
 * q5_0 / q5_1 / q8_0, expanded to f16 at load) */
OH_API float oh_model_load_ms(oh_context* ctx);
OH_API int oh_model_n_quantised(oh_context* ctx);

/* ---- pre-engine DSP on the device: what load_wav_file does between the WAV reader and the engine
 * (src/input/audio.rs:368-424) — samples to f32 (format 0: f32, 1: i16, 2: i32 carrying `bits`-bit integers, value / 2^(bits-1)),
 * mono = mean of a frame's channels, sinc resampling to 16 kHz with the parameters of resample_sinc (audio.rs:1007-1095:
 * 256 taps, cutoff 0.95, 256x oversampled, linear interpolation, Blackman-Harris^2, 1024-frame chunks, last chunk cut to
 * ceil(len * ratio)), zero padding to 1.1 s.  `pcm`: n_frames interleaved frames.  oh_pcm_16k_len gives the output length.
 * The host variant copies in and out; the device variant reads and writes device memory on the state's stream (its output
 * feeds oh_full_batch_device / oh_stream_submit(on_device) without touching the host). */
OH_API long long oh_pcm_16k_len(int sample_rate, long long n_frames);
OH_API int oh_load_pcm_16k(oh_context* ctx, oh_state* state, const void* pcm, int format, int bits, int channels, int sample_rate,
                           long long n_frames, float* out, long long cap, long long* out_n);
OH_API int oh_load_pcm_16k_device(oh_context* ctx, oh_state* state, const void* d_pcm, int format, int bits, int channels, int sample_rate,
                                  long long n_frames, float* d_out, long long cap, long long* out_n);

/* ---- optional level processing of 16 kHz clips on the device: the worker's hook between capture and
 * WhisperEngine::transcribe, src/queue/worker.rs:196-240 (preprocess_audio), in its order: AudioBuffer::normalize_rms
 * (src/input/audio.rs:108-122), ::compress (:139-198), ::limit (:200-239).  Field defaults are the reference's
 * (src/config.rs:1129-1159, 1239-1269); `preprocessing` itself is off by default there (config.rs:984), so nothing calls this
 * unless the configuration asks for it.  RNNoise denoising (worker.rs:200-206) is a third-party model and not provided. */
typedef struct oh_preprocess_params {
    int normalize;          /* NormalizationConfig.enabled */
    float target_db;        /* -18 */
    int compress;           /* CompressionConfig.enabled */
    float threshold_db;     /* -24 */
    float ratio;            /* 4 */
    float attack_ms;        /* 5 */
    float release_ms;       /* 50 */
    float makeup_gain_db;   /* 6 */
    int limit;              /* LimiterConfig.enabled */
    float ceiling_db;       /* -1 */
    float limiter_release_ms; /* 50 */
} oh_preprocess_params;
OH_API oh_preprocess_params oh_preprocess_default_params(void);
/* `batch` clips of host PCM (f32, 16 kHz), processed in place; rms_db (may be NULL) receives [batch][2] = RMS level in dB
 * before and after (-inf for a silent or empty clip, as AudioBuffer::rms_db) */
OH_API int oh_preprocess_pcm(oh_context* ctx, oh_state* state, float* const* pcm, const int* n_samples, int batch,
                             const oh_preprocess_params* params, float* rms_db);
/* same on device memory: d_pcm is one [batch][stride] array (the layout oh_full_batch_device / oh_stream_submit take), processed in
 * place on the state's stream; n_samples and rms_db are host arrays */
OH_API int oh_preprocess_pcm_device(oh_context* ctx, oh_state* state, float* d_pcm, long long stride, const int* n_samples, int batch,
                                    const oh_preprocess_params* params, float* rms_db);

/* RMS level in dB full scale of every `chunk`-sample block of a 16 kHz clip (host in, host out), one launch for the whole
 * clip: the per-chunk feature of the stand-in VAD engine behind the reference's VadEngine trait (src/vad/mod.rs:34-55; the
 * reference's own engine, Silero, is an ONNX model whose weights are not in its tree).  SURVEY.md section 8f N3. */
OH_API int oh_chunk_rms_db(oh_context* ctx, oh_state* state, const float* pcm, long long n_samples, int chunk, float* out, int n_chunks);

/* ---- stage entry points (whisper_pcm_to_mel / whisper_encode / whisper_decode / whisper_get_logits
 * in whisper.h; the reference never calls them, the parity tests and bench.py do) ---------------- */

/* log-mel of `batch` clips (host PCM).  out[i] receives n_mel * n_len[i] floats laid out [n_mel][n_len]
 * (n_len = (n + 480000) / 160), or may be NULL to only run the kernels.  K1 + K2. */
OH_API int oh_pcm_to_mel_batch(oh_context* ctx, oh_state* state, const float* const* samples, const int* n_samples, int batch,
                               float* const* out_mel, int* out_n_len);
/* same, PCM resident in device memory as one [batch][n_samples] array of equal-length clips; output
 * stays on the device ([batch][n_mel][n_len] f32).  Used for the roofline measurement of K1. */
OH_API int oh_pcm_to_mel_device(oh_context* ctx, oh_state* state, const float* d_samples, int n_samples, int batch, float* d_out_mel);

/* encoder forward of `batch` 30 s mel windows given as host [batch][n_mel][3000] f32; writes the encoder
 * states [batch][1500][d] f32 to out_enc (host, may be NULL) and fills the cross-attention K/V cache. */
OH_API int oh_encode_batch(oh_context* ctx, oh_state* state, const float* mel_windows, int batch, float* out_enc);
/* as oh_encode_batch with the mel windows already in device memory; no host output */
OH_API int oh_encode_device(oh_context* ctx, oh_state* state, const float* d_mel_windows, int batch);

/* decoder forward over `n_tokens` tokens per sequence starting at position n_past, for `batch` sequences
 * (tokens is [batch][n_tokens], host).  Appends to the self-attention KV cache and writes the logits of the
 * LAST position, [batch][n_vocab] f32, to out_logits (host). */
OH_API int oh_decode_batch(oh_context* ctx, oh_state* state, const int* tokens, int n_tokens, int n_past, int batch, float* out_logits);

/* timing / counters of the state's activity between the previous oh_full* / oh_stream_wait return and the last one */
typedef struct oh_timings {
    float mel_ms, encode_ms, decode_ms, total_ms; /* CUDA-event times on the state's stream */
    int n_windows, n_decode_steps, n_tokens, n_fallbacks;
    long long n_kernel_launches;
    long long h2d_bytes, d2h_bytes; /* host<->device traffic of the call (PCM in, decoder state / token records out) */
    long long n_prompt_steps;       /* row-steps spent feeding prompt tokens (one per step), not sampling */
    long long n_row_steps;          /* decode steps x resident rows */
    long long n_armed_row_steps;    /* of those, row-steps of rows that had a window to decode (the rest: no window was ready) */
} oh_timings;
OH_API oh_timings oh_get_timings(oh_state* state);

/* which mat-mul kernel the state uses: "tcgen05" (16-bit modes), "simt-f32" (parity mode) or "simt" (OH_GEMM=simt);
 * and the relative L2 error the tcgen05 kernel showed against the SIMT kernel in the start-up self-test (-1 if not run) */
OH_API const char* oh_gemm_backend(oh_state* state);
OH_API float oh_gemm_selftest_error(void);

/* per-kernel CUDA-event timing for bench.py's roofline block: oh_profile(state, 1) brackets every kernel launch made
 * by the calling thread with two events on the launching stream (CUDA graphs are bypassed meanwhile) until
 * oh_profile(state, 0); oh_profile_report sums them per kernel.  `work` = algorithmic FLOPs for the mat-mul kernels,
 * algorithmic bytes for the others (rows assumed active). */
typedef struct oh_kernel_stat {
    char name[32];
    int launches;
    float total_ms;
    double work;
} oh_kernel_stat;
OH_API int oh_profile(oh_state* state, int on);
OH_API int oh_profile_report(oh_state* state, oh_kernel_stat* out, int cap);

/* run the library on a caller-provided CUDA stream (cudaStream_t as void*); NULL = the state's own */
OH_API int oh_state_set_stream(oh_state* state, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* OH_WHISPER_H */
