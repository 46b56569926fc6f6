//! Raw declarations of `include/oh_whisper.h`, one item per C declaration, same order.
//! Each function replaces the whisper.h symbol named in the header's comment (the reference reaches them through
//! whisper-rs at /root/reference/src/engine/whisper.rs:156-169, :243-268, :271-280, :290).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int, c_longlong, c_void};

#[repr(C)]
pub struct oh_context {
    _private: [u8; 0],
}
#[repr(C)]
pub struct oh_state {
    _private: [u8; 0],
}

pub const OH_OK: c_int = 0;
pub const OH_ERR_MEL: c_int = -1;
pub const OH_ERR_INVALID: c_int = -2;
pub const OH_ERR_ENCODE: c_int = -6;
pub const OH_ERR_DECODE: c_int = -7;
pub const OH_ERR_CUDA: c_int = -100;
pub const OH_ERR_UNSUPPORTED: c_int = -101;

pub const OH_COMPUTE_F32: c_int = 0;
pub const OH_COMPUTE_F16: c_int = 1;
pub const OH_COMPUTE_BF16: c_int = 2;

pub const OH_SAMPLING_GREEDY: c_int = 0;
pub const OH_SAMPLING_BEAM_SEARCH: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oh_context_params {
    pub use_gpu: c_int,
    pub flash_attn: c_int,
    pub gpu_device: c_int,
    pub compute_type: c_int,
    pub max_batch: c_int,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oh_full_params {
    pub strategy: c_int,
    pub n_threads: c_int,
    pub n_max_text_ctx: c_int,
    pub offset_ms: c_int,
    pub duration_ms: c_int,
    pub translate: c_int,
    pub no_context: c_int,
    pub no_timestamps: c_int,
    pub single_segment: c_int,
    pub print_special: c_int,
    pub print_progress: c_int,
    pub print_realtime: c_int,
    pub print_timestamps: c_int,
    pub max_tokens: c_int,
    pub suppress_blank: c_int,
    pub suppress_non_speech_tokens: c_int,
    pub temperature: c_float,
    pub max_initial_ts: c_float,
    pub length_penalty: c_float,
    pub temperature_inc: c_float,
    pub entropy_thold: c_float,
    pub logprob_thold: c_float,
    pub no_speech_thold: c_float,
    pub greedy_best_of: c_int,
    pub beam_size: c_int,
    pub beam_patience: c_float,
    pub language: *const c_char,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct oh_timings {
    pub mel_ms: c_float,
    pub encode_ms: c_float,
    pub decode_ms: c_float,
    pub total_ms: c_float,
    pub n_windows: c_int,
    pub n_decode_steps: c_int,
    pub n_tokens: c_int,
    pub n_fallbacks: c_int,
    pub n_kernel_launches: c_longlong,
    pub h2d_bytes: c_longlong,
    pub d2h_bytes: c_longlong,
    pub n_prompt_steps: c_longlong,
    pub n_row_steps: c_longlong,
    pub n_armed_row_steps: c_longlong,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct oh_kernel_stat {
    pub name: [c_char; 32],
    pub launches: c_int,
    pub total_ms: c_float,
    pub work: f64,
}

/// Level processing between capture and the engine (`preprocess_audio`, src/queue/worker.rs:196-240); defaults are the
/// reference's `NormalizationConfig` / `CompressionConfig` / `LimiterConfig` (src/config.rs:1129-1159, 1239-1269).
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct oh_preprocess_params {
    pub normalize: c_int,
    pub target_db: c_float,
    pub compress: c_int,
    pub threshold_db: c_float,
    pub ratio: c_float,
    pub attack_ms: c_float,
    pub release_ms: c_float,
    pub makeup_gain_db: c_float,
    pub limit: c_int,
    pub ceiling_db: c_float,
    pub limiter_release_ms: c_float,
}

extern "C" {
    // ---- lifecycle
    pub fn oh_context_default_params() -> oh_context_params;
    pub fn oh_init_from_file(path_model: *const c_char, params: oh_context_params) -> *mut oh_context;
    pub fn oh_init_state(ctx: *mut oh_context) -> *mut oh_state;
    pub fn oh_free_state(state: *mut oh_state);
    pub fn oh_free(ctx: *mut oh_context);
    pub fn oh_last_error() -> *const c_char;
    // ---- the hot call
    pub fn oh_full_default_params(strategy: c_int) -> oh_full_params;
    pub fn oh_full(ctx: *mut oh_context, state: *mut oh_state, params: oh_full_params, samples: *const f32, n_samples: c_int) -> c_int;
    pub fn oh_full_batch(ctx: *mut oh_context, state: *mut oh_state, params: oh_full_params, samples: *const *const f32, n_samples: *const c_int,
                         batch: c_int) -> c_int;
    pub fn oh_full_batch_device(ctx: *mut oh_context, state: *mut oh_state, params: oh_full_params, d_samples: *const *const f32,
                                n_samples: *const c_int, batch: c_int) -> c_int;
    pub fn oh_stream_submit(ctx: *mut oh_context, state: *mut oh_state, params: oh_full_params, samples: *const *const f32,
                            n_samples: *const c_int, batch: c_int, on_device: c_int, out_ticket: *mut c_int) -> c_int;
    pub fn oh_stream_wait(ctx: *mut oh_context, state: *mut oh_state, ticket: c_int) -> c_int;
    pub fn oh_stream_pending(state: *mut oh_state) -> c_int;
    // ---- results
    pub fn oh_full_n_segments(state: *mut oh_state) -> c_int;
    pub fn oh_full_get_segment_text(state: *mut oh_state, i_segment: c_int) -> *const c_char;
    pub fn oh_full_get_segment_t0(state: *mut oh_state, i_segment: c_int) -> i64;
    pub fn oh_full_get_segment_t1(state: *mut oh_state, i_segment: c_int) -> i64;
    pub fn oh_full_n_tokens(state: *mut oh_state, i_segment: c_int) -> c_int;
    pub fn oh_full_get_token_id(state: *mut oh_state, i_segment: c_int, i_token: c_int) -> c_int;
    pub fn oh_full_get_token_p(state: *mut oh_state, i_segment: c_int, i_token: c_int) -> c_float;
    pub fn oh_full_lang_id(state: *mut oh_state) -> c_int;
    pub fn oh_full_item_rc(state: *mut oh_state, item: c_int) -> c_int;
    pub fn oh_full_n_segments_item(state: *mut oh_state, item: c_int) -> c_int;
    pub fn oh_full_get_segment_text_item(state: *mut oh_state, item: c_int, i_segment: c_int) -> *const c_char;
    pub fn oh_full_get_segment_t0_item(state: *mut oh_state, item: c_int, i_segment: c_int) -> i64;
    pub fn oh_full_get_segment_t1_item(state: *mut oh_state, item: c_int, i_segment: c_int) -> i64;
    pub fn oh_full_n_tokens_item(state: *mut oh_state, item: c_int, i_segment: c_int) -> c_int;
    pub fn oh_full_get_token_id_item(state: *mut oh_state, item: c_int, i_segment: c_int, i_token: c_int) -> c_int;
    pub fn oh_full_get_token_plog_item(state: *mut oh_state, item: c_int, i_segment: c_int, i_token: c_int) -> c_float;
    pub fn oh_full_get_token_tid_item(state: *mut oh_state, item: c_int, i_segment: c_int, i_token: c_int) -> c_int;
    pub fn oh_full_pack_records(state: *mut oh_state, out: *mut i64, cap: i64) -> i64;
    // ---- model introspection
    pub fn oh_n_vocab(ctx: *mut oh_context) -> c_int;
    pub fn oh_n_audio_ctx(ctx: *mut oh_context) -> c_int;
    pub fn oh_n_audio_state(ctx: *mut oh_context) -> c_int;
    pub fn oh_n_audio_layer(ctx: *mut oh_context) -> c_int;
    pub fn oh_n_text_ctx(ctx: *mut oh_context) -> c_int;
    pub fn oh_n_text_state(ctx: *mut oh_context) -> c_int;
    pub fn oh_n_text_layer(ctx: *mut oh_context) -> c_int;
    pub fn oh_n_mels(ctx: *mut oh_context) -> c_int;
    pub fn oh_is_multilingual(ctx: *mut oh_context) -> c_int;
    pub fn oh_token_eot(ctx: *mut oh_context) -> c_int;
    pub fn oh_token_sot(ctx: *mut oh_context) -> c_int;
    pub fn oh_token_beg(ctx: *mut oh_context) -> c_int;
    pub fn oh_token_not(ctx: *mut oh_context) -> c_int;
    pub fn oh_token_to_str(ctx: *mut oh_context, token: c_int) -> *const c_char;
    pub fn oh_lang_id(lang: *const c_char) -> c_int;
    pub fn oh_model_load_ms(ctx: *mut oh_context) -> c_float;
    pub fn oh_model_n_quantised(ctx: *mut oh_context) -> c_int;
    // ---- pre-engine DSP
    pub fn oh_pcm_16k_len(sample_rate: c_int, n_frames: c_longlong) -> c_longlong;
    pub fn oh_load_pcm_16k(ctx: *mut oh_context, state: *mut oh_state, pcm: *const c_void, format: c_int, bits: c_int, channels: c_int,
                           sample_rate: c_int, n_frames: c_longlong, out: *mut f32, cap: c_longlong, out_n: *mut c_longlong) -> c_int;
    pub fn oh_load_pcm_16k_device(ctx: *mut oh_context, state: *mut oh_state, d_pcm: *const c_void, format: c_int, bits: c_int, channels: c_int,
                                  sample_rate: c_int, n_frames: c_longlong, d_out: *mut f32, cap: c_longlong, out_n: *mut c_longlong) -> c_int;
    pub fn oh_chunk_rms_db(ctx: *mut oh_context, state: *mut oh_state, pcm: *const f32, n_samples: c_longlong, chunk: c_int, out: *mut f32,
                           n_chunks: c_int) -> c_int;
    pub fn oh_preprocess_default_params() -> oh_preprocess_params;
    pub fn oh_preprocess_pcm(ctx: *mut oh_context, state: *mut oh_state, pcm: *const *mut f32, n_samples: *const c_int, batch: c_int,
                             params: *const oh_preprocess_params, rms_db: *mut f32) -> c_int;
    pub fn oh_preprocess_pcm_device(ctx: *mut oh_context, state: *mut oh_state, d_pcm: *mut f32, stride: c_longlong, n_samples: *const c_int,
                                    batch: c_int, params: *const oh_preprocess_params, rms_db: *mut f32) -> c_int;
    // ---- stage entry points
    pub fn oh_pcm_to_mel_batch(ctx: *mut oh_context, state: *mut oh_state, samples: *const *const f32, n_samples: *const c_int, batch: c_int,
                               out_mel: *const *mut f32, out_n_len: *mut c_int) -> c_int;
    pub fn oh_pcm_to_mel_device(ctx: *mut oh_context, state: *mut oh_state, d_samples: *const f32, n_samples: c_int, batch: c_int,
                                d_out_mel: *mut f32) -> c_int;
    pub fn oh_encode_batch(ctx: *mut oh_context, state: *mut oh_state, mel_windows: *const f32, batch: c_int, out_enc: *mut f32) -> c_int;
    pub fn oh_encode_device(ctx: *mut oh_context, state: *mut oh_state, d_mel_windows: *const f32, batch: c_int) -> c_int;
    pub fn oh_decode_batch(ctx: *mut oh_context, state: *mut oh_state, tokens: *const c_int, n_tokens: c_int, n_past: c_int, batch: c_int,
                           out_logits: *mut f32) -> c_int;
    // ---- measurement
    pub fn oh_get_timings(state: *mut oh_state) -> oh_timings;
    pub fn oh_gemm_backend(state: *mut oh_state) -> *const c_char;
    pub fn oh_gemm_selftest_error() -> c_float;
    pub fn oh_profile(state: *mut oh_state, on: c_int) -> c_int;
    pub fn oh_profile_report(state: *mut oh_state, out: *mut oh_kernel_stat, cap: c_int) -> c_int;
    pub fn oh_state_set_stream(state: *mut oh_state, cuda_stream: *mut c_void) -> c_int;
}
