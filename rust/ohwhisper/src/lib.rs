//! The whisper-rs 0.13 surface that OpenHush's engine uses (/root/reference/src/engine/whisper.rs:10-12), implemented over
//! `ohwhisper-sys`.  Same type and method names, argument meaning and error variants, so the engine file is unchanged:
//!
//!   WhisperContextParameters::default() / .use_gpu(b)            whisper.rs:156-157
//!   WhisperContext::new_with_params(path, params)                 whisper.rs:159   (Err -> WhisperError::LoadFailed, :160)
//!   ctx.create_state()                                            whisper.rs:167
//!   FullParams::new(SamplingStrategy::Greedy { best_of: 1 })      whisper.rs:243
//!   params.set_language / set_translate / set_print_*             whisper.rs:246-263
//!   state.full(params, &samples)                                  whisper.rs:266
//!   state.full_n_segments() / full_get_segment_text(i)            whisper.rs:271-280
//!   state.full_lang_id_from_state()                               whisper.rs:290
//!
//! Additive (not in whisper-rs): `gpu_device`, `max_batch`, and the `BatchExt` trait (batched and streamed jobs for the
//! one-worker-per-GPU queue, /root/reference/src/queue/worker.rs:255-267).
use ohwhisper_sys as sys;
use std::ffi::{c_int, CStr, CString};
use std::marker::PhantomData;
use std::sync::Arc;

#[derive(Debug, Clone, PartialEq)]
pub enum WhisperError {
    InitError,
    NoSamples,
    UnableToCalculateSpectrogram,
    FailedToEncode,
    FailedToDecode,
    InvalidUtf8,
    NullPointer,
    GenericError(c_int),
}
impl std::fmt::Display for WhisperError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{self:?}")
    }
}
impl std::error::Error for WhisperError {}

fn last_error() -> String {
    let p = unsafe { sys::oh_last_error() };
    if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() }
}

// ------------------------------------------------------------------------------------------------ context
pub struct WhisperContextParameters {
    inner: sys::oh_context_params,
}
impl Default for WhisperContextParameters {
    fn default() -> Self {
        Self { inner: unsafe { sys::oh_context_default_params() } }
    }
}
impl WhisperContextParameters {
    pub fn new() -> Self {
        Self::default()
    }
    /// whisper.rs:157.  `false` makes `new_with_params` fail: the library has no CPU path.
    pub fn use_gpu(&mut self, use_gpu: bool) -> &mut Self {
        self.inner.use_gpu = use_gpu as c_int;
        self
    }
    pub fn flash_attn(&mut self, on: bool) -> &mut Self {
        self.inner.flash_attn = on as c_int;
        self
    }
    /// CUDA ordinal of this worker's GPU (`config.gpu.devices[i]`, /root/reference/src/config.rs:921-929).
    pub fn gpu_device(&mut self, device: c_int) -> &mut Self {
        self.inner.gpu_device = device;
        self
    }
    /// Decoder rows resident on the GPU: how many queued jobs a worker's batched call decodes side by side.
    pub fn max_batch(&mut self, rows: c_int) -> &mut Self {
        self.inner.max_batch = rows;
        self
    }
}

struct CtxPtr(*mut sys::oh_context);
unsafe impl Send for CtxPtr {}
unsafe impl Sync for CtxPtr {}
impl Drop for CtxPtr {
    fn drop(&mut self) {
        unsafe { sys::oh_free(self.0) }
    }
}

pub struct WhisperContext {
    ctx: Arc<CtxPtr>,
}
impl WhisperContext {
    pub fn new_with_params(path: &str, params: WhisperContextParameters) -> Result<Self, WhisperError> {
        let c = CString::new(path).map_err(|_| WhisperError::InitError)?;
        let ptr = unsafe { sys::oh_init_from_file(c.as_ptr(), params.inner) };
        if ptr.is_null() {
            eprintln!("ohwhisper: {}", last_error());
            return Err(WhisperError::InitError); // the engine formats it into LoadFailed, whisper.rs:160
        }
        Ok(Self { ctx: Arc::new(CtxPtr(ptr)) })
    }
    pub fn create_state(&self) -> Result<WhisperState, WhisperError> {
        let s = unsafe { sys::oh_init_state(self.ctx.0) };
        if s.is_null() {
            return Err(WhisperError::InitError);
        }
        Ok(WhisperState { ctx: self.ctx.clone(), state: s })
    }
}

// ------------------------------------------------------------------------------------------------ params
#[derive(Debug, Clone, Copy)]
pub enum SamplingStrategy {
    Greedy { best_of: c_int },
    BeamSearch { beam_size: c_int, patience: f32 },
}

pub struct FullParams<'a, 'b> {
    inner: sys::oh_full_params,
    language: Option<CString>,
    _lt: PhantomData<(&'a (), &'b ())>,
}
impl<'a, 'b> FullParams<'a, 'b> {
    pub fn new(strategy: SamplingStrategy) -> Self {
        let mut inner = unsafe {
            sys::oh_full_default_params(match strategy {
                SamplingStrategy::Greedy { .. } => sys::OH_SAMPLING_GREEDY,
                SamplingStrategy::BeamSearch { .. } => sys::OH_SAMPLING_BEAM_SEARCH,
            })
        };
        match strategy {
            SamplingStrategy::Greedy { best_of } => inner.greedy_best_of = best_of,
            SamplingStrategy::BeamSearch { beam_size, patience } => {
                inner.beam_size = beam_size;
                inner.beam_patience = patience;
            }
        }
        Self { inner, language: None, _lt: PhantomData }
    }
    /// whisper.rs:246-248 (only called when the configured language is not "auto").
    pub fn set_language(&mut self, language: Option<&'a str>) {
        self.language = language.and_then(|s| CString::new(s).ok());
        self.inner.language = self.language.as_ref().map_or(std::ptr::null(), |c| c.as_ptr());
    }
    /// Plain field assignment, like upstream (the engine passes `!config.translate`, whisper.rs:251-257).
    pub fn set_translate(&mut self, translate: bool) {
        self.inner.translate = translate as c_int;
    }
    pub fn set_no_context(&mut self, v: bool) {
        self.inner.no_context = v as c_int;
    }
    pub fn set_no_timestamps(&mut self, v: bool) {
        self.inner.no_timestamps = v as c_int;
    }
    pub fn set_single_segment(&mut self, v: bool) {
        self.inner.single_segment = v as c_int;
    }
    pub fn set_print_special(&mut self, v: bool) {
        self.inner.print_special = v as c_int;
    }
    pub fn set_print_progress(&mut self, v: bool) {
        self.inner.print_progress = v as c_int;
    }
    pub fn set_print_realtime(&mut self, v: bool) {
        self.inner.print_realtime = v as c_int;
    }
    pub fn set_print_timestamps(&mut self, v: bool) {
        self.inner.print_timestamps = v as c_int;
    }
    pub fn set_suppress_blank(&mut self, v: bool) {
        self.inner.suppress_blank = v as c_int;
    }
    pub fn set_temperature(&mut self, t: f32) {
        self.inner.temperature = t;
    }
    pub fn set_temperature_inc(&mut self, t: f32) {
        self.inner.temperature_inc = t;
    }
    pub fn set_max_tokens(&mut self, n: c_int) {
        self.inner.max_tokens = n;
    }
    pub fn set_n_threads(&mut self, n: c_int) {
        self.inner.n_threads = n; // accepted, unused
    }
}

fn map_rc(rc: c_int) -> Result<c_int, WhisperError> {
    match rc {
        0 => Ok(0),
        -1 => Err(WhisperError::UnableToCalculateSpectrogram),
        -6 | 7 => Err(WhisperError::FailedToEncode),
        -7 | 8 => Err(WhisperError::FailedToDecode),
        n => Err(WhisperError::GenericError(n)),
    }
}

// ------------------------------------------------------------------------------------------------ state
pub struct WhisperState {
    ctx: Arc<CtxPtr>, // the context outlives its states, like whisper-rs' Arc (whisper.rs:105-117)
    state: *mut sys::oh_state,
}
unsafe impl Send for WhisperState {} // created on the tokio thread, moved into the worker: daemon.rs:1803-1805
impl Drop for WhisperState {
    fn drop(&mut self) {
        unsafe { sys::oh_free_state(self.state) }
    }
}
impl WhisperState {
    /// whisper.rs:266.  Blocking; `data` is borrowed for the call only.
    pub fn full(&mut self, params: FullParams, data: &[f32]) -> Result<c_int, WhisperError> {
        if data.is_empty() {
            return Err(WhisperError::NoSamples);
        }
        map_rc(unsafe { sys::oh_full(self.ctx.0, self.state, params.inner, data.as_ptr(), data.len() as c_int) })
    }
    pub fn full_n_segments(&self) -> Result<c_int, WhisperError> {
        Ok(unsafe { sys::oh_full_n_segments(self.state) })
    }
    /// Err(InvalidUtf8) for a segment whose bytes are not UTF-8; the engine skips those (whisper.rs:277).
    pub fn full_get_segment_text(&self, segment: c_int) -> Result<String, WhisperError> {
        let p = unsafe { sys::oh_full_get_segment_text(self.state, segment) };
        if p.is_null() {
            return Err(WhisperError::NullPointer);
        }
        unsafe { CStr::from_ptr(p) }.to_str().map(str::to_owned).map_err(|_| WhisperError::InvalidUtf8)
    }
    pub fn full_get_segment_t0(&self, segment: c_int) -> Result<i64, WhisperError> {
        Ok(unsafe { sys::oh_full_get_segment_t0(self.state, segment) })
    }
    pub fn full_get_segment_t1(&self, segment: c_int) -> Result<i64, WhisperError> {
        Ok(unsafe { sys::oh_full_get_segment_t1(self.state, segment) })
    }
    pub fn full_n_tokens(&self, segment: c_int) -> Result<c_int, WhisperError> {
        Ok(unsafe { sys::oh_full_n_tokens(self.state, segment) })
    }
    pub fn full_get_token_id(&self, segment: c_int, token: c_int) -> Result<c_int, WhisperError> {
        Ok(unsafe { sys::oh_full_get_token_id(self.state, segment, token) })
    }
    pub fn full_lang_id_from_state(&self) -> Result<c_int, WhisperError> {
        Ok(unsafe { sys::oh_full_lang_id(self.state) })
    }
}

// ------------------------------------------------------------------------------------------------ additive: batches / streams
/// One transcribed job of a batched call: `(t0, t1, text)` per segment in 10 ms units, and the job's return code.
pub struct JobOutput {
    pub rc: c_int,
    pub segments: Vec<(i64, i64, Result<String, WhisperError>)>,
}

/// Batched and streamed jobs on one GPU worker.  A worker drains up to `max_batch` `TranscriptionJob`s
/// (/root/reference/src/queue/mod.rs:18-27) from the shared channel and hands them over in one call; with the stream pair
/// it submits the next drain before collecting the previous one, so the decoder rows never run empty.
pub trait BatchExt {
    fn full_batch(&mut self, params: &FullParams, jobs: &[&[f32]]) -> Result<Vec<JobOutput>, WhisperError>;
    fn stream_submit(&mut self, params: &FullParams, jobs: &[&[f32]]) -> Result<c_int, WhisperError>;
    fn stream_wait(&mut self, ticket: c_int, n_jobs: usize) -> Result<Vec<JobOutput>, WhisperError>;
}
impl WhisperState {
    fn collect(&self, n_jobs: usize) -> Vec<JobOutput> {
        (0..n_jobs as c_int)
            .map(|item| {
                let n = unsafe { sys::oh_full_n_segments_item(self.state, item) };
                let segments = (0..n)
                    .map(|i| {
                        let p = unsafe { sys::oh_full_get_segment_text_item(self.state, item, i) };
                        let text = if p.is_null() {
                            Err(WhisperError::NullPointer)
                        } else {
                            unsafe { CStr::from_ptr(p) }.to_str().map(str::to_owned).map_err(|_| WhisperError::InvalidUtf8)
                        };
                        (unsafe { sys::oh_full_get_segment_t0_item(self.state, item, i) }, unsafe { sys::oh_full_get_segment_t1_item(self.state, item, i) }, text)
                    })
                    .collect();
                JobOutput { rc: unsafe { sys::oh_full_item_rc(self.state, item) }, segments }
            })
            .collect()
    }
}
impl BatchExt for WhisperState {
    fn full_batch(&mut self, params: &FullParams, jobs: &[&[f32]]) -> Result<Vec<JobOutput>, WhisperError> {
        let ptrs: Vec<*const f32> = jobs.iter().map(|j| j.as_ptr()).collect();
        let lens: Vec<c_int> = jobs.iter().map(|j| j.len() as c_int).collect();
        map_rc(unsafe { sys::oh_full_batch(self.ctx.0, self.state, params.inner, ptrs.as_ptr(), lens.as_ptr(), jobs.len() as c_int) })?;
        Ok(self.collect(jobs.len()))
    }
    fn stream_submit(&mut self, params: &FullParams, jobs: &[&[f32]]) -> Result<c_int, WhisperError> {
        let ptrs: Vec<*const f32> = jobs.iter().map(|j| j.as_ptr()).collect();
        let lens: Vec<c_int> = jobs.iter().map(|j| j.len() as c_int).collect();
        let mut ticket: c_int = 0;
        map_rc(unsafe {
            sys::oh_stream_submit(self.ctx.0, self.state, params.inner, ptrs.as_ptr(), lens.as_ptr(), jobs.len() as c_int, 0, &mut ticket)
        })?;
        Ok(ticket)
    }
    fn stream_wait(&mut self, ticket: c_int, n_jobs: usize) -> Result<Vec<JobOutput>, WhisperError> {
        map_rc(unsafe { sys::oh_stream_wait(self.ctx.0, self.state, ticket) })?;
        Ok(self.collect(n_jobs))
    }
}

// ------------------------------------------------------------------------------------------------ additive: level processing
/// The worker's `preprocess_audio` hook (/root/reference/src/queue/worker.rs:196-240) on the GPU: RMS normalisation,
/// compression and the limiter of `AudioBuffer` (/root/reference/src/input/audio.rs:108-239) over the buffers of the jobs a
/// worker drained, in place, in one call.  `params` is filled from the reference's `AudioConfig`
/// (`normalization` / `compression` / `limiter`; see `LevelParams::from_config` in INTEGRATION.md); the caller keeps the
/// `if !config.preprocessing { return; }` of the reference.  RNNoise noise reduction stays where it is (host side).
pub type LevelParams = sys::oh_preprocess_params;

pub trait PreprocessExt {
    /// Returns the RMS level in dB before and after, per buffer (what the reference logs at worker.rs:210 and :236).
    fn preprocess_audio(&mut self, buffers: &mut [&mut [f32]], params: &LevelParams) -> Result<Vec<(f32, f32)>, WhisperError>;
}
impl PreprocessExt for WhisperState {
    fn preprocess_audio(&mut self, buffers: &mut [&mut [f32]], params: &LevelParams) -> Result<Vec<(f32, f32)>, WhisperError> {
        let ptrs: Vec<*mut f32> = buffers.iter_mut().map(|b| b.as_mut_ptr()).collect();
        let lens: Vec<c_int> = buffers.iter().map(|b| b.len() as c_int).collect();
        let mut rms = vec![0.0f32; 2 * buffers.len()];
        map_rc(unsafe {
            sys::oh_preprocess_pcm(self.ctx.0, self.state, ptrs.as_ptr(), lens.as_ptr(), buffers.len() as c_int, params, rms.as_mut_ptr())
        })?;
        Ok(rms.chunks(2).map(|c| (c[0], c[1])).collect())
    }
}

/// Defaults of the reference's configuration (src/config.rs:1129-1159, 1239-1269).
pub fn default_level_params() -> LevelParams {
    unsafe { sys::oh_preprocess_default_params() }
}
