"""Optional level processing between capture and the engine ("optional normalise / limit"):
the worker's preprocess_audio hook (/root/reference/src/queue/worker.rs:196-240) over AudioBuffer::normalize_rms / compress /
limit (/root/reference/src/input/audio.rs:89-239).  The oracle's restatement is pinned on the reference's own unit tests for
these functions (audio.rs:1189-1330, restated below with their inputs and bounds); the CUDA kernels are checked against the
oracle and against the same unit tests."""
import os
import re

import numpy as np
import pytest

from openhush_b200 import _abi
from openhush_b200 import engine as E
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sine(n=16000, amp=1.0, hz=440.0):
    # the reference's test signal: (2 pi 440 i / 16000).sin() evaluated in f32
    i = np.arange(n, dtype=np.float32)
    return (np.float32(amp) * np.sin((np.float32(2.0 * np.pi) * np.float32(hz) * i / np.float32(16000.0)).astype(np.float32))).astype(np.float32)


def _loud_quiet():
    return np.concatenate([_sine(16000, 0.8), _sine(16000, 0.1)])


def _reference_unit_tests(rms_db, normalize_rms, apply_gain, compress, limit):
    """audio.rs:1189-1330 with the reference's inputs and assertions; the five callables are the implementation under test."""
    # test_rms_db_silence
    assert np.isinf(rms_db(np.zeros(16000, np.float32)))
    # test_rms_db_full_scale: a full-scale sine is -3.01 dB
    assert abs(rms_db(_sine()) - (-3.01)) < 0.1
    # test_normalize_rms: a ~ -43 dB sine normalised to -18 dB
    assert abs(rms_db(normalize_rms(_sine(amp=0.01), -18.0)) - (-18.0)) < 0.5
    # test_apply_gain: +6.02 dB doubles the amplitude
    g = apply_gain(np.array([0.5, -0.5, 0.25, -0.25], np.float32), 6.02)
    assert abs(g[0] - 1.0) < 0.01 and abs(g[1] + 1.0) < 0.01
    # test_compress_reduces_dynamic_range: 4:1 above -20 dB, 5 ms attack, 50 ms release, no make-up gain
    x = _loud_quiet()
    y = compress(x, -20.0, 4.0, 5.0, 50.0, 0.0)
    loud_b, quiet_b = np.abs(x[:16000]).max(), np.abs(x[16000:]).max()
    loud_a, quiet_a = np.abs(y[:16000]).max(), np.abs(y[16000:]).max()
    assert loud_a < loud_b and loud_a / quiet_a < loud_b / quiet_b
    # test_limiter_prevents_clipping: ceiling -1 dB (~0.89)
    z = limit(np.array([0.5, 1.5, -1.2, 0.8, 2.0, -0.3], np.float32), -1.0, 50.0)
    assert np.all(np.abs(z) <= 10.0 ** (-1.0 / 20.0) + 0.01)
    # test_limiter_preserves_quiet_audio
    q = np.array([0.1, -0.2, 0.15, -0.05], np.float32)
    assert np.all(np.abs(limit(q, -1.0, 50.0) - q) < 0.001)


def test_reference_unit_tests_hold_for_the_oracle():
    _reference_unit_tests(O.rms_db, O.normalize_rms, O.apply_gain, O.compress, O.limit)


def test_oracle_stage_details():
    # an empty clip has no level (audio.rs:90-92); compress with ratio <= 1 and limit on an empty clip leave it alone (:147, :201)
    assert O.rms_db(np.zeros(0, np.float32)) == -np.inf
    x = _loud_quiet()
    assert np.array_equal(O.compress(x, -20.0, 1.0, 5.0, 50.0, 6.0), x)
    # a silent clip is not normalised (audio.rs:119-121)
    assert np.array_equal(O.normalize_rms(np.zeros(100, np.float32), -18.0), np.zeros(100, np.float32))
    # the limiter's gain recovers with the release constant after a peak: exp(-1 / (50 ms * 16 samples/ms)) per sample
    y = O.limit(np.array([2.0] + [0.5] * 3, np.float32), 0.0, 50.0)
    r = np.float32(np.exp(np.float32(-1.0 / 800.0)))
    g1 = r * np.float32(0.5) + (np.float32(1.0) - r) * np.float32(1.0)
    assert y[0] == np.float32(1.0) and abs(y[1] - np.float32(0.5) * g1) < 1e-7
    # worker.rs:214-234: normalise, then compress, then limit — the composition of the stages in that order
    p = _abi.default_preprocess_params()
    got, (before, after) = O.preprocess_audio(x, p)
    want = O.limit(O.compress(O.normalize_rms(x, p.target_db), p.threshold_db, p.ratio, p.attack_ms, p.release_ms, p.makeup_gain_db), p.ceiling_db,
                   p.limiter_release_ms)
    assert np.array_equal(got, want) and before == np.float32(O.rms_db(x)) and after == np.float32(O.rms_db(want))
    assert np.abs(got).max() <= 10.0 ** (-1.0 / 20.0) + 1e-6


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "oracle_levels.npz"))


def test_oracle_reproduces_its_frozen_outputs():
    g, p = _golden(), _abi.default_preprocess_params()
    for name in ("speech", "tone"):
        y, rms = O.preprocess_audio(g[name + "_in"], p)
        np.testing.assert_allclose(y, g[name + "_out"], rtol=1e-6, atol=1e-8)  # libm's last bits may differ between hosts
        np.testing.assert_allclose(np.array(rms, np.float32), g[name + "_rms"], atol=1e-4)
    assert np.abs(g["tone_out"]).max() <= 10.0 ** (-1.0 / 20.0) + 1e-6 < np.abs(g["tone_in"]).max()


def test_params_mirror_the_reference_defaults_and_the_header():
    from openhush_b200 import lib
    p, q = lib.load().oh_preprocess_default_params(), _abi.default_preprocess_params()
    for name, _ in _abi.PreprocessParams._fields_:
        assert getattr(p, name) == getattr(q, name), name
    cfg = E.AudioConfig()  # config.rs:984: preprocessing off; :1239-1269: stages enabled once it is on
    assert cfg.preprocessing is False and cfg.normalization.enabled and cfg.compression.enabled and cfg.limiter.enabled
    r = cfg.to_params()
    for name, _ in _abi.PreprocessParams._fields_:
        assert getattr(r, name) == getattr(q, name), name
    hdr = open(os.path.join(ROOT, "include", "oh_whisper.h")).read()
    body = re.search(r"typedef struct oh_preprocess_params \{(.*?)\} oh_preprocess_params;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    assert re.findall(r"(?:float|int)\s+(\w+);", body) == [n for n, _ in _abi.PreprocessParams._fields_]


def test_worker_hook_is_off_by_default_and_refuses_rnnoise():
    class _State:
        def preprocess(self, clips, params):
            raise AssertionError("must not run")

    buf = E.AudioBuffer(_sine())
    assert E.preprocess_audio(_State(), [buf], E.AudioConfig()) is None  # preprocessing = false: untouched
    with pytest.raises(NotImplementedError):
        E.preprocess_audio(_State(), [buf], E.AudioConfig(noise_reduction=E.NoiseReductionConfig(enabled=True)))


# ------------------------------------------------------------------------------------------------------------ GPU
def _close(got, want):
    # the recurrences are the same f32 operations in the same order; what differs is the last bits of log10f / powf / expf
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-7)


@pytest.fixture(scope="module")
def st(micro_path):
    from openhush_b200 import COMPUTE_F32, lib
    ctx = lib.Context(micro_path, compute_type=COMPUTE_F32, max_batch=4)
    s = lib.State(ctx)
    yield s
    s.close()
    ctx.close()


@pytest.mark.gpu
def test_device_level_processing_matches_oracle(st):
    from openhush_b200 import synth
    speech = synth.audio(3, 160000)
    clips = [speech, _loud_quiet() * np.float32(1.5), np.zeros(4000, np.float32), speech[:17601] * np.float32(0.02), np.zeros(0, np.float32)]
    p = _abi.default_preprocess_params()
    got, rms = st.preprocess(clips, p)
    for i, c in enumerate(clips):
        want, (before, after) = O.preprocess_audio(c, p)
        assert len(got[i]) == len(c)
        _close(got[i], want)
        for g, w in ((rms[i, 0], before), (rms[i, 1], after)):
            assert (np.isinf(w) and np.isinf(g) and g < 0) or abs(g - w) < 1e-3, (i, g, w)
    assert np.abs(np.concatenate(got)).max() <= 10.0 ** (-1.0 / 20.0) + 1e-6
    # every stage on its own, and pairs, against the oracle's functions
    for flags in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (0, 0, 0)):
        q = _abi.default_preprocess_params()
        q.normalize, q.compress, q.limit = flags
        g2, _ = st.preprocess(clips[:2], q)
        for c, g in zip(clips[:2], g2):
            _close(g, O.preprocess_audio(c, q)[0])
    # ratio <= 1 switches the compressor off, make-up gain included (audio.rs:147)
    q = _abi.default_preprocess_params()
    q.normalize, q.limit, q.ratio = 0, 0, 1.0
    assert np.array_equal(st.preprocess(clips[:1], q)[0][0], clips[0])


@pytest.mark.gpu
def test_device_level_processing_matches_golden(st):
    g, p = _golden(), _abi.default_preprocess_params()
    got, rms = st.preprocess([g["speech_in"], g["tone_in"]], p)
    for i, name in enumerate(("speech", "tone")):
        _close(got[i], g[name + "_out"])
        np.testing.assert_allclose(rms[i], g[name + "_rms"], atol=1e-3)


@pytest.mark.gpu
def test_reference_unit_tests_hold_for_the_device_kernels(st):
    def one(**kw):
        q = _abi.default_preprocess_params()
        q.normalize = q.compress = q.limit = 0
        for k, v in kw.items():
            setattr(q, k, v)
        return q

    def rms_db(x):
        return float(st.preprocess([x], one())[1][0, 0])

    def normalize_rms(x, target):
        return st.preprocess([x], one(normalize=1, target_db=target))[0][0]

    def apply_gain(x, gain_db):  # the make-up gain of a compressor whose threshold is never reached is apply_gain (audio.rs:186-188)
        return st.preprocess([x], one(compress=1, threshold_db=60.0, ratio=2.0, makeup_gain_db=gain_db))[0][0]

    def compress(x, thr, ratio, att, rel, makeup):
        return st.preprocess([x], one(compress=1, threshold_db=thr, ratio=ratio, attack_ms=att, release_ms=rel, makeup_gain_db=makeup))[0][0]

    def limit(x, ceil, rel):
        return st.preprocess([x], one(limit=1, ceiling_db=ceil, limiter_release_ms=rel))[0][0]

    _reference_unit_tests(rms_db, normalize_rms, apply_gain, compress, limit)


@pytest.mark.gpu
def test_device_resident_variant_and_worker_hook(st, micro_path):
    import torch
    from openhush_b200 import COMPUTE_F32, synth
    p = _abi.default_preprocess_params()
    a = np.stack([synth.audio(5, 48000), synth.audio(6, 48000) * np.float32(0.1)])
    ns = [48000, 40000]
    d = torch.from_numpy(a.copy()).cuda()
    rms = st.preprocess_device(d.data_ptr(), a.shape[1], ns, p)
    torch.cuda.synchronize()
    out = d.cpu().numpy()
    host, rms_h = st.preprocess([a[0][:ns[0]], a[1][:ns[1]]], p)
    for i in range(2):
        assert np.array_equal(out[i, :ns[i]], host[i]) and np.array_equal(out[i, ns[i]:], a[i, ns[i]:])  # same kernels; the tail is left alone
    assert np.array_equal(rms, rms_h)
    # the worker runs the hook before the engine sees the buffer (worker.rs:150-154): same text as transcribing the
    # processed clip, a different buffer than the raw one, and that buffer is the oracle's within the kernels' tolerance
    # (fresh engines on both sides: a state's first call is what sampled fallback tokens are defined against, oh_whisper.h)
    eng = E.WhisperEngine(micro_path, "en", False, True, compute_type=COMPUTE_F32, max_batch=2)
    ref = E.WhisperEngine(micro_path, "en", False, True, compute_type=COMPUTE_F32, max_batch=1)
    try:
        clip = synth.audio(2, 80000) * np.float32(0.05)
        want = ref.transcribe(E.AudioBuffer(ref.state.preprocess([clip], p)[0][0]))
        worker = E.TranscriptionWorker(eng, E.AudioConfig(preprocessing=True))
        job = E.TranscriptionJob(E.AudioBuffer(clip.copy()), sequence_id=0)
        r = worker.process_job(job)
        assert r.text == want.text and [s[:2] for s in r.segments] == [s[:2] for s in want.segments]
        assert not np.array_equal(job.buffer.samples, clip)
        _close(job.buffer.samples, O.preprocess_audio(clip, p)[0])
        two = worker.process_jobs([E.TranscriptionJob(E.AudioBuffer(clip.copy()), 1), E.TranscriptionJob(E.AudioBuffer(clip.copy()), 2)])
        assert [t.text for t in two] == [want.text, want.text]
    finally:
        eng.close()
        ref.close()
