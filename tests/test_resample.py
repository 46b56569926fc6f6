"""Pre-engine DSP: the oracle's restatement of the reference's load_wav_file steps
(/root/reference/src/input/audio.rs:368-424, resample_sinc :1007-1095) against the reference's own tests for them
(audio.rs:1166-1187), and the CUDA kernel against the oracle."""
import numpy as np
import pytest

from oracle import oracle as O


def _sine(rate, secs=1.0, hz=440.0, amp=1.0):
    return (amp * np.sin(2 * np.pi * hz * np.arange(int(rate * secs)) / rate)).astype(np.float32)


def test_reference_resample_tests_hold_for_the_oracle():
    # audio.rs:1166-1170 test_resample_same_rate: identity
    x = np.array([1.0, 2.0, 3.0, 4.0], np.float32)
    assert np.array_equal(O.load_pcm_16k(x, 16000, pad=False), x)
    # audio.rs:1180-1187 test_resample_downsample_sinc: 4800 samples 48 kHz -> roughly a third, within 100 of 1600
    y = O.load_pcm_16k(np.arange(4800, dtype=np.float32), 48000, pad=False)
    assert len(y) < 4800 and abs(len(y) - 1600) < 100
    # a ramp stays a ramp of three times the slope in the interior (unit DC gain of every interpolation table)
    assert np.allclose(np.diff(y[200:-200]), 3.0, atol=1e-2)


def test_oracle_resampler_properties():
    s = _sine(48000)
    y = O.load_pcm_16k(s, 48000, pad=False)
    assert abs(len(y) - 16000) < 100
    assert abs(np.sqrt(np.mean(y[500:-500] ** 2)) - 2 ** -0.5) < 1e-3         # pass-band gain 1 at 440 Hz
    # content above the new Nyquist is removed (cutoff 0.95 * 8 kHz), content below passes
    hi = O.load_pcm_16k(_sine(48000, hz=11000.0), 48000, pad=False)
    assert np.sqrt(np.mean(hi[500:-500] ** 2)) < 1e-3
    # stereo: mean of the channels (audio.rs:384-391); int16: s / 32768 (audio.rs:373-380)
    st = np.stack([s, 0.5 * s], 1)
    assert np.allclose(O.load_pcm_16k(st, 48000, pad=False), 0.75 * y, atol=1e-6)
    i16 = np.round(s * 32767).astype(np.int16)
    assert np.allclose(O.load_pcm_16k(i16, 48000, pad=False), O.load_pcm_16k(i16.astype(np.float32) / 32768.0, 48000, pad=False), atol=1e-7)
    # 44.1 kHz and upsampling from 8 kHz
    assert abs(len(O.load_pcm_16k(_sine(44100), 44100, pad=False)) - 16000) < 100
    up = O.load_pcm_16k(_sine(8000, hz=300.0), 8000, pad=False)
    assert abs(len(up) - 16000) < 600 and abs(np.sqrt(np.mean(up[1000:-1000] ** 2)) - 2 ** -0.5) < 2e-3
    # audio.rs:415-424: clips shorter than 1.1 s are padded with silence to 17600 samples
    short = O.load_pcm_16k(_sine(48000, secs=0.5), 48000)
    assert len(short) == 17600 and np.all(short[8100:] == 0.0)
    assert len(O.load_pcm_16k(_sine(16000, secs=2.0), 16000)) == 32000


@pytest.mark.gpu
def test_device_resampler_matches_oracle(micro_path):
    from openhush_b200 import COMPUTE_F32, default_full_params, lib
    ctx = lib.Context(micro_path, compute_type=COMPUTE_F32, max_batch=1)
    st = lib.State(ctx)
    rng = np.random.RandomState(7)
    cases = [
        (_sine(48000, secs=2.3) + 0.1 * rng.randn(int(48000 * 2.3)).astype(np.float32), 48000, 0),
        (np.stack([_sine(44100, secs=3.1, hz=700.0), _sine(44100, secs=3.1, hz=90.0)], 1), 44100, 0),
        (np.round(8000 * np.stack([_sine(44100, secs=1.7), _sine(44100, secs=1.7, hz=1234.5)], 1)).astype(np.int16), 44100, 0),
        (_sine(8000, secs=1.5, hz=300.0), 8000, 0),
        (np.round((2 ** 22) * np.stack([_sine(96000, secs=0.4, hz=h) for h in (200.0, 3000.0, 5000.0)], 1)).astype(np.int32), 96000, 24),
        (_sine(48000, secs=0.3), 48000, 0),                   # shorter than 1.1 s after resampling: padded
        (rng.randn(1000).astype(np.float32), 22050, 0),       # a single partial chunk
        (_sine(16000, secs=1.2), 16000, 0),                   # already 16 kHz: conversion only
        (np.round(20000 * _sine(16000, secs=1.2))[:, None].astype(np.int16).repeat(2, 1), 16000, 0),
    ]
    for frames, rate, bits in cases:
        ref = O.load_pcm_16k(frames, rate, bits)
        got = st.load_pcm_16k(frames, rate, bits)
        assert len(got) == len(ref) == st.lib.oh_pcm_16k_len(rate, frames.shape[0]), (rate, frames.shape)
        assert np.abs(got - ref).max() < 1e-5, (rate, frames.shape, np.abs(got - ref).max())
    # the resampled clip is what the engine then transcribes: same result from the oracle's and the device's clip
    from openhush_b200 import synth
    clip48 = np.repeat(synth.audio(3, 160000), 3)  # crude 48 kHz version of a 10 s clip
    a, b = st.load_pcm_16k(clip48, 48000), O.load_pcm_16k(clip48, 48000)
    p = default_full_params()
    p.temperature_inc = 0.0
    assert st.full(a, p) == 0
    ra = [(s["t0"], s["t1"], s["tokens"]) for s in st.segments(0)]
    assert st.full(b, p) == 0
    assert [(s["t0"], s["t1"], s["tokens"]) for s in st.segments(0)] == ra and len(ra) > 0  # (log-probs differ by the clips' 1e-5)
    st.close()
    ctx.close()
