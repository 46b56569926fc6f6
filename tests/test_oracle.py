"""CPU tests of the oracle (the checker) against known answers and the HF-generated golden vectors.
SURVEY.md §8(c): the reference pins nothing for this path, so these are the pins."""
import os

import numpy as np
import pytest

from openhush_b200 import default_full_params, synth
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_special_token_ids():
    # V = beg + 1501 in all three vocabularies; ids agree with OpenAI's tokenizers
    for arch, exp in (("tiny.en", (50256, 50257, 50357, 50358, 50359, 50360, 50361, 50362, 50363)),
                      ("micro", (50257, 50258, 50358, 50359, 50360, 50361, 50362, 50363, 50364)),
                      ("micro-v3", (50257, 50258, 50359, 50360, 50361, 50362, 50363, 50364, 50365))):
        if arch == "tiny.en":
            a = synth.Arch("micro.en", 128, 2, 2, 3, 80, 51864)
            p = "/tmp/openhush_b200_models/ggml-micro.en.bin"
            os.makedirs(os.path.dirname(p), exist_ok=True)
            if not os.path.exists(p):
                synth.write_model(p, a)
        else:
            p = synth.model_path(arch)
        m = O.OracleModel(p)
        got = (m.eot, m.sot, m.translate, m.transcribe, m.solm, m.prev, m.nosp, m.not_, m.beg)
        assert got == exp, (arch, got)
        assert m.n_vocab == m.beg + 1501
        assert m.blank == 220 and m.token_to_str(220) == b" "
        m.close()


def test_mel_filterbank_matches_transformers():
    from transformers.audio_utils import mel_filter_bank
    for n_mel in (80, 128):
        ref = mel_filter_bank(201, n_mel, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T.astype(np.float32)
        got = synth.mel_filters(n_mel)
        assert got.shape == ref.shape == (n_mel, 201)
        assert np.abs(got - ref).max() < 1e-7


def test_model_file_carries_filterbank(micro_oracle):
    assert np.array_equal(micro_oracle.mel_filters(), synth.mel_filters(80))


def test_mel_shapes_and_silence(micro_oracle):
    st = O.OracleState(micro_oracle)
    mel, n_len_org = st.pcm_to_mel(np.zeros(480000, np.float32))
    assert mel.shape == (80, 6000) and n_len_org == 2999
    assert np.all(mel == -1.5)  # log10(1e-10) = -10 everywhere; (−10+4)/4
    mel, n_len_org = st.pcm_to_mel(synth.audio(3, 17600))  # the reference's 1.1 s minimum, src/input/audio.rs:31-34
    assert mel.shape[1] == (17600 + 480000) // 160 and n_len_org == 109
    _, n_len_org = st.pcm_to_mel(synth.audio(3, 15999))
    assert n_len_org == 99  # < 100 frames: full() returns 0 segments
    st.close()


def test_mel_padding_floor(micro_oracle, clip1):
    st = O.OracleState(micro_oracle)
    mel, _ = st.pcm_to_mel(clip1)
    floor = mel[:, 3100:].max()
    assert np.all(mel[:, 3100:] == floor)  # zero-padding frames sit at the clamp floor ((max-8)+4)/4, not at -1.5
    assert abs(floor - (mel.max() - 2.0)) < 1e-6
    st.close()


def test_mel_matches_hf_feature_extractor(micro_oracle, clip1):
    g = np.load(os.path.join(GOLD, "hf_micro.npz"))
    st = O.OracleState(micro_oracle)
    mel, _ = st.pcm_to_mel(clip1)
    rows = g["hf_mel_rows"]
    d = np.abs(mel[rows, :2999] - g["hf_mel"][:, :2999]).max()
    assert d < 1e-4, d
    # frame 2999 differs by design: zero vs reflect right padding
    assert np.abs(mel[rows, 2999] - g["hf_mel"][:, 2999]).max() > 1e-3
    st.close()


def test_network_matches_hf(micro_oracle, clip1):
    g = np.load(os.path.join(GOLD, "hf_micro.npz"))
    st = O.OracleState(micro_oracle, gelu=O.GELU_ERF, act_round=O.ACT_NONE)
    mel, _ = st.pcm_to_mel(clip1)
    enc = st.encode(np.ascontiguousarray(mel[:, :3000]))
    ref = g["enc"]
    got = enc[g["enc_rows"]]
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-4
    assert np.abs(got - ref).max() < 1e-3
    prompt = g["prompt"]
    lg2 = st.decode(prompt[:3], 0)
    assert np.abs(lg2[g["logit_cols"]] - g["logits_pos2"]).max() < 2e-3
    lg = st.decode(prompt[3:], 3)  # continues on the KV cache
    assert np.abs(lg[g["logit_cols"]] - g["logits_last"]).max() < 2e-3
    assert int(lg.argmax()) == int(g["logits_last_argmax"])
    assert list(np.argsort(-lg)[:8]) == list(g["top_ids"][:8])
    st.close()


def test_ggml_numerics_knobs_are_small_perturbations(micro_oracle, clip1):
    st = O.OracleState(micro_oracle, gelu=O.GELU_ERF, act_round=O.ACT_NONE)
    mel, _ = st.pcm_to_mel(clip1)
    w = np.ascontiguousarray(mel[:, :3000])
    e0 = st.encode(w)
    st.set_options(O.GELU_TANH_F16LUT, O.ACT_F16)
    e1 = st.encode(w)
    rel = np.linalg.norm(e1 - e0) / np.linalg.norm(e0)
    assert 1e-5 < rel < 2e-2, rel
    st.close()


# ---- A.6 / A.7 known answers -------------------------------------------------------------------

def _flat(m, v=0.0):
    return np.full(m.n_vocab, v, np.float32)


def test_logits_always_suppressed(micro_oracle):
    m = micro_oracle
    lg = _flat(m)
    lg[m.beg:] = -30.0  # keep the timestamp-mass rule out of this case
    r = m.process_logits(lg, tokens_cur=[100])
    for t in (m.not_, m.sot, m.nosp, m.solm, m.translate, m.transcribe, m.prev, m.sot + 1, m.sot + 99):
        assert r["probs"][t] == 0.0 and r["logits"][t] == -np.inf
    assert r["probs"][m.eot] > 0


def test_suppress_blank_and_max_initial_ts(micro_oracle):
    m = micro_oracle
    lg = _flat(m)
    lg[m.beg + 50] = 5.0
    lg[m.beg + 51] = 50.0  # would win, but > 1.0 s on the first step
    lg[m.eot] = 40.0
    lg[m.blank] = 40.0
    r = m.process_logits(lg, tokens_cur=[])
    assert r["probs"][m.eot] == 0 and r["probs"][m.blank] == 0
    assert np.all(r["probs"][m.beg + 51:] == 0)
    assert r["id"] == m.beg + 50
    r = m.process_logits(lg, tokens_cur=[7])  # not initial any more
    assert r["id"] == m.beg + 51


def test_timestamp_pairing(micro_oracle):
    m = micro_oracle
    lg = _flat(m)
    lg[m.eot] = 20.0  # EOT is a "text" id (< beg): keep it above the timestamp mass so rule 10 does not mask it
    # text, ts -> next must be ts or EOT (all ids < eot masked)
    r = m.process_logits(lg, tokens_cur=[5, m.beg + 10], has_ts=True, seek_delta=20)
    assert np.all(r["probs"][:m.eot] == 0) and r["probs"][m.eot] > 0 and r["probs"][m.beg + 10] > 0
    assert np.all(r["probs"][m.beg:m.beg + 10] == 0)  # monotone: ids in [beg, beg + seek_delta/2) masked
    # ts, ts -> next cannot be a timestamp
    r = m.process_logits(lg, tokens_cur=[m.beg + 10, m.beg + 10], has_ts=True, seek_delta=20)
    assert np.all(r["probs"][m.beg:] == 0) and r["probs"][5] > 0
    # single token that is a ts: penultimate counts as ts (len < 2)
    r = m.process_logits(lg, tokens_cur=[m.beg + 3], has_ts=True, seek_delta=6)
    assert np.all(r["probs"][m.beg:] == 0)


def test_timestamp_mass_rule(micro_oracle):
    m = micro_oracle
    lg = _flat(m, -20.0)
    lg[100] = 0.0                       # best text token
    lg[m.beg + 60:m.beg + 70] = -1.5    # ten timestamps: mass 10*e^-1.5 = 2.23 > 1
    r = m.process_logits(lg, tokens_cur=[5])
    assert np.all(r["probs"][:m.beg] == 0) and r["id"] == m.beg + 60  # first index wins the tie
    lg[m.beg + 60:m.beg + 70] = -3.0    # mass 0.50 < 1
    r = m.process_logits(lg, tokens_cur=[5])
    assert r["id"] == 100 and r["tid"] == m.beg + 60
    assert abs(r["ptsum"] - r["probs"][m.beg:].sum()) < 1e-6


def test_argmax_first_index_tie_break_and_temperature(micro_oracle):
    m = micro_oracle
    lg = _flat(m)
    lg[m.beg:] = -30.0
    lg[300] = lg[200] = 3.0
    r = m.process_logits(lg, tokens_cur=[5])
    assert r["id"] == 200
    r2 = m.process_logits(lg, tokens_cur=[5], temperature=0.5)
    assert abs(r2["logits"][200] - 6.0) < 1e-6
    assert abs(np.exp(r["logprobs"][np.isfinite(r["logprobs"])]).sum() - 1.0) < 1e-3  # f32 sequential sum, as upstream


# ---- full() ------------------------------------------------------------------------------------

def test_full_short_audio_returns_no_segments(micro_oracle):
    st = O.OracleState(micro_oracle)
    assert st.full(synth.audio(2, 15999)) == 0 and st.segments() == []
    assert st.full(synth.audio(2, 100)) == -1
    st.close()


def test_full_golden(micro_oracle):
    """Token ids / timestamps of the oracle itself, frozen (tests/golden/make_oracle_golden.py)."""
    g = np.load(os.path.join(GOLD, "oracle_micro_full.npz"), allow_pickle=False)
    st = O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
    for k, seed in enumerate(g["seeds"]):
        p = default_full_params()
        p.temperature_inc = float(g["temperature_inc"])
        assert st.full(synth.audio(int(seed), int(g["n_samples"][k])), p) == 0
        segs = st.segments()
        toks = [t for s in segs for t in s["tokens"]]
        times = [x for s in segs for x in (s["t0"], s["t1"])]
        lo, hi = g["tok_off"][k], g["tok_off"][k + 1]
        assert toks == list(g["tokens"][lo:hi]), seed
        lo, hi = g["seg_off"][k], g["seg_off"][k + 1]
        assert times == list(g["times"][2 * lo:2 * hi]), seed
    st.close()


def test_full_segments_are_well_formed(micro_oracle, clip1):
    m = micro_oracle
    st = O.OracleState(m)
    p = default_full_params()
    assert st.full(clip1, p) == 0
    segs = st.segments()
    assert segs
    for s in segs:
        assert s["t0"] <= s["t1"]
        text = b"".join(m.token_to_str(t) for t in s["tokens"] if t < m.eot)
        assert text == s["text"]
    stats = st.stats()
    assert stats["n_encode"] == len(stats["seeks"]) >= 1 and stats["seeks"][0] == 0
    assert all(b > a for a, b in zip(stats["seeks"], stats["seeks"][1:]))
    st.close()


def test_beam_search_restatement(micro_oracle, clip1):
    """A.5 / A.7 beam: one beam is greedy decoding; five beams are deterministic, well-formed, and never score below the
    greedy sequence they contain as a candidate at step 0 (unpinned by the reference, which never selects beam search)."""
    ost = O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
    g = default_full_params()
    g.temperature_inc = 0.0
    assert ost.full(clip1, g) == 0
    greedy = [(s["t0"], s["t1"], s["tokens"]) for s in ost.segments()]
    b = default_full_params(1)
    b.temperature_inc = 0.0
    b.beam_size = 1
    assert ost.full(clip1, b) == 0
    assert [(s["t0"], s["t1"], s["tokens"]) for s in ost.segments()] == greedy
    b.beam_size = 5
    assert ost.full(clip1, b) == 0
    first = [(s["t0"], s["t1"], s["tokens"]) for s in ost.segments()]
    assert ost.full(clip1, b) == 0
    assert [(s["t0"], s["t1"], s["tokens"]) for s in ost.segments()] == first and len(first) > 0
    last = -1
    for t0, t1, toks in first:
        assert t0 >= 0 and t1 >= last
        last = t1
    ost.close()


def test_beam_golden(micro_oracle):
    """The frozen beam-search outputs (tests/golden/make_oracle_beam_golden.py): later edits of the oracle cannot drift."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_micro_beam.npz"))
    ost = O.OracleState(micro_oracle, gelu=O.GELU_TANH, act_round=O.ACT_NONE)
    p = default_full_params(1)
    p.temperature_inc = 0.0
    p.beam_size = int(g["beam_size"])
    for k, (seed, n) in enumerate(zip(g["seeds"], g["n_samples"])):
        assert ost.full(synth.audio(int(seed), int(n)), p) == 0
        segs = ost.segments()
        assert [t for s in segs for t in s["tokens"]] == list(g["tokens"][g["tok_off"][k]:g["tok_off"][k + 1]]), k
        assert [x for s in segs for x in (s["t0"], s["t1"])] == list(g["times"][2 * g["seg_off"][k]:2 * g["seg_off"][k + 1]]), k
    ost.close()
