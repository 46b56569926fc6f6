"""Generates tests/golden/hf_micro.npz — outputs of HuggingFace transformers' Whisper (an independent
implementation of the same published architecture) on the synthetic `micro` model and clip seed 1.

The reference's own arithmetic (whisper.cpp inside whisper-rs-sys 0.11.1) is not available offline, and
the reference holds no golden vectors for this path, so this is the pin the oracle is
checked against in tests/test_oracle_vs_hf.py:
  * log-mel: WhisperFeatureExtractor vs oracle A.3, frames 0..2998 (frame 2999 differs by design:
    HF reflect-pads the right edge, whisper.cpp zero-pads — SURVEY.md A.3 probe);
  * encoder states and decoder logits: WhisperForConditionalGeneration (exact-erf GELU, fp32) vs the
    oracle in gelu=ERF / act_round=NONE mode, on the ORACLE's mel window so the stages are checked separately.

Run from the repo root (CPU is enough, ~20 s):  python tests/golden/make_hf_golden.py
"""
import os
import struct
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openhush_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def read_ggml(path):
    """Minimal reader of the legacy ggml container -> (hparams dict, {name: np.ndarray in torch shape})."""
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def rd(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, buf, off)
        off += struct.calcsize(fmt)
        return v

    assert rd("<I")[0] == 0x67676D6C
    names = ["n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer", "n_text_ctx", "n_text_state",
             "n_text_head", "n_text_layer", "n_mels", "ftype"]
    hp = dict(zip(names, rd("<11i")))
    n_mel, n_fft = rd("<2i")
    off += 4 * n_mel * n_fft
    for _ in range(rd("<i")[0]):
        ln = rd("<I")[0]
        off += ln
    tensors = {}
    while off < len(buf):
        n_dims, name_len, ttype = rd("<3i")
        ne = rd(f"<{n_dims}i")
        name = buf[off:off + name_len].decode()
        off += name_len
        n = int(np.prod(ne))
        dt = np.float16 if ttype == 1 else np.float32
        arr = np.frombuffer(buf, dtype=dt, count=n, offset=off).astype(np.float32).reshape(tuple(reversed(ne)))
        off += n * (2 if ttype == 1 else 4)
        tensors[name] = arr
    return hp, tensors


def hf_model(hp, T):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    d = hp["n_audio_state"]
    cfg = WhisperConfig(vocab_size=hp["n_vocab"], num_mel_bins=hp["n_mels"], encoder_layers=hp["n_audio_layer"],
                        encoder_attention_heads=hp["n_audio_head"], decoder_layers=hp["n_text_layer"],
                        decoder_attention_heads=hp["n_text_head"], d_model=d, encoder_ffn_dim=4 * d, decoder_ffn_dim=4 * d,
                        max_source_positions=hp["n_audio_ctx"], max_target_positions=hp["n_text_ctx"], activation_function="gelu",
                        scale_embedding=False, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                        pad_token_id=50256, bos_token_id=50257, eos_token_id=50256, decoder_start_token_id=50257)
    model = WhisperForConditionalGeneration(cfg).eval().float()
    sd = {}

    def put(k, v):
        sd[k] = torch.from_numpy(np.ascontiguousarray(v))

    put("model.encoder.conv1.weight", T["encoder.conv1.weight"])
    put("model.encoder.conv1.bias", T["encoder.conv1.bias"].reshape(-1))
    put("model.encoder.conv2.weight", T["encoder.conv2.weight"])
    put("model.encoder.conv2.bias", T["encoder.conv2.bias"].reshape(-1))
    put("model.encoder.embed_positions.weight", T["encoder.positional_embedding"])
    put("model.decoder.embed_positions.weight", T["decoder.positional_embedding"])
    put("model.decoder.embed_tokens.weight", T["decoder.token_embedding.weight"])
    put("proj_out.weight", T["decoder.token_embedding.weight"])
    put("model.encoder.layer_norm.weight", T["encoder.ln_post.weight"])
    put("model.encoder.layer_norm.bias", T["encoder.ln_post.bias"])
    put("model.decoder.layer_norm.weight", T["decoder.ln.weight"])
    put("model.decoder.layer_norm.bias", T["decoder.ln.bias"])
    amap = {"query": "q_proj", "key": "k_proj", "value": "v_proj", "out": "out_proj"}
    for side, n_layer in (("encoder", hp["n_audio_layer"]), ("decoder", hp["n_text_layer"])):
        for l in range(n_layer):
            s, h = f"{side}.blocks.{l}", f"model.{side}.layers.{l}"
            for wb in ("weight", "bias"):
                put(f"{h}.self_attn_layer_norm.{wb}", T[f"{s}.attn_ln.{wb}"])
                put(f"{h}.final_layer_norm.{wb}", T[f"{s}.mlp_ln.{wb}"])
                put(f"{h}.fc1.{wb}", T[f"{s}.mlp.0.{wb}"])
                put(f"{h}.fc2.{wb}", T[f"{s}.mlp.2.{wb}"])
                if side == "decoder":
                    put(f"{h}.encoder_attn_layer_norm.{wb}", T[f"{s}.cross_attn_ln.{wb}"])
            for a, hname in (("attn", "self_attn"),) + ((("cross_attn", "encoder_attn"),) if side == "decoder" else ()):
                for k, v in amap.items():
                    put(f"{h}.{hname}.{v}.weight", T[f"{s}.{a}.{k}.weight"])
                    if k != "key":
                        put(f"{h}.{hname}.{v}.bias", T[f"{s}.{a}.{k}.bias"])
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("k_proj.bias" in m for m in missing) or not missing, missing
    return model


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    arch = "micro"
    path = synth.model_path(arch)
    hp, T = read_ggml(path)
    pcm = synth.audio(1)

    # 1. mel by HF's feature extractor
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor(feature_size=hp["n_mels"])
    hf_mel = fe(pcm, sampling_rate=16000, return_tensors="np")["input_features"][0].astype(np.float32)  # [n_mel, 3000]

    # 2. network on the ORACLE's mel window
    om = O.OracleModel(path)
    st = O.OracleState(om, gelu=O.GELU_ERF, act_round=O.ACT_NONE)
    omel, _ = st.pcm_to_mel(pcm)
    window = np.ascontiguousarray(omel[:, :3000])
    model = hf_model(hp, T)
    prompt = [om.sot, om.sot + 1, om.transcribe, 1541, 5163, om.beg + 40, om.beg + 40, 777]
    with torch.no_grad():
        enc = model.model.encoder(torch.from_numpy(window)[None]).last_hidden_state[0].numpy()
        out = model(input_features=torch.from_numpy(window)[None], decoder_input_ids=torch.tensor([prompt]))
        logits = out.logits[0].numpy()  # [len(prompt), V]
    rows = np.arange(0, 1500, 13)
    cols = np.arange(0, hp["n_vocab"], 17)
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "hf_micro.npz"),
        arch=arch, seed=1, model_seed=1234, prompt=np.array(prompt, dtype=np.int32),
        hf_mel_rows=np.arange(0, hp["n_mels"], 3), hf_mel=hf_mel[::3, :],
        enc_rows=rows, enc=enc[rows], logit_cols=cols, logits_last=logits[-1][cols], logits_pos2=logits[2][cols],
        logits_last_argmax=int(logits[-1].argmax()), logits_last_max=float(logits[-1].max()),
        top_ids=np.argsort(-logits[-1])[:16].astype(np.int32), top_vals=np.sort(logits[-1])[::-1][:16].copy())
    print("hf mel", hf_mel.shape, "enc", enc.shape, enc.std(), "logits", logits.shape, logits[-1].std(), logits[-1].argmax())
    # report agreement now (the test re-checks it from the fixture)
    o_enc = st.encode(window)
    print("mel   max|d| frames 0..2998:", np.abs(hf_mel[:, :2999] - omel[:, :2999]).max(), " frame 2999:", np.abs(hf_mel[:, 2999] - omel[:, 2999]).max())
    print("enc   max|d|:", np.abs(o_enc - enc).max(), "rel L2:", np.linalg.norm(o_enc - enc) / np.linalg.norm(enc))
    o_log = st.decode(prompt, 0)
    print("logit max|d|:", np.abs(o_log - logits[-1]).max(), "argmax", o_log.argmax(), logits[-1].argmax())


if __name__ == "__main__":
    main()
