// ORACLE (test infrastructure) — model-file reader.  Restates SURVEY.md Appendix A.1: the legacy
// "ggml" container that the reference names `ggml-<model>.bin` (/root/reference/src/engine/whisper.rs:71-79)
// and loads through whisper-rs at /root/reference/src/engine/whisper.rs:159.
#include "orc_common.h"

namespace orc {

namespace {

struct RawTensor {
    int n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    int ttype = 0;
    std::vector<uint8_t> data;
    int64_t nelem() const { return ne[0] * ne[1] * ne[2] * ne[3]; }
};

struct Reader {
    FILE* f;
    bool ok = true;
    template <class T>
    T rd() {
        T v{};
        if (fread(&v, sizeof(T), 1, f) != 1) ok = false;
        return v;
    }
};

std::vector<float> to_f32(const RawTensor& t) {
    std::vector<float> out((size_t)t.nelem());
    if (t.ttype == 0) std::memcpy(out.data(), t.data.data(), out.size() * 4);
    else {
        const uint16_t* h = (const uint16_t*)t.data.data();
        for (size_t i = 0; i < out.size(); ++i) out[i] = f16_to_f32(h[i]);
    }
    return out;
}

Mat to_mat(const RawTensor& t, int rows, int cols) {
    Mat m;
    m.rows = rows;
    m.cols = cols;
    m.is_f16 = t.ttype == 1;
    if (m.is_f16) {
        m.h.resize((size_t)rows * cols);
        std::memcpy(m.h.data(), t.data.data(), m.h.size() * 2);
    } else {
        m.f.resize((size_t)rows * cols);
        std::memcpy(m.f.data(), t.data.data(), m.f.size() * 4);
    }
    return m;
}

// conv weight: torch [out][C][3] -> [out][tap*C + c] so that a conv is one mat-mul over an im2col row
Mat conv_to_mat(const RawTensor& t, int out, int C) {
    Mat m;
    m.rows = out;
    m.cols = 3 * C;
    m.is_f16 = t.ttype == 1;
    if (m.is_f16) {
        m.h.resize((size_t)out * 3 * C);
        const uint16_t* s = (const uint16_t*)t.data.data();
        for (int o = 0; o < out; ++o)
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < 3; ++k) m.h[(size_t)o * 3 * C + k * C + c] = s[((size_t)o * C + c) * 3 + k];
    } else {
        m.f.resize((size_t)out * 3 * C);
        const float* s = (const float*)t.data.data();
        for (int o = 0; o < out; ++o)
            for (int c = 0; c < C; ++c)
                for (int k = 0; k < 3; ++k) m.f[(size_t)o * 3 * C + k * C + c] = s[((size_t)o * C + c) * 3 + k];
    }
    return m;
}

}  // namespace

static const char* k_lang_codes[] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
    "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
    "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
    "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue"};

// language table order: /root/reference/src/engine/whisper.rs:627-731 (0=en ... 98=su), +yue for 100-language vocabularies
int lang_id_from_code(const char* code) {
    for (int i = 0; i < 100; ++i)
        if (std::strcmp(code, k_lang_codes[i]) == 0) return i;
    return -1;
}

Model* load_model(const char* path, std::string* err) {
    auto fail = [&](const std::string& e) -> Model* {
        if (err) *err = e;
        return nullptr;
    };
    FILE* f = fopen(path, "rb");
    if (!f) return fail(std::string("cannot open ") + path);
    Reader r{f};
    std::unique_ptr<Model> m(new Model());
    if (r.rd<uint32_t>() != 0x67676d6cu) {
        fclose(f);
        return fail("bad magic");
    }
    HParams& hp = m->hp;
    hp.n_vocab = r.rd<int32_t>();
    hp.n_audio_ctx = r.rd<int32_t>();
    hp.n_audio_state = r.rd<int32_t>();
    hp.n_audio_head = r.rd<int32_t>();
    hp.n_audio_layer = r.rd<int32_t>();
    hp.n_text_ctx = r.rd<int32_t>();
    hp.n_text_state = r.rd<int32_t>();
    hp.n_text_head = r.rd<int32_t>();
    hp.n_text_layer = r.rd<int32_t>();
    hp.n_mels = r.rd<int32_t>();
    hp.ftype = r.rd<int32_t>();
    if (!r.ok || hp.n_audio_state <= 0 || hp.n_audio_state % hp.n_audio_head || hp.n_text_state % hp.n_text_head) {
        fclose(f);
        return fail("bad hparams");
    }

    int fm = r.rd<int32_t>(), ff = r.rd<int32_t>();
    if (fm != hp.n_mels || ff != 201) {
        fclose(f);
        return fail("bad mel filter dims");
    }
    m->mel_filters.resize((size_t)fm * ff);
    if (fread(m->mel_filters.data(), 4, m->mel_filters.size(), f) != m->mel_filters.size()) r.ok = false;

    // vocabulary
    Vocab& vc = m->vocab;
    vc.n_vocab = hp.n_vocab;
    int n_in_file = r.rd<int32_t>();
    vc.id_to_token.resize(hp.n_vocab);
    for (int i = 0; i < n_in_file && r.ok; ++i) {
        uint32_t len = r.rd<uint32_t>();
        std::string s(len, '\0');
        if (len && fread(&s[0], 1, len, f) != len) r.ok = false;
        if (i < hp.n_vocab) {
            vc.id_to_token[i] = s;
            vc.token_to_id[s] = i;
        }
    }
    if (vc.is_multilingual()) {
        vc.token_eot++;
        vc.token_sot++;
        const int dt = vc.num_languages() - 98;
        vc.token_translate += dt;
        vc.token_transcribe += dt;
        vc.token_solm += dt;
        vc.token_prev += dt;
        vc.token_nosp += dt;
        vc.token_not += dt;
        vc.token_beg += dt;
    }
    for (int i = n_in_file; i < hp.n_vocab; ++i) {
        char buf[64];
        if (i > vc.token_beg) snprintf(buf, sizeof buf, "[_TT_%d]", i - vc.token_beg);
        else if (i == vc.token_eot) snprintf(buf, sizeof buf, "[_EOT_]");
        else if (i == vc.token_sot) snprintf(buf, sizeof buf, "[_SOT_]");
        else if (i == vc.token_translate) snprintf(buf, sizeof buf, "[_TRANSLATE_]");
        else if (i == vc.token_transcribe) snprintf(buf, sizeof buf, "[_TRANSCRIBE_]");
        else if (i == vc.token_solm) snprintf(buf, sizeof buf, "[_SOLM_]");
        else if (i == vc.token_prev) snprintf(buf, sizeof buf, "[_PREV_]");
        else if (i == vc.token_nosp) snprintf(buf, sizeof buf, "[_NOSP_]");
        else if (i == vc.token_not) snprintf(buf, sizeof buf, "[_NOT_]");
        else if (i == vc.token_beg) snprintf(buf, sizeof buf, "[_BEG_]");
        else if (i > vc.token_sot && i <= vc.token_sot + vc.num_languages())
            snprintf(buf, sizeof buf, "[_LANG_%s]", k_lang_codes[std::min(i - vc.token_sot - 1, 99)]);
        else snprintf(buf, sizeof buf, "[_extra_token_%d]", i);
        vc.id_to_token[i] = buf;
        vc.token_to_id[buf] = i;
    }

    // tensors
    std::map<std::string, RawTensor> T;
    while (r.ok) {
        int32_t n_dims;
        if (fread(&n_dims, 4, 1, f) != 1) break;  // EOF
        int32_t name_len = r.rd<int32_t>();
        int32_t ttype = r.rd<int32_t>();
        if (!r.ok || n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 255 || (ttype != 0 && ttype != 1)) {
            fclose(f);
            return fail("bad tensor header (quantised ftypes are not supported)");
        }
        RawTensor t;
        t.n_dims = n_dims;
        t.ttype = ttype;
        for (int i = 0; i < n_dims; ++i) t.ne[i] = r.rd<int32_t>();
        std::string name(name_len, '\0');
        if (fread(&name[0], 1, name_len, f) != (size_t)name_len) r.ok = false;
        t.data.resize((size_t)t.nelem() * (ttype ? 2 : 4));
        if (fread(t.data.data(), 1, t.data.size(), f) != t.data.size()) r.ok = false;
        T[name] = std::move(t);
    }
    fclose(f);
    if (!r.ok) return fail("truncated file");

    std::string missing;
    auto need = [&](const std::string& n, int64_t nelem) -> const RawTensor* {
        auto it = T.find(n);
        if (it == T.end()) {
            missing = "missing tensor " + n;
            return nullptr;
        }
        if (it->second.nelem() != nelem) {
            missing = "wrong size for tensor " + n;
            return nullptr;
        }
        return &it->second;
    };
#define NEED(var, name, nelem)                 \
    const RawTensor* var = need(name, nelem);  \
    if (!var) return fail(missing);

    const int d = hp.n_audio_state, dt = hp.n_text_state, V = hp.n_vocab;
    {
        NEED(t, "encoder.positional_embedding", (int64_t)hp.n_audio_ctx * d);
        m->enc_pos = to_f32(*t);
    }
    {
        NEED(w1, "encoder.conv1.weight", (int64_t)d * hp.n_mels * 3);
        NEED(b1, "encoder.conv1.bias", d);
        NEED(w2, "encoder.conv2.weight", (int64_t)d * d * 3);
        NEED(b2, "encoder.conv2.bias", d);
        m->conv1 = conv_to_mat(*w1, d, hp.n_mels);
        m->conv2 = conv_to_mat(*w2, d, d);
        m->conv1_b = to_f32(*b1);
        m->conv2_b = to_f32(*b2);
    }
    auto load_attn = [&](const std::string& p, int dout, int din_kv, AttnW& a) -> bool {
        const RawTensor *q = need(p + ".query.weight", (int64_t)dout * dout), *qb = need(p + ".query.bias", dout);
        const RawTensor *k = need(p + ".key.weight", (int64_t)dout * din_kv);
        const RawTensor *v = need(p + ".value.weight", (int64_t)dout * din_kv), *vb = need(p + ".value.bias", dout);
        const RawTensor *o = need(p + ".out.weight", (int64_t)dout * dout), *ob = need(p + ".out.bias", dout);
        if (!q || !qb || !k || !v || !vb || !o || !ob) return false;
        a.q = to_mat(*q, dout, dout);
        a.k = to_mat(*k, dout, din_kv);
        a.v = to_mat(*v, dout, din_kv);
        a.o = to_mat(*o, dout, dout);
        a.qb = to_f32(*qb);
        a.vb = to_f32(*vb);
        a.ob = to_f32(*ob);
        return true;
    };
    auto load_ln = [&](const std::string& p, int dim, std::vector<float>& w, std::vector<float>& b) -> bool {
        const RawTensor *tw = need(p + ".weight", dim), *tb = need(p + ".bias", dim);
        if (!tw || !tb) return false;
        w = to_f32(*tw);
        b = to_f32(*tb);
        return true;
    };
    auto load_mlp = [&](const std::string& p, int dim, Mat& m0, std::vector<float>& b0, Mat& m2, std::vector<float>& b2) -> bool {
        const RawTensor *w0 = need(p + ".mlp.0.weight", (int64_t)4 * dim * dim), *bb0 = need(p + ".mlp.0.bias", 4 * dim);
        const RawTensor *w2 = need(p + ".mlp.2.weight", (int64_t)4 * dim * dim), *bb2 = need(p + ".mlp.2.bias", dim);
        if (!w0 || !bb0 || !w2 || !bb2) return false;
        m0 = to_mat(*w0, 4 * dim, dim);
        m2 = to_mat(*w2, dim, 4 * dim);
        b0 = to_f32(*bb0);
        b2 = to_f32(*bb2);
        return true;
    };

    m->enc.resize(hp.n_audio_layer);
    for (int l = 0; l < hp.n_audio_layer; ++l) {
        std::string p = "encoder.blocks." + std::to_string(l);
        EncLayer& L = m->enc[l];
        if (!load_ln(p + ".attn_ln", d, L.ln1_w, L.ln1_b) || !load_attn(p + ".attn", d, d, L.attn) ||
            !load_ln(p + ".mlp_ln", d, L.ln2_w, L.ln2_b) || !load_mlp(p, d, L.mlp0, L.mlp0_b, L.mlp2, L.mlp2_b))
            return fail(missing);
    }
    if (!load_ln("encoder.ln_post", d, m->ln_post_w, m->ln_post_b)) return fail(missing);

    {
        NEED(t, "decoder.positional_embedding", (int64_t)hp.n_text_ctx * dt);
        m->dec_pos = to_f32(*t);
        NEED(e, "decoder.token_embedding.weight", (int64_t)V * dt);
        m->tok_emb = to_mat(*e, V, dt);
    }
    m->dec.resize(hp.n_text_layer);
    for (int l = 0; l < hp.n_text_layer; ++l) {
        std::string p = "decoder.blocks." + std::to_string(l);
        DecLayer& L = m->dec[l];
        if (!load_ln(p + ".attn_ln", dt, L.ln1_w, L.ln1_b) || !load_attn(p + ".attn", dt, dt, L.attn) ||
            !load_ln(p + ".cross_attn_ln", dt, L.lnx_w, L.lnx_b) || !load_attn(p + ".cross_attn", dt, d, L.cross) ||
            !load_ln(p + ".mlp_ln", dt, L.ln2_w, L.ln2_b) || !load_mlp(p, dt, L.mlp0, L.mlp0_b, L.mlp2, L.mlp2_b))
            return fail(missing);
    }
    if (!load_ln("decoder.ln", dt, m->ln_w, m->ln_b)) return fail(missing);

    // GELU lookup table over all f16 bit patterns (A.4): tanh approximation, computed in f32, stored f16
    m->gelu_table.resize(65536);
    for (uint32_t i = 0; i < 65536; ++i) {
        float x = f16_to_f32((uint16_t)i);
        float y = 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
        m->gelu_table[i] = f32_to_f16(y);
    }
    return m.release();
}

}  // namespace orc
