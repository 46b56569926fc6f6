// Model loader: legacy-ggml file -> device tensors (see model.h for the reference citations).
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>

#include "../../include/oh_whisper.h"
#include "model.h"

namespace oh {

namespace {

static const char* k_lang_codes[100] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
    "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
    "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
    "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue"};

struct HostTensor {
    int n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    int ttype = 0;  // after load_model's read loop: 0 f32 or 1 f16 (quantised blocks are expanded while reading)
    std::vector<uint8_t> data;
    int64_t nelem() const { return ne[0] * ne[1] * ne[2] * ne[3]; }
    float at(int64_t i) const {
        if (ttype == 0) return ((const float*)data.data())[i];
        return __half2float(((const __half*)data.data())[i]);
    }
};

// ---- ggml block-quantised tensor types -------------------------------------------------------
// The reference downloads only f16 files (whisper.rs:71-79), but users drop quantised ggml-*.bin files into the model
// directory; whisper.cpp's quantize tool writes them in this same container with per-tensor types below.  Blocks of 32
// consecutive elements along the contiguous (K) dimension.  They are expanded to f16 at load (x = q * d [+ m] evaluated
// in f32, rounded once): the tensor-core path multiplies f16 operands, like ggml's own CUDA backend does for these
// types; ggml's CPU kernels instead quantise the activations to 8 bits per block, which this library does not imitate.
enum { GGML_F32 = 0, GGML_F16 = 1, GGML_Q4_0 = 2, GGML_Q4_1 = 3, GGML_Q5_0 = 6, GGML_Q5_1 = 7, GGML_Q8_0 = 8 };
constexpr int kQK = 32;
size_t ggml_block_bytes(int t) {
    switch (t) {
        case GGML_Q4_0: return 2 + 16;
        case GGML_Q4_1: return 2 + 2 + 16;
        case GGML_Q5_0: return 2 + 4 + 16;
        case GGML_Q5_1: return 2 + 2 + 4 + 16;
        case GGML_Q8_0: return 2 + 32;
        default: return 0;
    }
}
inline float f16_bits_to_f32(const uint8_t* p) {
    __half h;
    std::memcpy(&h, p, 2);
    return __half2float(h);
}
// one block -> 32 f16 values
void dequant_block(int t, const uint8_t* b, __half* out) {
    float v[kQK];
    switch (t) {
        case GGML_Q4_0: {
            const float d = f16_bits_to_f32(b);
            const uint8_t* qs = b + 2;
            for (int j = 0; j < 16; ++j) {
                v[j] = (float)((qs[j] & 0x0F) - 8) * d;
                v[j + 16] = (float)((qs[j] >> 4) - 8) * d;
            }
            break;
        }
        case GGML_Q4_1: {
            const float d = f16_bits_to_f32(b), m = f16_bits_to_f32(b + 2);
            const uint8_t* qs = b + 4;
            for (int j = 0; j < 16; ++j) {
                v[j] = (float)(qs[j] & 0x0F) * d + m;
                v[j + 16] = (float)(qs[j] >> 4) * d + m;
            }
            break;
        }
        case GGML_Q5_0: {
            const float d = f16_bits_to_f32(b);
            uint32_t qh;
            std::memcpy(&qh, b + 2, 4);
            const uint8_t* qs = b + 6;
            for (int j = 0; j < 16; ++j) {
                const uint8_t xh0 = ((qh >> j) << 4) & 0x10, xh1 = (qh >> (j + 12)) & 0x10;
                v[j] = (float)(((qs[j] & 0x0F) | xh0) - 16) * d;
                v[j + 16] = (float)(((qs[j] >> 4) | xh1) - 16) * d;
            }
            break;
        }
        case GGML_Q5_1: {
            const float d = f16_bits_to_f32(b), m = f16_bits_to_f32(b + 2);
            uint32_t qh;
            std::memcpy(&qh, b + 4, 4);
            const uint8_t* qs = b + 8;
            for (int j = 0; j < 16; ++j) {
                const uint8_t xh0 = ((qh >> j) << 4) & 0x10, xh1 = (qh >> (j + 12)) & 0x10;
                v[j] = (float)((qs[j] & 0x0F) | xh0) * d + m;
                v[j + 16] = (float)((qs[j] >> 4) | xh1) * d + m;
            }
            break;
        }
        case GGML_Q8_0: {
            const float d = f16_bits_to_f32(b);
            const int8_t* qs = (const int8_t*)(b + 2);
            for (int j = 0; j < kQK; ++j) v[j] = (float)qs[j] * d;
            break;
        }
        default: return;
    }
    for (int j = 0; j < kQK; ++j) out[j] = __float2half_rn(v[j]);
}

struct Loader {
    Model* m;
    std::map<std::string, HostTensor> T;
    std::string err;

    const HostTensor* need(const std::string& name, int64_t nelem) {
        auto it = T.find(name);
        if (it == T.end()) {
            if (err.empty()) err = "missing tensor " + name;
            return nullptr;
        }
        if (it->second.nelem() != nelem) {
            if (err.empty()) err = "wrong element count for tensor " + name;
            return nullptr;
        }
        return &it->second;
    }
    void* upload(const void* src, size_t bytes) {
        void* d = nullptr;
        if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess) {
            if (err.empty()) err = "cudaMalloc failed while loading the model";
            return nullptr;
        }
        m->allocs.push_back(d);
        m->bytes_on_device += bytes;
        if (bytes && cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            if (err.empty()) err = "cudaMemcpy failed while loading the model";
            return nullptr;
        }
        return d;
    }
    float* upload_f32(const std::vector<float>& v) { return (float*)upload(v.data(), v.size() * 4); }
    std::vector<float> vec_f32(const HostTensor& t) {
        std::vector<float> v((size_t)t.nelem());
        for (int64_t i = 0; i < t.nelem(); ++i) v[i] = t.at(i);
        return v;
    }
    // rows of 16-bit weights in the compute type's storage (f16 kept bit-exact; bf16 is a re-rounding)
    void* upload_w16(const std::vector<const HostTensor*>& parts, int k) {
        int64_t total = 0;
        for (auto* p : parts) total += p->nelem();
        std::vector<uint16_t> buf((size_t)total);
        size_t o = 0;
        for (auto* p : parts) {
            for (int64_t i = 0; i < p->nelem(); ++i, ++o) {
                if (m->compute_type == OH_COMPUTE_BF16) {
                    __nv_bfloat16 b = __float2bfloat16_rn(p->at(i));
                    std::memcpy(&buf[o], &b, 2);
                } else if (p->ttype == 1) {
                    buf[o] = ((const uint16_t*)p->data.data())[i];
                } else {
                    __half h = __float2half_rn(p->at(i));
                    std::memcpy(&buf[o], &h, 2);
                }
            }
        }
        (void)k;
        return upload(buf.data(), buf.size() * 2);
    }
    bool linear(Linear& L, const std::vector<std::string>& wnames, const std::vector<std::string>& bnames /* "" = zeros */, int n_each,
                int k) {
        std::vector<const HostTensor*> parts;
        for (auto& w : wnames) {
            const HostTensor* t = need(w, (int64_t)n_each * k);
            if (!t) return false;
            parts.push_back(t);
        }
        L.n = n_each * (int)wnames.size();
        L.k = k;
        L.w = upload_w16(parts, k);
        bool any_bias = false;
        for (auto& b : bnames) any_bias |= !b.empty();
        if (any_bias) {
            std::vector<float> bias((size_t)L.n, 0.f);
            for (size_t j = 0; j < bnames.size(); ++j) {
                if (bnames[j].empty()) continue;
                const HostTensor* t = need(bnames[j], n_each);
                if (!t) return false;
                for (int i = 0; i < n_each; ++i) bias[j * n_each + i] = t->at(i);
            }
            L.b = upload_f32(bias);
        }
        return L.w != nullptr && err.empty();
    }
    bool lnorm(LayerNormW& ln, const std::string& p, int d) {
        const HostTensor *w = need(p + ".weight", d), *b = need(p + ".bias", d);
        if (!w || !b) return false;
        ln.w = upload_f32(vec_f32(*w));
        ln.b = upload_f32(vec_f32(*b));
        return err.empty();
    }
    // conv weight torch [out][C][3] -> [out][tap*C + c]
    bool conv(Linear& L, const std::string& p, int out, int C) {
        const HostTensor *w = need(p + ".weight", (int64_t)out * C * 3), *b = need(p + ".bias", out);
        if (!w || !b) return false;
        HostTensor r;
        r.ttype = 0;
        r.n_dims = 2;
        r.ne[0] = 3 * C;
        r.ne[1] = out;
        r.data.resize((size_t)out * 3 * C * 4);
        float* dst = (float*)r.data.data();
        for (int o = 0; o < out; ++o)
            for (int c = 0; c < C; ++c)
                for (int t = 0; t < 3; ++t) dst[(size_t)o * 3 * C + t * C + c] = w->at(((int64_t)o * C + c) * 3 + t);
        // values are exactly representable in f16 when the file stores f16, so the f32 detour is lossless
        L.n = out;
        L.k = 3 * C;
        L.w = upload_w16({&r}, 3 * C);
        L.b = upload_f32(vec_f32(*b));
        return err.empty();
    }
};

template <class T>
bool rd(FILE* f, T& v) {
    return fread(&v, sizeof(T), 1, f) == 1;
}

}  // namespace

int lang_id_from_code(const char* code) {
    if (!code) return -1;
    for (int i = 0; i < 100; ++i)
        if (std::strcmp(code, k_lang_codes[i]) == 0) return i;
    return -1;
}
const char* lang_code_from_id(int id) { return (id >= 0 && id < 100) ? k_lang_codes[id] : nullptr; }

Model::~Model() {
    for (void* p : allocs) cudaFree(p);
}

static Model* load_model_impl(const char* path, int device, int compute_type);

Model* load_model(const char* path, int device, int compute_type) {
    try {  // no exception may cross the C ABI (std::bad_alloc / length_error on a hostile file)
        return load_model_impl(path, device, compute_type);
    } catch (const std::exception& e) {
        set_error(std::string("load_model(") + path + "): " + e.what());
    } catch (...) {
        set_error(std::string("load_model(") + path + "): unknown exception");
    }
    return nullptr;
}

static Model* load_model_impl(const char* path, int device, int compute_type) {
    const auto t_load0 = std::chrono::steady_clock::now();
    FILE* f = fopen(path, "rb");
    if (!f) {
        set_error(std::string("cannot open model file ") + path);
        return nullptr;
    }
    std::unique_ptr<Model> m(new Model());
    m->device = device;
    m->compute_type = compute_type;
    Loader L;
    L.m = m.get();
    auto fail = [&](const std::string& e) -> Model* {
        set_error("load_model(" + std::string(path) + "): " + e);
        if (f) fclose(f);
        f = nullptr;
        return nullptr;
    };
    struct FileGuard {  // an exception thrown while parsing must not leak the handle
        FILE*& f;
        ~FileGuard() {
            if (f) fclose(f);
        }
    } file_guard{f};

    long long file_size = 0;
    if (fseek(f, 0, SEEK_END) == 0) {
        file_size = ftell(f);
        fseek(f, 0, SEEK_SET);
    }
    uint32_t magic = 0;
    if (!rd(f, magic) || magic != 0x67676d6cu) return fail("bad magic (not a legacy ggml file)");
    HParams& hp = m->hp;
    if (fread(&hp, sizeof(int32_t), 11, f) != 11) return fail("truncated hparams");
    // every field is checked before it is used as a divisor, a vector size or a loop bound
    if (hp.n_audio_state <= 0 || hp.n_audio_state > 8192 || hp.n_audio_head <= 0 || hp.n_text_state <= 0 || hp.n_text_head <= 0 ||
        hp.n_audio_layer <= 0 || hp.n_audio_layer > 128 || hp.n_text_layer <= 0 || hp.n_text_layer > 128)
        return fail("unsupported hparams (non-positive or oversized dimension)");
    if (hp.n_audio_state % kHeadDim || hp.n_audio_state / hp.n_audio_head != kHeadDim || hp.n_text_state % kHeadDim ||
        hp.n_text_state / hp.n_text_head != kHeadDim || hp.n_audio_ctx != 1500 || hp.n_text_ctx != 448 ||
        hp.n_vocab < 51864 || hp.n_vocab > 51866 || (hp.n_mels != 80 && hp.n_mels != 128) || hp.n_audio_state != hp.n_text_state)
        return fail("unsupported hparams");
    // ftype is informational (whisper.cpp: 0 f32, 1 f16, 2 q4_0, 3 q4_1, 7 q8_0, 8 q5_0, 9 q5_1, possibly + 1000 x version);
    // what decides is every tensor's own type below

    int32_t fm = 0, ff = 0;
    if (!rd(f, fm) || !rd(f, ff) || fm != hp.n_mels || ff != kNBins) return fail("bad mel filter dims");
    m->mel.dense.resize((size_t)fm * ff);
    if (fread(m->mel.dense.data(), 4, m->mel.dense.size(), f) != m->mel.dense.size()) return fail("truncated mel filters");

    Vocab& vc = m->vocab;
    vc.n_vocab = hp.n_vocab;
    int32_t n_in_file = 0;
    if (!rd(f, n_in_file) || n_in_file < 0 || n_in_file > hp.n_vocab) return fail("bad vocabulary size");
    vc.id_to_token.resize(hp.n_vocab);
    for (int i = 0; i < n_in_file; ++i) {
        uint32_t len = 0;
        if (!rd(f, len) || len > 4096) return fail("bad vocabulary entry");
        std::string s(len, '\0');
        if (len && fread(&s[0], 1, len, f) != len) return fail("truncated vocabulary");
        vc.id_to_token[i] = s;
        if (s == " " && vc.token_blank < 0) vc.token_blank = i;
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
    }

    while (true) {
        int32_t n_dims = 0, name_len = 0, ttype = 0;
        if (!rd(f, n_dims)) break;  // EOF
        if (!rd(f, name_len) || !rd(f, ttype)) return fail("truncated tensor header");
        const size_t qblock = ggml_block_bytes(ttype);
        if (n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 255 || (ttype != GGML_F32 && ttype != GGML_F16 && qblock == 0))
            return fail("bad tensor header (unsupported ggml tensor type " + std::to_string(ttype) + ")");
        HostTensor t;
        t.n_dims = n_dims;
        t.ttype = ttype;
        for (int i = 0; i < n_dims; ++i) {
            int32_t ne = 0;
            if (!rd(f, ne) || ne <= 0) return fail("bad tensor dims");
            t.ne[i] = ne;
        }
        std::string name(name_len, '\0');
        if (fread(&name[0], 1, name_len, f) != (size_t)name_len) return fail("truncated tensor name");
        {   // the payload must fit in what is left of the file (also rules out overflowing dims)
            const long long here = ftell(f);
            long long ne_total = 1;
            for (int i = 0; i < n_dims; ++i) {
                if (t.ne[i] > (1LL << 40) / ne_total) return fail("tensor dims overflow for " + name);
                ne_total *= t.ne[i];
            }
            if (qblock && (t.ne[0] % kQK)) return fail("quantised tensor " + name + ": row length is not a multiple of 32");
            const long long payload = qblock ? ne_total / kQK * (long long)qblock : ne_total * (ttype ? 2 : 4);
            if (here < 0 || payload > file_size - here) return fail("tensor " + name + " is larger than the rest of the file");
        }
        if (qblock) {  // expand to f16 while reading, a few MB at a time
            const int64_t n_blocks = t.nelem() / kQK;
            t.data.resize((size_t)t.nelem() * 2);
            __half* out = (__half*)t.data.data();
            std::vector<uint8_t> raw;
            const int64_t chunk = 1 << 16;
            for (int64_t b0 = 0; b0 < n_blocks; b0 += chunk) {
                const int64_t nb = std::min(chunk, n_blocks - b0);
                raw.resize((size_t)nb * qblock);
                if (fread(raw.data(), 1, raw.size(), f) != raw.size()) return fail("truncated tensor data for " + name);
                for (int64_t b = 0; b < nb; ++b) dequant_block(ttype, raw.data() + (size_t)b * qblock, out + (size_t)(b0 + b) * kQK);
            }
            t.ttype = GGML_F16;
            m->n_quantised_tensors++;
        } else {
            t.data.resize((size_t)t.nelem() * (ttype ? 2 : 4));
            if (fread(t.data.data(), 1, t.data.size(), f) != t.data.size()) return fail("truncated tensor data for " + name);
        }
        L.T[name] = std::move(t);
    }
    fclose(f);
    f = nullptr;

    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
    const int d = hp.n_audio_state, V = hp.n_vocab;

    // mel filterbank in sparse (start, count, weights) form
    {
        MelBank& mb = m->mel;
        mb.n_mel = hp.n_mels;
        std::vector<int> start(mb.n_mel), count(mb.n_mel), offset(mb.n_mel);
        std::vector<float> w;
        for (int j = 0; j < mb.n_mel; ++j) {
            int lo = kNBins, hi = -1;
            for (int k = 0; k < kNBins; ++k)
                if (mb.dense[(size_t)j * kNBins + k] != 0.0f) {
                    lo = std::min(lo, k);
                    hi = std::max(hi, k);
                }
            start[j] = hi < 0 ? 0 : lo;
            count[j] = hi < 0 ? 0 : hi - lo + 1;
            offset[j] = (int)w.size();
            for (int k = 0; k < count[j]; ++k) w.push_back(mb.dense[(size_t)j * kNBins + start[j] + k]);
        }
        mb.n_weights = (int)w.size();
        if (mb.n_weights > kMaxMelFilterWeights)  // the mel kernel stages the sparse bank in shared memory (mel.cu)
            return fail("mel filterbank is not sparse enough (" + std::to_string(mb.n_weights) + " non-zero span weights, limit " +
                        std::to_string(kMaxMelFilterWeights) + ")");
        mb.start = (int*)L.upload(start.data(), start.size() * 4);
        mb.count = (int*)L.upload(count.data(), count.size() * 4);
        mb.offset = (int*)L.upload(offset.data(), offset.size() * 4);
        mb.weights = (float*)L.upload(w.data(), w.size() * 4);
    }

    {
        const HostTensor* t = L.need("encoder.positional_embedding", (int64_t)hp.n_audio_ctx * d);
        if (!t) return fail(L.err);
        m->enc_pos = L.upload_f32(L.vec_f32(*t));
        t = L.need("decoder.positional_embedding", (int64_t)hp.n_text_ctx * d);
        if (!t) return fail(L.err);
        m->dec_pos = L.upload_f32(L.vec_f32(*t));
    }
    if (!L.conv(m->conv1, "encoder.conv1", d, hp.n_mels) || !L.conv(m->conv2, "encoder.conv2", d, d)) return fail(L.err);

    m->enc.resize(hp.n_audio_layer);
    for (int l = 0; l < hp.n_audio_layer; ++l) {
        const std::string p = "encoder.blocks." + std::to_string(l);
        EncLayer& E = m->enc[l];
        bool ok = L.lnorm(E.ln1, p + ".attn_ln", d) && L.lnorm(E.ln2, p + ".mlp_ln", d) &&
                  L.linear(E.qkv, {p + ".attn.query.weight", p + ".attn.key.weight", p + ".attn.value.weight"},
                           {p + ".attn.query.bias", "", p + ".attn.value.bias"}, d, d) &&
                  L.linear(E.out, {p + ".attn.out.weight"}, {p + ".attn.out.bias"}, d, d) &&
                  L.linear(E.mlp0, {p + ".mlp.0.weight"}, {p + ".mlp.0.bias"}, 4 * d, d) &&
                  L.linear(E.mlp2, {p + ".mlp.2.weight"}, {p + ".mlp.2.bias"}, d, 4 * d);
        if (!ok) return fail(L.err);
    }
    if (!L.lnorm(m->ln_post, "encoder.ln_post", d)) return fail(L.err);

    if (!L.linear(m->tok_emb, {"decoder.token_embedding.weight"}, {""}, V, d)) return fail(L.err);
    m->dec.resize(hp.n_text_layer);
    for (int l = 0; l < hp.n_text_layer; ++l) {
        const std::string p = "decoder.blocks." + std::to_string(l);
        DecLayer& D = m->dec[l];
        bool ok = L.lnorm(D.ln1, p + ".attn_ln", d) && L.lnorm(D.lnx, p + ".cross_attn_ln", d) && L.lnorm(D.ln2, p + ".mlp_ln", d) &&
                  L.linear(D.qkv, {p + ".attn.query.weight", p + ".attn.key.weight", p + ".attn.value.weight"},
                           {p + ".attn.query.bias", "", p + ".attn.value.bias"}, d, d) &&
                  L.linear(D.out, {p + ".attn.out.weight"}, {p + ".attn.out.bias"}, d, d) &&
                  L.linear(D.xq, {p + ".cross_attn.query.weight"}, {p + ".cross_attn.query.bias"}, d, d) &&
                  L.linear(D.xkv, {p + ".cross_attn.key.weight", p + ".cross_attn.value.weight"}, {"", p + ".cross_attn.value.bias"}, d, d) &&
                  L.linear(D.xout, {p + ".cross_attn.out.weight"}, {p + ".cross_attn.out.bias"}, d, d) &&
                  L.linear(D.mlp0, {p + ".mlp.0.weight"}, {p + ".mlp.0.bias"}, 4 * d, d) &&
                  L.linear(D.mlp2, {p + ".mlp.2.weight"}, {p + ".mlp.2.bias"}, d, 4 * d);
        if (!ok) return fail(L.err);
    }
    if (!L.lnorm(m->ln, "decoder.ln", d)) return fail(L.err);

    // GELU table over every f16 bit pattern, computed with the host's tanhf as ggml does at start-up
    {
        std::vector<uint16_t> lut(65536);
        for (uint32_t i = 0; i < 65536; ++i) {
            uint16_t bits = (uint16_t)i;
            __half hx;
            std::memcpy(&hx, &bits, 2);
            float x = __half2float(hx);
            float y = 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));
            __half hy = __float2half_rn(y);
            std::memcpy(&lut[i], &hy, 2);
        }
        m->gelu_lut = (uint16_t*)L.upload(lut.data(), lut.size() * 2);
    }
    if (!L.err.empty()) return fail(L.err);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail("device error while loading");
    m->load_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_load0).count();
    return m.release();
}

}  // namespace oh
