// Device-resident model: parsed from the reference's model file format and laid out for the kernels.
// File format: legacy "ggml" container named ggml-<model>.bin by the reference
// (/root/reference/src/engine/whisper.rs:71-79; loaded at whisper.rs:159); layout per SURVEY.md Appendix A.1.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace oh {

struct HParams {
    int32_t n_vocab = 0, n_audio_ctx = 0, n_audio_state = 0, n_audio_head = 0, n_audio_layer = 0;
    int32_t n_text_ctx = 0, n_text_state = 0, n_text_head = 0, n_text_layer = 0, n_mels = 0, ftype = 0;
};

struct Vocab {
    int n_vocab = 0;
    std::vector<std::string> id_to_token;
    int token_eot = 50256, token_sot = 50257;
    int token_translate = 50357, token_transcribe = 50358, token_solm = 50359, token_prev = 50360;
    int token_nosp = 50361, token_not = 50362, token_beg = 50363;
    int token_blank = -1;  // id of " " (suppress_blank), -1 if the vocabulary has none
    bool is_multilingual() const { return n_vocab >= 51865; }
    int num_languages() const { return n_vocab - 51765 - (is_multilingual() ? 1 : 0); }
};

// y[N] = W[N][K] x + b.  W is 16-bit (f16 as stored in the file, or bf16 in OH_COMPUTE_BF16), K contiguous.
struct Linear {
    void* w = nullptr;
    float* b = nullptr;  // may be null
    int n = 0, k = 0;
};
struct LayerNormW {
    float* w = nullptr;
    float* b = nullptr;
};
struct EncLayer {
    LayerNormW ln1, ln2;
    Linear qkv, out, mlp0, mlp2;  // qkv: rows [Wq; Wk; Wv], bias [bq; 0; bv]
};
struct DecLayer {
    LayerNormW ln1, lnx, ln2;
    Linear qkv, out, xq, xkv, xout, mlp0, mlp2;  // xkv: rows [Wk_cross; Wv_cross], bias [0; bv]
};

struct MelBank {  // sparse form of the file's [n_mel][201] filterbank
    int n_mel = 0;
    int* start = nullptr;    // [n_mel] first non-zero bin
    int* count = nullptr;    // [n_mel] number of bins
    int* offset = nullptr;   // [n_mel] offset into weights
    float* weights = nullptr;
    int n_weights = 0;
    std::vector<float> dense;  // host copy [n_mel][201]
};

struct Model {
    HParams hp;
    Vocab vocab;
    int compute_type = 1;
    int device = 0;
    MelBank mel;
    float* enc_pos = nullptr;  // [1500][d] f32
    Linear conv1, conv2;       // [d][3*n_mel], [d][3*d]; column = tap * C + c
    std::vector<EncLayer> enc;
    LayerNormW ln_post;
    float* dec_pos = nullptr;  // [448][d] f32
    Linear tok_emb;            // [V][d], tied input embedding and logit projection
    std::vector<DecLayer> dec;
    LayerNormW ln;
    uint16_t* gelu_lut = nullptr;  // [65536] f16 -> f16
    size_t bytes_on_device = 0;
    int n_quantised_tensors = 0;   // tensors stored block-quantised in the file (expanded to f16 at load)
    float load_ms = 0.0f;          // file parse + upload time of load_model
    std::vector<void*> allocs;
    ~Model();
};

// returns nullptr and sets the thread's error string on failure
Model* load_model(const char* path, int device, int compute_type);

int lang_id_from_code(const char* code);
const char* lang_code_from_id(int id);

}  // namespace oh
