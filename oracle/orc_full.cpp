// ORACLE (test infrastructure) — logit filters, sampling, the full() window loop and segmenting.
// Restates SURVEY.md Appendix A.5-A.8 (whisper.cpp `whisper_full_with_state`, `whisper_process_logits`,
// `whisper_sample_token`, `whisper_sequence_score`; out of tree, reached from
// /root/reference/src/engine/whisper.rs:266-280).
#include <algorithm>

#include "orc_common.h"

namespace orc {

// A.6 — order matters
void process_logits(const Model& m, const FullParams& p, Decoder& dec, float temperature, bool /*unused*/) {
    const Vocab& vc = m.vocab;
    const int n_logits = vc.n_vocab;
    const auto& tokens_cur = dec.tokens;
    const bool is_initial = tokens_cur.empty();
    std::vector<float>& logits = dec.logits;
    std::vector<float>& logprobs = dec.logprobs;
    std::vector<float>& probs = dec.probs;
    logprobs.resize(n_logits);
    probs.resize(n_logits);

    if (temperature > 0.0f)
        for (int i = 0; i < n_logits; ++i) logits[i] /= temperature;

    logits[vc.token_not] = -INFINITY;
    if (p.no_timestamps)
        for (int i = vc.token_beg; i < n_logits; ++i) logits[i] = -INFINITY;

    if (p.suppress_blank && is_initial) {
        logits[vc.token_eot] = -INFINITY;
        auto it = vc.token_to_id.find(" ");
        if (it != vc.token_to_id.end()) logits[it->second] = -INFINITY;
    }
    logits[vc.token_sot] = -INFINITY;
    logits[vc.token_nosp] = -INFINITY;
    logits[vc.token_solm] = -INFINITY;  // tinydiarize off
    logits[vc.token_translate] = -INFINITY;
    logits[vc.token_transcribe] = -INFINITY;
    logits[vc.token_prev] = -INFINITY;
    for (int i = 0; i < vc.num_languages(); ++i) logits[vc.token_sot + 1 + i] = -INFINITY;

    // timestamps come in pairs, except directly before EOT
    {
        const bool last_was_ts = !tokens_cur.empty() && tokens_cur.back().id >= vc.token_beg;
        const bool penult_was_ts = tokens_cur.size() < 2 || tokens_cur[tokens_cur.size() - 2].id >= vc.token_beg;
        if (last_was_ts) {
            if (penult_was_ts)
                for (int i = vc.token_beg; i < n_logits; ++i) logits[i] = -INFINITY;
            else
                for (int i = 0; i < vc.token_eot; ++i) logits[i] = -INFINITY;
        }
    }
    // the first timestamp may not exceed max_initial_ts
    if (is_initial && p.max_initial_ts > 0.0f) {
        const float precision = 30.0f / (float)m.hp.n_audio_ctx;
        const int tid0 = (int)std::round(p.max_initial_ts / precision);
        for (int i = vc.token_beg + tid0 + 1; i < n_logits; ++i) logits[i] = -INFINITY;
    }
    // timestamps are monotone
    if (dec.has_ts) {
        const int tid0 = dec.seek_delta / 2;
        for (int i = vc.token_beg; i < vc.token_beg + tid0 && i < n_logits; ++i) logits[i] = -INFINITY;
    }
    // log-softmax
    {
        const float logit_max = *std::max_element(logits.begin(), logits.end());
        float lse = 0.0f;
        for (int i = 0; i < n_logits; ++i)
            if (logits[i] > -INFINITY) lse += expf(logits[i] - logit_max);
        lse = logf(lse) + logit_max;
        for (int i = 0; i < n_logits; ++i) logprobs[i] = logits[i] > -INFINITY ? logits[i] - lse : -INFINITY;
    }
    // if the probability mass on timestamps exceeds every text token, force a timestamp
    {
        float ts_logprob = -INFINITY;
        {
            float lse = 0.0f;
            const float lp_max = *std::max_element(logprobs.begin() + vc.token_beg, logprobs.end());
            for (int i = vc.token_beg; i < n_logits; ++i)
                if (logprobs[i] > -INFINITY) lse += expf(logprobs[i] - lp_max);
            if (lse > 0.0f) ts_logprob = logf(lse) + lp_max;
        }
        const float max_text = *std::max_element(logprobs.begin(), logprobs.begin() + vc.token_beg);
        if (ts_logprob > max_text)
            for (int i = 0; i < vc.token_beg; ++i) {
                logits[i] = -INFINITY;
                logprobs[i] = -INFINITY;
            }
    }
    for (int i = 0; i < n_logits; ++i) probs[i] = logits[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
}

// A.7
TokenData sample_token(const Model& m, Decoder& dec, bool best) {
    const Vocab& vc = m.vocab;
    const int n_logits = vc.n_vocab;
    const auto& probs = dec.probs;
    const auto& logprobs = dec.logprobs;
    TokenData r;
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = vc.token_beg; i < n_logits; ++i) {
            if (probs[i] == -INFINITY) continue;
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                r.tid = i;
            }
        }
        r.pt = (float)(max_ts / (sum_ts + 1e-10));
        r.ptsum = (float)sum_ts;
    }
    if (best) {
        for (int i = 0; i < n_logits; ++i)
            if (r.p < probs[i]) {  // strict: lowest index wins ties
                r.id = i;
                r.p = probs[i];
                r.plog = logprobs[i];
            }
    } else {
        std::discrete_distribution<> dist(probs.begin(), probs.end());
        r.id = dist(dec.rng);
        r.p = probs[r.id];
        r.plog = logprobs[r.id];
    }
    if (r.id >= vc.token_beg) {
        r.tid = r.id;
        r.pt = r.p;
    }
    return r;
}

// A.7, beam search: the k most probable ids of the filtered distribution, best first (ties: lower id first — a
// deterministic order, the same one the product uses).  tid / pt / ptsum as in sample_token.
// UNVERIFIED OFFLINE (SURVEY.md A.7 "I believe"): upstream versions differ in whether this takes the top-k or draws k
// samples; the reference never selects beam search (/root/reference/src/engine/whisper.rs:243), so nothing pins it.
std::vector<TokenData> sample_token_topk(const Model& m, Decoder& dec, int k) {
    const Vocab& vc = m.vocab;
    const int n_logits = vc.n_vocab;
    const auto& probs = dec.probs;
    const auto& logprobs = dec.logprobs;
    TokenData base;
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = vc.token_beg; i < n_logits; ++i) {
            if (probs[i] == -INFINITY) continue;
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                base.tid = i;
            }
        }
        base.pt = (float)(max_ts / (sum_ts + 1e-10));
        base.ptsum = (float)sum_ts;
    }
    std::vector<int> ids(n_logits);
    for (int i = 0; i < n_logits; ++i) ids[i] = i;
    k = std::min(k, n_logits);
    std::partial_sort(ids.begin(), ids.begin() + k, ids.end(), [&](int a, int b) {
        if (logprobs[a] != logprobs[b]) return logprobs[a] > logprobs[b];
        return a < b;
    });
    std::vector<TokenData> out;
    for (int c = 0; c < k; ++c) {
        TokenData r = base;
        r.id = ids[c];
        r.p = probs[r.id];
        r.plog = logprobs[r.id];
        if (r.id >= vc.token_beg) {
            r.tid = r.id;
            r.pt = r.p;
        }
        out.push_back(r);
    }
    return out;
}

// A.8
static void sequence_score(const FullParams& p, Decoder& d) {
    const int n = d.result_len;
    double result = 0.0;
    for (int i = 0; i < n; ++i) result += d.tokens[i].plog;
    d.sum_logprobs = result;
    d.avg_logprobs = n > 0 ? result / n : -INFINITY;
    double penalty = n;
    if (p.length_penalty > 0.0f) penalty = std::pow((5.0 + n) / 6.0, (double)p.length_penalty);
    d.score = n > 0 ? result / penalty : -INFINITY;
    {
        const int w = 32;
        int cnt = 0;
        std::map<int, int> counts;
        for (int i = std::max(0, n - w); i < n; ++i) {
            counts[d.tokens[i].id]++;
            cnt++;
        }
        double entropy = 0.0;
        for (auto& kv : counts) {
            double pr = (double)kv.second / cnt;
            entropy -= pr * std::log(pr);
        }
        d.entropy = entropy;
    }
}

// A.5
int full(State& st, const FullParams& p, const float* pcm, int n) {
    const Model& m = *st.model;
    const Vocab& vc = m.vocab;
    st.result_all.clear();
    st.n_encode = st.n_decode_tokens = st.n_fallbacks = 0;
    st.window_seeks.clear();
    st.window_temps.clear();
    st.window_tokens.clear();
    if (n <= 200) return -1;  // cannot reflect-pad: "failed to compute log mel spectrogram"
    if (p.strategy != 0 && p.strategy != 1) return -101;
    const bool beam = p.strategy == 1;

    log_mel_spectrogram(m, pcm, n, st.mel, st.opt.n_threads);

    const int seek_start = 0;
    const int seek_end = st.mel.n_len_org;
    if (seek_end < seek_start + 100) return 0;  // < 1 s of audio: no segments (the reason for the reference's
                                                 // 1.1 s padding, /root/reference/src/input/audio.rs:31-34)

    std::vector<float> temperatures;
    if (p.temperature_inc > 0.0f) {
        for (float t = p.temperature; t < 1.0f + 1e-6f; t += p.temperature_inc) temperatures.push_back(t);
    } else {
        temperatures.push_back(p.temperature);
    }

    if (p.no_context) st.prompt_past.clear();

    std::vector<int> prompt_init = {vc.token_sot};
    bool no_timestamps = p.no_timestamps;
    if (vc.is_multilingual()) {
        int lang_id = lang_id_from_code(p.language.c_str());
        if (lang_id < 0 || lang_id >= vc.num_languages()) lang_id = 0;
        st.lang_id = lang_id;
        prompt_init.push_back(vc.token_sot + 1 + lang_id);
        prompt_init.push_back(p.translate ? vc.token_translate : vc.token_transcribe);
    } else {
        st.lang_id = 0;
    }
    if (m.hp.n_text_layer == 2 && m.hp.n_vocab != 51866) no_timestamps = true;  // "distilled" heuristic (A.1)
    if (no_timestamps) prompt_init.push_back(vc.token_not);
    FullParams pp = p;
    pp.no_timestamps = no_timestamps;

    const int n_decoders = std::max(1, beam ? std::max(p.beam_size, p.best_of) : p.best_of);
    if ((int)st.decoders.size() < n_decoders) st.decoders.resize(n_decoders);

    int seek = seek_start;
    const int T2 = 2 * m.hp.n_audio_ctx;  // 3000 mel frames per window
    std::vector<float> window((size_t)m.hp.n_mels * T2);
    std::vector<int> prompt;

    while (true) {
        if (seek + 100 >= seek_end) break;

        // encoder input: columns [seek, seek+3000) of the mel, zero past n_len
        for (int j = 0; j < m.hp.n_mels; ++j)
            for (int i = 0; i < T2; ++i) {
                int c = seek + i;
                window[(size_t)j * T2 + i] = c < st.mel.n_len ? st.mel.data[(size_t)j * st.mel.n_len + c] : 0.0f;
            }
        encode(m, st.opt, window.data(), st.enc_out, nullptr);
        cross_kv(m, st.opt, st.enc_out, st.ck, st.cv);
        st.n_encode++;
        st.window_seeks.push_back(seek);

        // a very short tail left: drop the prompt (upstream behaviour recalled beyond Appendix A)
        if (seek > seek_start && seek + 500 >= seek_end) st.prompt_past.clear();

        int best_decoder_id = 0;
        float t_accepted = 0.f;
        for (int it = 0; it < (int)temperatures.size(); ++it) {
            const float t_cur = temperatures[it];
            t_accepted = t_cur;
            int n_decoders_cur = beam ? p.beam_size : 1;  // A.5: GREEDY t==0 ? 1 : best_of | BEAM t==0 ? beam_size : best_of
            if (t_cur > 0.0f) n_decoders_cur = p.best_of;
            n_decoders_cur = std::max(1, n_decoders_cur);
            const bool beam_step = beam && t_cur <= 0.0f;

            for (int j = 0; j < n_decoders_cur; ++j) {
                Decoder& d = st.decoders[j];
                d.tokens.clear();
                d.result_len = 0;
                d.seek_delta = 100 * 30;
                d.has_ts = d.failed = d.completed = false;
                d.sum_logprobs_all = 0.0;
                d.sum_logprobs = -INFINITY;
                d.avg_logprobs = -INFINITY;
                d.entropy = 0.0;
                d.score = -INFINITY;
            }

            prompt.clear();
            if (!st.prompt_past.empty() && t_cur < 0.5f && p.n_max_text_ctx > 0) {
                int n_take = std::min(std::min(p.n_max_text_ctx, m.hp.n_text_ctx / 2), (int)st.prompt_past.size());
                prompt.push_back(vc.token_prev);
                prompt.insert(prompt.end(), st.prompt_past.end() - n_take, st.prompt_past.end());
            }
            prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());

            {
                Decoder& d0 = st.decoders[0];
                d0.logits.resize(vc.n_vocab);
                for (int i = 0; i < (int)prompt.size(); ++i)
                    decode_token(m, st.opt, st.ck, st.cv, d0.kv, prompt[i], i, i + 1 == (int)prompt.size() ? d0.logits.data() : nullptr, nullptr);
                st.n_decode_tokens += (int)prompt.size();
                process_logits(m, pp, d0, t_cur, true);
                for (int j = 1; j < n_decoders_cur; ++j) {
                    Decoder& d = st.decoders[j];
                    d.kv = d0.kv;
                    d.logits = d0.logits;
                    d.logprobs = d0.logprobs;
                    d.probs = d0.probs;
                }
            }

            const int n_max = m.hp.n_text_ctx / 2 - 4;
            for (int i = 0; i < n_max; ++i) {
                if (!beam_step) {
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        Decoder& d = st.decoders[j];
                        if (d.completed || d.failed) continue;
                        TokenData tok = sample_token(m, d, t_cur < 1e-6f);
                        d.tokens.push_back(tok);
                        d.sum_logprobs_all += tok.plog;
                    }
                } else {
                    // A.7 beam: every live decoder proposes its top beam_size continuations; all candidates are ranked by
                    // sum_logprobs_all (ties: proposing decoder, then rank inside its proposal) and handed to the live
                    // decoders in order, skipping a candidate whose token sequence equals the previous pick (i > 0);
                    // a decoder that takes a candidate of another decoder also takes that decoder's KV cache
                    struct Cand {
                        int decoder_idx, seek_delta;
                        bool has_ts;
                        int result_len;
                        std::vector<TokenData> tokens;
                        double sum_logprobs_all;
                    };
                    std::vector<Cand> cands;
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        Decoder& d = st.decoders[j];
                        if (d.completed || d.failed) continue;
                        for (const TokenData& tok : sample_token_topk(m, d, p.beam_size)) {
                            Cand c{j, d.seek_delta, d.has_ts, d.result_len, d.tokens, d.sum_logprobs_all + tok.plog};
                            c.tokens.push_back(tok);
                            cands.push_back(std::move(c));
                        }
                    }
                    std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
                        if (a.sum_logprobs_all != b.sum_logprobs_all) return a.sum_logprobs_all > b.sum_logprobs_all;
                        return a.decoder_idx < b.decoder_idx;
                    });
                    auto same_tokens = [](const Cand& a, const Cand& b) {
                        if (a.tokens.size() != b.tokens.size()) return false;
                        for (size_t q = 0; q < a.tokens.size(); ++q)
                            if (a.tokens[q].id != b.tokens[q].id) return false;
                        return true;
                    };
                    std::vector<KVCache> kv_src(n_decoders_cur);
                    std::vector<char> kv_saved(n_decoders_cur, 0);
                    std::vector<int> take(n_decoders_cur, -1);
                    size_t cur_c = 0;
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        Decoder& d = st.decoders[j];
                        if (d.completed || d.failed) continue;
                        if (cur_c >= cands.size()) cur_c = 0;
                        const size_t pick = cur_c++;
                        while (cur_c < cands.size() && i > 0 && same_tokens(cands[cur_c], cands[pick])) ++cur_c;
                        take[j] = (int)pick;
                    }
                    for (int j = 0; j < n_decoders_cur; ++j)  // caches that another decoder is about to inherit: keep a copy
                        if (take[j] >= 0 && cands[take[j]].decoder_idx != j && !kv_saved[cands[take[j]].decoder_idx]) {
                            kv_src[cands[take[j]].decoder_idx] = st.decoders[cands[take[j]].decoder_idx].kv;
                            kv_saved[cands[take[j]].decoder_idx] = 1;
                        }
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        if (take[j] < 0) continue;
                        Decoder& d = st.decoders[j];
                        const Cand& c = cands[take[j]];
                        if (c.decoder_idx != j) d.kv = kv_src[c.decoder_idx];
                        d.tokens = c.tokens;
                        d.seek_delta = c.seek_delta;
                        d.has_ts = c.has_ts;
                        d.result_len = c.result_len;
                        d.sum_logprobs_all = c.sum_logprobs_all;
                    }
                }
                for (int j = 0; j < n_decoders_cur; ++j) {
                    Decoder& d = st.decoders[j];
                    if (d.completed || d.failed) continue;
                    const TokenData& token = d.tokens.back();
                    if (token.id > vc.token_beg) {
                        const int seek_delta_new = 2 * (token.id - vc.token_beg);
                        if (d.has_ts && d.seek_delta > seek_delta_new && d.result_len < i) {
                            d.failed = true;  // went back in time
                            continue;
                        }
                        d.seek_delta = seek_delta_new;
                        d.result_len = i + 1;
                        d.has_ts = true;
                    }
                    if (token.id == vc.token_eot || (p.max_tokens > 0 && i >= p.max_tokens) ||
                        (d.has_ts && seek + d.seek_delta + 100 >= seek_end)) {
                        if (d.result_len == 0 && !pp.no_timestamps) {
                            if (seek + d.seek_delta + 100 >= seek_end) d.result_len = i + 1;
                            else {
                                d.failed = true;
                                continue;
                            }
                        }
                        if (p.single_segment || pp.no_timestamps) {
                            d.result_len = i + 1;
                            d.seek_delta = 100 * 30;
                        }
                        d.completed = true;
                        continue;
                    }
                    if (i == n_max - 1 && (d.result_len == 0 || d.seek_delta < 100 * 30 / 2)) {
                        d.failed = true;  // repetition-loop guard
                        continue;
                    }
                }
                bool all_done = true;
                for (int j = 0; j < n_decoders_cur; ++j)
                    if (!st.decoders[j].completed && !st.decoders[j].failed) all_done = false;
                if (all_done) break;

                for (int j = 0; j < n_decoders_cur; ++j) {
                    Decoder& d = st.decoders[j];
                    if (d.completed || d.failed) continue;
                    const int n_past = (int)prompt.size() + i;
                    decode_token(m, st.opt, st.ck, st.cv, d.kv, d.tokens.back().id, n_past, d.logits.data(), nullptr);
                    st.n_decode_tokens++;
                    process_logits(m, pp, d, t_cur, false);
                }
            }

            // rank
            {
                double best_score = -INFINITY;
                for (int j = 0; j < n_decoders_cur; ++j) {
                    Decoder& d = st.decoders[j];
                    if (d.failed) continue;
                    d.tokens.resize(d.result_len);
                    sequence_score(p, d);
                    if (d.result_len > 32 && d.entropy < p.entropy_thold) {
                        d.failed = true;
                        continue;
                    }
                    if (best_score < d.score) {
                        best_score = d.score;
                        best_decoder_id = j;
                    }
                }
            }
            bool success = true;
            if (it != (int)temperatures.size() - 1) {
                const Decoder& d = st.decoders[best_decoder_id];
                if (d.failed || d.avg_logprobs < p.logprob_thold) success = false;
            }
            if (success) break;
            st.n_fallbacks++;
        }
        st.window_temps.push_back(t_accepted);
        {
            std::vector<int> ids;
            for (const TokenData& t : st.decoders[best_decoder_id].tokens) ids.push_back(t.id);
            st.window_tokens.push_back(std::move(ids));
        }

        // output
        {
            const Decoder& best = st.decoders[best_decoder_id];
            const int seek_delta = best.seek_delta;
            const int result_len = best.result_len;
            const auto& tokens_cur = best.tokens;

            std::vector<int> new_past;
            if (!prompt.empty() && prompt.front() == vc.token_prev)
                new_past.insert(new_past.end(), prompt.begin() + 1, prompt.end() - prompt_init.size());
            for (int i = 0; i < result_len && i < (int)tokens_cur.size(); ++i) new_past.push_back(tokens_cur[i].id);
            st.prompt_past.swap(new_past);

            if (!tokens_cur.empty()) {
                int i0 = 0;
                int64_t t0 = seek + 2 * (tokens_cur.front().tid - vc.token_beg);
                std::string text;
                for (int i = 0; i < (int)tokens_cur.size(); ++i) {
                    if (tokens_cur[i].id < vc.token_eot) text += vc.id_to_token[tokens_cur[i].id];
                    if (tokens_cur[i].id > vc.token_beg && !p.single_segment) {
                        const int64_t t1 = seek + 2 * (tokens_cur[i].tid - vc.token_beg);
                        if (!text.empty()) {
                            Segment s;
                            s.t0 = t0;
                            s.t1 = t1;
                            s.text = text;
                            for (int j = i0; j <= i; ++j) s.tokens.push_back(tokens_cur[j]);
                            st.result_all.push_back(std::move(s));
                        }
                        text.clear();
                        while (i < (int)tokens_cur.size() && tokens_cur[i].id > vc.token_beg) i++;
                        i--;
                        t0 = t1;
                        i0 = i + 1;
                    }
                }
                if (!text.empty()) {
                    Segment s;
                    s.t0 = t0;
                    s.t1 = seek + seek_delta;
                    s.text = text;
                    for (int j = i0; j < (int)tokens_cur.size(); ++j) s.tokens.push_back(tokens_cur[j]);
                    st.result_all.push_back(std::move(s));
                }
            }
            seek += seek_delta;
        }
    }
    return 0;
}

}  // namespace orc
