// The full() driver: the device-side replacement of whisper.cpp's `whisper_full_with_state`,
// reached from /root/reference/src/engine/whisper.rs:266, re-designed as a continuous-batching scheduler.
//
// Every item of a call is one independent reference call (one TranscriptionJob of /root/reference/src/queue/mod.rs:18-27);
// windows of one item are sequentially dependent (A.5: seek from the last timestamp, prompt from the previous window),
// items are independent.  The scheduler keeps three pipelines busy at once:
//   front-end stream  PCM upload + log-mel of a new submission
//   encoder stream    encoder passes over waiting windows, cross K/V straight into free K/V slots (look-ahead: there are
//                     more slots than decoder rows, so a window is usually encoded before a row is free for it)
//   decode stream     greedy decode steps (one CUDA graph per token) over the resident rows under the device state machine
// and every `check_steps` steps the host looks at the rows' `active` flags: a finished row is scored (A.8), its segments
// are emitted, its slot is freed, its item's next window queues for the encoder, and the row is re-armed with an encoded
// window at once.  Submissions stay live until they are waited for (oh_stream_submit / oh_stream_wait), so the rows stay
// full across calls; oh_full / oh_full_batch are submit + wait.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "decode_rules.h"
#include "engine.h"

namespace oh {

using clk = std::chrono::steady_clock;
static inline float ms_since(clk::time_point t0) { return std::chrono::duration<float, std::milli>(clk::now() - t0).count(); }

namespace {

struct PassRow {  // one decoder row of a lock-step pass (temperature ladder)
    Submission* sub;
    int item, slot;
    std::vector<int> prompt;
};
struct PassResult {
    std::vector<TokenData> tokens;
    int result_len = 0, seek_delta = 0;
    bool has_ts = false, failed = false, completed = false;
    double sum_logprobs = 0, avg_logprobs = 0, entropy = 0, score = 0;
};

void sequence_score(const oh_full_params& p, PassResult& r) {  // A.8
    const int n = r.result_len;
    double sum = 0.0;
    for (int i = 0; i < n && i < (int)r.tokens.size(); ++i) sum += r.tokens[i].plog;
    r.sum_logprobs = sum;
    r.avg_logprobs = n > 0 ? sum / n : -INFINITY;
    double penalty = n;
    if (p.length_penalty > 0.0f) penalty = std::pow((5.0 + n) / 6.0, (double)p.length_penalty);
    r.score = n > 0 ? sum / penalty : -INFINITY;
    std::map<int, int> counts;
    int cnt = 0;
    for (int i = std::max(0, n - 32); i < n; ++i) {
        counts[r.tokens[i].id]++;
        cnt++;
    }
    double ent = 0.0;
    for (auto& kv : counts) {
        const double q = (double)kv.second / cnt;
        ent -= q * std::log(q);
    }
    r.entropy = ent;
}

void emit_segments(const Model& m, const oh_full_params& p, Item& it, const PassResult& best, const std::vector<int>& prompt, int n_prompt_init) {
    const Vocab& vc = m.vocab;
    const auto& tk = best.tokens;
    std::vector<int> new_past;
    if (!prompt.empty() && prompt.front() == vc.token_prev) new_past.insert(new_past.end(), prompt.begin() + 1, prompt.end() - n_prompt_init);
    for (int i = 0; i < best.result_len && i < (int)tk.size(); ++i) new_past.push_back(tk[i].id);
    it.prompt_past.swap(new_past);
    const int seek = it.seek;
    if (!tk.empty()) {
        int i0 = 0;
        int64_t t0 = seek + 2 * (tk.front().tid - vc.token_beg);
        std::string text;
        for (int i = 0; i < (int)tk.size(); ++i) {
            if (p.print_special || tk[i].id < vc.token_eot) text += vc.id_to_token[tk[i].id];
            if (tk[i].id > vc.token_beg && !p.single_segment) {
                const int64_t t1 = seek + 2 * (tk[i].tid - vc.token_beg);
                if (!text.empty()) {
                    Segment s;
                    s.t0 = t0;
                    s.t1 = t1;
                    s.text = text;
                    for (int j = i0; j <= i; ++j) s.tokens.push_back(tk[j]);
                    it.segments.push_back(std::move(s));
                }
                text.clear();
                while (i < (int)tk.size() && tk[i].id > vc.token_beg) i++;
                i--;
                t0 = t1;
                i0 = i + 1;
            }
        }
        if (!text.empty()) {
            Segment s;
            s.t0 = t0;
            s.t1 = seek + best.seek_delta;
            s.text = text;
            for (int j = i0; j < (int)tk.size(); ++j) s.tokens.push_back(tk[j]);
            it.segments.push_back(std::move(s));
        }
    }
    it.seek += best.seek_delta;
}

// tid / pt / ptsum of a distribution, as recorded with every sampled token (A.7)
TokenData timestamp_stats(const Vocab& vc, const float* probs) {
    TokenData r;
    double sum_ts = 0.0, max_ts = 0.0;
    for (int i = vc.token_beg; i < vc.n_vocab; ++i) {
        sum_ts += probs[i];
        if (max_ts < probs[i]) {
            max_ts = probs[i];
            r.tid = i;
        }
    }
    r.pt = (float)(max_ts / (sum_ts + 1e-10));
    r.ptsum = (float)sum_ts;
    return r;
}

// host side of one sampled token (A.7, t > 0): std::discrete_distribution over the row's probabilities
TokenData sample_on_host(const Vocab& vc, const float* probs, const float* logprobs, std::mt19937& rng) {
    TokenData r = timestamp_stats(vc, probs);
    std::discrete_distribution<> dist(probs, probs + vc.n_vocab);
    r.id = dist(rng);
    r.p = probs[r.id];
    r.plog = logprobs[r.id];
    if (r.id >= vc.token_beg) {
        r.tid = r.id;
        r.pt = r.p;
    }
    return r;
}

void build_prompt(const Submission& sub, const HParams& hp, const Vocab& vc, const Item& it, float t_cur, std::vector<int>& prompt) {
    prompt.clear();
    if (!it.prompt_past.empty() && t_cur < 0.5f && sub.p.n_max_text_ctx > 0) {
        const int n_take = std::min(std::min(sub.p.n_max_text_ctx, hp.n_text_ctx / 2), (int)it.prompt_past.size());
        prompt.push_back(vc.token_prev);
        prompt.insert(prompt.end(), it.prompt_past.end() - n_take, it.prompt_past.end());
    }
    prompt.insert(prompt.end(), sub.prompt_init.begin(), sub.prompt_init.end());
}

}  // namespace

DecodeRules make_rules(const Model& m, const oh_full_params& p, bool no_timestamps, float temperature) {
    const Vocab& vc = m.vocab;
    DecodeRules R;
    R.n_vocab = vc.n_vocab; R.eot = vc.token_eot; R.sot = vc.token_sot; R.translate = vc.token_translate; R.transcribe = vc.token_transcribe;
    R.solm = vc.token_solm; R.prev = vc.token_prev; R.nosp = vc.token_nosp; R.tok_not = vc.token_not; R.beg = vc.token_beg;
    R.blank = vc.token_blank; R.n_lang = vc.num_languages();
    R.no_timestamps = no_timestamps; R.suppress_blank = p.suppress_blank; R.single_segment = p.single_segment; R.max_tokens = p.max_tokens;
    R.n_max = m.hp.n_text_ctx / 2 - 4;
    R.tid0_initial = p.max_initial_ts > 0.0f ? (int)std::round(p.max_initial_ts / (30.0f / (float)m.hp.n_audio_ctx)) : -1;
    R.temperature = temperature;
    return R;
}

// ---------------------------------------------------------------------------------------------------- lock-step pass
// Rows 0..n-1 decode one window each, in lock step, at one temperature: the temperature ladder of A.5 (t > 0: the
// distributions come back to the host and are sampled with the item's std::mt19937, A.7) and the t = 0 pass of a
// non-default base temperature.  Overwrites the whole row-state block, so it runs only while no greedy row is armed.
static int decode_pass(oh_state* st, const oh_full_params& p, bool no_timestamps, const std::vector<PassRow>& rows_in, float temperature,
                       std::vector<PassResult>& out) {
    const Model& m = *st->ctx->model;
    const Vocab& vc = m.vocab;
    const int rows = (int)rows_in.size();
    const int n_max = m.hp.n_text_ctx / 2 - 4;
    SeqLayout lay(st->max_batch);
    int32_t* hb = (int32_t*)st->pinned;
    DecodeSeqState hs = seq_view(hb, lay);
    memset(hb, 0, lay.total_words * 4);
    int max_prompt = 0;
    for (int r = 0; r < rows; ++r) {
        const PassRow& pr = rows_in[r];
        const Item& it = pr.sub->items[pr.item];
        const int pl = (int)pr.prompt.size();
        if (pl > kMaxPrompt || pl < 1) {
            set_error("prompt too long");
            return OH_ERR_DECODE;
        }
        max_prompt = std::max(max_prompt, pl);
        hs.slot[r] = pr.slot;
        hs.sslot[r] = r;
        hs.active[r] = 1;
        hs.pos[r] = 0;
        hs.cur_token[r] = pr.prompt[0];
        hs.prompt_len[r] = pl;
        for (int i = 0; i < pl; ++i) hs.prompt[(size_t)r * kMaxPrompt + i] = pr.prompt[i];
        hs.seek_delta[r] = 100 * 30;
        hs.seek[r] = it.seek;
        hs.seek_end[r] = it.seek_end;
    }
    *hs.n_active_total = rows;
    const DecodeRules R = make_rules(m, p, no_timestamps, temperature);
    cudaStream_t s = st->stream;
    OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
    OH_CUDA_CHECK(cudaMemcpyAsync(st->rules_dev.p, &R, sizeof R, cudaMemcpyHostToDevice, s));
    st->timings.h2d_bytes += (long long)lay.total_words * 4 + (long long)sizeof R;
    OH_CUDA_CHECK(cudaStreamSynchronize(s));  // R and hb are reused below
    st->sched.rules_valid = false;

    const int max_steps = max_prompt + n_max - 1;
    int rc;
    if (temperature < 1e-6f) {
        int* counter = (int*)((char*)st->pinned + lay.total_words * 4 + 64);
        for (int step = 0; step < max_steps; ++step) {
            if ((rc = run_decode_step_graph(st, rows))) return rc;
            st->timings.n_decode_steps++;
            if ((step & 3) == 3 || step == max_steps - 1) {
                OH_CUDA_CHECK(cudaMemcpyAsync(counter, st->seq.n_active_total, 4, cudaMemcpyDeviceToHost, s));
                st->timings.d2h_bytes += 4;
                OH_CUDA_CHECK(cudaStreamSynchronize(s));
                if (*counter <= 0) break;
            }
        }
        OH_CUDA_CHECK(cudaMemcpyAsync(hb, st->seq_dev.p, lay.total_words * 4, cudaMemcpyDeviceToHost, s));
        st->timings.d2h_bytes += (long long)lay.total_words * 4;
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
    } else {
        // t > 0: probabilities come back to the host and are sampled with the same libstdc++ calls as upstream
        std::vector<float> probs((size_t)rows * st->ld_logits), logprobs((size_t)rows * st->ld_logits);
        int n_live = rows;
        for (int step = 0; step < max_steps && n_live > 0; ++step) {
            if ((rc = run_decode_step(st, rows, true))) return rc;
            st->timings.n_decode_steps++;
            OH_CUDA_CHECK(cudaMemcpyAsync(probs.data(), st->dump_p.p, probs.size() * 4, cudaMemcpyDeviceToHost, s));
            OH_CUDA_CHECK(cudaMemcpyAsync(logprobs.data(), st->dump_lp.p, logprobs.size() * 4, cudaMemcpyDeviceToHost, s));
            st->timings.d2h_bytes += (long long)probs.size() * 8;
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
            for (int r = 0; r < rows; ++r) {
                if (!hs.active[r]) continue;
                const int pos = hs.pos[r];
                if (pos + 1 < hs.prompt_len[r]) {  // mirrors the kernel's prompt feeding
                    hs.pos[r] = pos + 1;
                    hs.cur_token[r] = hs.prompt[(size_t)r * kMaxPrompt + pos + 1];
                    continue;
                }
                TokenData t = sample_on_host(vc, probs.data() + (size_t)r * st->ld_logits, logprobs.data() + (size_t)r * st->ld_logits,
                                             (*rows_in[r].sub->rngs)[rows_in[r].item]);
                const int i = hs.n_gen[r];
                const size_t o = (size_t)r * kMaxGen + i;
                hs.tok_id[o] = t.id; hs.tok_tid[o] = t.tid; hs.tok_p[o] = t.p; hs.tok_plog[o] = t.plog; hs.tok_pt[o] = t.pt; hs.tok_ptsum[o] = t.ptsum;
                hs.n_gen[r] = i + 1;
                SeqFlags f{hs.has_ts[r], hs.seek_delta[r], hs.result_len[r], 0, 0};
                const bool live = apply_token_rules(f, t.id, i, hs.seek[r], hs.seek_end[r], vc.token_beg, vc.token_eot, p.max_tokens, no_timestamps,
                                                    p.single_segment, n_max);
                hs.has_ts[r] = f.has_ts; hs.seek_delta[r] = f.seek_delta; hs.result_len[r] = f.result_len;
                hs.failed[r] = f.failed; hs.completed[r] = f.completed;
                if (live) {
                    hs.cur_token[r] = t.id;
                    hs.pos[r] = pos + 1;
                } else {
                    hs.active[r] = 0;
                    n_live--;
                }
            }
            // scalars + token ids (the kernel reads the last two sampled ids)
            OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, (lay.tok_off[1]) * 4, cudaMemcpyHostToDevice, s));
            st->timings.h2d_bytes += (long long)lay.tok_off[1] * 4;
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
        }
    }
    out.assign(rows, PassResult());
    for (int r = 0; r < rows; ++r) {
        PassResult& pr = out[r];
        const int ng = hs.n_gen[r];
        pr.tokens.resize(ng);
        for (int i = 0; i < ng; ++i) {
            const size_t o = (size_t)r * kMaxGen + i;
            pr.tokens[i] = TokenData{hs.tok_id[o], hs.tok_tid[o], hs.tok_p[o], hs.tok_plog[o], hs.tok_pt[o], hs.tok_ptsum[o]};
        }
        pr.result_len = hs.result_len[r];
        pr.seek_delta = hs.seek_delta[r];
        pr.has_ts = hs.has_ts[r];
        pr.failed = hs.failed[r];
        pr.completed = hs.completed[r];
        st->timings.n_tokens += ng;
    }
    return OH_OK;
}

// ------------------------------------------------------------------------------------------------ beam search
// BASELINE config 4 / A.7 "Beam".  The reference never selects it (/root/reference/src/engine/whisper.rs:243 hard-codes
// Greedy) and the pinned upstream's candidate sampler cannot be verified offline, so this follows SURVEY.md A.5/A.7
// as written — per live decoder the top beam_size ids of the filtered distribution, candidates ranked by
// sum_logprobs_all, assigned to the decoders in order skipping repeats, KV cache inherited from the proposing decoder
// — decision for decision the same procedure the test oracle restates.
//
// Batched: the resident rows are cut into groups of beam_size rows; a group decodes one window, its rows are the window's
// decoders.  The decoders of a window share ONE cross-attention K/V slot (the window is encoded once and its K/V is
// streamed by beam_size rows that run side by side, so after the first row the reads hit L2) and own one self-attention
// cache line each; a line follows its candidate by index and is copied (one batched kernel per step) only when two
// decoders inherit it.  K9 runs in top-k mode: every row reports its beam_size best ids with logprob / prob and the
// timestamp statistics; the host ranks the candidates of each group and writes the rows' next state.  Windows of all live
// items decode at once (one host look-in per step), finished groups are re-armed from the encoder's look-ahead.
namespace {

}  // namespace


// --------------------------------------------------------------------------------------------------- scheduler
namespace {

cudaEvent_t take_event(Sched& sc) {
    if (!sc.event_pool.empty()) {
        cudaEvent_t e = sc.event_pool.back();
        sc.event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}

// device buffers of finished submissions are kept and handed to the next ones (cudaFree would synchronise the device)
DevBuf take_buf(Sched& sc, size_t bytes) {
    int best = -1;
    for (int i = 0; i < (int)sc.buf_pool.size(); ++i)
        if (sc.buf_pool[i].bytes >= bytes && (best < 0 || sc.buf_pool[i].bytes < sc.buf_pool[best].bytes)) best = i;
    DevBuf b;
    if (best >= 0) {
        b = sc.buf_pool[best];
        sc.buf_pool.erase(sc.buf_pool.begin() + best);
    }
    return b;  // empty: the caller's ensure() allocates
}
void give_buf(Sched& sc, DevBuf& b) {
    if (b.p) sc.buf_pool.push_back(b);
    b = DevBuf();
}

struct Pump {
    oh_state* st;
    Sched& sc;
    const Model& m;
    const HParams& hp;
    const Vocab& vc;
    const int R, n_max;
    SeqLayout lay;
    int32_t* hb;
    DecodeSeqState hs;
    size_t scalar_words, upload_words, token_words;

    explicit Pump(oh_state* s)
        : st(s), sc(s->sched), m(*s->ctx->model), hp(m.hp), vc(m.vocab), R(s->max_batch), n_max(hp.n_text_ctx / 2 - 4), lay(s->max_batch),
          hb((int32_t*)s->pinned), hs(seq_view(hb, lay)) {
        scalar_words = (size_t)seqlay::N_SCALARS * lay.rows;
        upload_words = lay.prompt_off + (size_t)lay.rows * kMaxPrompt;  // scalars + prompts
        token_words = lay.counter_off - lay.tok_off[0];
    }

    // ---- encoder side ------------------------------------------------------------------------------------------
    int inflight_windows() const {
        int n = 0;
        for (const EncoderPass& p : sc.passes) n += (int)p.wins.size();
        return n;
    }
    // collect finished encoder passes (their windows become ready); a device fault surfaces here instead of spinning
    int poll_passes(bool block_on_oldest) {
        while (!sc.passes.empty()) {
            EncoderPass& p = sc.passes.front();
            cudaError_t q = block_on_oldest ? cudaEventSynchronize(p.done) : cudaEventQuery(p.done);
            block_on_oldest = false;
            if (q == cudaErrorNotReady) {
                cudaGetLastError();
                break;
            }
            if (q != cudaSuccess) {
                set_error(std::string("encoder pass failed: ") + cudaGetErrorString(q));
                return OH_ERR_CUDA;
            }
            for (const WinRef& w : p.wins) sc.ready.push_back(w);
            sc.event_pool.push_back(p.done);
            sc.passes.pop_front();
        }
        return OH_OK;
    }
    int launch_pass(int k) {
        const auto t0 = clk::now();
        cudaStream_t es = st->async_encoder ? st->enc_stream : st->stream;
        EncoderPass pass;
        pass.done = take_event(sc);
        std::sort(sc.free_slots.begin(), sc.free_slots.end(), std::greater<int>());  // pop_back yields ascending ids: runs of
                                                                                       // consecutive slots share one cross-K/V mat-mul
        const int area = sc.desc_cursor++ % 4;
        WindowDesc* hd = st->win_desc_host + (size_t)area * st->enc_chunk;
        WindowDesc* dd = (WindowDesc*)st->win_desc.p + (size_t)area * st->enc_chunk;
        std::vector<int> slots(k);
        std::set<Submission*> subs_seen;
        for (int i = 0; i < k; ++i) {
            WinRef w = sc.need_window.front();
            sc.need_window.pop_front();
            w.slot = sc.free_slots.back();
            sc.free_slots.pop_back();
            slots[i] = w.slot;
            const Item& it = w.sub->items[w.item];
            hd[i].raw = (const float*)w.sub->mel_raw.p + (size_t)w.item * hp.n_mels * w.sub->mel_ld;
            hd[i].max_key = clip_max_key(w.sub->clip_meta, (int)w.sub->items.size(), w.item);
            hd[i].ld = w.sub->mel_ld;
            hd[i].n_active = it.n_active;
            hd[i].n_len = it.n_len;
            hd[i].seek = it.seek;
            pass.wins.push_back(w);
            subs_seen.insert(w.sub);
        }
        for (Submission* sub : subs_seen) OH_CUDA_CHECK(cudaStreamWaitEvent(es, sub->ev_mel, 0));  // the log-mel ran on the front-end stream
        OH_CUDA_CHECK(cudaMemcpyAsync(dd, hd, (size_t)k * sizeof(WindowDesc), cudaMemcpyHostToDevice, es));
        int rc;
        if ((rc = launch_window_gather(dd, hp.n_mels, k, st->xin.p, st->ctx->act_type, es))) return rc == OH_ERR_CUDA ? rc : OH_ERR_ENCODE;
        if ((rc = run_encoder(st, k, 0, nullptr, slots.data(), es))) return rc == OH_ERR_CUDA ? rc : OH_ERR_ENCODE;
        OH_CUDA_CHECK(cudaEventRecord(pass.done, es));
        sc.passes.push_back(std::move(pass));
        st->timings.n_windows += k;
        st->timings.encode_ms += ms_since(t0);  // host time to issue the pass (the pass itself overlaps the decode steps)
        return OH_OK;
    }
    // Encoder passes for waiting windows.  Two passes may be in flight; a pass is worth issuing when it is at least half a
    // chunk, or when rows would otherwise go unfed (then whatever is waiting goes).
    int launch_passes() {
        while ((int)sc.passes.size() < 2 && !sc.need_window.empty() && !sc.free_slots.empty() && sc.parked.empty()) {
            const int avail = (int)std::min<size_t>(std::min(sc.need_window.size(), sc.free_slots.size()), (size_t)st->enc_chunk);
            const int fed = sc.n_armed + (int)sc.ready.size() + inflight_windows();
            const bool starving = fed < R;
            // full passes run the mat-muls at their best; a half pass goes when the stock of encoded windows is getting low
            const bool low_stock = (int)sc.ready.size() + inflight_windows() < st->enc_chunk;
            if (!starving && avail < st->enc_chunk && !(low_stock && avail >= std::max(1, st->enc_chunk / 2))) break;
            int rc = launch_pass(avail);
            if (rc) return rc;
        }
        return OH_OK;
    }

    // ---- window results ------------------------------------------------------------------------------------------
    // accept a finished window: segments, prompt carry-over, seek advance, next window or done; the K/V slot is free again
    void finish_window(Submission* sub, int item, int slot, const std::vector<int>& prompt, const PassResult& res) {
        Item& it = sub->items[item];
        emit_segments(m, sub->p, it, res, prompt, (int)sub->prompt_init.size());
        if (it.seek + 100 >= it.seek_end) {
            it.done = true;
            sub->n_done++;
        } else {
            sc.need_window.push_front(WinRef{sub, item, -1});  // continuation windows first: items in progress finish sooner
        }
        sc.free_slots.push_back(slot);
    }
    // score one result; returns true if it is accepted at temperature index ti
    bool evaluate(const Submission* sub, PassResult& pr, size_t ti) {
        if (!pr.failed) {
            pr.tokens.resize(std::min<size_t>(pr.tokens.size(), (size_t)pr.result_len));
            sequence_score(sub->p, pr);
            if (pr.result_len > 32 && pr.entropy < sub->p.entropy_thold) pr.failed = true;
        }
        if (ti + 1 != sub->temperatures.size() && (pr.failed || pr.avg_logprobs < sub->p.logprob_thold)) return false;
        return true;
    }

    // ---- rows ----------------------------------------------------------------------------------------------------
    // Arm free rows with encoded windows.  Only called at a look-in (no step in flight, host mirror current).
    int arm_rows() {
        if (sc.ready.empty()) return OH_OK;
        for (int r = 0; r < R && !sc.ready.empty(); ++r) {
            RowInfo& ri = sc.rows[r];
            if (ri.armed) continue;
            const bool base_t_positive = sc.ready.front().sub->temperatures[0] >= 1e-6f;
            if (!sc.parked.empty() && !base_t_positive) break;  // rows drain for the temperature ladder
            if (base_t_positive && (int)sc.parked.size() >= R) break;
            WinRef w = sc.ready.front();
            sc.ready.pop_front();
            Item& it = w.sub->items[w.item];
            if (it.seek > 0 && it.seek + 500 >= it.seek_end) it.prompt_past.clear();  // short tail: drop the prompt (upstream behaviour)
            std::vector<int> prompt;
            build_prompt(*w.sub, hp, vc, it, w.sub->temperatures[0], prompt);
            const int pl = (int)prompt.size();
            if (pl > kMaxPrompt) {
                set_error("prompt too long");
                return OH_ERR_DECODE;
            }
            if (base_t_positive) {  // a non-zero base temperature samples on the host from the first pass on: lock-step path
                sc.parked.push_back(ParkedRow{w.sub, w.item, w.slot, std::move(prompt)});
                --r;
                continue;
            }
            ri.sub = w.sub;
            ri.item = w.item;
            ri.slot = w.slot;
            ri.armed = true;
            ri.prompt = std::move(prompt);
            hs.slot[r] = w.slot;
            hs.sslot[r] = r;
            hs.active[r] = 1;
            hs.pos[r] = 0;
            hs.cur_token[r] = ri.prompt[0];
            hs.prompt_len[r] = pl;
            for (int j = 0; j < pl; ++j) hs.prompt[(size_t)r * kMaxPrompt + j] = ri.prompt[j];
            hs.n_gen[r] = 0;
            hs.has_ts[r] = 0;
            hs.seek_delta[r] = 100 * 30;
            hs.result_len[r] = 0;
            hs.failed[r] = 0;
            hs.completed[r] = 0;
            hs.seek[r] = it.seek;
            hs.seek_end[r] = it.seek_end;
            st->timings.n_prompt_steps += pl - 1;
            sc.n_armed++;
            sc.mirror_dirty = true;
        }
        return OH_OK;
    }

    // temperature ladder for the parked windows (A.5), lock-step per temperature, up to R rows at a time
    int run_ladder() {
        const auto t0 = clk::now();
        cudaStream_t s = st->stream;
        while (!sc.parked.empty()) {
            std::vector<ParkedRow> batch;
            const size_t take = std::min<size_t>(sc.parked.size(), (size_t)R);
            batch.assign(std::make_move_iterator(sc.parked.begin()), std::make_move_iterator(sc.parked.begin() + take));
            sc.parked.erase(sc.parked.begin(), sc.parked.begin() + take);
            Submission* sub0 = batch[0].sub;  // live submissions share their params (sched_submit drains on a change)
            const size_t ti0 = sub0->temperatures[0] >= 1e-6f ? 0 : 1;
            std::vector<PassRow> pending;
            for (ParkedRow& pr : batch) pending.push_back(PassRow{pr.sub, pr.item, pr.slot, std::move(pr.prompt)});
            for (size_t ti = ti0; ti < sub0->temperatures.size() && !pending.empty(); ++ti) {
                for (PassRow& pr : pending) build_prompt(*pr.sub, hp, vc, pr.sub->items[pr.item], pr.sub->temperatures[ti], pr.prompt);
                std::vector<PassResult> res;
                int rc = decode_pass(st, sub0->p, sub0->no_timestamps, pending, sub0->temperatures[ti], res);
                if (rc) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
                std::vector<PassRow> next;
                for (size_t r = 0; r < pending.size(); ++r) {
                    if (evaluate(pending[r].sub, res[r], ti)) finish_window(pending[r].sub, pending[r].item, pending[r].slot, pending[r].prompt, res[r]);
                    else {
                        st->timings.n_fallbacks++;
                        next.push_back(std::move(pending[r]));
                    }
                }
                pending.swap(next);
            }
        }
        // decode_pass reused the row-state block: the greedy rows start from a clean slate
        memset(hb, 0, lay.total_words * 4);
        OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
        sc.mirror_dirty = false;
        sc.rules_valid = false;
        st->timings.decode_ms += ms_since(t0);
        return OH_OK;
    }

    int ensure_rules(const Submission* sub) {
        const DecodeRules Rn = make_rules(m, sub->p, sub->no_timestamps, sub->temperatures[0]);
        if (sc.rules_valid && memcmp(&Rn, &sc.rules, sizeof Rn) == 0) return OH_OK;
        sc.rules = Rn;  // the source of the async copy must stay put: it lives in the scheduler
        OH_CUDA_CHECK(cudaMemcpyAsync(st->rules_dev.p, &sc.rules, sizeof sc.rules, cudaMemcpyHostToDevice, st->stream));
        OH_CUDA_CHECK(cudaStreamSynchronize(st->stream));
        sc.rules_valid = true;
        return OH_OK;
    }

    // `check_steps` greedy steps over every row (inactive rows exit at once); asynchronous
    clk::time_point t_steps;
    int launch_steps() {
        t_steps = clk::now();
        cudaStream_t s = st->stream;
        int rc;
        Submission* any = nullptr;
        for (int r = 0; r < R && !any; ++r)
            if (sc.rows[r].armed) any = sc.rows[r].sub;
        if ((rc = ensure_rules(any))) return rc;
        if (sc.mirror_dirty) {
            OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, upload_words * 4, cudaMemcpyHostToDevice, s));
            st->timings.h2d_bytes += (long long)upload_words * 4;
            sc.mirror_dirty = false;
        }
        if (st->trace) {  // OH_SCHED_TRACE=<file>: one line per look-in (decode step count, rows armed, stock, queue, passes, tickets)
            fprintf(st->trace, "%lld,%d,%zu,%zu,%zu,%zu,%zu\n", st->trace_steps, sc.n_armed, sc.ready.size(), sc.need_window.size(), sc.passes.size(),
                    sc.free_slots.size(), sc.subs.size());
            st->trace_steps += st->check_steps;
        }
        g_active_rows_hint = sc.n_armed;
        for (int step = 0; step < st->check_steps; ++step) {
            rc = run_decode_step_graph(st, R);
            if (rc) g_active_rows_hint = -1;
            if (rc) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
            st->timings.n_decode_steps++;
            st->timings.n_row_steps += R;
            st->timings.n_armed_row_steps += sc.n_armed;
        }
        g_active_rows_hint = -1;
        return OH_OK;
    }
    // the look-in: wait for the steps, read the rows' scalars, collect the rows that finished
    int look() {
        const auto t0 = t_steps;
        cudaStream_t s = st->stream;
        OH_CUDA_CHECK(cudaMemcpyAsync(hb, st->seq_dev.p, scalar_words * 4, cudaMemcpyDeviceToHost, s));
        OH_CUDA_CHECK(cudaStreamSynchronize(s));
        st->timings.d2h_bytes += (long long)scalar_words * 4;
        bool any_finished = false;
        for (int r = 0; r < R; ++r)
            if (sc.rows[r].armed && !hs.active[r]) any_finished = true;
        if (any_finished) {
            OH_CUDA_CHECK(cudaMemcpyAsync(hb + lay.tok_off[0], (int32_t*)st->seq_dev.p + lay.tok_off[0], token_words * 4, cudaMemcpyDeviceToHost, s));
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
            st->timings.d2h_bytes += (long long)token_words * 4;
            for (int r = 0; r < R; ++r) {
                RowInfo& ri = sc.rows[r];
                if (!ri.armed || hs.active[r]) continue;
                PassResult pr;
                const int ng = hs.n_gen[r];
                pr.tokens.resize(ng);
                for (int i = 0; i < ng; ++i) {
                    const size_t o = (size_t)r * kMaxGen + i;
                    pr.tokens[i] = TokenData{hs.tok_id[o], hs.tok_tid[o], hs.tok_p[o], hs.tok_plog[o], hs.tok_pt[o], hs.tok_ptsum[o]};
                }
                pr.result_len = hs.result_len[r];
                pr.seek_delta = hs.seek_delta[r];
                pr.has_ts = hs.has_ts[r];
                pr.failed = hs.failed[r];
                pr.completed = hs.completed[r];
                st->timings.n_tokens += ng;
                ri.armed = false;
                sc.n_armed--;
                if (evaluate(ri.sub, pr, 0)) finish_window(ri.sub, ri.item, ri.slot, ri.prompt, pr);
                else {
                    st->timings.n_fallbacks++;
                    sc.parked.push_back(ParkedRow{ri.sub, ri.item, ri.slot, ri.prompt});
                }
                ri.sub = nullptr;
            }
        }
        st->timings.decode_ms += ms_since(t0);
        return OH_OK;
    }

    // ---- beam search: groups of beam_size rows (state lives in the scheduler: windows outlive a wait) -------------------
    std::vector<BeamGroup>& groups = sc.beam_groups;
    std::vector<int4>& kv_copies = sc.beam_kv_copies;  // (src line, dst line, positions) queued for the next step

    void beam_set_idle_rows(int g, int nb) {
        for (int j = 0; j < nb; ++j) hs.active[g * nb + j] = 0;
    }
    int beam_start(int g, int nb, const WinRef& w) {
        BeamGroup& G = groups[g];
        Item& it = w.sub->items[w.item];
        if (it.seek > 0 && it.seek + 500 >= it.seek_end) it.prompt_past.clear();  // short tail: drop the prompt (upstream behaviour)
        std::vector<int> prompt;
        build_prompt(*w.sub, hp, vc, it, w.sub->temperatures[0], prompt);
        const int pl = (int)prompt.size();
        if (pl > kMaxPrompt || pl < 1) {
            set_error("prompt too long");
            return OH_ERR_DECODE;
        }
        if (w.sub->temperatures[0] > 0.0f) {  // A.5: at t > 0 beam search runs best_of sampled decoders: the lock-step ladder path
            sc.parked.push_back(ParkedRow{w.sub, w.item, w.slot, std::move(prompt)});
            return OH_OK;
        }
        G = BeamGroup();
        G.busy = true;
        G.sub = w.sub;
        G.item = w.item;
        G.xslot = w.slot;
        G.prompt = std::move(prompt);
        G.dec.assign(nb, BeamDec());
        for (int j = 0; j < nb; ++j) {
            const int r = g * nb + j;
            hs.slot[r] = w.slot;
            hs.sslot[r] = r;
            hs.active[r] = j == 0 ? 1 : 0;
            hs.pos[r] = 0;
            hs.cur_token[r] = G.prompt[0];
            hs.prompt_len[r] = pl;
            for (int q = 0; q < pl; ++q) hs.prompt[(size_t)r * kMaxPrompt + q] = G.prompt[q];
            hs.n_gen[r] = 0;
            hs.has_ts[r] = 0;
            hs.seek_delta[r] = 100 * 30;
            hs.result_len[r] = 0;
            hs.failed[r] = 0;
            hs.completed[r] = 0;
            hs.seek[r] = it.seek;
            hs.seek_end[r] = it.seek_end;
            G.dec[j].line = r;
        }
        st->timings.n_prompt_steps += pl - 1;
        return OH_OK;
    }
    // candidates of row r from the step's top-k block (beam_size best; a distribution with fewer live ids is padded with
    // the lowest absent ids at logprob -inf, which is what a partial sort over the whole vocabulary yields)
    void beam_read_cands(const int32_t* tk, int r, int nb, std::vector<TokenData>& out) {
        const size_t Rr = (size_t)R;
        const int32_t* ids = tk + (size_t)r * kTopK;
        const float* lps = (const float*)(tk + Rr * kTopK) + (size_t)r * kTopK;
        const float* ps = (const float*)(tk + 2 * Rr * kTopK) + (size_t)r * kTopK;
        TokenData base;
        base.tid = tk[3 * Rr * kTopK + r];
        base.pt = ((const float*)(tk + 3 * Rr * kTopK + Rr))[r];
        base.ptsum = ((const float*)(tk + 3 * Rr * kTopK + 2 * Rr))[r];
        out.clear();
        for (int c = 0; c < nb; ++c) {
            TokenData t = base;
            if (ids[c] >= 0) {
                t.id = ids[c];
                t.p = ps[c];
                t.plog = lps[c];
            } else {
                int cand = 0;
                auto used = [&](int id) {
                    for (const TokenData& u : out)
                        if (u.id == id) return true;
                    return false;
                };
                while (used(cand)) ++cand;
                t.id = cand;
                t.p = 0.0f;
                t.plog = -INFINITY;
            }
            if (t.id >= vc.token_beg) {
                t.tid = t.id;
                t.pt = t.p;
            }
            out.push_back(t);
        }
    }
    // one A.5 iteration for group g after a step: rank candidates, hand them out, move / copy cache lines, apply the
    // per-token rules; returns true when the window is finished (result in `best`)
    int beam_select(int g, int nb, bool& finished, PassResult& best) {
        BeamGroup& G = groups[g];
        std::vector<BeamDec>& dec = G.dec;
        const Item& it = G.sub->items[G.item];
        const oh_full_params& p = G.sub->p;
        const int pl = (int)G.prompt.size(), i = G.i;
        finished = false;
        std::vector<BeamCand> cands;
        for (int j = 0; j < nb; ++j) {
            BeamDec& d = dec[j];
            if (d.completed || d.failed) continue;
            for (const TokenData& tok : d.cand) {
                BeamCand c{j, d.seek_delta, d.result_len, d.has_ts, d.tokens, d.sum_logprobs_all + tok.plog};
                c.tokens.push_back(tok);
                cands.push_back(std::move(c));
            }
        }
        std::stable_sort(cands.begin(), cands.end(), [](const BeamCand& a, const BeamCand& b) {
            if (a.sum_logprobs_all != b.sum_logprobs_all) return a.sum_logprobs_all > b.sum_logprobs_all;
            return a.decoder_idx < b.decoder_idx;
        });
        auto same_tokens = [](const BeamCand& a, const BeamCand& b) {
            if (a.tokens.size() != b.tokens.size()) return false;
            for (size_t q = 0; q < a.tokens.size(); ++q)
                if (a.tokens[q].id != b.tokens[q].id) return false;
            return true;
        };
        std::vector<int> take(nb, -1);
        size_t cur_c = 0;
        for (int j = 0; j < nb; ++j) {
            if (dec[j].completed || dec[j].failed) continue;
            if (cur_c >= cands.size()) cur_c = 0;
            const size_t pick = cur_c++;
            while (cur_c < cands.size() && i > 0 && same_tokens(cands[cur_c], cands[pick])) ++cur_c;
            take[j] = (int)pick;
        }
        // ---- cache lines follow the candidates: a line moves with its index; only a line that two decoders inherit (or that
        // its owner keeps while another decoder inherits it) is copied, into the line of a decoder that nobody follows
        std::vector<int> old_line(nb), new_line(nb, -1);
        std::vector<char> claimed(nb, 0);  // indexed by line - g * nb
        const int l0 = g * nb;
        for (int j = 0; j < nb; ++j) old_line[j] = dec[j].line;
        for (int j = 0; j < nb; ++j)
            if (take[j] >= 0 && cands[take[j]].decoder_idx == j) {
                new_line[j] = old_line[j];
                claimed[old_line[j] - l0] = 1;
            }
        for (int j = 0; j < nb; ++j)
            if (take[j] >= 0 && new_line[j] < 0) {
                const int src = old_line[cands[take[j]].decoder_idx];
                if (!claimed[src - l0]) {
                    new_line[j] = src;
                    claimed[src - l0] = 1;
                }
            }
        for (int j = 0; j < nb; ++j)
            if (take[j] >= 0 && new_line[j] < 0) {
                int f = -1;
                for (int q = 0; q < nb; ++q)
                    if (!claimed[old_line[q] - l0]) {
                        f = old_line[q];
                        break;
                    }
                if (f < 0) {
                    set_error("beam search: no free cache line");
                    return OH_ERR_DECODE;
                }
                kv_copies.push_back(make_int4(old_line[cands[take[j]].decoder_idx], f, pl + i, 0));
                new_line[j] = f;
                claimed[f - l0] = 1;
            }
        {   // decoders that are finished keep distinct (unused) lines
            int q = 0;
            for (int j = 0; j < nb; ++j)
                if (new_line[j] < 0) {
                    while (claimed[old_line[q] - l0]) ++q;
                    new_line[j] = old_line[q];
                    claimed[old_line[q] - l0] = 1;
                }
        }
        for (int j = 0; j < nb; ++j) {
            dec[j].line = new_line[j];
            if (take[j] < 0) continue;
            const BeamCand& c = cands[take[j]];
            dec[j].tokens = c.tokens;
            dec[j].seek_delta = c.seek_delta;
            dec[j].has_ts = c.has_ts;
            dec[j].result_len = c.result_len;
            dec[j].sum_logprobs_all = c.sum_logprobs_all;
        }
        // ---- A.5 bookkeeping for the token each live decoder ended up with
        bool all_done = true;
        for (int j = 0; j < nb; ++j) {
            BeamDec& d = dec[j];
            if (d.completed || d.failed) continue;
            SeqFlags f{d.has_ts ? 1 : 0, d.seek_delta, d.result_len, 0, 0};
            apply_token_rules(f, d.tokens.back().id, i, it.seek, it.seek_end, vc.token_beg, vc.token_eot, p.max_tokens, G.sub->no_timestamps,
                              p.single_segment, n_max);
            d.has_ts = f.has_ts != 0;
            d.seek_delta = f.seek_delta;
            d.result_len = f.result_len;
            d.failed = f.failed != 0;
            d.completed = f.completed != 0;
            if (!d.completed && !d.failed) all_done = false;
        }
        if (all_done || i + 1 == n_max) {
            // ---- rank the decoders (A.5 / A.8): first best score among the ones that did not fail
            double best_score = -INFINITY;
            int best_id = 0;
            std::vector<PassResult> res(nb);
            for (int j = 0; j < nb; ++j) {
                PassResult& r = res[j];
                const BeamDec& d = dec[j];
                r.tokens = d.tokens;
                r.result_len = d.result_len;
                r.seek_delta = d.seek_delta;
                r.has_ts = d.has_ts;
                r.failed = d.failed;
                r.completed = d.completed;
                st->timings.n_tokens += (int)d.tokens.size();
                if (r.failed) continue;
                r.tokens.resize(std::min<size_t>(r.tokens.size(), (size_t)r.result_len));
                sequence_score(p, r);
                if (r.result_len > 32 && r.entropy < p.entropy_thold) {
                    r.failed = true;
                    continue;
                }
                if (best_score < r.score) {
                    best_score = r.score;
                    best_id = j;
                }
            }
            best = res[best_id];
            finished = true;
            return OH_OK;
        }
        // ---- one token per live decoder
        for (int j = 0; j < nb; ++j) {
            const BeamDec& d = dec[j];
            const int r = g * nb + j;
            const bool live = !d.completed && !d.failed;
            hs.active[r] = live ? 1 : 0;
            hs.sslot[r] = d.line;
            if (!live) continue;
            hs.pos[r] = pl + i;
            hs.cur_token[r] = d.tokens.back().id;
            hs.n_gen[r] = (int)d.tokens.size();
            hs.has_ts[r] = d.has_ts ? 1 : 0;
            hs.seek_delta[r] = d.seek_delta;
            hs.result_len[r] = d.result_len;
            for (size_t q = 0; q < d.tokens.size(); ++q) hs.tok_id[(size_t)r * kMaxGen + q] = d.tokens[q].id;
        }
        G.i = i + 1;
        return OH_OK;
    }

    int run_beam(const Submission* target, int nb) {
        const int NG = R / nb, rows_used = NG * nb;
        cudaStream_t s = st->stream;
        const size_t tk_words = (size_t)R * (3 * kTopK + 3);
        std::vector<int32_t> tk(tk_words);
        int rc;
        if ((int)groups.size() != NG || sc.beam_nb != nb) {
            // first beam pass of this policy (a policy change drains the scheduler first, so nothing is in flight): the
            // groups start idle and whatever the row-state block held belongs to nobody
            groups.assign(NG, BeamGroup());
            kv_copies.clear();
            sc.beam_nb = nb;
            memset(hb, 0, lay.total_words * 4);
        }
        if ((rc = ensure(st->beam_pairs, (size_t)R * sizeof(int4)))) return rc;
        while (target ? !target->finished() : true) {
            if ((rc = poll_passes(false))) return rc;
            if ((rc = launch_passes())) return rc;
            int n_busy = 0;
            for (int g = 0; g < NG; ++g) {
                if (!groups[g].busy && !sc.ready.empty() && sc.parked.empty()) {
                    WinRef w = sc.ready.front();
                    sc.ready.pop_front();
                    if ((rc = beam_start(g, nb, w))) return rc;
                }
                n_busy += groups[g].busy ? 1 : 0;
            }
            if (n_busy > 0 && NG - n_busy >= std::max(1, st->starve_rows / nb) && sc.ready.empty() && !sc.passes.empty() && sc.parked.empty()) {
                if ((rc = poll_passes(true))) return rc;  // decode yields to the encoder pass in flight when groups run empty (see run())
                continue;
            }
            if (n_busy == 0) {
                if (!sc.parked.empty()) {
                    if ((rc = run_ladder())) return rc;
                    memset(hb, 0, lay.total_words * 4);
                    continue;
                }
                if (!sc.ready.empty()) continue;
                if (!sc.passes.empty()) {
                    if ((rc = poll_passes(true))) return rc;
                    continue;
                }
                if (!sc.need_window.empty()) {
                    if (sc.free_slots.empty()) {
                        set_error("scheduler: windows waiting, nothing running and no free K/V slot");
                        return OH_ERR_DECODE;
                    }
                    continue;
                }
                break;
            }
            const auto t0 = clk::now();
            if ((rc = ensure_rules(groups[std::find_if(groups.begin(), groups.end(), [](const BeamGroup& G) { return G.busy; }) - groups.begin()].sub))) return rc;
            // rows' state for this step (the host is authoritative in beam mode), queued cache-line copies, the step, the candidates back
            OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.tok_off[1] * 4, cudaMemcpyHostToDevice, s));
            st->timings.h2d_bytes += (long long)lay.tok_off[1] * 4;
            if (!kv_copies.empty()) {
                OH_CUDA_CHECK(cudaMemcpyAsync(st->beam_pairs.p, kv_copies.data(), kv_copies.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
                if ((rc = launch_copy_self_lines(st->kself.p, st->vself.p, st->esz, (const int4*)st->beam_pairs.p, (int)kv_copies.size(), hp.n_text_layer, R,
                                                 hp.n_text_state / kHeadDim, hp.n_text_ctx, s)))
                    return rc;
            }
            g_active_rows_hint = n_busy * nb;
            rc = run_decode_step_graph(st, rows_used, K9_TOPK);
            g_active_rows_hint = -1;
            if (rc) return rc == OH_ERR_CUDA ? rc : OH_ERR_DECODE;
            st->timings.n_decode_steps++;
            st->timings.n_row_steps += rows_used;
            st->timings.n_armed_row_steps += n_busy * nb;
            OH_CUDA_CHECK(cudaMemcpyAsync(tk.data(), st->topk_dev.p, tk_words * 4, cudaMemcpyDeviceToHost, s));
            OH_CUDA_CHECK(cudaStreamSynchronize(s));  // also: kv_copies (pageable) was consumed
            st->timings.d2h_bytes += (long long)tk_words * 4;
            kv_copies.clear();
            for (int g = 0; g < NG; ++g) {
                BeamGroup& G = groups[g];
                if (!G.busy) continue;
                const int r0 = g * nb, pl = (int)G.prompt.size();
                if (G.in_prompt) {
                    G.fed++;
                    if (G.fed < pl) {  // A.5: "decode(prompt)" through decoder 0, one token per step
                        hs.pos[r0] = G.fed;
                        hs.cur_token[r0] = G.prompt[G.fed];
                        continue;
                    }
                    // the prompt is in: decoder 0's distribution and cache go to every decoder ("copy logits/probs/kv")
                    G.in_prompt = false;
                    beam_read_cands(tk.data(), r0, nb, G.dec[0].cand);
                    for (int j = 1; j < nb; ++j) {
                        G.dec[j].cand = G.dec[0].cand;
                        kv_copies.push_back(make_int4(r0, r0 + j, pl, 0));
                    }
                } else {
                    for (int j = 0; j < nb; ++j)
                        if (!G.dec[j].completed && !G.dec[j].failed) beam_read_cands(tk.data(), r0 + j, nb, G.dec[j].cand);
                }
                bool finished = false;
                PassResult best;
                const size_t copies_mark = kv_copies.size();
                const bool first_selection = G.i == 0;
                if ((rc = beam_select(g, nb, finished, best))) return rc;
                if (first_selection)  // every line of the group is (about to be) a copy of decoder 0's: read the original, the
                    for (size_t q = copies_mark; q < kv_copies.size(); ++q) kv_copies[q].x = r0;  // copies of one batch run unordered
                if (finished) {
                    // cache-line copies queued by this very selection are moot
                    kv_copies.erase(std::remove_if(kv_copies.begin(), kv_copies.end(), [&](const int4& c) { return c.y >= r0 && c.y < r0 + nb; }),
                                    kv_copies.end());
                    // A.5: fall back to the temperature ladder when the best beam fails the checks
                    const Submission* sub = G.sub;
                    const bool accept = sub->temperatures.size() == 1 || !(best.failed || best.avg_logprobs < sub->p.logprob_thold);
                    if (accept) finish_window(G.sub, G.item, G.xslot, G.prompt, best);
                    else {
                        st->timings.n_fallbacks++;
                        sc.parked.push_back(ParkedRow{G.sub, G.item, G.xslot, G.prompt});
                    }
                    beam_set_idle_rows(g, nb);
                    G.busy = false;
                }
            }
            st->timings.decode_ms += ms_since(t0);
        }
        bool any_busy = false;
        for (const BeamGroup& G : groups) any_busy |= G.busy;
        if (!any_busy) {  // idle: leave a clean row-state block behind (the greedy path relies on it)
            groups.clear();
            kv_copies.clear();
            memset(hb, 0, lay.total_words * 4);
            OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, hb, lay.total_words * 4, cudaMemcpyHostToDevice, s));
            OH_CUDA_CHECK(cudaStreamSynchronize(s));
            sc.mirror_dirty = false;
        }
        return OH_OK;
    }

    // run until `target` has finished (nullptr: until nothing is left)
    int run(const Submission* target) {
        int rc;
        // live submissions share one decode policy (sched_submit): beam search has its own loop over row groups
        for (auto& sb : sc.subs)
            if (!sb->finished() && sb->p.strategy == OH_SAMPLING_BEAM_SEARCH) return run_beam(target, std::max(1, sb->p.beam_size));
        while (target ? !target->finished() : true) {
            if ((rc = poll_passes(false))) return rc;
            if (sc.n_armed < R && sc.ready.empty() && sc.passes.empty() && (rc = launch_passes())) return rc;  // nothing encoded, nothing in flight
            if ((rc = arm_rows())) return rc;
            // Decode yields to the encoder when rows run empty.  An encoder pass issued next to the decode steps only gets
            // the SMs the steps leave idle (its persistent mat-mul needs a whole SM's shared memory, and the steps' CTAs keep
            // every SM partly occupied), so without this the rows settle at ~60 % occupancy; with it the pass runs alone
            // for ~100 ms exactly when windows are short, and the look-ahead slots keep that rare in steady state.
            if (sc.n_armed > 0 && R - sc.n_armed >= st->starve_rows && sc.ready.empty() && !sc.passes.empty() && sc.parked.empty()) {
                if ((rc = poll_passes(true))) return rc;
                if ((rc = launch_passes())) return rc;
                if ((rc = arm_rows())) return rc;
            }
            if (sc.n_armed == 0) {
                if (!sc.parked.empty()) {
                    if ((rc = run_ladder())) return rc;
                    continue;
                }
                if (!sc.ready.empty()) continue;  // arm_rows was held back by the ladder that just ran
                if (!sc.passes.empty()) {
                    if ((rc = poll_passes(true))) return rc;
                    continue;
                }
                if (!sc.need_window.empty()) {
                    if (sc.free_slots.empty()) {
                        set_error("scheduler: windows waiting, nothing running and no free K/V slot");
                        return OH_ERR_DECODE;
                    }
                    if ((rc = launch_passes())) return rc;  // nothing is fed, so any size goes
                    continue;
                }
                break;  // idle
            }
            // the steps go first; the host then issues the next encoder passes (hundreds of launches) while the GPU decodes
            if ((rc = launch_steps())) return rc;
            if ((rc = launch_passes())) return rc;
            if ((rc = look())) return rc;
        }
        return OH_OK;
    }
};

bool same_policy(const oh_full_params& a, const oh_full_params& b, const std::string& la, const std::string& lb) {
    return a.strategy == b.strategy && a.n_max_text_ctx == b.n_max_text_ctx && a.translate == b.translate && a.no_context == b.no_context &&
           a.no_timestamps == b.no_timestamps && a.single_segment == b.single_segment && a.print_special == b.print_special &&
           a.max_tokens == b.max_tokens && a.suppress_blank == b.suppress_blank && a.temperature == b.temperature &&
           a.max_initial_ts == b.max_initial_ts && a.length_penalty == b.length_penalty && a.temperature_inc == b.temperature_inc &&
           a.entropy_thold == b.entropy_thold && a.logprob_thold == b.logprob_thold && a.beam_size == b.beam_size && la == lb;
}

}  // namespace

int sched_pending(const oh_state* st) { return (int)st->sched.subs.size(); }

int sched_submit(oh_state* st, const oh_full_params& p, const float* const* pcm, const int* n, int batch, bool pcm_on_device, bool persistent_rngs,
                 int* ticket, bool carry_context) {
    const Model& m = *st->ctx->model;
    const Vocab& vc = m.vocab;
    const HParams& hp = m.hp;
    Sched& sc = st->sched;
    if (batch < 1) {
        set_error("empty batch");
        return OH_ERR_INVALID;
    }
    if (p.strategy != OH_SAMPLING_GREEDY && p.strategy != OH_SAMPLING_BEAM_SEARCH) {
        set_error("unknown sampling strategy");
        return OH_ERR_UNSUPPORTED;
    }
    if (p.suppress_non_speech_tokens || p.offset_ms != 0 || p.duration_ms != 0) {
        set_error("unsupported oh_full_params (suppress_non_speech_tokens / offset_ms / duration_ms)");
        return OH_ERR_UNSUPPORTED;
    }
    if (p.strategy == OH_SAMPLING_GREEDY && p.greedy_best_of > 1) {
        set_error("greedy.best_of > 1 is not supported (the reference always passes best_of = 1, whisper.rs:243)");
        return OH_ERR_UNSUPPORTED;
    }
    const auto t_begin = clk::now();
    int rc;
    const std::string lang = (p.language && strcmp(p.language, "auto") != 0 && p.language[0]) ? p.language : "en";
    // one decode policy at a time on the device (the rules block is per step): a submission with different params waits
    // for the live ones to finish
    if (p.strategy == OH_SAMPLING_BEAM_SEARCH && (p.beam_size < 1 || p.beam_size > kTopK || p.beam_size > st->max_batch)) {
        set_error("beam search needs 1 <= beam_size <= " + std::to_string(kTopK) + " and oh_context_params.max_batch >= beam_size (one decoder row per beam)");
        return OH_ERR_INVALID;
    }
    bool drain = false;
    for (auto& s : sc.subs)
        if (!s->finished() && !same_policy(s->p, p, s->language, lang)) drain = true;
    if (drain) {
        Pump pump(st);
        if ((rc = pump.run(nullptr))) return rc;
    }

    if (sc.subs.empty()) {
        // starting from idle: the stage entry points (oh_decode_batch ...) may have used the row-state block meanwhile
        SeqLayout lay(st->max_batch);
        memset(st->pinned, 0, lay.total_words * 4);
        OH_CUDA_CHECK(cudaMemcpyAsync(st->seq_dev.p, st->pinned, lay.total_words * 4, cudaMemcpyHostToDevice, st->stream));
        OH_CUDA_CHECK(cudaStreamSynchronize(st->stream));
        sc.mirror_dirty = false;
        sc.rules_valid = false;
    }
    std::unique_ptr<Submission> sub(new Submission());
    sub->ticket = sc.next_ticket++;
    sub->p = p;
    sub->language = lang;
    sub->p.language = nullptr;  // owned copy in `language`
    sub->t_begin = t_begin;
    if (p.temperature_inc > 0.0f) {
        for (float t = p.temperature; t < 1.0f + 1e-6f; t += p.temperature_inc) sub->temperatures.push_back(t);
    } else {
        sub->temperatures.push_back(p.temperature);
    }
    sub->prompt_init = {vc.token_sot};
    sub->no_timestamps = p.no_timestamps != 0;
    int lang_id = 0;
    if (vc.is_multilingual()) {
        lang_id = lang_id_from_code(lang.c_str());
        if (lang_id < 0 || lang_id >= vc.num_languages()) lang_id = 0;
        sub->prompt_init.push_back(vc.token_sot + 1 + lang_id);
        sub->prompt_init.push_back(p.translate ? vc.token_translate : vc.token_transcribe);
    }
    if (hp.n_text_layer == 2 && hp.n_vocab != 51866) sub->no_timestamps = true;  // "distilled" heuristic (A.1)
    if (sub->no_timestamps) sub->prompt_init.push_back(vc.token_not);

    // the prompts of earlier calls survive only with no_context = false (blocking calls: item i continues item i)
    std::vector<std::vector<int>> carried(batch);
    if (!p.no_context && carry_context)
        for (int i = 0; i < batch && i < (int)st->items.size(); ++i) carried[i] = st->items[i].prompt_past;

    // ---- front end: PCM upload + log-mel on the front-end stream
    MelPlan plan = plan_mel(sub->items, n, batch);
    sub->mel_ld = plan.ld;
    sub->pcm_stage = pcm_on_device ? DevBuf() : take_buf(sc, std::max<size_t>(plan.stage_floats, 4) * 4);
    sub->mel_raw = take_buf(sc, (size_t)batch * hp.n_mels * plan.ld * 4);
    sub->clip_meta = take_buf(sc, (size_t)batch * 20);
    sub->ev_mel = take_event(sc);
    rc = launch_mel(st, plan, sub->items, pcm, pcm_on_device, sub->pcm_stage, sub->mel_raw, sub->clip_meta, st->fe_stream, &st->timings.h2d_bytes);
    if (rc == OH_OK && cudaEventRecord(sub->ev_mel, st->fe_stream) != cudaSuccess) rc = OH_ERR_CUDA;
    if (rc == OH_OK && cudaEventSynchronize(sub->ev_mel) != cudaSuccess) rc = OH_ERR_CUDA;  // the caller's PCM is borrowed for the call only
    if (rc) {
        if (rc == OH_ERR_CUDA) set_error(std::string("front end (PCM upload / log-mel): ") + cudaGetErrorString(cudaGetLastError()));
        cudaStreamSynchronize(st->fe_stream);
        give_buf(sc, sub->pcm_stage);
        give_buf(sc, sub->mel_raw);
        give_buf(sc, sub->clip_meta);
        sc.event_pool.push_back(sub->ev_mel);
        return rc;
    }
    st->timings.mel_ms += ms_since(t_begin);

    if (persistent_rngs) {
        if ((int)st->rngs.size() < batch) st->rngs.resize(batch, std::mt19937(0));
        sub->rngs = &st->rngs;
    } else {
        sub->own_rngs.assign(batch, std::mt19937(0));
        sub->rngs = &sub->own_rngs;
    }
    Submission* sp = sub.get();
    for (int i = 0; i < batch; ++i) {
        Item& it = sp->items[i];
        it.lang_id = lang_id;
        it.prompt_past = carried[i];
        it.seek = 0;
        it.seek_end = it.n_len_org;
        if (!it.done && it.seek_end < 100) it.done = true;  // < 1 s: no segments, rc 0 (A.3 step 7)
        if (!it.done && it.seek + 100 >= it.seek_end) it.done = true;
        if (it.done) sp->n_done++;
        else sc.need_window.push_back(WinRef{sp, i, -1});
    }
    sc.subs.push_back(std::move(sub));
    if (ticket) *ticket = sp->ticket;
    return OH_OK;
}

int sched_wait(oh_state* st, int ticket) {
    Sched& sc = st->sched;
    Submission* sub = nullptr;
    size_t idx = 0;
    for (size_t i = 0; i < sc.subs.size(); ++i)
        if (sc.subs[i]->ticket == ticket) {
            sub = sc.subs[i].get();
            idx = i;
        }
    if (!sub) {
        set_error("unknown ticket");
        return OH_ERR_INVALID;
    }
    Pump pump(st);
    int rc = pump.run(sub);
    if (rc) return rc;
    st->items = std::move(sub->items);
    give_buf(sc, sub->pcm_stage);
    give_buf(sc, sub->mel_raw);
    give_buf(sc, sub->clip_meta);
    if (sub->ev_mel) sc.event_pool.push_back(sub->ev_mel);
    st->timings.total_ms = ms_since(sub->t_begin);
    st->timings.n_kernel_launches = g_launch_count - st->launches_mark;
    st->launches_mark = g_launch_count;
    st->timings_out = st->timings;
    st->timings = oh_timings{};
    sc.subs.erase(sc.subs.begin() + idx);
    return OH_OK;
}

int run_full(oh_state* st, const oh_full_params& p, const float* const* pcm, const int* n, int batch, bool pcm_on_device, bool persistent_rng) {
    int ticket = 0;
    int rc = sched_submit(st, p, pcm, n, batch, pcm_on_device, persistent_rng, &ticket, /*carry_context=*/true);
    if (rc == OH_OK) rc = sched_wait(st, ticket);
    if (rc && rc != OH_ERR_INVALID && rc != OH_ERR_UNSUPPORTED) sched_abort(st);  // a device-side failure: nothing half-done stays behind
    return rc;
}

// drop every live submission and reset the rows / slots (after an error inside the scheduler)
void sched_abort(oh_state* st) {
    Sched& sc = st->sched;
    cudaStreamSynchronize(st->stream);
    if (st->enc_stream) cudaStreamSynchronize(st->enc_stream);
    if (st->fe_stream) cudaStreamSynchronize(st->fe_stream);
    for (EncoderPass& p : sc.passes) sc.event_pool.push_back(p.done);
    sc.passes.clear();
    sc.ready.clear();
    sc.need_window.clear();
    sc.parked.clear();
    for (auto& s : sc.subs) {
        give_buf(sc, s->pcm_stage);
        give_buf(sc, s->mel_raw);
        give_buf(sc, s->clip_meta);
        if (s->ev_mel) sc.event_pool.push_back(s->ev_mel);
    }
    sc.subs.clear();
    sc.rows.assign(st->max_batch, RowInfo());
    sc.n_armed = 0;
    sc.beam_groups.clear();
    sc.beam_kv_copies.clear();
    sc.free_slots.clear();
    for (int i = st->n_slots - 1; i >= 0; --i) sc.free_slots.push_back(i);
    SeqLayout lay(st->max_batch);
    memset(st->pinned, 0, lay.total_words * 4);
    cudaMemcpyAsync(st->seq_dev.p, st->pinned, lay.total_words * 4, cudaMemcpyHostToDevice, st->stream);
    cudaStreamSynchronize(st->stream);
    sc.mirror_dirty = false;
    sc.rules_valid = false;
    cudaGetLastError();
}

void sched_release(oh_state* st) {
    Sched& sc = st->sched;
    for (EncoderPass& p : sc.passes) sc.event_pool.push_back(p.done);
    sc.passes.clear();
    for (auto& s : sc.subs) {
        give_buf(sc, s->pcm_stage);
        give_buf(sc, s->mel_raw);
        give_buf(sc, s->clip_meta);
        if (s->ev_mel) sc.event_pool.push_back(s->ev_mel);
    }
    sc.subs.clear();
    for (cudaEvent_t e : sc.event_pool) cudaEventDestroy(e);
    sc.event_pool.clear();
    for (DevBuf& b : sc.buf_pool) release(b);
    sc.buf_pool.clear();
}

}  // namespace oh
