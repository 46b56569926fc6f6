// Per-token decoder bookkeeping of whisper's window loop (SURVEY.md Appendix A.5 "per decoder, with tok =
// last sampled"), shared by the device state machine (decoder.cu, greedy path: no host round trip per token)
// and the host path used when a temperature fallback samples with std::mt19937 (A.7).
#pragma once

#ifdef __CUDACC__
#define OH_RULE_HD __host__ __device__ __forceinline__
#else
#define OH_RULE_HD inline
#endif

namespace oh {

struct SeqFlags {
    int has_ts, seek_delta, result_len, failed, completed;
};

// Applies the rules for the token sampled at index i (0-based).  Returns true while the sequence stays live.
OH_RULE_HD bool apply_token_rules(SeqFlags& f, int id, int i, int seek, int seek_end, int beg, int eot, int max_tokens, int no_timestamps,
                                  int single_segment, int n_max) {
    if (id > beg) {
        const int seek_delta_new = 2 * (id - beg);
        if (f.has_ts && f.seek_delta > seek_delta_new && f.result_len < i) {
            f.failed = 1;  // went back in time
            return false;
        }
        f.seek_delta = seek_delta_new;
        f.result_len = i + 1;
        f.has_ts = 1;
    }
    if (id == eot || (max_tokens > 0 && i >= max_tokens) || (f.has_ts && seek + f.seek_delta + 100 >= seek_end)) {
        if (f.result_len == 0 && !no_timestamps) {
            if (seek + f.seek_delta + 100 >= seek_end) f.result_len = i + 1;
            else {
                f.failed = 1;
                return false;
            }
        }
        if (single_segment || no_timestamps) {
            f.result_len = i + 1;
            f.seek_delta = 100 * 30;
        }
        f.completed = 1;
        return false;
    }
    if (i == n_max - 1 && (f.result_len == 0 || f.seek_delta < 100 * 30 / 2)) {
        f.failed = 1;  // repetition-loop guard
        return false;
    }
    return i + 1 < n_max;
}

}  // namespace oh
