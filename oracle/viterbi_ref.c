/* oracle/viterbi_ref.c -- TEST INFRASTRUCTURE ONLY (CPU checker + CPU baseline), never shipped.
 *
 * Plain-C restatement of pytorch-crf's Viterbi decode, the routine the reference reaches through
 * `self.crf.decode(emissions, mask=output_mask)` (Cross_Modal_Interaction_Module.py:1051, :1056).
 * PARITY UNPINNED: the algorithm lives in the third-party package `pytorch-crf` (import name
 * torchcrf), absent from /root/reference and from this image; see oracle/crf_ref.py.
 *
 * fp32 operation order is the one that decides ties and therefore the tags:
 *     cand = (score[i] + trans[i][j]) + e[t][j]     -- two roundings, in this order
 *     argmax over i takes the FIRST maximal index (strict > while scanning i upward)
 *     score[j] is replaced only where mask[t] is on; back-pointers are recorded for every t
 *     len = sum(mask);  last = first argmax_j (score[j] + end[j]);  walk bp[len-2 .. 0]
 * Build with -ffp-contract=off so no FMA sneaks in (there are only adds, but be explicit).
 *
 * C ABI (all row-major, batch-first):
 *   emissions [B,S,T] f32, mask [B,S] u8 (may be NULL = all on), start[T], end[T], trans[T,T]
 *   tags_out [B,S] i32 (positions >= len are set to -1), lens_out [B] i32
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ICKA_ORACLE_MAX_T 64

int icka_oracle_viterbi(const float* emissions, const uint8_t* mask,
                        const float* start, const float* end, const float* trans,
                        int32_t* tags_out, int32_t* lens_out, int B, int S, int T)
{
    if (T <= 0 || T > ICKA_ORACLE_MAX_T || S <= 0 || B < 0) return -1;
    uint8_t* bp = (uint8_t*)malloc((size_t)(S > 1 ? S - 1 : 1) * (size_t)T);
    if (!bp) return -2;
    for (int b = 0; b < B; ++b) {
        const float* e = emissions + (size_t)b * S * T;
        const uint8_t* m = mask ? mask + (size_t)b * S : NULL;
        float score[ICKA_ORACLE_MAX_T], nxt[ICKA_ORACLE_MAX_T];
        for (int j = 0; j < T; ++j) score[j] = start[j] + e[j];
        int len = m ? (m[0] != 0) : 1;
        for (int t = 1; t < S; ++t) {
            const float* et = e + (size_t)t * T;
            uint8_t* bpt = bp + (size_t)(t - 1) * T;
            for (int j = 0; j < T; ++j) {
                float best = (score[0] + trans[j]) + et[j];
                int arg = 0;
                for (int i = 1; i < T; ++i) {
                    float c = (score[i] + trans[(size_t)i * T + j]) + et[j];
                    if (c > best) { best = c; arg = i; }
                }
                nxt[j] = best;
                bpt[j] = (uint8_t)arg;
            }
            int on = m ? (m[t] != 0) : 1;
            if (on) { memcpy(score, nxt, sizeof(float) * (size_t)T); }
            len += on;
        }
        float best = score[0] + end[0];
        int last = 0;
        for (int j = 1; j < T; ++j) {
            float c = score[j] + end[j];
            if (c > best) { best = c; last = j; }
        }
        int32_t* out = tags_out + (size_t)b * S;
        for (int t = 0; t < S; ++t) out[t] = -1;
        lens_out[b] = len;
        if (len >= 1) {
            out[len - 1] = last;
            for (int k = len - 2; k >= 0; --k) {
                last = bp[(size_t)k * T + last];
                out[k] = last;
            }
        }
    }
    free(bp);
    return 0;
}
