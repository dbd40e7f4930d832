// Tag post-processing + chunk-F1 counters on the device (SURVEY 8f "next" row 2).
//
// Replaces the host loops that consume the Viterbi tags in the reference driver:
//   My_cross_attention.py:879-903 (dev) / 1052-1077 (test)  walk every sentence while its mask is on, keep the
//       positions whose GOLD label is not one of X, <s>, </s>, [CLS], [SEP]
//   ner_evaluate.py:4-48 (get_chunks), 64-110 (evaluate)     token accuracy + |gold chunks|, |pred chunks|,
//       |gold & pred| over the sets of (type, start, end) triples of the FILTERED sequences
//
// All integer work: the five counters are exact, so precision / recall / F1 formed from them on the host are
// the reference's own floating-point expressions on identical integers.
//
// Mapping: one warp per sentence.  Filtering is a stream compaction by warp ballots into shared memory; a chunk
// is described by two bit masks over the filtered sequence -- `start` (a chunk begins here) and `boundary`
// (a chunk cannot continue past here: a start, an O token, or the end) -- so a gold chunk beginning at i is
// also a predicted chunk iff both sequences have a start at i with the same type and the next boundary after i
// is the same position in both.
#include "common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kMaxWords = 32;             // S <= 1024
constexpr uint16_t kSkip = 1u << 8;       // gold label is filtered out by the driver
constexpr uint16_t kOutside = 1u << 9;    // the 'O' tag (tags['O'] of get_chunks)
constexpr uint16_t kBegin = 1u << 10;     // tag class == 'B'

__device__ __forceinline__ int next_boundary(const uint32_t* bnd, int pos) {
  int w = pos >> 5;
  const int bit = pos & 31;
  uint32_t m = (bit == 31) ? 0u : (bnd[w] & (~0u << (bit + 1)));
  while (m == 0) m = bnd[++w];            // the end sentinel (bit n) always terminates the walk
  return w * 32 + __ffs(m) - 1;
}

__global__ void __launch_bounds__(kWarps * 32)
ner_counts_kernel(const int32_t* __restrict__ pred, const int64_t* __restrict__ gold, const uint8_t* __restrict__ mask,
                  const uint16_t* __restrict__ info, int n_ids, unsigned long long* __restrict__ totals,
                  int32_t* __restrict__ per_sentence, int B, int S, int words) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int block_tot[6];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 6) block_tot[threadIdx.x] = 0;
  __syncthreads();

  // per-warp slices: filtered gold / pred ids (u8), then 4 mask arrays of (words + 1) u32
  const size_t seq_bytes = ((size_t)S + 3) / 4 * 4;
  const size_t warp_bytes = 2 * seq_bytes + 4 * (size_t)(words + 1) * sizeof(uint32_t);
  unsigned char* base = smem_raw + warp * warp_bytes;
  uint8_t* fg = base;
  uint8_t* fp = base + seq_bytes;
  uint32_t* g_start = reinterpret_cast<uint32_t*>(base + 2 * seq_bytes);
  uint32_t* g_bnd = g_start + (words + 1);
  uint32_t* p_start = g_bnd + (words + 1);
  uint32_t* p_bnd = p_start + (words + 1);

  const int b = blockIdx.x * kWarps + warp;
  if (b < B) {
    const int32_t* prow = pred + (size_t)b * S;
    const int64_t* grow = gold + (size_t)b * S;
    const uint8_t* mrow = mask ? mask + (size_t)b * S : nullptr;
    const uint32_t lt = (1u << lane) - 1u;

    // ---- filter (MCA:881-903): the walk stops at the first position whose mask is off
    int n = 0, n_ok = 0, n_bad = 0;
    bool alive = true;
    for (int w = 0; w < words && alive; ++w) {
      const int pos = w * 32 + lane;
      const bool in = pos < S;
      const bool on = in && (mrow ? mrow[pos] != 0 : true);
      uint32_t run = __ballot_sync(0xffffffffu, on);
      const uint32_t off = ~run;
      if (off) {                          // first zero (or the end of the row) inside this word
        run &= (1u << (__ffs(off) - 1)) - 1u;
        alive = false;
      }
      const bool walked = (run >> lane) & 1u;
      long long g = 0;
      int p = 0;
      if (walked) {
        g = grow[pos];
        p = prow[pos];
      }
      const bool g_okay = g >= 0 && g < n_ids;
      const bool p_okay = p >= 0 && p < n_ids;
      const bool keep = walked && g_okay && !(info[g_okay ? g : 0] & kSkip);
      const bool bad = walked && (!g_okay || (keep && !p_okay));
      const uint32_t kb = __ballot_sync(0xffffffffu, keep && p_okay);
      n_bad += __popc(__ballot_sync(0xffffffffu, bad));
      if (keep && p_okay) {
        const int idx = n + __popc(kb & lt);
        fg[idx] = (uint8_t)g;
        fp[idx] = (uint8_t)p;
      }
      n_ok += __popc(__ballot_sync(0xffffffffu, keep && p_okay && (long long)p == g));
      n += __popc(kb);
    }
    __syncwarp();

    // ---- chunk masks (ner_evaluate.py:27-48)
    const int fw = (n >> 5) + 1;          // words that hold positions 0..n (n = end sentinel)
    int n_gold = 0, n_pred = 0;
    for (int w = 0; w < fw; ++w) {
      const int i = w * 32 + lane;
      bool gs = false, gb = false, ps = false, pb = false;
      if (i < n) {
        const uint16_t gi = info[fg[i]], pi = info[fp[i]];
        const bool g_out = gi & kOutside, p_out = pi & kOutside;
        uint16_t gprev = kOutside, pprev = kOutside;
        if (i > 0) {
          gprev = info[fg[i - 1]];
          pprev = info[fp[i - 1]];
        }
        gs = !g_out && ((gprev & kOutside) || (gprev & 0xff) != (gi & 0xff) || (gi & kBegin));
        ps = !p_out && ((pprev & kOutside) || (pprev & 0xff) != (pi & 0xff) || (pi & kBegin));
        gb = gs || g_out;
        pb = ps || p_out;
      } else if (i == n) {
        gb = pb = true;
      }
      const uint32_t m_gs = __ballot_sync(0xffffffffu, gs), m_gb = __ballot_sync(0xffffffffu, gb);
      const uint32_t m_ps = __ballot_sync(0xffffffffu, ps), m_pb = __ballot_sync(0xffffffffu, pb);
      if (lane == 0) {
        g_start[w] = m_gs;
        g_bnd[w] = m_gb;
        p_start[w] = m_ps;
        p_bnd[w] = m_pb;
      }
      n_gold += __popc(m_gs);
      n_pred += __popc(m_ps);
    }
    __syncwarp();

    // ---- |gold & pred|: same start, same type, same end
    int n_correct = 0;
    for (int w = 0; w < fw; ++w) {
      const int i = w * 32 + lane;
      bool hit = false;
      if (i < n && ((g_start[w] & p_start[w]) >> lane & 1u)) {
        if ((info[fg[i]] & 0xff) == (info[fp[i]] & 0xff)) hit = next_boundary(g_bnd, i) == next_boundary(p_bnd, i);
      }
      n_correct += __popc(__ballot_sync(0xffffffffu, hit));
    }

    if (lane == 0) {
      if (per_sentence) {
        int32_t* o = per_sentence + (size_t)b * 5;
        o[0] = n;
        o[1] = n_ok;
        o[2] = n_correct;
        o[3] = n_pred;
        o[4] = n_gold;
      }
      atomicAdd(&block_tot[0], n);
      atomicAdd(&block_tot[1], n_ok);
      atomicAdd(&block_tot[2], n_correct);
      atomicAdd(&block_tot[3], n_pred);
      atomicAdd(&block_tot[4], n_gold);
      atomicAdd(&block_tot[5], n_bad);
    }
  }
  __syncthreads();
  if (threadIdx.x < 6 && block_tot[threadIdx.x] != 0)
    atomicAdd(&totals[threadIdx.x], (unsigned long long)block_tot[threadIdx.x]);
}

}  // namespace

extern "C" int icka_ner_chunk_counts(icka_handle* h, const int32_t* pred, const int64_t* gold, const uint8_t* mask,
                                     const uint16_t* label_info, int n_ids, unsigned long long* totals,
                                     int32_t* per_sentence, int B, int S, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1, "ner_chunk_counts: bad shape B=%d S=%d", B, S);
  ICKA_REQUIRE(S <= 32 * kMaxWords, "ner_chunk_counts: S=%d > %d not supported", S, 32 * kMaxWords);
  ICKA_REQUIRE(n_ids >= 1 && n_ids <= 256, "ner_chunk_counts: %d label ids (1..256 supported)", n_ids);
  ICKA_REQUIRE(pred && gold && label_info && totals, "ner_chunk_counts: null pointer");
  if (B == 0) return ICKA_OK;
  const int words = (S + 31) / 32;
  const size_t seq_bytes = ((size_t)S + 3) / 4 * 4;
  const size_t smem = kWarps * (2 * seq_bytes + 4 * (size_t)(words + 1) * sizeof(uint32_t));
  ner_counts_kernel<<<(B + kWarps - 1) / kWarps, kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      pred, gold, mask, label_info, n_ids, totals, per_sentence, B, S, words);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
