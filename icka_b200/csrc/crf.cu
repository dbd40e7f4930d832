// CRF kernels: batched Viterbi decode and per-sentence log-likelihood.
//
// Replaces torchcrf.CRF.decode / CRF.forward as called from Cross_Modal_Interaction_Module.py:1047-1057.
// The reference runs a Python loop of S-1 steps x ~5 ATen launches plus B x len .item() syncs; here one
// launch handles the whole batch and the only output is [B,S] int32 tags + [B] lengths.
//
// Mapping: a group of LPS lanes (16 when T <= 16, else 32) owns one sentence; lane j holds score[j]
// and column j of the transition matrix in registers.  The sentence's emission slab (contiguous
// S*T fp32 in HBM) is streamed into shared memory with 16-byte cp.async in four commit groups so the
// serial chain starts while the rest is in flight; back-pointers live in shared memory only.
// Bit-exactness: cand = (score[i] + trans[i][j]) + e[t][j], strict '>' scanning i upward (first index
// wins ties) -- the exact fp32 op order of pytorch-crf's _viterbi_decode.
// Prefix masks (the only kind ICKA builds, My_cross_attention.py:369-383) stop after len steps: later
// steps cannot change `score` and their back-pointers are never read.  Masks with holes run all S steps.
#include "common.cuh"

#include <stdlib.h>

namespace {

constexpr int kThreads = 128;

// Gold tag ids index start[], trans[][], e[t][] and shared-memory accumulators: ids outside [0, T) (e.g. an ignore-index
// of -100 on padding) are clamped so no access leaves its array.  pytorch-crf raises IndexError for them; the Python
// mirror (icka_b200/crf.py) checks the range before the launch, so a clamped id never reaches a result.
__device__ __forceinline__ int gold_tag(int64_t y, int T) { return (int)(y < 0 ? 0 : (y >= T ? T - 1 : y)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct SeqSmem {
  float* score;    // [2][16] score exchange buffers (viterbi16_kernel)
  float* em;       // [S*T] (rounded up to 4)
  uint8_t* bp;     // [(S-1)*LPS]
  uint8_t* mask;   // [S]
  uint8_t* tagb;   // [S]
};

__host__ __device__ inline size_t seq_smem_bytes(int S, int T, int LPS) {
  size_t em = ((size_t)S * T + 3) / 4 * 16;
  size_t bp = ((size_t)(S > 1 ? S - 1 : 1) * LPS + 15) / 16 * 16;
  size_t mk = ((size_t)S + 15) / 16 * 16;
  return 128 + em + bp + 2 * mk;
}

template <int LPS>
__device__ __forceinline__ SeqSmem carve(uint8_t* base, int S, int T) {
  SeqSmem s;
  size_t em = ((size_t)S * T + 3) / 4 * 16;
  size_t bp = ((size_t)(S > 1 ? S - 1 : 1) * LPS + 15) / 16 * 16;
  size_t mk = ((size_t)S + 15) / 16 * 16;
  s.score = reinterpret_cast<float*>(base);
  s.em = reinterpret_cast<float*>(base + 128);
  s.bp = base + 128 + em;
  s.mask = s.bp + bp;
  s.tagb = s.mask + mk;
  return s;
}

// Loads the mask row into smem, returns len = sum(mask) and whether the mask is a prefix of ones.
template <int LPS>
__device__ __forceinline__ void load_mask(const uint8_t* __restrict__ mrow, uint8_t* smask, int S, int gl,
                                          unsigned gmask, int gshift, int& len, bool& prefix) {
  len = 0;
  prefix = true;
  bool seen_zero = false;
  for (int t0 = 0; t0 < S; t0 += LPS) {
    int t = t0 + gl;
    uint8_t m = 0;
    if (t < S) {
      m = mrow ? (mrow[t] != 0) : 1;
      smask[t] = m;
    }
    unsigned bits = (__ballot_sync(gmask, m != 0) >> gshift) & (LPS == 32 ? 0xffffffffu : 0xffffu);
    int c = __popc(bits);
    if (bits != 0 && seen_zero) prefix = false;
    if (bits != (c == 32 ? 0xffffffffu : ((1u << c) - 1u))) prefix = false;
    int valid = min(LPS, S - t0);
    if (c < valid) seen_zero = true;
    len += c;
  }
}

// Streams n_floats of the emission slab to smem in 4 commit groups.  Returns per-group float boundaries.
template <int LPS>
__device__ __forceinline__ void load_emissions(const float* __restrict__ g, float* s, int n_floats, int gl,
                                               bool vec_ok, int bound[4]) {
  if (vec_ok) {
    int nvec = (n_floats + 3) / 4;   // slab is padded in smem; reading <=3 floats past n_floats stays inside
                                      // the sentence's own S*T slab or the next sentence's (still mapped)
    for (int c = 0; c < 4; ++c) {
      int v0 = (int)((long long)nvec * c / 4), v1 = (int)((long long)nvec * (c + 1) / 4);
      for (int v = v0 + gl; v < v1; v += LPS) cp_async16(s + 4 * v, g + 4 * v);
      cp_async_commit();
      bound[c] = min(n_floats, 4 * v1);
    }
  } else {
    for (int i = gl; i < n_floats; i += LPS) s[i] = g[i];
    for (int c = 0; c < 4; ++c) {
      cp_async_commit();   // empty groups keep the wait logic uniform
      bound[c] = n_floats;
    }
  }
}

__device__ __forceinline__ void wait_chunk(int c, unsigned gmask) {
  if (c == 0) cp_async_wait<3>();
  else if (c == 1) cp_async_wait<2>();
  else if (c == 2) cp_async_wait<1>();
  else cp_async_wait<0>();
  __syncwarp(gmask);
}

template <int LPS>
__global__ void __launch_bounds__(kThreads) viterbi_kernel(
    const float* __restrict__ emissions, const uint8_t* __restrict__ mask, const float* __restrict__ start,
    const float* __restrict__ end, const float* __restrict__ trans, int32_t* __restrict__ tags_out,
    int32_t* __restrict__ lens_out, int B, int S, int T, size_t per_seq_bytes) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int kSeqPerBlock = kThreads / LPS;
  const int g_in_block = threadIdx.x / LPS;
  const int gl = threadIdx.x % LPS;
  const int gshift = (threadIdx.x % 32) / LPS * LPS;
  const unsigned gmask = (LPS == 32) ? 0xffffffffu : (0xffffu << gshift);
  const int b = blockIdx.x * kSeqPerBlock + g_in_block;
  if (b >= B) return;   // whole group leaves together; the other group of the warp syncs on its own gmask

  SeqSmem sm = carve<LPS>(smem_raw + (size_t)g_in_block * per_seq_bytes, S, T);
  const float* e_g = emissions + (size_t)b * S * T;

  int len;
  bool prefix;
  load_mask<LPS>(mask ? mask + (size_t)b * S : nullptr, sm.mask, S, gl, gmask, gshift, len, prefix);
  const int n_steps = (prefix && len >= 1) ? len : S;   // steps t = 0 .. n_steps-1 are executed

  const bool vec_ok = ((((size_t)S * T) % 4) == 0) && ((reinterpret_cast<uintptr_t>(emissions) % 16) == 0);
  int bound[4];
  load_emissions<LPS>(e_g, sm.em, n_steps * T, gl, vec_ok, bound);

  const bool active = gl < T;
  float tr[LPS];
#pragma unroll
  for (int i = 0; i < LPS; ++i) tr[i] = (active && i < T) ? trans[i * T + gl] : 0.0f;
  const float st = active ? start[gl] : 0.0f;
  const float en = active ? end[gl] : 0.0f;

  int chunk = 0;
  wait_chunk(0, gmask);
  float score = active ? st + sm.em[gl] : -INFINITY;

  for (int t = 1; t < n_steps; ++t) {
    while ((t + 1) * T > bound[chunk]) { ++chunk; wait_chunk(chunk, gmask); }
    const float e = active ? sm.em[t * T + gl] : 0.0f;
    float best = (__shfl_sync(gmask, score, 0, LPS) + tr[0]) + e;
    int arg = 0;
#pragma unroll
    for (int i = 1; i < LPS; ++i) {
      if (i < T) {
        const float c = (__shfl_sync(gmask, score, i, LPS) + tr[i]) + e;
        if (c > best) { best = c; arg = i; }
      }
    }
    sm.bp[(t - 1) * LPS + gl] = (uint8_t)arg;
    if (active && sm.mask[t]) score = best;
  }
  while (chunk < 3) { ++chunk; wait_chunk(chunk, gmask); }   // drain outstanding copies before exit

  // last = first argmax_j (score[j] + end[j])
  float fin = active ? score + en : -INFINITY;
  int idx = gl;
#pragma unroll
  for (int off = LPS / 2; off >= 1; off >>= 1) {
    const float ov = __shfl_xor_sync(gmask, fin, off, LPS);
    const int oi = __shfl_xor_sync(gmask, idx, off, LPS);
    if (ov > fin || (ov == fin && oi < idx)) { fin = ov; idx = oi; }
  }
  __syncwarp(gmask);   // back-pointer writes visible to lane 0
  if (gl == 0 && len >= 1) {
    int cur = idx;
    sm.tagb[len - 1] = (uint8_t)cur;
    for (int k = len - 2; k >= 0; --k) {
      cur = sm.bp[k * LPS + cur];
      sm.tagb[k] = (uint8_t)cur;
    }
  }
  __syncwarp(gmask);
  int32_t* out = tags_out + (size_t)b * S;
  for (int t = gl; t < S; t += LPS) out[t] = (t < len) ? (int32_t)sm.tagb[t] : -1;
  if (gl == 0) lens_out[b] = len;
}

// ---- T <= 16: two sentences per warp, both halves in lock-step ------------------------------------------
// The generic kernel above syncs each half-warp on its own member mask, which makes the two halves of a
// warp run the serial chain one after the other, and scans the 15 predecessors with a 15-deep compare
// chain.  Here every shuffle uses the full mask (the warp runs max(len_a, len_b) steps; a half that is
// already past its own length just keeps its score), lane j publishes score[j] in shared memory and reads all
// 16 back as four broadcast 16-byte loads (the shuffle unit, 16 SHFL per step, was the limiter: ~4 cycles
// per SHFL per SM), the 16 candidates of a step are formed independently
// and reduced by a 4-level (value, index) tournament whose left operand always holds the lower indices, so
// "first index wins ties" is preserved exactly, and the next step's emission is fetched from shared memory
// one step ahead of the chain.
struct Cand {
  float v;
  int i;
};
__device__ __forceinline__ Cand take_first_max(const Cand a, const Cand b) {   // a holds the lower indices
  Cand r;
  r.v = fmaxf(a.v, b.v);          // value chain: one FMNMX per level (inputs are NaN-free; a signed-zero
  r.i = (b.v > a.v) ? b.i : a.i;  // difference cannot change a later comparison); the index trails it
  return r;
}

constexpr int kV16Threads = 128;   // maximum; long sequences launch fewer warps per block so the per-sentence slabs fit

__global__ void __launch_bounds__(kV16Threads) viterbi16_kernel(
    const float* __restrict__ emissions, const uint8_t* __restrict__ mask, const float* __restrict__ start,
    const float* __restrict__ end, const float* __restrict__ trans, int32_t* __restrict__ tags_out,
    int32_t* __restrict__ lens_out, int B, int S, int T, size_t per_seq_bytes) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr unsigned kFull = 0xffffffffu;
  const int g_in_block = threadIdx.x >> 4;
  const int gl = threadIdx.x & 15;
  const int gshift = threadIdx.x & 16;
  const int seqs_per_block = blockDim.x >> 4;
  const int b_raw = blockIdx.x * seqs_per_block + g_in_block;
  if (blockIdx.x * seqs_per_block + (g_in_block & ~1) >= B) return;   // whole warp out of range
  const bool seq_ok = b_raw < B;
  const int b = seq_ok ? b_raw : B - 1;   // an odd tail half shadows the last sentence and stores nothing

  SeqSmem sm = carve<16>(smem_raw + (size_t)g_in_block * per_seq_bytes, S, T);
  const float* e_g = emissions + (size_t)b * S * T;
  const uint8_t* mrow = mask ? mask + (size_t)b * S : nullptr;

  // The sentence is a chain of dependent global round trips (mask -> length -> emissions) in front of a 127-step chain:
  // at 1024 sentences there are ~4 warps per SM and nothing hides them.  So: the first 16 steps of the emission slab are
  // requested before anything else (they do not depend on the length), the mask row is fetched with all its loads in
  // flight at once, and the transition column / start / end loads are issued before the mask is reduced.
  const int head_steps = min(S, 16);
  const int head_floats = head_steps * T;
  const bool vec_ok = ((((size_t)S * T) % 4) == 0) && ((reinterpret_cast<uintptr_t>(emissions) % 16) == 0);
  const int hvec = (head_floats + 3) / 4;                  // <= 3 floats of slack stay inside S*T
  if (vec_ok) {
    for (int v = gl; v < hvec; v += 16) cp_async16(sm.em + 4 * v, e_g + 4 * v);
  } else {
    for (int i = gl; i < head_floats; i += 16) sm.em[i] = e_g[i];
  }
  cp_async_commit();

  const bool active = gl < T;
  float tr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) tr[i] = (i < T) ? (active ? __ldg(trans + i * T + gl) : 0.0f) : -INFINITY;
  const float st = active ? __ldg(start + gl) : 0.0f;
  const float en = active ? __ldg(end + gl) : 0.0f;

  // mask row -> smem, len = sum(mask), prefix = (mask is len ones followed by zeros); eight 16-position pieces per batch
  int len = 0;
  bool prefix = true, seen_zero = false;
  for (int tb = 0; tb < S; tb += 128) {
    uint8_t mv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = tb + 16 * i + gl;
      mv[i] = (t < S) ? (mrow ? (uint8_t)(__ldg(mrow + t) != 0) : (uint8_t)1) : (uint8_t)0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t0 = tb + 16 * i;
      if (t0 < S) {                                       // uniform across the warp
        const int t = t0 + gl;
        if (t < S) sm.mask[t] = mv[i];
        const unsigned bits = (__ballot_sync(kFull, mv[i] != 0) >> gshift) & 0xffffu;
        const int c = __popc(bits);
        if (bits != 0 && seen_zero) prefix = false;
        if (bits != ((1u << c) - 1u)) prefix = false;
        if (c < min(16, S - t0)) seen_zero = true;
        len += c;
      }
    }
  }
  const bool fast = prefix && len >= 1;
  const int n_steps = fast ? len : S;                                      // this sentence runs t = 0 .. n_steps-1
  const int n_warp = max(n_steps, __shfl_xor_sync(kFull, n_steps, 16));    // the warp's trip count
  const bool all_fast = __all_sync(kFull, fast);

  // the rest of the emission slab (steps 16 .. n_steps-1) as a second cp.async group
  const int n_floats = n_steps * T;
  if (vec_ok) {
    const int nvec = (n_floats + 3) / 4;
    for (int v = hvec + gl; v < nvec; v += 16) cp_async16(sm.em + 4 * v, e_g + 4 * v);
  } else {
    for (int i = head_floats + gl; i < n_floats; i += 16) sm.em[i] = e_g[i];
  }
  cp_async_commit();


  cp_async_wait<1>();
  __syncwarp();
  float score = active ? st + sm.em[gl] : 0.0f;   // idle lanes carry 0; their tr[] row is never selected
  sm.score[gl] = score;
  __syncwarp();
  const float* e_ptr = sm.em + gl;
  uint8_t* bp_ptr = sm.bp + gl;
  // reads past a sentence's own n_steps hit unwritten (but allocated) smem; the result is discarded
  float e_next = (n_warp > 1) ? e_ptr[T] : 0.0f;
  for (int t = 1; t < n_warp; ++t) {
    if (t == 15) {   // the prefetch below leaves the head group
      cp_async_wait<0>();
      __syncwarp();
    }
    const float e = e_next;
    e_next = e_ptr[min(t + 1, S - 1) * T];
    // all 16 scores of the previous step: four broadcast 16-byte shared loads (one address per half-warp)
    const float4* sp = reinterpret_cast<const float4*>(sm.score + ((t - 1) & 1) * 16);
    const float4 s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3];
    const float sv[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w,
                          s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
    Cand c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      c[i].v = (sv[i] + tr[i]) + e;
      c[i].i = i;
    }
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
      for (int k = 0; k < w; ++k) c[k] = take_first_max(c[2 * k], c[2 * k + 1]);
    // note: level order above pairs (0,1)(2,3).. then (01,23).. so the left operand is always the lower range
    bp_ptr[(t - 1) * 16] = (uint8_t)c[0].i;
    bool upd = active && t < n_steps;
    if (!all_fast) upd = upd && sm.mask[t] != 0;
    score = upd ? c[0].v : score;
    sm.score[(t & 1) * 16 + gl] = score;
    __syncwarp();   // step t+1 reads this buffer and overwrites the other one, which every lane has finished reading
  }
  cp_async_wait<0>();   // nothing may be in flight at exit

  // last = first argmax_j (score[j] + end[j])
  float fin = active ? score + en : -INFINITY;
  int idx = gl;
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const float ov = __shfl_xor_sync(kFull, fin, off, 16);
    const int oi = __shfl_xor_sync(kFull, idx, off, 16);
    if (ov > fin || (ov == fin && oi < idx)) { fin = ov; idx = oi; }
  }
  __syncwarp();   // back-pointer writes visible to the tracing lane
  if (gl == 0 && len >= 1) {
    int cur = idx;
    sm.tagb[len - 1] = (uint8_t)cur;
    for (int k = len - 2; k >= 0; --k) {
      cur = sm.bp[k * 16 + cur];
      sm.tagb[k] = (uint8_t)cur;
    }
  }
  __syncwarp();
  if (!seq_ok) return;
  int32_t* out = tags_out + (size_t)b * S;
  if ((S % 4) == 0 && (reinterpret_cast<uintptr_t>(tags_out) % 16) == 0) {
    for (int t = 4 * gl; t < S; t += 64) {
      int4 v;
      v.x = (t + 0 < len) ? (int)sm.tagb[t + 0] : -1;
      v.y = (t + 1 < len) ? (int)sm.tagb[t + 1] : -1;
      v.z = (t + 2 < len) ? (int)sm.tagb[t + 2] : -1;
      v.w = (t + 3 < len) ? (int)sm.tagb[t + 3] : -1;
      *reinterpret_cast<int4*>(out + t) = v;
    }
  } else {
    for (int t = gl; t < S; t += 16) out[t] = (t < len) ? (int32_t)sm.tagb[t] : -1;
  }
  if (gl == 0) lens_out[b] = len;
}

// Coalesced copy of one sentence's [S*T] fp32 slab into shared memory by the LPS lanes of its group.
template <int LPS>
__device__ __forceinline__ void stage_slab(const float* __restrict__ g, float* s, int n, int gl) {
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int nv = n / 4;
    for (int v = gl; v < nv; v += LPS) reinterpret_cast<float4*>(s)[v] = reinterpret_cast<const float4*>(g)[v];
    for (int i = nv * 4 + gl; i < n; i += LPS) s[i] = g[i];
  } else {
    for (int i = gl; i < n; i += LPS) s[i] = g[i];
  }
}

// Per-sentence log-likelihood: gold-path score minus log-partition (forward algorithm), fp32.
// Same lane mapping as Viterbi; logsumexp over predecessors i is max-shifted like torch.logsumexp.
template <int LPS>
__global__ void __launch_bounds__(kThreads) crf_llh_kernel(
    const float* __restrict__ emissions, const int64_t* __restrict__ tags, const uint8_t* __restrict__ mask,
    const float* __restrict__ start, const float* __restrict__ end, const float* __restrict__ trans,
    float* __restrict__ llh_out, int B, int S, int T) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int g_in_block = threadIdx.x / LPS;
  const int gl = threadIdx.x % LPS;
  const int gshift = (threadIdx.x % 32) / LPS * LPS;
  const unsigned gmask = (LPS == 32) ? 0xffffffffu : (0xffffu << gshift);
  const int b = blockIdx.x * (kThreads / LPS) + g_in_block;
  if (b >= B) return;
  const uint8_t* m_g = mask ? mask + (size_t)b * S : nullptr;
  const int64_t* y_g = tags + (size_t)b * S;
  const bool active = gl < T;
  // the serial chain must not wait on HBM every step: stage the sentence's emission slab in shared memory
  // (nor on a per-step global read of the mask byte that decides the branch)
  const size_t e_floats = ((size_t)S * T + 3) / 4 * 4;
  const size_t seq_bytes = e_floats * sizeof(float) + ((size_t)S + 15) / 16 * 16;
  float* e_g = reinterpret_cast<float*>(smem_raw + (size_t)g_in_block * seq_bytes);
  uint8_t* m_s = reinterpret_cast<uint8_t*>(e_g + e_floats);
  stage_slab<LPS>(emissions + (size_t)b * S * T, e_g, S * T, gl);
  for (int t = gl; t < S; t += LPS) m_s[t] = m_g ? (m_g[t] != 0) : 1;
  __syncwarp(gmask);

  float tr[LPS];
#pragma unroll
  for (int i = 0; i < LPS; ++i) tr[i] = (active && i < T) ? trans[i * T + gl] : 0.0f;

  // ---- normalizer (pytorch-crf _compute_normalizer) ----
  float alpha = active ? start[gl] + e_g[gl] : -INFINITY;
  for (int t = 1; t < S; ++t) {
    if (!m_s[t]) continue;   // group-uniform
    const float e = active ? e_g[t * T + gl] : 0.0f;
    float c[LPS];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < LPS; ++i) {
      if (i < T) {
        c[i] = (__shfl_sync(gmask, alpha, i, LPS) + tr[i]) + e;
        mx = fmaxf(mx, c[i]);
      }
    }
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < LPS; ++i)
      if (i < T) sum += expf(c[i] - mx);
    if (active) alpha = mx + logf(sum);
  }
  float fin = active ? alpha + end[gl] : -INFINITY;
  float mx = fin;
#pragma unroll
  for (int off = LPS / 2; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, off, LPS));
  float ex = active ? expf(fin - mx) : 0.0f;
#pragma unroll
  for (int off = LPS / 2; off >= 1; off >>= 1) ex += __shfl_xor_sync(gmask, ex, off, LPS);
  const float logz = mx + logf(ex);

  // ---- gold path score (pytorch-crf _compute_score); lanes take strided time steps, then reduce ----
  float s = 0.0f;
  int len = 0;
  for (int t = gl; t < S; t += LPS) {
    const int y = gold_tag(y_g[t], T);
    const bool on = m_g ? (m_g[t] != 0) : true;
    if (t == 0) {
      s += start[y] + e_g[y];
      len += on;
    } else if (on) {
      s += trans[gold_tag(y_g[t - 1], T) * T + y] + e_g[t * T + y];
      len += 1;
    }
  }
#pragma unroll
  for (int off = LPS / 2; off >= 1; off >>= 1) {
    s += __shfl_xor_sync(gmask, s, off, LPS);
    len += __shfl_xor_sync(gmask, len, off, LPS);
  }
  if (gl == 0) {
    const int last = gold_tag(y_g[max(len - 1, 0)], T);
    llh_out[b] = (s + end[last]) - logz;
  }
}

// Gradient of sum_b w[b] * llh[b]: forward-backward marginals minus the gold path (the autograd of
// pytorch-crf's forward()).  Same lane mapping as crf_llh_kernel (lane j <-> tag j); the chain runs over the
// "on" steps only (step 0 always counts, exactly as in _compute_normalizer where masked steps leave alpha
// untouched); the alphas of all on-steps are kept in shared memory for the backward sweep.  Parameter
// gradients are gathered per block in shared memory and added to the global vectors once per block.
template <int LPS>
__global__ void __launch_bounds__(kThreads) crf_llh_bwd_kernel(
    const float* __restrict__ emissions, const int64_t* __restrict__ tags, const uint8_t* __restrict__ mask,
    const float* __restrict__ start, const float* __restrict__ end, const float* __restrict__ trans,
    const float* __restrict__ w, float* __restrict__ d_emissions, float* __restrict__ d_start,
    float* __restrict__ d_end, float* __restrict__ d_trans, int B, int S, int T, size_t per_seq_bytes) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int kSeqs = kThreads / LPS;
  float* s_dtrans = reinterpret_cast<float*>(smem_raw);      // [T*T]
  float* s_dstart = s_dtrans + T * T;                        // [T]
  float* s_dend = s_dstart + T;                              // [T]
  const size_t head = (((size_t)T * T + 2 * T) * sizeof(float) + 15) / 16 * 16;
  const int g_in_block = threadIdx.x / LPS;
  const int gl = threadIdx.x % LPS;
  const int gshift = (threadIdx.x % 32) / LPS * LPS;
  const unsigned gmask = (LPS == 32) ? 0xffffffffu : (0xffffu << gshift);
  const int b = blockIdx.x * kSeqs + g_in_block;
  for (int i = threadIdx.x; i < T * T + 2 * T; i += kThreads) s_dtrans[i] = 0.0f;
  __syncthreads();

  if (b < B) {
    float* alpha = reinterpret_cast<float*>(smem_raw + head + (size_t)g_in_block * per_seq_bytes);   // [S][LPS]
    float* e_g = alpha + (size_t)S * LPS;                                                            // [S*T] (16-B aligned)
    int* on_idx = reinterpret_cast<int*>(e_g + ((size_t)S * T + 3) / 4 * 4);                         // [S]
    int* y_s = on_idx + S;                                                                           // [S]
    stage_slab<LPS>(emissions + (size_t)b * S * T, e_g, S * T, gl);
    const uint8_t* m_g = mask ? mask + (size_t)b * S : nullptr;
    const int64_t* y_g = tags + (size_t)b * S;
    float* de_g = d_emissions + (size_t)b * S * T;
    const bool active = gl < T;
    const float wb = w[b];

    // on-steps (t = 0 always), len = sum(mask) as pytorch-crf counts it
    int n_on = 0, len = 0;
    for (int t0 = 0; t0 < S; t0 += LPS) {
      const int t = t0 + gl;
      const bool m = (t < S) && (m_g ? (m_g[t] != 0) : true);
      const bool on = (t < S) && (m || t == 0);
      const unsigned bits_on = (__ballot_sync(gmask, on) >> gshift) & (LPS == 32 ? 0xffffffffu : 0xffffu);
      const unsigned bits_m = (__ballot_sync(gmask, m) >> gshift) & (LPS == 32 ? 0xffffffffu : 0xffffu);
      if (on) on_idx[n_on + __popc(bits_on & ((1u << gl) - 1u))] = t;
      n_on += __popc(bits_on);
      len += __popc(bits_m);
    }
    for (int i = gl; i < S * T; i += LPS) de_g[i] = 0.0f;
    for (int t = gl; t < S; t += LPS) y_s[t] = gold_tag(y_g[t], T);   // the sweeps below must not wait on HBM per step
    __syncwarp(gmask);

    float tr[LPS], trr[LPS];   // column gl and row gl of the transition matrix
#pragma unroll
    for (int i = 0; i < LPS; ++i) {
      tr[i] = (active && i < T) ? trans[i * T + gl] : 0.0f;
      trr[i] = (active && i < T) ? trans[gl * T + i] : 0.0f;
    }

    // ---- forward sweep: alpha[n][j] ----
    float a = active ? start[gl] + e_g[gl] : -INFINITY;
    alpha[gl] = a;
    for (int n = 1; n < n_on; ++n) {
      const int t = on_idx[n];
      const float e = active ? e_g[t * T + gl] : 0.0f;
      float c[LPS];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < LPS; ++i) {
        if (i < T) {
          c[i] = (__shfl_sync(gmask, a, i, LPS) + tr[i]) + e;
          mx = fmaxf(mx, c[i]);
        }
      }
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < LPS; ++i)
        if (i < T) sum += expf(c[i] - mx);
      if (active) a = mx + logf(sum);
      alpha[n * LPS + gl] = a;
    }
    float fin = active ? a + end[gl] : -INFINITY;
    float mx = fin;
#pragma unroll
    for (int off = LPS / 2; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, off, LPS));
    float ex = active ? expf(fin - mx) : 0.0f;
#pragma unroll
    for (int off = LPS / 2; off >= 1; off >>= 1) ex += __shfl_xor_sync(gmask, ex, off, LPS);
    const float logz = mx + logf(ex);
    __syncwarp(gmask);

    // ---- backward sweep ----
    float acc_tr[LPS];   // sum over steps of P(y_prev = i, y_cur = gl)
#pragma unroll
    for (int i = 0; i < LPS; ++i) acc_tr[i] = 0.0f;
    float beta = active ? end[gl] : -INFINITY;
    const int y_last = y_s[max(len - 1, 0)];
    for (int n = n_on - 1; n >= 0; --n) {
      const int t = on_idx[n];
      const float an = alpha[n * LPS + gl];
      const float p = active ? expf(an + beta - logz) : 0.0f;
      const int y = y_s[t];
      if (active) {
        de_g[t * T + gl] = wb * ((y == gl ? 1.0f : 0.0f) - p);
        if (n == n_on - 1) atomicAdd(s_dend + gl, wb * ((y_last == gl ? 1.0f : 0.0f) - p));
        if (n == 0) atomicAdd(s_dstart + gl, wb * ((y == gl ? 1.0f : 0.0f) - p));
      }
      if (n == 0) break;
      const float e = active ? e_g[t * T + gl] : 0.0f;
      const float eb = active ? e + beta : -INFINITY;     // e[t][j] + beta_n[j], lane j
      // pairwise marginals of (on-step n-1 -> n), column gl
#pragma unroll
      for (int i = 0; i < LPS; ++i) {
        if (i < T) {
          const float ai = alpha[(n - 1) * LPS + i];
          if (active) acc_tr[i] += expf(((ai + tr[i]) + eb) - logz);
        }
      }
      // beta_{n-1}[i] = logsumexp_j (trans[i][j] + e[t][j] + beta_n[j]), lane i
      float c[LPS];
      float bm = -INFINITY;
#pragma unroll
      for (int j = 0; j < LPS; ++j) {
        if (j < T) {
          c[j] = trr[j] + __shfl_sync(gmask, eb, j, LPS);
          bm = fmaxf(bm, c[j]);
        }
      }
      float bs = 0.0f;
#pragma unroll
      for (int j = 0; j < LPS; ++j)
        if (j < T) bs += expf(c[j] - bm);
      beta = active ? bm + logf(bs) : -INFINITY;
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < LPS; ++i)
        if (i < T) atomicAdd(s_dtrans + i * T + gl, -wb * acc_tr[i]);
    }
    // gold transitions: pytorch-crf's _compute_score pairs every on-step t >= 1 with position t-1
    for (int t = 1 + gl; t < S; t += LPS) {
      const bool on = m_g ? (m_g[t] != 0) : true;
      if (on) atomicAdd(s_dtrans + y_s[t - 1] * T + y_s[t], wb);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += kThreads) atomicAdd(d_trans + i, s_dtrans[i]);
  for (int i = threadIdx.x; i < T; i += kThreads) {
    atomicAdd(d_start + i, s_dstart[i]);
    atomicAdd(d_end + i, s_dend[i]);
  }
}

// ---- T <= 16: lock-step log-likelihood kernels ------------------------------------------------------------
// Like viterbi16_kernel: two sentences per warp, every warp-level primitive with the full mask (a per-half
// member mask makes the two halves run one after the other and costs ~10 extra instructions per shuffle),
// the 16 values a step needs from its predecessor read back from shared memory as four broadcast 16-byte
// loads, and exp / log through ex2.approx / lg2.approx (arguments are max-shifted, so in [-inf, 0]; the
// absolute error of ~1e-7 per term is far inside the 1e-5 relative gate on the log-likelihood).
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float fast_log(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.6931471805599453f;
}
__device__ __forceinline__ void load16(const float* p, float (&v)[16]) {
  const float4* q = reinterpret_cast<const float4*>(p);
  const float4 a = q[0], b = q[1], c = q[2], d = q[3];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
}
// logsumexp of 16 candidates (entries that must not count hold -inf; at least one is finite)
__device__ __forceinline__ float lse16(const float (&c)[16]) {
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(c[2 * i], c[2 * i + 1]);
  const float mx = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
  float e[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) e[i] = fast_exp(c[i] - mx);
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) e[i] += e[i + w];
  return mx + fast_log(e[0]);
}
__device__ __forceinline__ float group16_lse(float v, bool active) {   // over the 16 lanes of a half-warp
  float mx = v;
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off, 16));
  float ex = active ? fast_exp(v - mx) : 0.0f;
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) ex += __shfl_xor_sync(0xffffffffu, ex, off, 16);
  return mx + fast_log(ex);
}

constexpr int kL16Threads = 128;   // maximum; long sequences launch fewer warps per block so the per-sentence slabs fit

// shared memory per sentence (16-byte aligned pieces): alpha[S][16] | e[S*T] | on_idx[S] | y[S] | xch[2][16]
__host__ __device__ inline size_t llh16_seq_bytes(int S, int T, bool with_alpha) {
  size_t n = with_alpha ? (size_t)S * 16 * sizeof(float) : 0;
  n += (((size_t)S * T + 3) / 4 * 4) * sizeof(float);
  n += (((size_t)2 * S + 3) / 4 * 4) * sizeof(int);
  n += 32 * sizeof(float);
  return n;
}

template <bool kBackward>
__global__ void __launch_bounds__(kL16Threads) crf_llh16_kernel(
    const float* __restrict__ emissions, const int64_t* __restrict__ tags, const uint8_t* __restrict__ mask,
    const float* __restrict__ start, const float* __restrict__ end, const float* __restrict__ trans,
    const float* __restrict__ w, float* __restrict__ llh_out, float* __restrict__ d_emissions,
    float* __restrict__ d_start, float* __restrict__ d_end, float* __restrict__ d_trans, int B, int S, int T) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr unsigned kFull = 0xffffffffu;
  float* s_dtrans = reinterpret_cast<float*>(smem_raw);      // [T*T] | d_start [T] | d_end [T]   (backward only)
  float* s_dstart = s_dtrans + T * T;
  float* s_dend = s_dstart + T;
  const size_t head = kBackward ? (((size_t)T * T + 2 * T) * sizeof(float) + 15) / 16 * 16 : 0;
  const int g_in_block = threadIdx.x >> 4;
  const int gl = threadIdx.x & 15;
  const int gshift = threadIdx.x & 16;
  if (kBackward) {
    for (int i = threadIdx.x; i < T * T + 2 * T; i += blockDim.x) s_dtrans[i] = 0.0f;
    __syncthreads();
  }
  const int seqs_per_block = blockDim.x >> 4;
  const int b_raw = blockIdx.x * seqs_per_block + g_in_block;
  const bool warp_ok = blockIdx.x * seqs_per_block + (g_in_block & ~1) < B;
  if (warp_ok) {
    const bool seq_ok = b_raw < B;
    const int b = seq_ok ? b_raw : B - 1;   // an odd tail half shadows the last sentence and stores nothing
    uint8_t* base = smem_raw + head + (size_t)g_in_block * llh16_seq_bytes(S, T, kBackward);
    float* alpha = reinterpret_cast<float*>(base);                                   // [S][16] (backward only)
    float* e_s = reinterpret_cast<float*>(base) + (kBackward ? (size_t)S * 16 : 0);  // [S*T]
    int* on_idx = reinterpret_cast<int*>(e_s + ((size_t)S * T + 3) / 4 * 4);        // [S]
    int* y_s = on_idx + S;                                                            // [S]
    float* xch = reinterpret_cast<float*>(on_idx + ((size_t)2 * S + 3) / 4 * 4);    // [2][16]
    const uint8_t* m_g = mask ? mask + (size_t)b * S : nullptr;
    const int64_t* y_g = tags + (size_t)b * S;
    const bool active = gl < T;

    stage_slab<16>(emissions + (size_t)b * S * T, e_s, S * T, gl);
    for (int t = gl; t < S; t += 16) y_s[t] = gold_tag(y_g[t], T);
    // on-steps (t = 0 always counts, as in _compute_normalizer); len = sum(mask) as pytorch-crf counts it
    int n_on = 0, len = 0;
    for (int t0 = 0; t0 < S; t0 += 16) {
      const int t = t0 + gl;
      const bool m = (t < S) && (m_g ? (m_g[t] != 0) : true);
      const bool on = (t < S) && (m || t == 0);
      const unsigned bits_on = (__ballot_sync(kFull, on) >> gshift) & 0xffffu;
      const unsigned bits_m = (__ballot_sync(kFull, m) >> gshift) & 0xffffu;
      if (on) on_idx[n_on + __popc(bits_on & ((1u << gl) - 1u))] = t;
      n_on += __popc(bits_on);
      len += __popc(bits_m);
    }
    const int n_warp = max(n_on, __shfl_xor_sync(kFull, n_on, 16));
    __syncwarp();

    float tr[16];   // column gl of the transition matrix; rows >= T are -inf so they never count
#pragma unroll
    for (int i = 0; i < 16; ++i) tr[i] = (i < T) ? (active ? trans[i * T + gl] : 0.0f) : -INFINITY;

    // ---- forward sweep over the on-steps: a[j] = logsumexp_i((a[i] + trans[i][j]) + e[t][j]) ----
    float a = active ? start[gl] + e_s[gl] : 0.0f;   // idle lanes carry 0
    float* pub = kBackward ? alpha : xch;
    pub[gl] = a;
    __syncwarp();
    for (int n = 1; n < n_warp; ++n) {
      const bool mine = n < n_on;
      const int t = on_idx[mine ? n : 0];
      const float e = e_s[t * T + (active ? gl : 0)];
      float prev[16], c[16];
      load16(kBackward ? alpha + (n - 1) * 16 : xch + ((n - 1) & 1) * 16, prev);
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = (prev[i] + tr[i]) + e;
      const float nxt = lse16(c);
      a = (mine && active) ? nxt : a;
      if (kBackward) { if (mine) alpha[n * 16 + gl] = a; }
      else xch[(n & 1) * 16 + gl] = a;
      __syncwarp();
    }
    const float en = active ? end[gl] : 0.0f;
    const float logz = group16_lse(active ? a + en : -INFINITY, active);

    if (!kBackward) {
      // ---- gold path score (pytorch-crf _compute_score): lanes take strided time steps, then reduce ----
      float sc = 0.0f;
      for (int t = gl; t < S; t += 16) {
        const int y = y_s[t];
        const bool on = m_g ? (m_g[t] != 0) : true;
        if (t == 0) sc += start[y] + e_s[y];
        else if (on) sc += trans[y_s[t - 1] * T + y] + e_s[t * T + y];
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) sc += __shfl_xor_sync(kFull, sc, off, 16);
      if (gl == 0 && seq_ok) llh_out[b] = (sc + end[y_s[max(len - 1, 0)]]) - logz;
    } else {
      // ---- backward sweep: marginals, pairwise marginals, beta ----
      float* de_g = d_emissions + (size_t)b * S * T;
      if (seq_ok)
        for (int i = gl; i < S * T; i += 16) de_g[i] = 0.0f;
      __syncwarp();
      const float wb = seq_ok ? w[b] : 0.0f;
      float trr[16];   // row gl of the transition matrix; columns >= T are -inf
#pragma unroll
      for (int j = 0; j < 16; ++j) trr[j] = (j < T) ? (active ? trans[gl * T + j] : 0.0f) : -INFINITY;
      float acc_tr[16];   // sum over steps of P(y_prev = i, y_cur = gl)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc_tr[i] = 0.0f;
      float beta = en;
      const int y_last = y_s[max(len - 1, 0)];
      for (int kk = 0; kk < n_warp; ++kk) {
        const int n = n_on - 1 - kk;         // this half's on-step (counted from its own end)
        const bool mine = n >= 0;
        const int nn = mine ? n : 0;
        const int t = on_idx[nn];
        const float an = alpha[nn * 16 + gl];
        const float e = e_s[t * T + (active ? gl : 0)];
        const float p = fast_exp(an + beta - logz);
        const int y = y_s[t];
        if (mine && active && seq_ok) {
          de_g[t * T + gl] = wb * ((y == gl ? 1.0f : 0.0f) - p);
          if (n == n_on - 1) atomicAdd(s_dend + gl, wb * ((y_last == gl ? 1.0f : 0.0f) - p));
          if (n == 0) atomicAdd(s_dstart + gl, wb * ((y == gl ? 1.0f : 0.0f) - p));
        }
        // publish e[t][j] + beta_n[j]; idle lanes publish -inf so they never count
        xch[(kk & 1) * 16 + gl] = active ? e + beta : -INFINITY;
        __syncwarp();
        const bool step = n >= 1;            // a transition (on-step n-1 -> n) exists
        float eb[16], c[16], prev[16];
        load16(xch + (kk & 1) * 16, eb);
        load16(alpha + (step ? n - 1 : 0) * 16, prev);
        const float ebm = active ? e + beta : 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float pr = fast_exp(((prev[i] + tr[i]) + ebm) - logz);   // tr[i] = -inf for i >= T -> 0
          acc_tr[i] += (step && active) ? pr : 0.0f;
          c[i] = trr[i] + eb[i];
        }
        const float nb = lse16(c);
        beta = (step && active) ? nb : beta;
      }
      if (active && seq_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < T) atomicAdd(s_dtrans + i * T + gl, -wb * acc_tr[i]);
      }
      // gold transitions: pytorch-crf's _compute_score pairs every on-step t >= 1 with position t-1
      if (seq_ok) {
        for (int t = 1 + gl; t < S; t += 16) {
          const bool on = m_g ? (m_g[t] != 0) : true;
          if (on) atomicAdd(s_dtrans + y_s[t - 1] * T + y_s[t], wb);
        }
      }
    }
  }
  if (kBackward) {
    __syncthreads();
    for (int i = threadIdx.x; i < T * T; i += blockDim.x) atomicAdd(d_trans + i, s_dtrans[i]);
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      atomicAdd(d_start + i, s_dstart[i]);
      atomicAdd(d_end + i, s_dend[i]);
    }
  }
}

}  // namespace

extern "C" int icka_viterbi_decode(icka_handle* h, const float* emissions, const uint8_t* mask,
                                   const float* start, const float* end, const float* trans,
                                   int32_t* tags_out, int32_t* lens_out, int B, int S, int T, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && T >= 1, "viterbi: bad shape B=%d S=%d T=%d", B, S, T);
  ICKA_REQUIRE(T <= 32, "viterbi: num_tags %d > 32 not supported by this kernel", T);
  ICKA_REQUIRE(emissions && start && end && trans && tags_out && lens_out, "viterbi: null pointer");
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int LPS = (T <= 16) ? 16 : 32;
  const size_t per_seq = seq_smem_bytes(S, T, LPS);
  int spb = kThreads / LPS;
  // long sequences (S >~ 470 at T = 15): fewer sentences per block (whole warps: 2 sentences each) until the slabs fit
  if (LPS == 16)
    while (spb > 2 && per_seq * spb > h->smem_optin) spb >>= 1;
  const size_t smem = per_seq * spb;
  if (smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "viterbi: S=%d T=%d needs %zu B shared memory per block (max %zu)", S, T,
              smem, h->smem_optin);
  const int grid = (B + spb - 1) / spb;
  // (A one-sentence-per-warp variant -- 32 lanes share the 16 x 16 candidates of a step, half the instructions, one
  // shuffle pair to join the halves -- was built and measured: 31.7 vs 29.8 us at 1024 full-length sentences, 72 vs 65 us
  // at 4096.  The step is a ~250-cycle chain of shared-memory round trip + two adds + four FMNMX levels + ~140 issue
  // slots of one warp per scheduler; halving the arithmetic does not shorten it.  tools/viterbi_steps.py: 0.19 us per
  // step, 2.5 us fixed.)
  if (LPS == 16) {
    ICKA_CUDA(cudaFuncSetAttribute(viterbi16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    viterbi16_kernel<<<grid, spb * 16, smem, st>>>(emissions, mask, start, end, trans, tags_out, lens_out, B, S,
                                                   T, per_seq);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(viterbi_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    viterbi_kernel<32><<<grid, kThreads, smem, st>>>(emissions, mask, start, end, trans, tags_out, lens_out, B,
                                                      S, T, per_seq);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_crf_llh_fwd(icka_handle* h, const float* emissions, const int64_t* tags, const uint8_t* mask,
                                const float* start, const float* end, const float* trans, float* llh_out,
                                int B, int S, int T, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && T >= 1, "crf_llh: bad shape B=%d S=%d T=%d", B, S, T);
  ICKA_REQUIRE(T <= 32, "crf_llh: num_tags %d > 32 not supported by this kernel", T);
  ICKA_REQUIRE(emissions && tags && start && end && trans && llh_out, "crf_llh: null pointer");
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int LPS = (T <= 16) ? 16 : 32;
  const int spb = kThreads / LPS;
  const int grid = (B + spb - 1) / spb;
  const size_t smem = ((((size_t)S * T + 3) / 4 * 4) * sizeof(float) + ((size_t)S + 15) / 16 * 16) * spb;
  if (LPS != 16 && smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "crf_llh: S=%d T=%d needs %zu B shared memory per block (max %zu)", S, T, smem,
              h->smem_optin);
  if (LPS == 16) {
    int seqs = kL16Threads / 16;
    while (seqs > 2 && llh16_seq_bytes(S, T, false) * seqs > h->smem_optin) seqs >>= 1;
    const size_t smem16 = llh16_seq_bytes(S, T, false) * seqs;
    if (smem16 > h->smem_optin)
      ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "crf_llh: S=%d T=%d needs %zu B shared memory per block", S, T, smem16);
    ICKA_CUDA(cudaFuncSetAttribute(crf_llh16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    crf_llh16_kernel<false><<<(B + seqs - 1) / seqs, seqs * 16, smem16, st>>>(
        emissions, tags, mask, start, end, trans, nullptr, llh_out, nullptr, nullptr, nullptr, nullptr, B, S, T);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(crf_llh_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    crf_llh_kernel<32><<<grid, kThreads, smem, st>>>(emissions, tags, mask, start, end, trans, llh_out, B, S, T);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_crf_llh_bwd(icka_handle* h, const float* emissions, const int64_t* tags, const uint8_t* mask,
                                const float* start, const float* end, const float* trans, const float* w,
                                float* d_emissions, float* d_start, float* d_end, float* d_trans, int B, int S, int T,
                                void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && T >= 1, "crf_llh_bwd: bad shape B=%d S=%d T=%d", B, S, T);
  ICKA_REQUIRE(T <= 32, "crf_llh_bwd: num_tags %d > 32 not supported by this kernel", T);
  ICKA_REQUIRE(emissions && tags && start && end && trans && w && d_emissions && d_start && d_end && d_trans,
               "crf_llh_bwd: null pointer");
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int LPS = (T <= 16) ? 16 : 32;
  const int spb = kThreads / LPS;
  const size_t head = (((size_t)T * T + 2 * T) * sizeof(float) + 15) / 16 * 16;
  const size_t per_seq =
      ((size_t)S * LPS * sizeof(float) + 2 * (size_t)S * sizeof(int) + (((size_t)S * T + 3) / 4 * 4) * sizeof(float) + 15) /
      16 * 16;
  const size_t smem = head + per_seq * spb;
  if (LPS != 16 && smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "crf_llh_bwd: S=%d T=%d needs %zu B shared memory per block (max %zu)", S, T, smem,
              h->smem_optin);
  const int grid = (B + spb - 1) / spb;
  if (LPS == 16) {
    // S = 256 (the hi-res shape) at 8 sentences per block needs ~271 KB: halve the sentences per block until it fits
    int seqs = kL16Threads / 16;
    while (seqs > 2 && head + llh16_seq_bytes(S, T, true) * seqs > h->smem_optin) seqs >>= 1;
    const size_t smem16 = head + llh16_seq_bytes(S, T, true) * seqs;
    if (smem16 > h->smem_optin)
      ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "crf_llh_bwd: S=%d T=%d needs %zu B shared memory per block", S, T, smem16);
    ICKA_CUDA(cudaFuncSetAttribute(crf_llh16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
    crf_llh16_kernel<true><<<(B + seqs - 1) / seqs, seqs * 16, smem16, st>>>(
        emissions, tags, mask, start, end, trans, w, nullptr, d_emissions, d_start, d_end, d_trans, B, S, T);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(crf_llh_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    crf_llh_bwd_kernel<32><<<grid, kThreads, smem, st>>>(emissions, tags, mask, start, end, trans, w, d_emissions,
                                                         d_start, d_end, d_trans, B, S, T, per_seq);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
