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

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct SeqSmem {
  float* em;       // [S*T] (rounded up to 4)
  uint8_t* bp;     // [(S-1)*LPS]
  uint8_t* mask;   // [S]
  uint8_t* tagb;   // [S]
};

__host__ __device__ inline size_t seq_smem_bytes(int S, int T, int LPS) {
  size_t em = ((size_t)S * T + 3) / 4 * 16;
  size_t bp = ((size_t)(S > 1 ? S - 1 : 1) * LPS + 15) / 16 * 16;
  size_t mk = ((size_t)S + 15) / 16 * 16;
  return em + bp + 2 * mk;
}

template <int LPS>
__device__ __forceinline__ SeqSmem carve(uint8_t* base, int S, int T) {
  SeqSmem s;
  size_t em = ((size_t)S * T + 3) / 4 * 16;
  size_t bp = ((size_t)(S > 1 ? S - 1 : 1) * LPS + 15) / 16 * 16;
  size_t mk = ((size_t)S + 15) / 16 * 16;
  s.em = reinterpret_cast<float*>(base);
  s.bp = base + em;
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

// Per-sentence log-likelihood: gold-path score minus log-partition (forward algorithm), fp32.
// Same lane mapping as Viterbi; logsumexp over predecessors i is max-shifted like torch.logsumexp.
template <int LPS>
__global__ void __launch_bounds__(kThreads) crf_llh_kernel(
    const float* __restrict__ emissions, const int64_t* __restrict__ tags, const uint8_t* __restrict__ mask,
    const float* __restrict__ start, const float* __restrict__ end, const float* __restrict__ trans,
    float* __restrict__ llh_out, int B, int S, int T) {
  const int g_in_block = threadIdx.x / LPS;
  const int gl = threadIdx.x % LPS;
  const int gshift = (threadIdx.x % 32) / LPS * LPS;
  const unsigned gmask = (LPS == 32) ? 0xffffffffu : (0xffffu << gshift);
  const int b = blockIdx.x * (kThreads / LPS) + g_in_block;
  if (b >= B) return;
  const float* e_g = emissions + (size_t)b * S * T;
  const uint8_t* m_g = mask ? mask + (size_t)b * S : nullptr;
  const int64_t* y_g = tags + (size_t)b * S;
  const bool active = gl < T;

  float tr[LPS];
#pragma unroll
  for (int i = 0; i < LPS; ++i) tr[i] = (active && i < T) ? trans[i * T + gl] : 0.0f;

  // ---- normalizer (pytorch-crf _compute_normalizer) ----
  float alpha = active ? start[gl] + e_g[gl] : -INFINITY;
  for (int t = 1; t < S; ++t) {
    const bool on = m_g ? (m_g[t] != 0) : true;   // group-uniform
    if (!on) continue;
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
    const int y = (int)y_g[t];
    const bool on = m_g ? (m_g[t] != 0) : true;
    if (t == 0) {
      s += start[y] + e_g[y];
      len += on;
    } else if (on) {
      s += trans[(int)y_g[t - 1] * T + y] + e_g[t * T + y];
      len += 1;
    }
  }
#pragma unroll
  for (int off = LPS / 2; off >= 1; off >>= 1) {
    s += __shfl_xor_sync(gmask, s, off, LPS);
    len += __shfl_xor_sync(gmask, len, off, LPS);
  }
  if (gl == 0) {
    const int last = (int)y_g[max(len - 1, 0)];
    llh_out[b] = (s + end[last]) - logz;
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
  const int spb = kThreads / LPS;
  const size_t smem = per_seq * spb;
  if (smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "viterbi: S=%d T=%d needs %zu B shared memory per block (max %zu)", S, T,
              smem, h->smem_optin);
  const int grid = (B + spb - 1) / spb;
  if (LPS == 16) {
    ICKA_CUDA(cudaFuncSetAttribute(viterbi_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    viterbi_kernel<16><<<grid, kThreads, smem, st>>>(emissions, mask, start, end, trans, tags_out, lens_out, B,
                                                      S, T, per_seq);
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
  if (LPS == 16)
    crf_llh_kernel<16><<<grid, kThreads, 0, st>>>(emissions, tags, mask, start, end, trans, llh_out, B, S, T);
  else
    crf_llh_kernel<32><<<grid, kThreads, 0, st>>>(emissions, tags, mask, start, end, trans, llh_out, B, S, T);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
