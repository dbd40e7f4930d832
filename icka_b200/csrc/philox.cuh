// Counter-based dropout masks (Philox4x32-10, Salmon et al. 2011): every element's keep/drop decision is a pure
// function of (seed, site, element index), so forward and backward kernels -- and the test oracle, through
// icka_dropout_mask -- regenerate the same mask without storing it.  The reference's nn.Dropout sites on the path:
// attention probabilities (CMIM:616) and the two dense outputs (CMIM:563, 534).
#pragma once
#include <stdint.h>

namespace icka_rng {

enum Site : uint32_t { kSiteHidden = 0x48, kSiteAttention = 0x41 };

__host__ __device__ inline uint32_t keep_threshold(float p_drop) {
  // element kept iff random word < threshold; threshold = (1 - p) * 2^32 (p = 0 keeps everything)
  const double t = (1.0 - (double)p_drop) * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}

__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint32_t site, uint64_t index) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint4 c = make_uint4((uint32_t)index, (uint32_t)(index >> 32), site, 0u);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// Seed of a launch = the by-value seed of the call + an optional device-resident base (icka_set_seed_base): a training step
// captured into a CUDA graph replays with frozen kernel arguments, so the part of the seed that must change from step to
// step lives in device memory, advanced by a kernel of the graph itself.
__device__ __forceinline__ uint64_t effective_seed(uint64_t seed, const unsigned long long* base) {
  return base ? seed + (uint64_t)__ldg(base) : seed;
}

// Keep bits (bit j = element 4*group + j kept) of one group of four consecutive elements.
__device__ __forceinline__ uint32_t keep_bits4(uint64_t seed, uint32_t site, uint64_t group, uint32_t thresh) {
  const uint4 r = philox4x32_10(seed, site, group);
  return (r.x < thresh ? 1u : 0u) | (r.y < thresh ? 2u : 0u) | (r.z < thresh ? 4u : 0u) | (r.w < thresh ? 8u : 0u);
}

// Attention probabilities: row = ((b * nh + h) * Sq + q), groups of four keys inside the row.
__device__ __forceinline__ uint64_t attn_group(uint64_t row, int Skv, int key) {
  return row * (uint64_t)((Skv + 3) / 4) + (uint64_t)(key >> 2);
}

}  // namespace icka_rng
