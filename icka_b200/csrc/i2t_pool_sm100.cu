// Single-query (image -> text) attention pool in folded form on tcgen05 / TMEM (sm_100a), S <= 128 keys (KB = 1) or
// S <= 256 keys as two 128-key blocks per sentence (KB = 2; the description below is for one block).
//
// Same contract as i2t_pool_kernel (i2t_pool.cu): per sentence, U [nh x H] (the folded queries) against the text
// states X [S x H]:  scores[h][s] = U_h . x_s / 8 + mask[s],  p = softmax_s,  xbar_h = sum_s p[h][s] x_s.
// The mma.sync kernel is bound by shared-memory traffic and barriers (ncu: MIO throttle, short scoreboard, 36 %
// issue slots): every warp re-reads fragments through ldmatrix and eight partial score tiles are summed through
// shared memory.  Here both contractions run as SS tcgen05.mma straight from the TMA-written tiles, transposed
// so that the 128 keys (then the 128 hidden dims of a chunk pair) are the M dimension and the 16 heads the N
// dimension:
//   pass 1   S^T[128 keys x 16 heads] += X_c[128 x 64] . U_c[16 x 64]^T           c = 0 .. H/64-1   (K-major A, B)
//   softmax  thread = key: 16 head scores out of TMEM, per-head max / sum across the 128 threads (warp
//            butterflies + one shared-memory hop), P^T written as the [16 heads][128 keys] K-major B tile
//   pass 2   O^T[128 dims x 16 heads] = X_pair[128 keys x 128 dims]^T . P^T       pair = 0 .. H/128-1
//            (the SAME X tiles, now read as the MN-major A operand: no transposed copy)
//   epilogue thread = hidden dim: 16 x H/128 values out of TMEM -> bf16 -> xbar[h][dim]
// X streams twice through a 10-stage ring of 16-KB chunk tiles (the second pass hits L2), software-pipelined by one
// sentence (pass 1 of sentence n+1 runs while sentence n is in its softmax), U and P^T are double-buffered, two
// softmax/epilogue groups alternate sentences; one persistent CTA per SM.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace {

using namespace sm100;

constexpr int kKeys = 128;
constexpr int kHeads = 16;
constexpr int kChunkBytes = kKeys * 128;   // [128 keys][64 dims] bf16
constexpr int kThreads = 320;

template <int H, int KB = 1>     // KB: 128-key blocks per sentence (1: S <= 128; 2: S <= 256, the 448-px configuration)
struct PoolCfg {
  static constexpr int kChunks = H / 64;
  static constexpr int kPairs = H / 128;
  static constexpr int kUBytes = kChunks * kHeads * 128;           // [chunk][16 heads][64 dims]
  static constexpr int kPTBytes = KB * 2 * kHeads * 128;           // [2 KB key chunks][16 heads][64 keys]
  static constexpr int kStages = (H <= 768 && KB == 1) ? 10 : 8;   // even: pass 2 consumes stages in pairs
  static constexpr int kSlotCols = 32 + ((kPairs * 16 + 31) / 32) * 32;   // S^T at +0 (16 cols per key block), O^T at +32
  static_assert(KB >= 1 && KB <= 2, "the score slot holds two key blocks");
  static constexpr int kTmemCols = (2 * kSlotCols <= 256) ? 256 : 512;
  static constexpr size_t kSmemBytes = (size_t)kStages * kChunkBytes + 2 * kUBytes + 2 * kPTBytes + 2 * 2 * 4 * kHeads * 4 +
                                       1024 /*align*/ + 512 /*barriers*/;
};

struct PoolArgs {
  const float* mask_add;   // [B, S] or null
  __nv_bfloat16* xbar;     // [B, nh * H]
  int B, S, nh;
};

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <int H, int KB>
__global__ void __launch_bounds__(kThreads, 1)
i2t_pool_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_u,
                        const PoolArgs args) {
  using Cfg = PoolCfg<H, KB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* x_ring = smem;                                           // kStages x 16 KB
  uint8_t* u_base = x_ring + (size_t)Cfg::kStages * kChunkBytes;    // 2 x kUBytes
  uint8_t* pt_base = u_base + 2 * Cfg::kUBytes;                     // 2 x kPTBytes
  float* red = reinterpret_cast<float*>(pt_base + 2 * Cfg::kPTBytes);   // [2 groups][2][4 warps][16 heads]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * 2 * 4 * kHeads);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* u_full = empty_bar + Cfg::kStages;
  uint64_t* u_empty = u_full + 2;
  uint64_t* s_full = u_empty + 2;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_free = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_u);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&u_full[s], 1);
      mbar_init(&u_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // The first pass asks L2 to KEEP the lines (evict_last), the second releases them (evict_first): between the two reads
      // of a chunk every SM streams another sentence (58 MB in flight machine-wide, against two 63 MB L2 partitions), and
      // with the default policy 18 % of the second pass came from DRAM again (ncu: 281 MB per launch vs 239 MB algorithmic).
      const uint64_t pol_keep = l2_policy_evict_last(), pol_done = l2_policy_evict_first();
      auto x_tile = [&](int b, int kb, int c, uint64_t pol) {   // chunk c (64 dims) of key block kb of sentence b
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], kChunkBytes);
        tma_load_2d_hint(x_ring + (size_t)stage * kChunkBytes, &tmap_x, &full_bar[stage], c * 64, b * args.S + kb * kKeys, pol);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      };
      // pass 1 walks (key block, chunk); pass 2 walks (pair, key block, the pair's two chunks): the MMA warp consumes in
      // exactly this order
      auto x_chunks = [&](int b, uint64_t pol, bool second) {
        if (!second) {
          for (int kb = 0; kb < KB; ++kb)
            for (int c = 0; c < Cfg::kChunks; ++c) x_tile(b, kb, c, pol);
        } else {
          for (int pr = 0; pr < Cfg::kPairs; ++pr)
            for (int kb = 0; kb < KB; ++kb) {
              x_tile(b, kb, 2 * pr, pol);
              x_tile(b, kb, 2 * pr + 1, pol);
            }
        }
      };
      // Software-pipelined by one sentence: pass 1 of sentence n+1 is streamed (from HBM) BEFORE pass 2 of
      // sentence n (L2 hits), so the HBM stream never waits for a softmax.  The MMA warp consumes in this order.
      int n = 0, prev_b = -1;
      for (int b = blockIdx.x;; b += gridDim.x, ++n) {
        const bool live = b < args.B;
        if (live) {
          const int slot = n & 1;
          mbar_wait(&u_empty[slot], ((n >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&u_full[slot], Cfg::kUBytes);
          for (int c = 0; c < Cfg::kChunks; ++c)
            tma_load_2d(u_base + slot * Cfg::kUBytes + c * (kHeads * 128), &tmap_u, &u_full[slot], c * 64, b * args.nh);
          x_chunks(b, pol_keep, false);            // pass 1 of sentence n
        }
        if (prev_b >= 0) x_chunks(prev_b, pol_done, true);   // pass 2 of sentence n-1
        if (!live) break;
        prev_b = b;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(kKeys, kHeads);                 // X_c K-major, U_c K-major
      constexpr uint32_t idesc_o = make_idesc_bf16_f32(128, kHeads, true, false);      // X pair MN-major, P^T K-major
      int stage = 0;
      uint32_t phase = 0;
      auto pass1 = [&](int n) {   // S^T += X_c . U_c^T for the n-th sentence of this CTA
        const int slot = n & 1;
        const uint32_t tmem_s = tmem_base + (uint32_t)(slot * Cfg::kSlotCols);
        mbar_wait(&u_full[slot], (n >> 1) & 1);
        const uint32_t u_addr = smem_u32(u_base + slot * Cfg::kUBytes);
        for (int kb = 0; kb < KB; ++kb)
          for (int c = 0; c < Cfg::kChunks; ++c) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t x_addr = smem_u32(x_ring + (size_t)stage * kChunkBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_s + (uint32_t)(kb * kHeads), make_kmajor_sw128_desc(x_addr + k * 32),
                        make_kmajor_sw128_desc(u_addr + c * (kHeads * 128) + k * 32), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        umma_commit(&u_empty[slot]);
        umma_commit(&s_full[slot]);
      };
      auto pass2 = [&](int n) {   // O^T[pair] = X_pair^T . P^T
        const int slot = n & 1;
        const uint32_t tmem_o = tmem_base + (uint32_t)(slot * Cfg::kSlotCols + 32);
        mbar_wait(&p_full[slot], (n >> 1) & 1);
        if (n >= 2) mbar_wait(&o_free[slot], ((n >> 1) - 1) & 1);
        const uint32_t pt_addr = smem_u32(pt_base + slot * Cfg::kPTBytes);
        for (int pr = 0; pr < Cfg::kPairs; ++pr)
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            mbar_wait(&full_bar[stage + 1], phase);      // kStages is even and pairs start on even stages
            tc_fence_after();
            const uint32_t x_addr = smem_u32(x_ring + (size_t)stage * kChunkBytes);
#pragma unroll
            for (int k = 0; k < kKeys / 16; ++k)
              umma_bf16(tmem_o + (uint32_t)(pr * 16), make_mnmajor_sw128_desc(x_addr + k * (16 * 128), kChunkBytes),
                        make_kmajor_sw128_desc(pt_addr + (kb * 2 + (k >> 2)) * (kHeads * 128) + (k & 3) * 32), idesc_o,
                        (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            umma_commit(&empty_bar[stage + 1]);
            stage += 2;
            if (stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        umma_commit(&o_full[slot]);
      };
      // same order as the producer: P1(0), P1(1), P2(0), P1(2), P2(1), ...  (S^T of slot (n+1)&1 is free: its last
      // reader signalled p_full(n-1), which pass2(n-1) waited for earlier in this sequence)
      int n = 0;
      for (int b = blockIdx.x;; b += gridDim.x, ++n) {
        const bool live = b < args.B;
        if (live) pass1(n);
        if (n >= 1) pass2(n - 1);
        if (!live) break;
      }
    }
  } else {
    // ===================== softmax + epilogue groups (warps 2-5: even sentences, warps 6-9: odd) =====================
    const int group = (warp - 2) >> 2;
    const int gw = (warp - 2) & 3;                     // warp index inside the group
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int key = quad * 32 + lane;                  // == TMEM lane: key in pass 1, hidden dim of the pair in pass 2
    uint8_t* pt_buf = pt_base + group * Cfg::kPTBytes;
    float* red_max = red + group * (2 * 4 * kHeads);
    float* red_sum = red_max + 4 * kHeads;
    const uint32_t tmem_s = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(group * Cfg::kSlotCols);
    const uint32_t tmem_o = tmem_s + 32;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kScale = 0.125f * kLog2e;
    int n = 0;
    for (int b = blockIdx.x; b < args.B; b += gridDim.x, ++n) {
      if ((n & 1) != group) continue;
      const uint32_t par = (n >> 1) & 1;
      float mk[KB];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const int gk = kb * kKeys + key;
        mk[kb] = (gk < args.S) ? (args.mask_add ? args.mask_add[(size_t)b * args.S + gk] * kLog2e : 0.0f) : -INFINITY;
      }
      mbar_wait(&s_full[group], par);
      tc_fence_after();
      float v[KB][16], m[16];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        uint32_t sr[16];
        tmem_ld_32x32b_x16(tmem_s + (uint32_t)(kb * kHeads), sr);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          v[kb][h] = fmaf(__uint_as_float(sr[h]), kScale, mk[kb]);
          m[h] = kb == 0 ? v[kb][h] : fmaxf(m[h], v[kb][h]);
        }
      }
      // per-head max over the 128 keys: butterflies inside the warp, then one hop through shared memory
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
        for (int h = 0; h < 16; ++h) m[h] = fmaxf(m[h], __shfl_xor_sync(0xffffffffu, m[h], off));
      if (lane == 0) {   // every lane holds all 16 maxima after the butterflies
#pragma unroll
        for (int h = 0; h < 16; ++h) red_max[gw * kHeads + h] = m[h];
      }
      named_bar_sync(1 + group, 128);
      float l[16];
#pragma unroll
      for (int h = 0; h < 16; ++h) {
        const float mh = fmaxf(fmaxf(red_max[h], red_max[kHeads + h]), fmaxf(red_max[2 * kHeads + h], red_max[3 * kHeads + h]));
        l[h] = 0.0f;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          v[kb][h] = ex2(v[kb][h] - mh);
          l[h] += v[kb][h];
        }
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
        for (int h = 0; h < 16; ++h) l[h] += __shfl_xor_sync(0xffffffffu, l[h], off);
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < 16; ++h) red_sum[gw * kHeads + h] = l[h];
      }
      named_bar_sync(1 + group, 128);
      // P^T[h][key] (bf16) into the K-major 128-B-swizzled [16 heads][64 keys] x 2 tile
      {
        const int c16 = (key & 63) >> 3;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const float lh = (red_sum[h] + red_sum[kHeads + h]) + (red_sum[2 * kHeads + h] + red_sum[3 * kHeads + h]);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {     // key chunk (64 keys) index: 2 kb + (key >> 6)
            uint8_t* col = pt_buf + (kb * 2 + (key >> 6)) * (kHeads * 128) + (key & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(col + h * 128 + ((c16 ^ (h & 7)) << 4)) = __float2bfloat16_rn(v[kb][h] / lh);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[group]);

      // ---- epilogue: this thread's hidden dim of every pair, all heads ----
      mbar_wait(&o_full[group], par);
      tc_fence_after();
      __nv_bfloat16* out = args.xbar + (size_t)b * args.nh * H + key;
#pragma unroll
      for (int pr = 0; pr < Cfg::kPairs; ++pr) {
        uint32_t orr[16];
        tmem_ld_32x32b_x16(tmem_o + (uint32_t)(pr * 16), orr);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 16; ++h)
          if (h < args.nh) out[(size_t)h * H + pr * 128] = __float2bfloat16_rn(__uint_as_float(orr[h]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[group]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int H, int KB>
int launch_pool(icka_handle* h, const void* U, const void* X, const float* mask_add, void* xbar, int B, int S, int nh,
                cudaStream_t st) {
  using Cfg = PoolCfg<H, KB>;
  if (h->smem_optin < Cfg::kSmemBytes) return 1;
  CUtensorMap tx, tu;
  int rc = icka_make_tmap_bf16(h, &tx, X, (int64_t)B * S, H, H, kKeys);          // box {64 dims, 128 keys}
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tu, U, (int64_t)B * nh, H, H, kHeads);            // box {64 dims, 16 heads}
  if (rc) return rc;
  PoolArgs args{mask_add, static_cast<__nv_bfloat16*>(xbar), B, S, nh};
  ICKA_CUDA(cudaFuncSetAttribute(i2t_pool_tcgen05_kernel<H, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
  const int grid = B < h->sm_count ? B : h->sm_count;
  i2t_pool_tcgen05_kernel<H, KB><<<grid, kThreads, Cfg::kSmemBytes, st>>>(tx, tu, args);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

}  // namespace

// ICKA_OK after launching; > 0 when the shape is outside this kernel's envelope (caller uses the mma.sync kernel).
int icka_i2t_pool_tcgen05_launch(icka_handle* h, const void* U, const void* X, const float* mask_add, void* xbar, int B,
                                 int S, int H, int nh, cudaStream_t st) {
  if (S > 2 * kKeys || nh > kHeads) return 1;
  if (S <= kKeys) {
    if (H == 768) return launch_pool<768, 1>(h, U, X, mask_add, xbar, B, S, nh, st);
    if (H == 1024) return launch_pool<1024, 1>(h, U, X, mask_add, xbar, B, S, nh, st);
  } else {   // two 128-key blocks per sentence (S = 256: the 448-px configuration)
    if (H == 768) return launch_pool<768, 2>(h, U, X, mask_add, xbar, B, S, nh, st);
    if (H == 1024) return launch_pool<1024, 2>(h, U, X, mask_add, xbar, B, S, nh, st);
  }
  return 1;
}
