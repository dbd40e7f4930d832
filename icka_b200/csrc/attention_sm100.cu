// Cross-attention core on tcgen05 / TMEM for sm_100a (bf16; Skv <= 64: the 49-region text->image attention, and a
// wide variant for Skv <= 224: the 196-region grid of the 448-px configuration).
//
// Same contract as cross_attn_mma_kernel (attention.cu; BertCoAttention, CMIM:598-623):
//   P = softmax(Q_h K_h^T / 8 + mask),  ctx_h = P V_h,  heads merged in place.
// The mma.sync kernel spends its time in instruction issue (ncu: 54 % issue-slot utilisation, 4 CTAs / SM
// at 128 registers): fragments, quad shuffles for the row statistics, staging.  Here the two GEMMs of a
// (sentence, head, 128-query) item run on the 5th-generation tensor cores and the softmax works on whole
// rows:
//   S[128 x 64]  = Q[128 x 64] . K[64 x 64]^T     tcgen05.mma, operands by TMA (SWIZZLE_128B), fp32 in TMEM
//   row softmax : each of 128 threads pulls ITS row of S out of TMEM (tcgen05.ld, lane = row), so max / sum
//                 / exp2 need no cross-lane traffic; P is rounded to bf16 and written to shared memory in
//                 the K-major 128-B-swizzled layout the UMMA A descriptor expects
//   O[128 x 64]  = P[128 x 64] . V[64 x 64]        tcgen05.mma, V read in place as the MN-major B operand
//   epilogue    : thread = row again: O from TMEM, times 1/rowsum, bf16, one contiguous 128-B store per row
// One persistent CTA per SM, 320 threads: warp 0 TMA producer (3-stage ring of {Q, K, V} tiles), warp 1
// MMA issuer, warps 2-5 and 6-9 two softmax/epilogue groups that alternate items (each item owns one of
// two TMEM {S, O} slots and one of two P buffers), so the loads, the two MMAs, the softmax of one item and
// the epilogue of another overlap.  Keys beyond Skv inside the 64-row K / V box belong to the next sentence
// (or are TMA zero-fill): they get probability exactly 0.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "philox.cuh"

#include <stdlib.h>

namespace {

using namespace sm100;

constexpr int kD = 64;          // head dim
constexpr int kRows = 128;      // query rows per item
constexpr int kKeys = 64;       // keys per item (Skv <= 64)
constexpr int kStages = 5;
constexpr int kQBytes = kRows * kD * 2, kKBytes = kKeys * kD * 2, kVBytes = kKeys * kD * 2;
constexpr int kStageBytes = kQBytes + kKBytes + kVBytes;   // 32 KB
constexpr int kPBytes = kRows * kKeys * 2;                  // 16 KB
constexpr int kThreads = 320;
constexpr int kTmemCols = 256;                              // 2 slots x {S: 64, O: 64}
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 2 * kPBytes + 4 * kKeys * sizeof(float) + 1024 + 256;

struct AttnArgs {
  const float* mask_add;   // [B, Skv] or null
  __nv_bfloat16* ctx;
  int64_t ldc;
  int B, Sq, Skv, nh, q_tiles;
  uint32_t drop_thresh;   // attention-probability dropout (CMIM:616), training only; 0 = off
  float drop_scale;
  uint64_t seed;
  const unsigned long long* seed_base;   // device-resident seed base (icka_set_seed_base) or null
  int debug;              // developer probe (ICKA_ATTN_DEBUG=1): CTA 0 prints the cycle stamps of its first items (wide2 kernel)
};

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 1)
cross_attn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                          const __grid_constant__ CUtensorMap tmap_v, const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B: 1024-B alignment
  uint8_t* stage_base = smem;
  uint8_t* p_base = smem + (size_t)kStages * kStageBytes;                         // 2 x [128][64] bf16
  float* mask_s = reinterpret_cast<float*>(p_base + 2 * kPBytes);                  // 2 groups x 2 buffers x [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(mask_s + 4 * kKeys);
  uint64_t* full_bar = bars;                   // [kStages] TMA bytes landed
  uint64_t* empty_bar = bars + kStages;        // [kStages] both MMAs of the item retired
  uint64_t* s_full = bars + 2 * kStages;       // [2] S = QK^T complete in TMEM slot
  uint64_t* p_full = s_full + 2;               // [2] P written to shared memory (and S consumed)
  uint64_t* o_full = p_full + 2;               // [2] O = PV complete in TMEM slot
  uint64_t* o_free = o_full + 2;               // [2] O consumed by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int items = args.B * args.nh * args.q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);    // one arrive per warp of the group
      mbar_init(&o_full[s], 1);
      mbar_init(&o_free[s], 4);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
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
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int qt = it % args.q_tiles, bh = it / args.q_tiles;
        const int h = bh % args.nh, b = bh / args.nh;
        uint8_t* st = stage_base + (size_t)stage * kStageBytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_2d(st, &tmap_q, &full_bar[stage], h * kD, b * args.Sq + qt * kRows);
        tma_load_2d(st + kQBytes, &tmap_k, &full_bar[stage], h * kD, b * args.Skv);
        tma_load_2d(st + kQBytes + kKBytes, &tmap_v, &full_bar[stage], h * kD, b * args.Skv);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(kRows, kKeys);              // Q, K both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16_f32(kRows, kD, false, true);    // P K-major, V MN-major
      int stage = 0, prev_stage = 0;
      uint32_t phase = 0;
      int n = 0;
      auto issue_pv = [&](int m, int st_idx) {   // O = P V for the m-th item of this CTA
        const int slot = m & 1;
        const uint32_t par = (m >> 1) & 1;
        mbar_wait(&p_full[slot], par);
        if (m >= 2) mbar_wait(&o_free[slot], ((m >> 1) - 1) & 1);   // the epilogue drained this slot's O
        tc_fence_after();
        const uint32_t p_addr = smem_u32(p_base + slot * kPBytes);
        const uint32_t v_addr = smem_u32(stage_base + (size_t)st_idx * kStageBytes + kQBytes + kKBytes);
        const uint32_t tmem_o = tmem_base + (uint32_t)(slot * 128 + 64);
#pragma unroll
        for (int k = 0; k < kKeys / 16; ++k) {
          const uint64_t da = make_kmajor_sw128_desc(p_addr + k * 32);
          const uint64_t db = make_mnmajor_sw128_desc(v_addr + k * (16 * 128), kKeys * 128);
          umma_bf16(tmem_o, da, db, idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[st_idx]);   // Q, K, V of this stage are no longer read
        umma_commit(&o_full[slot]);
      };
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        const int slot = n & 1;
        mbar_wait(&full_bar[stage], phase);
        // S of this slot was consumed when p_full of item n-2 completed, which issue_pv(n-2) waited for.
        tc_fence_after();
        const uint32_t q_addr = smem_u32(stage_base + (size_t)stage * kStageBytes);
        const uint32_t k_addr = q_addr + kQBytes;
        const uint32_t tmem_s = tmem_base + (uint32_t)(slot * 128);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(tmem_s, make_kmajor_sw128_desc(q_addr + k * 32), make_kmajor_sw128_desc(k_addr + k * 32), idesc_s,
                    k > 0 ? 1u : 0u);
        umma_commit(&s_full[slot]);
        if (n >= 1) issue_pv(n - 1, prev_stage);
        prev_stage = stage;
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (n >= 1) issue_pv(n - 1, prev_stage);
    }
  } else {
    // ===================== softmax + epilogue groups (warps 2-5: even items, warps 6-9: odd items) =====================
    const int group = (warp - 2) >> 2;                 // == TMEM slot == P buffer
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                  // query row inside the item
    const int gtid = (warp - 2 - group * 4) * 32 + lane;
    uint8_t* p_buf = p_base + group * kPBytes;
    float* mk2 = mask_s + group * 2 * kKeys;           // double-buffered: the next item's mask is fetched early
    const uint32_t tmem_s = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(group * 128);
    const uint32_t tmem_o = tmem_s + 64;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kScale = 0.125f * kLog2e;
    // additive mask of a sentence (log2 domain); keys >= Skv get -inf.  Written by the first 64 threads of the group.
    auto fetch_mask = [&](int item, float* dst) {
      if (gtid < kKeys && item < items) {
        const int b = item / (args.q_tiles * args.nh);
        dst[gtid] = (gtid < args.Skv) ? (args.mask_add ? args.mask_add[(size_t)b * args.Skv + gtid] * kLog2e : 0.0f)
                                      : -INFINITY;
      }
    };
    fetch_mask((int)blockIdx.x + group * (int)gridDim.x, mk2);
    int n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      if ((n & 1) != group) continue;
      const uint32_t par = (n >> 1) & 1;
      const int qt = it % args.q_tiles, bh = it / args.q_tiles;
      const int h = bh % args.nh, b = bh / args.nh;
      const float* mk = mk2 + ((n >> 1) & 1) * kKeys;
      named_bar_sync(1 + group, 128);      // this item's mask is in place; the other buffer is no longer read
      fetch_mask(it + 2 * (int)gridDim.x, mk2 + (((n >> 1) & 1) ^ 1) * kKeys);   // latency hides behind this item
      mbar_wait(&s_full[group], par);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_32x32b_x32(tmem_s, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld_32x32b_x32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float s = fmaf(__uint_as_float(sr[j]), kScale, mk[j]);
        sr[j] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      float l = 0.0f;
      uint8_t* prow = p_buf + row * 128;
      const uint64_t drow = ((uint64_t)b * args.nh + h) * (uint64_t)args.Sq + (uint64_t)(qt * kRows + row);
#pragma unroll
      for (int c = 0; c < 8; ++c) {          // 8 keys = one 16-byte chunk of the P row
        float p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          p[j] = ex2(__uint_as_float(sr[8 * c + j]) - mx);
          l += p[j];
        }
        if (args.drop_thresh) {              // the normaliser keeps the undropped sum; P' = P keep / (1 - p) feeds P.V
#pragma unroll
          for (int g4 = 0; g4 < 2; ++g4) {
            const uint32_t keep = icka_rng::keep_bits4(icka_rng::effective_seed(args.seed, args.seed_base), icka_rng::kSiteAttention,
                                                       icka_rng::attn_group(drow, args.Skv, 8 * c + 4 * g4), args.drop_thresh);
#pragma unroll
            for (int j = 0; j < 4; ++j) p[4 * g4 + j] = (keep >> j & 1u) ? p[4 * g4 + j] * args.drop_scale : 0.0f;
          }
        }
        uint4 u;
        u.x = pack_bf16x2(p[0], p[1]);
        u.y = pack_bf16x2(p[2], p[3]);
        u.z = pack_bf16x2(p[4], p[5]);
        u.w = pack_bf16x2(p[6], p[7]);
        *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = u;
      }
      fence_proxy_async();        // generic-proxy stores -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[group]);

      // ---- epilogue: O row from TMEM, normalise, one 128-byte row store ----
      mbar_wait(&o_full[group], par);
      tc_fence_after();
      uint32_t orr[64];
      tmem_ld_32x32b_x32(tmem_o, *reinterpret_cast<uint32_t(*)[32]>(&orr[0]));
      tmem_ld_32x32b_x32(tmem_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&orr[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[group]);
      // normalise, round to bf16 and park the row in this group's P buffer (PV has retired: o_full), then write
      // out with 16 bytes per lane and 128 contiguous bytes per 8 lanes (4 full rows per warp instruction)
      const float inv = 1.0f / l;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(orr[8 * c]) * inv, __uint_as_float(orr[8 * c + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(orr[8 * c + 2]) * inv, __uint_as_float(orr[8 * c + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(orr[8 * c + 4]) * inv, __uint_as_float(orr[8 * c + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(orr[8 * c + 6]) * inv, __uint_as_float(orr[8 * c + 7]) * inv);
        *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = u;
      }
      __syncwarp();
      {
        const int cch = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = quad * 32 + i * 4 + (lane >> 3);
          const int q_row = qt * kRows + r;
          if (q_row < args.Sq) {
            const uint4 u = *reinterpret_cast<const uint4*>(p_buf + r * 128 + ((cch ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(args.ctx + ((size_t)b * args.Sq + q_row) * args.ldc + (size_t)h * kD + cch * 8) = u;
          }
        }
      }
      __syncwarp();   // the rows are re-used by this warp's next softmax
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Wide variant: 64 < Skv <= 224 keys (the 196-region grid of the 448-px configuration; 128-key problems).
// Same structure with ONE {S, O} TMEM slot and one softmax group: S[128 x KEYS] is a single tcgen05.mma of
// N = KEYS, the row softmax reads its TMEM row in 32-column pieces twice (max, then exp / sum / P), P is KEYS wide
// (64-key K-major chunks), O = P V runs KEYS/16 k-steps over the in-place MN-major V tile.
// ------------------------------------------------------------------------------------------------------------
template <int KEYS>
struct WideCfg {
  static constexpr int kKVBytes = KEYS * 128;
  static constexpr int kStageBytes = kQBytes + 2 * kKVBytes;
  static constexpr int kStages = (KEYS <= 128) ? 3 : 2;
  static constexpr int kPChunks = (KEYS + 63) / 64;
  static constexpr int kPBytes = kPChunks * kRows * 128;
  static constexpr int kThreads = 192;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kPBytes + 2 * KEYS * sizeof(float) + 1024 + 256;
};

template <int KEYS>
__global__ void __launch_bounds__(WideCfg<KEYS>::kThreads, 1)
cross_attn_tcgen05_wide_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                               const __grid_constant__ CUtensorMap tmap_v, const AttnArgs args) {
  using Cfg = WideCfg<KEYS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* p_buf = smem + (size_t)Cfg::kStages * Cfg::kStageBytes;
  float* mask_s = reinterpret_cast<float*>(p_buf + Cfg::kPBytes);                  // 2 buffers x [KEYS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(mask_s + 2 * KEYS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* s_full = bars + 2 * Cfg::kStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint64_t* o_free = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  constexpr uint32_t kOCol = 256;              // S at columns [0, KEYS), O at [256, 320)

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int items = args.B * args.nh * args.q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int qt = it % args.q_tiles, bh = it / args.q_tiles;
        const int h = bh % args.nh, b = bh / args.nh;
        uint8_t* st = stage_base + (size_t)stage * Cfg::kStageBytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        tma_load_2d(st, &tmap_q, &full_bar[stage], h * kD, b * args.Sq + qt * kRows);
        tma_load_2d(st + kQBytes, &tmap_k, &full_bar[stage], h * kD, b * args.Skv);
        tma_load_2d(st + kQBytes + Cfg::kKVBytes, &tmap_v, &full_bar[stage], h * kD, b * args.Skv);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(kRows, KEYS);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32(kRows, kD, false, true);
      int stage = 0, prev_stage = 0;
      uint32_t phase = 0;
      int n = 0;
      auto issue_pv = [&](int m, int st_idx) {
        mbar_wait(p_full, m & 1);
        if (m >= 1) mbar_wait(o_free, (m - 1) & 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(p_buf);
        const uint32_t v_addr = smem_u32(stage_base + (size_t)st_idx * Cfg::kStageBytes + kQBytes + Cfg::kKVBytes);
#pragma unroll
        for (int k = 0; k < KEYS / 16; ++k)
          umma_bf16(tmem_base + kOCol, make_kmajor_sw128_desc(p_addr + (k >> 2) * (kRows * 128) + (k & 3) * 32),
                    make_mnmajor_sw128_desc(v_addr + k * (16 * 128), Cfg::kKVBytes), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(&empty_bar[st_idx]);
        umma_commit(o_full);
      };
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        mbar_wait(&full_bar[stage], phase);
        if (n >= 1) issue_pv(n - 1, prev_stage);      // also guarantees S of item n-1 has been read (p_full)
        tc_fence_after();
        const uint32_t q_addr = smem_u32(stage_base + (size_t)stage * Cfg::kStageBytes);
        const uint32_t k_addr = q_addr + kQBytes;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(tmem_base, make_kmajor_sw128_desc(q_addr + k * 32), make_kmajor_sw128_desc(k_addr + k * 32), idesc_s,
                    k > 0 ? 1u : 0u);
        umma_commit(s_full);
        prev_stage = stage;
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (n >= 1) issue_pv(n - 1, prev_stage);
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int gtid = (warp - 2) * 32 + lane;
    const uint32_t tmem_s = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tmem_o = tmem_s + kOCol;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kScale = 0.125f * kLog2e;
    auto fetch_mask = [&](int item, float* dst) {
      if (item < items) {
        const int b = item / (args.q_tiles * args.nh);
        for (int j = gtid; j < KEYS; j += 128)
          dst[j] = (j < args.Skv) ? (args.mask_add ? args.mask_add[(size_t)b * args.Skv + j] * kLog2e : 0.0f) : -INFINITY;
      }
    };
    fetch_mask((int)blockIdx.x, mask_s);
    int n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const uint32_t par = n & 1;
      const int qt = it % args.q_tiles, bh = it / args.q_tiles;
      const int h = bh % args.nh, b = bh / args.nh;
      const float* mk = mask_s + (n & 1) * KEYS;
      named_bar_sync(1, 128);
      fetch_mask(it + (int)gridDim.x, mask_s + ((n & 1) ^ 1) * KEYS);
      mbar_wait(s_full, par);
      tc_fence_after();
      // ---- sweep 1: row maximum ----
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < KEYS; c0 += 32) {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(tmem_s + (uint32_t)c0, sr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(sr[j]), kScale, mk[c0 + j]));
      }
      // ---- sweep 2: P = exp2(s - max) (bf16, dropped if training) into the K-major P tile; l = undropped row sum ----
      float l = 0.0f;
      const uint64_t drow = ((uint64_t)b * args.nh + h) * (uint64_t)args.Sq + (uint64_t)(qt * kRows + row);
#pragma unroll 1
      for (int c0 = 0; c0 < KEYS; c0 += 32) {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(tmem_s + (uint32_t)c0, sr);
        tmem_ld_wait();
        uint8_t* prow = p_buf + (c0 >> 6) * (kRows * 128) + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {      // 8 keys = one 16-byte chunk
          float p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            p[j] = ex2(fmaf(__uint_as_float(sr[8 * c + j]), kScale, mk[c0 + 8 * c + j]) - mx);
            l += p[j];
          }
          if (args.drop_thresh) {
#pragma unroll
            for (int g4 = 0; g4 < 2; ++g4) {
              const uint32_t keep = icka_rng::keep_bits4(icka_rng::effective_seed(args.seed, args.seed_base), icka_rng::kSiteAttention,
                                                         icka_rng::attn_group(drow, args.Skv, c0 + 8 * c + 4 * g4), args.drop_thresh);
#pragma unroll
              for (int j = 0; j < 4; ++j) p[4 * g4 + j] = (keep >> j & 1u) ? p[4 * g4 + j] * args.drop_scale : 0.0f;
            }
          }
          uint4 u;
          u.x = pack_bf16x2(p[0], p[1]);
          u.y = pack_bf16x2(p[2], p[3]);
          u.z = pack_bf16x2(p[4], p[5]);
          u.w = pack_bf16x2(p[6], p[7]);
          const int chunk = ((c0 & 63) >> 3) + c;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = u;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);

      // ---- epilogue ----
      mbar_wait(o_full, par);
      tc_fence_after();
      uint32_t orr[64];
      tmem_ld_32x32b_x32(tmem_o, *reinterpret_cast<uint32_t(*)[32]>(&orr[0]));
      tmem_ld_32x32b_x32(tmem_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&orr[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      const float inv = 1.0f / l;
      uint8_t* prow0 = p_buf + row * 128;      // first P chunk is free: P V has retired (o_full)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(orr[8 * c]) * inv, __uint_as_float(orr[8 * c + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(orr[8 * c + 2]) * inv, __uint_as_float(orr[8 * c + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(orr[8 * c + 4]) * inv, __uint_as_float(orr[8 * c + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(orr[8 * c + 6]) * inv, __uint_as_float(orr[8 * c + 7]) * inv);
        *reinterpret_cast<uint4*>(prow0 + ((c ^ (row & 7)) << 4)) = u;
      }
      __syncwarp();
      {
        const int cch = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = quad * 32 + i * 4 + (lane >> 3);
          const int q_row = qt * kRows + r;
          if (q_row < args.Sq) {
            const uint4 u = *reinterpret_cast<const uint4*>(p_buf + r * 128 + ((cch ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(args.ctx + ((size_t)b * args.Sq + q_row) * args.ldc + (size_t)h * kD + cch * 8) = u;
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Wide variant 2 (the default for 64 < Skv <= 224): TWO softmax groups of EIGHT warps alternate items (two warps per
// TMEM lane quadrant: a row's keys are split between two threads, which exchange the partial row maximum and sum through
// shared memory -- the row softmax is a chain of TMEM loads, exponentials and packs that four warps per scheduler hide
// far better than two), and the probabilities never touch shared memory: a thread overwrites its fp32 score row in TMEM with the bf16 probabilities (tcgen05.st, two
// per 32-bit column) and O = P V takes its A operand straight from tensor memory (tcgen05.mma with [a_tmem]).
//   TMEM  S/P slot g at columns [g * KEYS, (g + 1) * KEYS)  (P = the first KEYS / 2 columns of the slot), O at [2 KEYS, +64)
//   smem  two rings with different lifetimes: {Q 128 x 64, K KEYS x 64} x 2 -- dead as soon as QK^T has run, so the next
//         items' Q and K are always in place early -- and V KEYS x 64 x 3, held until P V has run; one 16 KB output
//         staging tile per group (row -> coalesced 16-byte stores).  (With one ring of whole {Q, K, V} stages, 72 KB each,
//         only three fit and the loads sat on the critical path: 2200 of the 9500 cycles of an item were spent waiting
//         for the scores, tools/attn_one.py with ICKA_ATTN_DEBUG=1.)
// The single O slot serialises only the short P V + read-out of consecutive items; QK^T of item n + 1 and the softmax
// of item n overlap, as do one group's softmax and the other's epilogue.
// ------------------------------------------------------------------------------------------------------------
template <int KEYS>
struct Wide2Cfg {
  static constexpr int kKVBytes = KEYS * 128;
  static constexpr int kQKBytes = kQBytes + kKVBytes;
  static constexpr int kQKStages = 2, kVStages = 3;
  static constexpr int kOutBytes = kRows * 128;             // 128 rows x 64 bf16
  static constexpr int kThreads = 64 + 512;                  // TMA warp, MMA warp, 2 groups x 8 softmax warps
  static constexpr int kSplit = ((KEYS / 32 + 1) / 2) * 32; // keys [0, kSplit) belong to the first warp of a quadrant pair
  static constexpr uint32_t kOCol = 2 * KEYS;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr int kXchBytes = 2 * 2 * 4 * 2 * 32 * 4;   // {max, sum} x group x quadrant x half x lane
  static constexpr size_t kSmemBytes = (size_t)kQKStages * kQKBytes + (size_t)kVStages * kKVBytes + 2 * kOutBytes +
                                       4 * KEYS * sizeof(float) + kXchBytes + 1024 + 256;
  static_assert(2 * KEYS + 64 <= 512, "two score slots and one output slot must fit the 512 TMEM columns");
  static_assert(KEYS % 32 == 0, "score rows are swept in 32-column pieces");
};

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[tmem: lane = row, two bf16 per 32-bit column] . B[smem], issued by ONE thread
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// developer timeline (ICKA_ATTN_DEBUG=1): cycle stamps of CTA 0, items [kTraceFrom, kTraceFrom + kTraceItems), read back and
// printed by the launcher after the kernel: [item][0..1] = QK issued, PV issued (MMA thread); [2..7] = group thread: start,
// S ready, max known, P written, O ready, done
constexpr int kTraceFrom = 8, kTraceItems = 16;
__device__ long long g_wide2_trace[kTraceItems][8];

template <int KEYS>
__global__ void __launch_bounds__(Wide2Cfg<KEYS>::kThreads, 1)
cross_attn_tcgen05_wide2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                                const __grid_constant__ CUtensorMap tmap_v, const AttnArgs args) {
  using Cfg = Wide2Cfg<KEYS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* qk_base = smem;                                                       // [kQKStages] {Q, K}
  uint8_t* v_base = smem + (size_t)Cfg::kQKStages * Cfg::kQKBytes;              // [kVStages] V
  uint8_t* out_buf = v_base + (size_t)Cfg::kVStages * Cfg::kKVBytes;            // [group] 16 KB
  float* mask_s = reinterpret_cast<float*>(out_buf + 2 * Cfg::kOutBytes);       // [group][2 buffers][KEYS]
  float* xch = mask_s + 4 * KEYS;                                               // [max | sum][group][quadrant][half][lane]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + Cfg::kXchBytes / 4);
  uint64_t* qk_full = bars;
  uint64_t* qk_empty = qk_full + Cfg::kQKStages;
  uint64_t* v_full = qk_empty + Cfg::kQKStages;
  uint64_t* v_empty = v_full + Cfg::kVStages;
  uint64_t* s_full = v_empty + Cfg::kVStages;   // [2] S = QK^T of the group's item is in its slot
  uint64_t* p_full = s_full + 2;                // [2] the group has replaced S by P
  uint64_t* o_full = p_full + 2;                // [2] O = P V of the group's item is complete
  uint64_t* o_free = o_full + 2;                // [2] the group has read O
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int items = args.B * args.nh * args.q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < Cfg::kQKStages; ++s) {
      mbar_init(&qk_full[s], 1);
      mbar_init(&qk_empty[s], 1);
    }
    for (int s = 0; s < Cfg::kVStages; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 8);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 8);
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
    if (lane == 0) {
      int sq = 0, sv = 0;
      uint32_t pq = 0, pv = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int qt = it % args.q_tiles, bh = it / args.q_tiles;
        const int h = bh % args.nh, b = bh / args.nh;
        uint8_t* qk = qk_base + (size_t)sq * Cfg::kQKBytes;
        mbar_wait(&qk_empty[sq], pq ^ 1);
        mbar_arrive_expect_tx(&qk_full[sq], Cfg::kQKBytes);
        tma_load_2d(qk, &tmap_q, &qk_full[sq], h * kD, b * args.Sq + qt * kRows);
        tma_load_2d(qk + kQBytes, &tmap_k, &qk_full[sq], h * kD, b * args.Skv);
        if (++sq == Cfg::kQKStages) { sq = 0; pq ^= 1; }
        mbar_wait(&v_empty[sv], pv ^ 1);
        mbar_arrive_expect_tx(&v_full[sv], Cfg::kKVBytes);
        tma_load_2d(v_base + (size_t)sv * Cfg::kKVBytes, &tmap_v, &v_full[sv], h * kD, b * args.Skv);
        if (++sv == Cfg::kVStages) { sv = 0; pv ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(kRows, KEYS);
      constexpr uint32_t idesc_o = make_idesc_bf16_f32(kRows, kD, false, true);
      int sq = 0;
      uint32_t pq = 0;
      int n = 0;
      auto issue_pv = [&](int m) {
        const int g = m & 1;
        const int sv = m % Cfg::kVStages;
        mbar_wait(&v_full[sv], (uint32_t)((m / Cfg::kVStages) & 1));
        mbar_wait(&p_full[g], (m >> 1) & 1);
        if (m >= 1) mbar_wait(&o_free[(m - 1) & 1], ((m - 1) >> 1) & 1);     // the previous item's O has been read
        tc_fence_after();
        const uint32_t p_tmem = tmem_base + (uint32_t)(g * KEYS);
        const uint32_t v_addr = smem_u32(v_base + (size_t)sv * Cfg::kKVBytes);
#pragma unroll
        for (int k = 0; k < KEYS / 16; ++k) {    // 16 keys = 8 packed TMEM columns of P, 16 rows of the in-place V tile
          // each half of a row's keys keeps its probabilities at the start of ITS OWN score columns
          const uint32_t pcol = (16 * k < Cfg::kSplit) ? (uint32_t)(8 * k) : (uint32_t)(Cfg::kSplit + 8 * k - Cfg::kSplit / 2);
          umma_bf16_ts(tmem_base + Cfg::kOCol, p_tmem + pcol, make_mnmajor_sw128_desc(v_addr + k * (16 * 128), Cfg::kKVBytes),
                       idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[sv]);
        umma_commit(&o_full[g]);
      };
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
        mbar_wait(&qk_full[sq], pq);
        tc_fence_after();
        // slot n & 1 is free: P V of item n - 2 was issued before this point and the tensor pipe runs in order
        const uint32_t q_addr = smem_u32(qk_base + (size_t)sq * Cfg::kQKBytes);
        const uint32_t k_addr = q_addr + kQBytes;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)((n & 1) * KEYS), make_kmajor_sw128_desc(q_addr + k * 32),
                    make_kmajor_sw128_desc(k_addr + k * 32), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&qk_empty[sq]);            // Q and K of this item are dead once the scores exist
        umma_commit(&s_full[n & 1]);
        const bool tr = args.debug && blockIdx.x == 0;
        if (tr && n >= kTraceFrom && n < kTraceFrom + kTraceItems) g_wide2_trace[n - kTraceFrom][0] = clock64();
        if (n >= 1) issue_pv(n - 1);
        if (tr && n - 1 >= kTraceFrom && n - 1 < kTraceFrom + kTraceItems) g_wide2_trace[n - 1 - kTraceFrom][1] = clock64();
        if (++sq == Cfg::kQKStages) { sq = 0; pq ^= 1; }
      }
      if (n >= 1) issue_pv(n - 1);
    }
  } else {
    const int sw = warp - 2;                           // 0..15
    const int g = sw >> 3;                             // softmax group == score slot
    const int half = (sw >> 2) & 1;                    // which part of the row's keys / of the output columns
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int gtid = (sw & 7) * 32 + lane;             // 0..255 inside the group
    const int key0 = half ? Cfg::kSplit : 0, key1 = half ? KEYS : Cfg::kSplit;
    const uint32_t tmem_s = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * KEYS + key0);
    const uint32_t tmem_o = tmem_base + ((uint32_t)(quad * 32) << 16) + Cfg::kOCol + (uint32_t)(32 * half);
    uint8_t* obuf = out_buf + g * Cfg::kOutBytes;
    float* gmask = mask_s + g * 2 * KEYS;
    float* xmax = xch + ((g * 4 + quad) * 2) * 32;                    // [half][lane]
    float* xsum = xch + 2 * 4 * 2 * 32 + ((g * 4 + quad) * 2) * 32;
    const int pair_bar = 3 + g * 4 + quad;             // named barrier of the two warps that share this quadrant's rows
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kScale = 0.125f * kLog2e;
    static_assert(KEYS <= 256, "one mask value per thread of a group");
    // additive key mask of an item: loaded into a register at the start of the PREVIOUS item of the group, written to the
    // other shared-memory buffer at its end (the global latency is off the critical path)
    auto load_mask = [&](int item) -> float {
      float v = -INFINITY;
      if (item < items && gtid < args.Skv) {
        const int b = item / (args.q_tiles * args.nh);
        v = args.mask_add ? __ldg(args.mask_add + (size_t)b * args.Skv + gtid) * kLog2e : 0.0f;
      }
      return v;
    };
    const int stride = 2 * (int)gridDim.x;
    float mreg = load_mask((int)blockIdx.x + g * (int)gridDim.x);
    if (gtid < KEYS) gmask[gtid] = mreg;
    const uint64_t scale2 = f2_splat(kScale);
    int m = 0;                                         // this group's item counter
    for (int it = blockIdx.x + g * gridDim.x; it < items; it += stride, ++m) {
      const uint32_t par = m & 1;
      const int qt = it % args.q_tiles, bh = it / args.q_tiles;
      const int h = bh % args.nh, b = bh / args.nh;
      const float* mk = gmask + (m & 1) * KEYS + key0;
      const bool trace = args.debug && blockIdx.x == 0 && quad == 0 && half == 0 && lane == 0 && 2 * m + g >= kTraceFrom &&
                         2 * m + g < kTraceFrom + kTraceItems;
      long long tk[6];
      if (trace) tk[0] = clock64();
      named_bar_sync(1 + g, 256);            // the mask of this item is in place (and the other buffer is free)
      mreg = load_mask(it + stride);
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      if (trace) tk[1] = clock64();
      // ---- sweep 1: maximum over this thread's part of the row (scores scaled and masked two at a time, FFMA2) ----
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c0 = 0; c0 < key1 - key0; c0 += 32) {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(tmem_s + (uint32_t)c0, sr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 mq = *reinterpret_cast<const float4*>(mk + c0 + j);
          float t0, t1, t2, t3;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(sr[j]), __uint_as_float(sr[j + 1])), scale2, f2_pack(mq.x, mq.y)), t0, t1);
          f2_unpack(f2_fma(f2_pack(__uint_as_float(sr[j + 2]), __uint_as_float(sr[j + 3])), scale2, f2_pack(mq.z, mq.w)), t2, t3);
          mx4[(j >> 2) & 3] = fmaxf(fmaxf(mx4[(j >> 2) & 3], fmaxf(t0, t1)), fmaxf(t2, t3));
        }
      }
      xmax[half * 32 + lane] = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      named_bar_sync(pair_bar, 64);
      const float mx = fmaxf(xmax[lane], xmax[32 + lane]);
      if (trace) tk[2] = clock64();
      // ---- sweep 2: P = exp2(s - max) as bf16 pairs over the consumed part of this thread's score columns ----
      const uint64_t nmx2 = f2_splat(-mx);
      uint64_t l2[2] = {f2_splat(0.0f), f2_splat(0.0f)};
      const uint64_t drow = ((uint64_t)b * args.nh + h) * (uint64_t)args.Sq + (uint64_t)(qt * kRows + row);
#pragma unroll 1
      for (int c0 = 0; c0 < key1 - key0; c0 += 32) {
        uint32_t sr[32];
        tmem_ld_32x32b_x32(tmem_s + (uint32_t)c0, sr);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {      // 8 keys at a time
          const float4 ma = *reinterpret_cast<const float4*>(mk + c0 + 8 * c);
          const float4 mb = *reinterpret_cast<const float4*>(mk + c0 + 8 * c + 4);
          float p[8];
          f2_unpack(f2_add(f2_fma(f2_pack(__uint_as_float(sr[8 * c]), __uint_as_float(sr[8 * c + 1])), scale2, f2_pack(ma.x, ma.y)), nmx2), p[0], p[1]);
          f2_unpack(f2_add(f2_fma(f2_pack(__uint_as_float(sr[8 * c + 2]), __uint_as_float(sr[8 * c + 3])), scale2, f2_pack(ma.z, ma.w)), nmx2), p[2], p[3]);
          f2_unpack(f2_add(f2_fma(f2_pack(__uint_as_float(sr[8 * c + 4]), __uint_as_float(sr[8 * c + 5])), scale2, f2_pack(mb.x, mb.y)), nmx2), p[4], p[5]);
          f2_unpack(f2_add(f2_fma(f2_pack(__uint_as_float(sr[8 * c + 6]), __uint_as_float(sr[8 * c + 7])), scale2, f2_pack(mb.z, mb.w)), nmx2), p[6], p[7]);
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = ex2(p[j]);
          l2[0] = f2_add(l2[0], f2_add(f2_pack(p[0], p[1]), f2_pack(p[2], p[3])));
          l2[1] = f2_add(l2[1], f2_add(f2_pack(p[4], p[5]), f2_pack(p[6], p[7])));
          if (args.drop_thresh) {
#pragma unroll
            for (int g4 = 0; g4 < 2; ++g4) {
              const uint32_t keep = icka_rng::keep_bits4(icka_rng::effective_seed(args.seed, args.seed_base), icka_rng::kSiteAttention,
                                                         icka_rng::attn_group(drow, args.Skv, key0 + c0 + 8 * c + 4 * g4), args.drop_thresh);
#pragma unroll
              for (int j = 0; j < 4; ++j) p[4 * g4 + j] = (keep >> j & 1u) ? p[4 * g4 + j] * args.drop_scale : 0.0f;
            }
          }
          pk[4 * c] = pack_bf16x2(p[0], p[1]);
          pk[4 * c + 1] = pack_bf16x2(p[2], p[3]);
          pk[4 * c + 2] = pack_bf16x2(p[4], p[5]);
          pk[4 * c + 3] = pack_bf16x2(p[6], p[7]);
        }
        // packed columns [c0/2, c0/2 + 16) of this thread's part end at or below c0 + 32: those scores are consumed
        tmem_st_32x32b_x16(tmem_s + (uint32_t)(c0 >> 1), pk);
      }
      {
        float la, lb, lc, ld;
        f2_unpack(l2[0], la, lb);
        f2_unpack(l2[1], lc, ld);
        xsum[half * 32 + lane] = (la + lb) + (lc + ld);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      if (trace) tk[3] = clock64();

      // ---- epilogue: this warp's 32 of the 64 output columns ----
      mbar_wait(&o_full[g], par);
      tc_fence_after();
      if (trace) tk[4] = clock64();
      uint32_t orr[32];
      tmem_ld_32x32b_x32(tmem_o, orr);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[g]);
      named_bar_sync(pair_bar, 64);                    // both partial row sums are visible; the staging rows are free
      const float inv = 1.0f / (xsum[lane] + xsum[32 + lane]);
      uint8_t* prow0 = obuf + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(orr[8 * c]) * inv, __uint_as_float(orr[8 * c + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(orr[8 * c + 2]) * inv, __uint_as_float(orr[8 * c + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(orr[8 * c + 4]) * inv, __uint_as_float(orr[8 * c + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(orr[8 * c + 6]) * inv, __uint_as_float(orr[8 * c + 7]) * inv);
        *reinterpret_cast<uint4*>(prow0 + (((4 * half + c) ^ (row & 7)) << 4)) = u;
      }
      named_bar_sync(pair_bar, 64);                    // the quadrant's 32 rows are complete: each warp copies 16 of them out
      {
        const int cch = lane & 7;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = quad * 32 + (4 * half + i) * 4 + (lane >> 3);
          const int q_row = qt * kRows + r;
          if (q_row < args.Sq) {
            const uint4 u = *reinterpret_cast<const uint4*>(obuf + r * 128 + ((cch ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(args.ctx + ((size_t)b * args.Sq + q_row) * args.ldc + (size_t)h * kD + cch * 8) = u;
          }
        }
      }
      if (gtid < KEYS) gmask[((m & 1) ^ 1) * KEYS + gtid] = mreg;
      if (trace) {
        tk[5] = clock64();
#pragma unroll
        for (int i = 0; i < 6; ++i) g_wide2_trace[2 * m + g - kTraceFrom][2 + i] = tk[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int KEYS>
int launch_wide2(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const AttnArgs& args,
                 cudaStream_t st) {
  using Cfg = Wide2Cfg<KEYS>;
  if (h->smem_optin < Cfg::kSmemBytes) return 1;
  CUtensorMap tq, tk, tv;
  int rc = icka_make_tmap_bf16(h, &tq, q, (int64_t)args.B * args.Sq, (int64_t)args.nh * kD, ldq, kRows);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tk, k, (int64_t)args.B * args.Skv, (int64_t)args.nh * kD, ldkv, KEYS);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tv, v, (int64_t)args.B * args.Skv, (int64_t)args.nh * kD, ldkv, KEYS);
  if (rc) return rc;
  const int items = args.B * args.nh * args.q_tiles;
  ICKA_CUDA(cudaFuncSetAttribute(cross_attn_tcgen05_wide2_kernel<KEYS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg::kSmemBytes));
  const int grid = items < h->sm_count ? items : h->sm_count;
  cross_attn_tcgen05_wide2_kernel<KEYS><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tq, tk, tv, args);
  ICKA_LAUNCHED(h);
  if (args.debug) {      // developer timeline: synchronises, prints cycles relative to the first traced stamp
    long long tr[kTraceItems][8];
    ICKA_CUDA(cudaStreamSynchronize(st));
    ICKA_CUDA(cudaMemcpyFromSymbol(tr, g_wide2_trace, sizeof(tr)));
    const long long t0 = tr[0][2];
    fprintf(stderr, "wide2 timeline (cycles, CTA 0): item grp | QK issued  PV issued | start  S ready  max  P written  O ready  done\n");
    for (int i = 0; i < kTraceItems; ++i)
      fprintf(stderr, "  %3d  g%d | %8lld %8lld | %8lld %8lld %8lld %8lld %8lld %8lld\n", i + kTraceFrom, (i + kTraceFrom) & 1,
              tr[i][0] - t0, tr[i][1] - t0, tr[i][2] - t0, tr[i][3] - t0, tr[i][4] - t0, tr[i][5] - t0, tr[i][6] - t0, tr[i][7] - t0);
  }
  return ICKA_OK;
}

template <int KEYS>
int launch_wide(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const AttnArgs& args,
                cudaStream_t st) {
  using Cfg = WideCfg<KEYS>;
  if (h->smem_optin < Cfg::kSmemBytes) return 1;
  CUtensorMap tq, tk, tv;
  int rc = icka_make_tmap_bf16(h, &tq, q, (int64_t)args.B * args.Sq, (int64_t)args.nh * kD, ldq, kRows);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tk, k, (int64_t)args.B * args.Skv, (int64_t)args.nh * kD, ldkv, KEYS);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tv, v, (int64_t)args.B * args.Skv, (int64_t)args.nh * kD, ldkv, KEYS);
  if (rc) return rc;
  const int items = args.B * args.nh * args.q_tiles;
  ICKA_CUDA(cudaFuncSetAttribute(cross_attn_tcgen05_wide_kernel<KEYS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)Cfg::kSmemBytes));
  const int grid = items < h->sm_count ? items : h->sm_count;
  cross_attn_tcgen05_wide_kernel<KEYS><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tq, tk, tv, args);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

}  // namespace

// Returns ICKA_OK after launching, or a positive value when the shape is outside this kernel's envelope
// (the caller then uses the mma.sync kernel).
int icka_attn_tcgen05_launch(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                             const float* mask_add, void* ctx, int64_t ldc, int B, int Sq, int Skv, int nh,
                             uint32_t drop_thresh, float drop_scale, uint64_t seed, const unsigned long long* seed_base,
                             cudaStream_t st) {
  extern int g_attn_mode;
  if (Skv > kKeys) {
    // 64 < Skv <= 224 (the 196-region grid): two softmax groups with the probabilities in tensor memory (wide2) by
    // default; icka_set_attn_mode(2) selects the first wide variant (one group, P through shared memory: no faster than
    // the mma.sync kernel, 0.46 vs 0.48 ms at 512 hi-res sentences; wide2: 0.28 ms)
    // (a single query row per sentence -- the image->text encoders in fp32 / training mode -- would fill 1 of the 128 rows
    // of an item: the mma.sync kernel is faster there, 122 vs 160 us at 1024 sentences)
    if (Skv > 224 || (Sq < 32 && g_attn_mode == 0)) return 1;
    AttnArgs wargs{mask_add, static_cast<__nv_bfloat16*>(ctx), ldc, B, Sq, Skv, nh, (Sq + kRows - 1) / kRows,
                   drop_thresh, drop_scale, seed, seed_base, getenv("ICKA_ATTN_DEBUG") ? 1 : 0};
    if (g_attn_mode == 2)
      return Skv <= 128 ? launch_wide<128>(h, q, ldq, k, v, ldkv, wargs, st) : launch_wide<224>(h, q, ldq, k, v, ldkv, wargs, st);
    return Skv <= 128 ? launch_wide2<128>(h, q, ldq, k, v, ldkv, wargs, st) : launch_wide2<224>(h, q, ldq, k, v, ldkv, wargs, st);
  }
  if (h->smem_optin < kSmemBytes) return 1;
  CUtensorMap tq, tk, tv;
  int rc = icka_make_tmap_bf16(h, &tq, q, (int64_t)B * Sq, (int64_t)nh * kD, ldq, kRows);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tk, k, (int64_t)B * Skv, (int64_t)nh * kD, ldkv, kKeys);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tv, v, (int64_t)B * Skv, (int64_t)nh * kD, ldkv, kKeys);
  if (rc) return rc;
  AttnArgs args{mask_add, static_cast<__nv_bfloat16*>(ctx), ldc, B, Sq, Skv, nh, (Sq + kRows - 1) / kRows,
                drop_thresh, drop_scale, seed, seed_base, 0};
  const int items = B * nh * args.q_tiles;
  ICKA_CUDA(cudaFuncSetAttribute(cross_attn_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const int grid = items < h->sm_count ? items : h->sm_count;
  cross_attn_tcgen05_kernel<<<grid, kThreads, kSmemBytes, st>>>(tq, tk, tv, args);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
