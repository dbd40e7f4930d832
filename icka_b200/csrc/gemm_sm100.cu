// bf16 tensor-core path of icka_linear_fwd: a persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) (+ residual[M,N])       A, W bf16 (K-major), fp32 accumulate
//
// This is the kernel behind every dense contraction of the fusion path (>99 % of its FLOPs):
// region projection (CMIM:958), Q and K|V projections (CMIM:592-594), attention out-proj (CMIM:562),
// FFN up + erf-GELU (CMIM:549-550) and FFN down (CMIM:533).  nn.Linear weights are [out, in] row-major,
// i.e. already the K-major B operand; activations are [rows, in] row-major = K-major A.
//
// Structure (one CTA per SM, 320 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor 2-D boxes {64 x 128} of A and {64 x BN} of W into a
//              kStages-deep shared-memory ring, 128-byte swizzle, completion on "full" mbarriers
//   warp 1     allocates TMEM, then one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x 4 per
//              k-block; tcgen05.commit releases the smem slot ("empty") and, after the last k-block,
//              publishes the accumulator ("tmem_full")
//   warps 2-9  epilogue: tcgen05.ld the 128 x BN fp32 accumulator (lane = row) 32 columns at a time,
//              transpose each 32x32 block through a swizzled shared-memory tile so that global traffic is
//              row-contiguous (lane = column pair): coalesced bias / fp32-residual loads and stores, with
//              erf-GELU and the bf16 conversion applied on the way; the accumulator is handed back
//              ("tmem_empty") as soon as its last column block has been read
// Two accumulators (2 x BN TMEM columns) let the epilogue of tile i overlap the mainloop of tile i+1.
// Tiles are walked n-fastest so the CTAs that share an A row-block run together and A is read from
// HBM once (weights are a few MB and stay in the 126 MB L2).
// Ragged edges: TMA zero-fills out-of-bounds rows/columns (M, N and K tails), the epilogue masks
// rows >= M and columns >= N.
//
// Operand layouts (template parameter MAJOR) -- the same pipeline serves the backward GEMMs of training
// without transposed copies of weights or activations, by switching the UMMA descriptors to MN-major:
//   MAJOR 0  forward   out = A[M,K] . W[N,K]^T          A, W contraction-contiguous (K-major)
//   MAJOR 1  dgrad     dX  = dY[M,K'] . W[K',N]         second operand read as stored by nn.Linear ([out,in]):
//                                                        its N (output) index is contiguous -> MN-major B
//   MAJOR 2  wgrad     dW += dY[T,M]^T . X[T,N]         both operands token-major: MN-major A and B, the token
//                                                        range is split over CTAs (split-K) and partial tiles
//                                                        are added to dW with fp32 red.global
// MN-major tiles are fetched by one 3-D TMA box {64 elements, 64 contraction rows, tile/64 chunks} each.
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <stdlib.h>

namespace {

using namespace sm100;

constexpr int kBM = 128;
constexpr int kBK = 64;            // 64 bf16 = 128 B = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;           // two per TMEM lane quadrant; they split the column blocks
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
__device__ __forceinline__ void tma_store_2d_out(const CUtensorMap* tm, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

constexpr int kActRuntime = 100;   // template ACT value: relu / swish (config.hidden_act, CMIM:43) chosen by GemmArgs::act_rt

// CTAS == 1: one CTA computes a 128 x BN tile.  CTAS == 2: a CTA pair (cluster of 2, tcgen05 cta_group::2)
// computes a 256 x BN tile; each CTA stages its own 128 rows of A and HALF of the weight tile (BN/2 rows),
// so the same shared memory holds 6 stages instead of 4 and the weight tile crosses the L2->SM fabric once
// per pair.  The leader CTA (cluster rank 0) issues the MMAs for both.
template <int BN, int CTAS = 1>
struct GemmCfg {
  static constexpr int kStages = (BN == 256 && CTAS == 1) ? 4 : 6;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / CTAS) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;   // two accumulators (per CTA: 128 lanes x BN columns each)
  static constexpr int kStagingBytes = kEpiWarps * 32 * 128;   // one 32 x 32 fp32 transpose tile per epilogue warp
  static constexpr size_t kSmemBytes =
      (size_t)kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmArgs {
  const float* bias;
  const float* residual;   // [M, N] fp32 or null
  void* out;
  int64_t ldo;
  int M, N, K;
  int debug;   // 0 = normal; developer probes: 1 = epilogue only releases the accumulator, 2 = no global stores
  const __nv_bfloat16* aux_in;   // ACT_GELU_ERF_BWD: pre-activation u [M, N] (pitch ld_aux)
  __nv_bfloat16* aux_out;        // ACT_GELU_ERF: optional copy of the pre-activation (acc + bias) [M, N]
  int64_t ld_aux;
  int splits;                    // number of contraction ranges (split-K); 1 otherwise
  int kb_per_split;
  int64_t split_stride;          // forward split-K: split s writes its partial tile to out + s * split_stride
  // LayerNorm-fused epilogue (LNF): out = LN(acc + bias + residual) in fp32 plus a bf16 copy
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  __nv_bfloat16* out16;          // [M, N] bf16 copy of the normalised rows (pitch N) or null
  int act_rt;                    // ACT == kActRuntime: ICKA_ACT_RELU or ICKA_ACT_SWISH
  // bf16 outputs without residual / aux copy (Q, K||V, FFN-up): the epilogue packs rows straight out of TMEM and a TMA store
  // per 32 x 32 block writes them (tmap_out: box {32 columns, 32 rows}, SWIZZLE_64B); 0 = the transposing epilogue
  int tma_epi;
  alignas(64) CUtensorMap tmap_out;
};

constexpr int kChunkBytes = kBK * 128;   // one 64-element-wide MN-major chunk of a stage: 64 contraction rows x 128 B

// Tile walk.  Default: tile t = group + i * groups, n fastest, so the CTAs that share an A row-block run together.
// LNF (LayerNorm-fused epilogue): a CTA owns WHOLE output rows -- it walks all n-tiles of row-block
// group + j * groups before moving on -- because the row statistics span every n-tile.
template <bool LNF>
__device__ __forceinline__ bool tile_at(int i, int group_id, int num_groups, int m_tiles, int n_tiles, int num_tiles,
                                        int& mn, int& split) {
  if (LNF) {
    const int mb = group_id + (i / n_tiles) * num_groups;
    if (mb >= m_tiles) return false;
    mn = mb * n_tiles + i % n_tiles;
    split = 0;
    return true;
  }
  const int tile = group_id + i * num_groups;
  if (tile >= num_tiles) return false;
  const int mn_tiles = m_tiles * n_tiles;
  mn = tile % mn_tiles;
  split = tile / mn_tiles;
  return true;
}

template <int BN, int ACT, bool OUT_BF16, int CTAS, int MAJOR, bool LNF = false, bool TMAE = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ GemmArgs args) {
  static_assert(!LNF || (MAJOR == 0 && CTAS == 1 && !OUT_BF16 && ACT == ICKA_ACT_NONE), "LNF: plain fp32 forward only");
  static_assert(ACT != kActRuntime || MAJOR == 0, "runtime activations are forward-only");
  static_assert(!TMAE || (OUT_BF16 && MAJOR == 0 && !LNF), "TMA-store epilogue: bf16 forward outputs only");
  static_assert(MAJOR == 0 || CTAS == 1, "MN-major operands are built for single-CTA tiles only");
  static_assert(MAJOR != 2 || (!OUT_BF16 && ACT == ICKA_ACT_NONE), "wgrad accumulates plain fp32");
  using Cfg = GemmCfg<BN, CTAS>;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const int group_id = (CTAS == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // tile-walking unit
  const int num_groups = (CTAS == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kTileM = kBM * CTAS;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)Cfg::kStages * Cfg::kABytes;
  uint8_t* smem_stage = smem + (size_t)Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int M = args.M, N = args.N, K = args.K;
  const int m_tiles = (M + kTileM - 1) / kTileM;
  const int n_tiles = (N + BN - 1) / BN;
  const int mn_tiles = m_tiles * n_tiles;
  const int num_tiles = mn_tiles * args.splits;   // split-major: concurrent CTAs share a contraction range
  const int num_kb = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kEpiWarps * CTAS);   // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    if (CTAS == 2) {
      tmem_alloc_pair(tmem_base_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();   // peer barriers must be initialised before use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0;; ++i) {
        int mn, split;
        if (!tile_at<LNF>(i, group_id, num_groups, m_tiles, n_tiles, num_tiles, mn, split)) break;
        const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
        const int a_row = m_blk * kTileM + (int)cta_rank * kBM;
        const int b_row = n_blk * BN + (int)cta_rank * (BN / CTAS);
        const int kb0 = split * args.kb_per_split, kb1 = min(num_kb, kb0 + args.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (CTAS == 2) {
            // both CTAs' bytes are credited to the LEADER's full barrier, which it arms for the pair
            const uint32_t full_leader = map_to_cta(&full_bar[stage], 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * 2);
            tma_load_2d_pair(smem_a + (size_t)stage * Cfg::kABytes, &tmap_a, full_leader, kb * kBK, a_row);
            tma_load_2d_pair(smem_b + (size_t)stage * Cfg::kBBytes, &tmap_b, full_leader, kb * kBK, b_row);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            if (MAJOR == 2)
              tma_load_3d(smem_a + (size_t)stage * Cfg::kABytes, &tmap_a, &full_bar[stage], 0, kb * kBK, a_row / 64);
            else
              tma_load_2d(smem_a + (size_t)stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * kBK, a_row);
            if (MAJOR >= 1)
              tma_load_3d(smem_b + (size_t)stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], 0, kb * kBK, b_row / 64);
            else
              tma_load_2d(smem_b + (size_t)stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * kBK, b_row);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread; the leader CTA of a pair) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(kTileM, BN, MAJOR == 2, MAJOR >= 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (;; ++it) {
        int mn, split;
        if (!tile_at<LNF>(it, group_id, num_groups, m_tiles, n_tiles, num_tiles, mn, split)) break;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // epilogue(s) have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        const int kb0 = split * args.kb_per_split, kb1 = min(num_kb, kb0 + args.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);             // TMA bytes (of both CTAs) have landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + (size_t)stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // K-major: 16 contraction elements = 32 B inside the 128-B swizzle row;
            // MN-major: 16 contraction rows = 2 swizzle atoms of 8 rows x 128 B
            const uint64_t da = (MAJOR == 2) ? make_mnmajor_sw128_desc(a_addr + k * (kUmmaK * 128), kChunkBytes)
                                             : make_kmajor_sw128_desc(a_addr + k * (kUmmaK * 2));
            const uint64_t db = (MAJOR >= 1) ? make_mnmajor_sw128_desc(b_addr + k * (kUmmaK * 128), kChunkBytes)
                                             : make_kmajor_sw128_desc(b_addr + k * (kUmmaK * 2));
            if (CTAS == 2) umma_bf16_pair(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot free (in both CTAs) once these MMAs retire
          if (CTAS == 2) umma_commit_pair(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete: wake the epilogue warps of both CTAs
        if (CTAS == 2) umma_commit_pair(&tmem_full_bar[acc], 3); else umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue (8 warps: 4 lane quadrants x 2 column halves) =====================
    const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                     // even / odd 32-column blocks
    uint8_t* stage_tile = smem_stage + (warp - 2) * (32 * 128);
    const int sub = lane >> 4;                            // row parity handled in the transposed phase
    const int cp = lane & 15;                             // column pair handled in the transposed phase
    // Staging tile: 32 rows x 128 B; the 16-byte slot s of row r is stored at slot s ^ (r & 7).
    //   write side (lane = row):  slot offsets for s = 0..7
    //   read side  (row = 2i+sub, 8 bytes at column pair cp): (2i+sub)&7 == (2(i&3)) ^ sub, so the slot is
    //   ((cp>>1) ^ sub) ^ 2(i&3): four lane-dependent base pointers, everything else is an immediate.
    uint8_t* wr_row = stage_tile + lane * 128;
    int wr_off[8];
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) wr_off[sl] = (sl ^ (lane & 7)) << 4;
    const uint8_t* rd_base[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      rd_base[k] = stage_tile + sub * 128 + ((((cp >> 1) ^ sub) ^ (2 * k)) << 4) + ((cp & 1) << 3);

    int it = 0;
    const uint32_t empty_leader[2] = {CTAS == 2 ? map_to_cta(&tmem_empty_bar[0], 0) : 0u,
                                      CTAS == 2 ? map_to_cta(&tmem_empty_bar[1], 0) : 0u};
    int tma_blocks = 0;     // TMA-store epilogue: blocks this warp has issued (staging half = parity)
    float rs[16], rq[16];   // LNF: per-thread partial row sums / sums of squares of rows 2i+sub over this warp's columns
    // LNF keeps its pre-LayerNorm rows in L2 between the two passes: they are written and re-read with evict_last,
    // everything that streams (residual reads, final fp32 / bf16 stores) goes through with evict_first
    const uint64_t pol_keep = LNF ? l2_policy_evict_last() : 0, pol_stream = LNF ? l2_policy_evict_first() : 0;
    for (;; ++it) {
      int mn, split_idx;
      if (!tile_at<LNF>(it, group_id, num_groups, m_tiles, n_tiles, num_tiles, mn, split_idx)) break;
      const int m_blk = mn / n_tiles, n_blk = mn % n_tiles;
      if (LNF && n_blk == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) rs[i] = rq[i] = 0.0f;
      }
      const bool lead = args.splits == 1 || MAJOR == 2;   // split-K forward: bias / residual are added by the reduce pass
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile0 = n_blk * BN;
      const int nchunks = min(BN / 32, (N - n_tile0 + 31) / 32);
      const int row_base = m_blk * kTileM + (int)cta_rank * kBM + quad * 32;
      const int rows_left = M - row_base - sub;           // row 2i+sub of this warp's block is valid iff 2i < rows_left
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      // Forward split-K: every split stores its partial tile to its own slab of the workspace (plain stores);
      // splitk_reduce_kernel then adds the slabs in split order + bias + residual, so results are reproducible.
      const bool slabs = (MAJOR != 2) && args.splits > 1;
      uint32_t r[32];
      bool released = false;
      if constexpr (TMAE) {
        {
          // ---- bf16 rows straight out of TMEM (thread = row, 32 consecutive columns): bias (+ activation), pack, four
          //      16-byte shared-memory stores into a 32 x 64 B tile (SWIZZLE_64B: chunk ^= (row >> 1) & 3, conflict-free
          //      for lane = row) and ONE TMA store per block -- no transpose through shared memory, no per-thread global
          //      stores; the two 2 KB halves of the warp's staging tile alternate, so a store drains while the next
          //      block is computed.  TMA clips rows >= M.  (host: N % 32 == 0, no residual / aux copy / split-K)
          for (int c = half; c < nchunks; c += 2) {
            tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            if (c + 2 >= nchunks) {   // this warp's last TMEM read of the tile: hand the accumulator back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(empty_leader[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
              released = true;
            }
            const int col0 = n_tile0 + c * 32;
            uint32_t pk[16];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4 bb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
              if (args.bias) bb = __ldg(reinterpret_cast<const float4*>(args.bias + col0) + j4);   // warp-uniform address
              float x0 = __uint_as_float(r[4 * j4]) + bb.x, x1 = __uint_as_float(r[4 * j4 + 1]) + bb.y;
              float x2 = __uint_as_float(r[4 * j4 + 2]) + bb.z, x3 = __uint_as_float(r[4 * j4 + 3]) + bb.w;
              if (ACT == ICKA_ACT_GELU_ERF) {
                const float2 g0 = gelu_erf_tanh2(x0, x1), g1 = gelu_erf_tanh2(x2, x3);
                x0 = g0.x; x1 = g0.y; x2 = g1.x; x3 = g1.y;
              }
              if (ACT == kActRuntime) {
                if (args.act_rt == ICKA_ACT_RELU) {
                  x0 = act_relu(x0); x1 = act_relu(x1); x2 = act_relu(x2); x3 = act_relu(x3);
                } else {
                  x0 = act_swish_fast(x0); x1 = act_swish_fast(x1); x2 = act_swish_fast(x2); x3 = act_swish_fast(x3);
                }
              }
              if (ACT == ICKA_ACT_TANH) {
                asm("tanh.approx.f32 %0, %1;" : "=f"(x0) : "f"(x0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(x1) : "f"(x1));
                asm("tanh.approx.f32 %0, %1;" : "=f"(x2) : "f"(x2));
                asm("tanh.approx.f32 %0, %1;" : "=f"(x3) : "f"(x3));
              }
              pk[2 * j4] = pack_bf16x2(x0, x1);
              pk[2 * j4 + 1] = pack_bf16x2(x2, x3);
            }
            uint8_t* buf = stage_tile + (tma_blocks & 1) * 2048;
            if (lane == 0 && tma_blocks >= 2) bulk_wait_group_read1();   // the store that used this half has read it
            __syncwarp();
            uint8_t* rowp = buf + lane * 64;
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int sl = 0; sl < 4; ++sl)
              *reinterpret_cast<uint4*>(rowp + ((sl ^ sw) << 4)) = make_uint4(pk[4 * sl], pk[4 * sl + 1], pk[4 * sl + 2], pk[4 * sl + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d_out(&args.tmap_out, buf, col0, row_base);
              bulk_commit_group();
            }
            ++tma_blocks;
          }
          if (!released) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(empty_leader[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
          }
          continue;
        }
      }
      if (args.debug == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(empty_leader[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
        continue;
      }
      if (half < nchunks) tmem_ld_32x32b_x32(taddr0 + (uint32_t)(half * 32), r);
#pragma unroll 1
      for (int c = half; c < nchunks; c += 2) {
        tmem_ld_wait();
#pragma unroll
        for (int sl = 0; sl < 8; ++sl)
          *reinterpret_cast<uint4*>(wr_row + wr_off[sl]) = make_uint4(r[4 * sl], r[4 * sl + 1], r[4 * sl + 2], r[4 * sl + 3]);
        __syncwarp();
        if (c + 2 < nchunks) {
          tmem_ld_32x32b_x32(taddr0 + (uint32_t)((c + 2) * 32), r);   // overlaps the phase below
        } else {
          // this warp's last TMEM read has landed: hand the accumulator back before the global stores
          tc_fence_before();
          if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(empty_leader[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
          released = true;
        }
        const int col = n_tile0 + c * 32 + 2 * cp;
        const bool col_ok = col < N && args.debug != 2;
        float b0 = 0.0f, b1 = 0.0f;
        if (args.bias && col_ok && lead) {
          const float2 bb = __ldg(reinterpret_cast<const float2*>(args.bias + col));
          b0 = bb.x;
          b1 = bb.y;
        }
        const bool full = (rows_left >= 32) && (n_tile0 + c * 32 + 32 <= N) && args.debug != 2;   // warp-uniform
        const size_t row0 = (size_t)(row_base + sub);
        float2 res[16];
        const bool use_res = args.residual != nullptr && lead;
        if (use_res) {
          const float* rp = args.residual + row0 * N + col;
          const size_t rstep = (size_t)2 * N;
#pragma unroll
          for (int i = 0; i < 16; ++i, rp += rstep)
            res[i] = (full || (col_ok && 2 * i < rows_left))
                         ? (LNF ? ldg_f2_hint(rp, pol_stream) : __ldg(reinterpret_cast<const float2*>(rp)))
                         : make_float2(0.0f, 0.0f);
        }
        uint32_t aux[16];
        if (ACT == ICKA_ACT_GELU_ERF_BWD) {
          const __nv_bfloat16* ap = args.aux_in + row0 * args.ld_aux + col;
          const size_t astep = (size_t)2 * args.ld_aux;
#pragma unroll
          for (int i = 0; i < 16; ++i, ap += astep)
            aux[i] = (full || (col_ok && 2 * i < rows_left)) ? __ldg(reinterpret_cast<const uint32_t*>(ap)) : 0u;
        }
        float2 x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = *reinterpret_cast<const float2*>(rd_base[i & 3] + i * 256);
        __syncwarp();   // staging tile may be rewritten (next column block) once every lane has read it
        if ((ACT == ICKA_ACT_GELU_ERF || ACT == kActRuntime) && args.aux_out != nullptr) {   // training: keep the pre-activation
          __nv_bfloat16* up = args.aux_out + row0 * args.ld_aux + col;
          const size_t ustep = (size_t)2 * args.ld_aux;
#pragma unroll
          for (int i = 0; i < 16; ++i, up += ustep)
            if (full || (col_ok && 2 * i < rows_left))
              *reinterpret_cast<uint32_t*>(up) = pack_bf16x2(x[i].x + b0, x[i].y + b1);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0 = x[i].x + b0, x1 = x[i].y + b1;
          if (ACT == ICKA_ACT_GELU_ERF_BWD) {
            x0 *= gelu_erf_grad_fast(__uint_as_float(aux[i] << 16));
            x1 *= gelu_erf_grad_fast(__uint_as_float(aux[i] & 0xffff0000u));
          }
          if (ACT == ICKA_ACT_GELU_ERF) {
            if (OUT_BF16) {
              const float2 gg = gelu_erf_tanh2(x0, x1);
              x0 = gg.x;
              x1 = gg.y;
            } else {
              x0 = gelu_erf(x0);
              x1 = gelu_erf(x1);
            }
          }
          if (ACT == kActRuntime) {
            if (args.act_rt == ICKA_ACT_RELU) {
              x0 = act_relu(x0);
              x1 = act_relu(x1);
            } else if (OUT_BF16) {
              x0 = act_swish_fast(x0);
              x1 = act_swish_fast(x1);
            } else {
              x0 = act_swish(x0);
              x1 = act_swish(x1);
            }
          }
          if (ACT == ICKA_ACT_TANH) {
            if (OUT_BF16) {   // bf16 output: MUFU.TANH (2^-11 relative) is below the rounding of the store
              asm("tanh.approx.f32 %0, %1;" : "=f"(x0) : "f"(x0));
              asm("tanh.approx.f32 %0, %1;" : "=f"(x1) : "f"(x1));
            } else {
              x0 = tanhf(x0);
              x1 = tanhf(x1);
            }
          }
          if (use_res) {
            x0 += res[i].x;
            x1 += res[i].y;
          }
          x[i] = make_float2(x0, x1);
          if (LNF && col_ok) {   // columns beyond N hold zeros of the TMA fill + nothing: keep them out of the statistics
            rs[i] += x0 + x1;
            rq[i] = fmaf(x0, x0, fmaf(x1, x1, rq[i]));
          }
        }
        const size_t ostep = (size_t)2 * args.ldo;
        if (OUT_BF16) {
          __nv_bfloat16* op = static_cast<__nv_bfloat16*>(args.out) + row0 * args.ldo + col;
#pragma unroll
          for (int i = 0; i < 16; ++i, op += ostep)
            if (full || (col_ok && 2 * i < rows_left)) *reinterpret_cast<uint32_t*>(op) = pack_bf16x2(x[i].x, x[i].y);
        } else if (MAJOR == 2) {   // wgrad: partial tiles are added onto the running gradient with red.global
          float* op = static_cast<float*>(args.out) + row0 * args.ldo + col;
#pragma unroll
          for (int i = 0; i < 16; ++i, op += ostep)
            if (full || (col_ok && 2 * i < rows_left)) atomicAdd(reinterpret_cast<float2*>(op), x[i]);
        } else if (slabs) {
          float* op = static_cast<float*>(args.out) + (size_t)split_idx * args.split_stride + row0 * args.ldo + col;
#pragma unroll
          for (int i = 0; i < 16; ++i, op += ostep)
            if (full || (col_ok && 2 * i < rows_left)) *reinterpret_cast<float2*>(op) = x[i];
        } else if (LNF) {
          float* op = static_cast<float*>(args.out) + row0 * args.ldo + col;
#pragma unroll
          for (int i = 0; i < 16; ++i, op += ostep)
            if (full || (col_ok && 2 * i < rows_left)) stg_f2_hint(op, x[i], pol_keep);
        } else {
          float* op = static_cast<float*>(args.out) + row0 * args.ldo + col;
#pragma unroll
          for (int i = 0; i < 16; ++i, op += ostep)
            if (full || (col_ok && 2 * i < rows_left)) *reinterpret_cast<float2*>(op) = x[i];
        }
      }
      if (!released) {   // warp had no column block in this (ragged) tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(empty_leader[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
      }
      if (LNF && n_blk == n_tiles - 1) {
        // ---- LayerNorm over the rows this CTA has just completed (BertLayerNorm, CMIM:518-522) ----
        // (1) transpose-reduce the 16 per-thread row partials over the 16 lanes that share `sub`: after the four
        //     exchange steps lane (sub, cp) holds the total of row 2*cp + sub over this warp's columns
        float tsum, tsq;
        {
          float a[16], q[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { a[i] = rs[i]; q[i] = rq[i]; }
#pragma unroll
          for (int w = 8; w >= 1; w >>= 1) {
            const bool up = (cp & w) != 0;
#pragma unroll
            for (int k = 0; k < w; ++k) {
              const float sa = up ? a[k] : a[k + w], sq = up ? q[k] : q[k + w];
              const float ra = __shfl_xor_sync(0xffffffffu, sa, w), rq2 = __shfl_xor_sync(0xffffffffu, sq, w);
              a[k] = (up ? a[k + w] : a[k]) + ra;
              q[k] = (up ? q[k + w] : q[k]) + rq2;
            }
          }
          tsum = a[0];
          tsq = q[0];
        }
        // (2) the two warps of a lane quadrant own interleaved column blocks of the same 32 rows: combine through
        //     shared memory (the idle transpose tile of the quadrant's first warp), then every thread fetches the
        //     statistics of its 16 rows; the second barrier frees the tile for the next column block
        float2* st = reinterpret_cast<float2*>(smem_stage + ((quad + 2) & 3) * (32 * 128));   // [half][32 rows]
        asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quad) : "memory");   // the tile's owner is done transposing through it
        st[half * 32 + 2 * cp + sub] = make_float2(tsum, tsq);
        asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quad) : "memory");
        const float inv_n = 1.0f / (float)N;
        float mean[16], rstd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 p0 = st[2 * i + sub], p1 = st[32 + 2 * i + sub];
          const float mu = (p0.x + p1.x) * inv_n;
          const float var = fmaxf((p0.y + p1.y) * inv_n - mu * mu, 0.0f);
          mean[i] = mu;
          rstd[i] = rsqrtf(var + args.ln_eps);
        }
        asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quad) : "memory");
        // (3) second pass over this thread's own stores (L2-resident): normalise in place, emit the bf16 copy
        const int row_base2 = m_blk * kTileM + quad * 32;
        const int rows_left2 = M - row_base2 - sub;
        for (int nb = 0; nb < n_tiles; ++nb) {
          const int nch = min(BN / 32, (N - nb * BN + 31) / 32);
          for (int c = half; c < nch; c += 2) {
            const int col = nb * BN + c * 32 + 2 * cp;
            if (col >= N) continue;
            const float2 g = __ldg(reinterpret_cast<const float2*>(args.ln_gamma + col));
            const float2 bt = __ldg(reinterpret_cast<const float2*>(args.ln_beta + col));
            float* op = static_cast<float*>(args.out) + (size_t)(row_base2 + sub) * args.ldo + col;
            __nv_bfloat16* hp = args.out16 ? args.out16 + (size_t)(row_base2 + sub) * N + col : nullptr;
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v[i] = (2 * i < rows_left2) ? ldg_f2_hint(op + (size_t)(2 * i) * args.ldo, pol_keep) : make_float2(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (2 * i < rows_left2) {
                const float y0 = g.x * ((v[i].x - mean[i]) * rstd[i]) + bt.x;
                const float y1 = g.y * ((v[i].y - mean[i]) * rstd[i]) + bt.y;
                stg_f2_hint(op + (size_t)(2 * i) * args.ldo, make_float2(y0, y1), pol_stream);
                if (hp) stg_u32_hint(hp + (size_t)(2 * i) * N, pack_bf16x2(y0, y1), pol_stream);
              }
            }
          }
        }
      }
    }
    if (tma_blocks > 0 && lane == 0) bulk_wait_group_all();   // TMA-store epilogue: writes performed before the CTA retires
  }

  // ---- teardown ----
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();   // the peer's smem / barriers stay valid until both are done
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// out[m, n] = sum_s part[s][m][n] (s ascending) + bias[n] + residual[m][n]
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int64_t split_stride,
                                                            int splits, const float* __restrict__ bias,
                                                            const float* __restrict__ residual, float* __restrict__ out,
                                                            int64_t total, int N) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += stride) {
    float4 a = __ldcg(reinterpret_cast<const float4*>(part + i));
    for (int s = 1; s < splits; ++s) {
      const float4 b = __ldcg(reinterpret_cast<const float4*>(part + (size_t)s * split_stride + i));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (bias) {
      const float4 b = *reinterpret_cast<const float4*>(bias + (int)(i % N));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + i);
      a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    }
    *reinterpret_cast<float4*>(out + i) = a;
  }
}

// The same reduction followed by the row LayerNorm (two-pass statistics, as layernorm_kernel): one warp per row, the row
// in registers.  The single-query encoders end every block with  LN(split-K GEMM + bias + residual)  on a few hundred to a few
// thousand rows; as two launches the fp32 pre-LayerNorm rows made a round trip and the second launch cost more than its work.
constexpr int kRedLnMaxVec = 8;   // N <= 1024
__global__ void __launch_bounds__(256) splitk_reduce_ln_kernel(const float* __restrict__ part, int64_t split_stride, int splits,
                                                               const float* __restrict__ bias,
                                                               const float* __restrict__ residual,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float eps, float* __restrict__ y32,
                                                               __nv_bfloat16* __restrict__ y16, int M, int N) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = N / 4;
  float4 v[kRedLnMaxVec];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kRedLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const size_t off = (size_t)row * N + 4 * c;
      float4 a = __ldcg(reinterpret_cast<const float4*>(part + off));
      for (int sp = 1; sp < splits; ++sp) {
        const float4 b = __ldcg(reinterpret_cast<const float4*>(part + (size_t)sp * split_stride + off));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      if (residual) {
        const float4 r = *reinterpret_cast<const float4*>(residual + off);
        a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
      }
      v[i] = a;
      sum += (a.x + a.y) + (a.z + a.w);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)N;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < kRedLnMaxVec; ++i)
    if (lane + 32 * i < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.0f / sqrtf(sq / (float)N + eps);
#pragma unroll
  for (int i = 0; i < kRedLnMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = g.x * ((v[i].x - mean) * rstd) + bt.x;
      o.y = g.y * ((v[i].y - mean) * rstd) + bt.y;
      o.z = g.z * ((v[i].z - mean) * rstd) + bt.z;
      o.w = g.w * ((v[i].w - mean) * rstd) + bt.w;
      reinterpret_cast<float4*>(y32 + (size_t)row * N)[c] = o;
      if (y16) reinterpret_cast<uint2*>(y16 + (size_t)row * N)[c] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// 2-D bf16 tensor [rows, cols] with row pitch `ld` elements; box = {64 cols, box_rows}, 128-byte swizzle.
int icka_make_tmap_bf16(icka_handle* h, CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    ICKA_FAIL(ICKA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_rows);
  return ICKA_OK;
}

// MN-major operand: matrix [rows = contraction index, cols = M or N index] with row pitch `ld`, seen as a 3-D
// tensor {64 elements, rows, cols / 64}; box = {64, 64 rows, box_chunks}, 128-byte swizzle.  cols % 64 == 0.
int icka_make_tmap_bf16_mn(icka_handle* h, CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                           int box_chunks) {
  const cuuint64_t gdim[3] = {64, (cuuint64_t)rows, (cuuint64_t)(cols / 64)};
  const cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, 128};
  const cuuint32_t box[3] = {64, (cuuint32_t)kBK, (cuuint32_t)box_chunks};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    ICKA_FAIL(ICKA_ERR_CUDA, "cuTensorMapEncodeTiled(3-D) failed (%d) rows=%lld cols=%lld ld=%lld chunks=%d", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_chunks);
  return ICKA_OK;
}

namespace {

template <int BN, int ACT, bool OUT_BF16, int CTAS, int MAJOR, bool LNF = false, bool TMAE = false>
int launch_gemm(icka_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CTAS>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, ACT, OUT_BF16, CTAS, MAJOR, LNF, TMAE>;
  ICKA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
  const int m_tiles = (args.M + kBM * CTAS - 1) / (kBM * CTAS), n_tiles = (args.N + BN - 1) / BN;
  const int tiles = LNF ? m_tiles : m_tiles * n_tiles * args.splits;   // LNF: a CTA owns whole row-blocks
  const int groups_max = h->sm_count / CTAS;
  const int groups = tiles < groups_max ? tiles : groups_max;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * CTAS);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ICKA_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, args));
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

}  // namespace

// Output tensor map of the TMA-store epilogue: [rows, cols] bf16, box {32 columns, 32 rows}, 64-byte swizzle.
static int icka_make_tmap_bf16_out(icka_handle* h, CUtensorMap* tm, void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {32u, 32u};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    ICKA_FAIL(ICKA_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
  return ICKA_OK;
}
// developer knob ICKA_GEMM_TMA_EPI=0: bf16 outputs through the transposing epilogue as well
static const bool g_gemm_tma_epi = [] {
  const char* e = getenv("ICKA_GEMM_TMA_EPI");
  return !(e && e[0] == '0');
}();

// icka_gemm_mode: 0 = choose per shape (default), 1 = force single-CTA tiles, 2 = force CTA pairs.
static int g_gemm_mode = 0;
static int g_gemm_debug = 0;
extern "C" int icka_set_gemm_mode(int mode) {
  if (mode < 0 || (mode & 15) > 2) ICKA_FAIL(ICKA_ERR_INVALID, "gemm mode %d not in 0..2", mode);
  g_gemm_mode = mode & 15;
  g_gemm_debug = mode >> 4;   // developer probes, see GemmArgs::debug
  return ICKA_OK;
}

namespace {
// split count for a skinny fp32-out problem (0 / 1: do not split)
int plan_splits(const icka_handle* h, int M, int N, int K, int BN) {
  const int tiles = ((M + kBM - 1) / kBM) * ((N + BN - 1) / BN);
  const int num_kb = (K + kBK - 1) / kBK;
  int splits = h->sm_count / tiles;
  if (splits > num_kb / 8) splits = num_kb / 8;
  if (splits < 2 || (size_t)splits * M * N * sizeof(float) > ICKA_WORKSPACE_BYTES || h->workspace == nullptr || N % 4 != 0)
    return 0;
  return splits;
}
// per-split partial products into the handle's workspace; args.splits / split_stride describe the slabs afterwards
int launch_split_partials(icka_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, GemmArgs& args, int M, int N, int K,
                          int BN, int splits, cudaStream_t st) {
  const int num_kb = (K + kBK - 1) / kBK;
  const int kbps = (num_kb + splits - 1) / splits;
  args.splits = (num_kb + kbps - 1) / kbps;
  args.kb_per_split = kbps;
  args.split_stride = (int64_t)M * N;
  args.out = h->workspace;
  args.ldo = N;
  args.bias = nullptr;
  args.residual = nullptr;
  return (BN == 256) ? launch_gemm<256, ICKA_ACT_NONE, false, 1, 0>(h, ta, tb, args, st)
                     : launch_gemm<128, ICKA_ACT_NONE, false, 1, 0>(h, ta, tb, args, st);
}
}  // namespace

// LayerNorm(A . W^T + bias + residual) for skinny problems: split-K partials, then ONE reduce + LayerNorm pass.
// Returns 0 = launched, < 0 = error, > 0 = the shape does not split (the caller picks another route).
int icka_gemm_bf16_splitk_ln_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                    const float* residual, const float* gamma, const float* beta, float eps, float* out32,
                                    void* out16, int M, int N, int K, cudaStream_t st) {
  if (lda % 8 != 0 || ldw % 8 != 0 || N % 8 != 0 || N > 128 * kRedLnMaxVec || !icka_aligned(A, 16) || !icka_aligned(W, 16) ||
      !icka_aligned(out32, 16) || !icka_aligned(out16, 8) || !icka_aligned(residual, 16) || !icka_aligned(bias, 16) ||
      !icka_aligned(gamma, 16) || !icka_aligned(beta, 16))
    return 1;
  const int BN = (N > 128) ? 256 : 128;
  const int splits = plan_splits(h, M, N, K, BN);
  if (splits < 2) return 1;
  CUtensorMap ta, tb;
  int rc = icka_make_tmap_bf16(h, &ta, A, M, K, lda, kBM);
  if (rc) return rc < 0 ? rc : -1;
  rc = icka_make_tmap_bf16(h, &tb, W, N, K, ldw, BN);
  if (rc) return rc < 0 ? rc : -1;
  GemmArgs args{nullptr, nullptr, h->workspace, N, M, N, K, g_gemm_debug, nullptr, nullptr, (int64_t)N, 1, (K + kBK - 1) / kBK, 0};
  args.act_rt = ICKA_ACT_NONE;
  rc = launch_split_partials(h, ta, tb, args, M, N, K, BN, splits, st);
  if (rc) return rc < 0 ? rc : -1;
  splitk_reduce_ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(static_cast<const float*>(h->workspace), args.split_stride, args.splits,
                                                       bias, residual, gamma, beta, eps, out32,
                                                       static_cast<__nv_bfloat16*>(out16), M, N);
  ICKA_LAUNCHED(h);
  return 0;
}

// Forward GEMM  out[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual); aux_out (optional, with ACT_GELU_ERF)
// receives the bf16 pre-activation for the backward pass.
int icka_gemm_bf16_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                          const float* residual, void* out, int64_t ldo, int out_dtype, int M, int N, int K, int act,
                          void* aux_out, cudaStream_t st) {
  ICKA_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "linear(bf16): lda=%lld, ldw=%lld must be multiples of 8",
               (long long)lda, (long long)ldw);
  ICKA_REQUIRE(icka_aligned(A, 16) && icka_aligned(W, 16) && icka_aligned(out, 16),
               "linear(bf16): A, W, out must be 16-byte aligned");
  ICKA_REQUIRE(N % 8 == 0 && ldo % 8 == 0, "linear(bf16): N=%d and ldo=%lld must be multiples of 8", N, (long long)ldo);
  ICKA_REQUIRE(!residual || icka_aligned(residual, 16), "linear(bf16): residual must be 16-byte aligned");
  const bool rt = act == ICKA_ACT_RELU || act == ICKA_ACT_SWISH;
  ICKA_REQUIRE(!aux_out || ((act == ICKA_ACT_GELU_ERF || rt) && icka_aligned(aux_out, 16)),
               "linear(bf16): aux_out needs an FFN activation (gelu / relu / swish) and 16-byte alignment");
  constexpr size_t kNeedSmem = GemmCfg<256, 1>::kSmemBytes;
  ICKA_REQUIRE(h->smem_optin >= kNeedSmem, "linear(bf16): device offers too little shared memory");
  const int BN = (N > 128) ? 256 : 128;
  // CTA pairs pay off once there are enough 256-row tiles to fill the machine
  bool pair = (BN == 256) && ((long long)((M + 255) / 256) * ((N + 255) / 256) >= h->sm_count / 2);
  if (g_gemm_mode == 1 || act == ICKA_ACT_TANH) pair = false;
  if (g_gemm_mode == 2) pair = (BN == 256);
  CUtensorMap ta, tb;
  int rc = icka_make_tmap_bf16(h, &ta, A, M, K, lda, kBM);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tb, W, N, K, ldw, pair ? BN / 2 : BN);
  if (rc) return rc;
  GemmArgs args{bias, residual, out, ldo, M, N, K, g_gemm_debug, nullptr, static_cast<__nv_bfloat16*>(aux_out),
                (int64_t)N, 1, (K + kBK - 1) / kBK, 0};
  args.act_rt = act;
  const bool bf = out_dtype == ICKA_BF16;
  const bool gelu = act == ICKA_ACT_GELU_ERF;
  if (bf && !residual && !aux_out && (act == ICKA_ACT_NONE || gelu) && N % 32 == 0 && ldo % 8 == 0 && g_gemm_tma_epi &&
      (!bias || icka_aligned(bias, 16))) {
    rc = icka_make_tmap_bf16_out(h, &args.tmap_out, out, M, N, ldo);
    if (rc) return rc;
    args.tma_epi = 1;
  }
  if (act == ICKA_ACT_TANH) {   // prompt mapping networks (CMIM:914-930): M = batch, single-CTA tiles
    if (BN == 256) {
      if (bf) return launch_gemm<256, ICKA_ACT_TANH, true, 1, 0>(h, ta, tb, args, st);
      return launch_gemm<256, ICKA_ACT_TANH, false, 1, 0>(h, ta, tb, args, st);
    }
    if (bf) return launch_gemm<128, ICKA_ACT_TANH, true, 1, 0>(h, ta, tb, args, st);
    return launch_gemm<128, ICKA_ACT_TANH, false, 1, 0>(h, ta, tb, args, st);
  }
  // Skinny problems (the single-query encoders: M = batch, N = 768, K = 3072 / 9216) have only a few dozen output
  // tiles for 148 SMs: split the contraction over CTAs into per-split slabs of the handle's workspace, then add the
  // slabs in split order (+ bias, residual) with a small reduce pass -- bit-reproducible, unlike atomics.
  if (!bf && !gelu && !rt && !pair && ldo == N) {
    const int splits = plan_splits(h, M, N, K, BN);
    if (splits >= 2) {
      int rc2 = launch_split_partials(h, ta, tb, args, M, N, K, BN, splits, st);
      if (rc2) return rc2;
      const int64_t total = (int64_t)M * N;
      int blocks = (int)((total / 4 + 255) / 256);
      if (blocks > 4 * h->sm_count) blocks = 4 * h->sm_count;
      splitk_reduce_kernel<<<blocks, 256, 0, st>>>(static_cast<const float*>(h->workspace), args.split_stride, args.splits,
                                                   bias, residual, static_cast<float*>(out), total, N);
      ICKA_LAUNCHED(h);
      return ICKA_OK;
    }
  }
  if (args.tma_epi && !rt) {   // bf16 output, no residual / aux copy: TMA-store epilogue (own instantiations)
#define ICKA_GEMM_T(BN_, ACT_, C_) return launch_gemm<BN_, ACT_, true, C_, 0, false, true>(h, ta, tb, args, st)
    if (pair)           { if (gelu) ICKA_GEMM_T(256, ICKA_ACT_GELU_ERF, 2); else ICKA_GEMM_T(256, ICKA_ACT_NONE, 2); }
    else if (BN == 256) { if (gelu) ICKA_GEMM_T(256, ICKA_ACT_GELU_ERF, 1); else ICKA_GEMM_T(256, ICKA_ACT_NONE, 1); }
    else                { if (gelu) ICKA_GEMM_T(128, ICKA_ACT_GELU_ERF, 1); else ICKA_GEMM_T(128, ICKA_ACT_NONE, 1); }
#undef ICKA_GEMM_T
  }
#define ICKA_GEMM(BN_, ACT_, BF_, C_) return launch_gemm<BN_, ACT_, BF_, C_, 0>(h, ta, tb, args, st)
  if (rt) {   // non-default config.hidden_act
    if (pair)           { if (bf) ICKA_GEMM(256, kActRuntime, true, 2); else ICKA_GEMM(256, kActRuntime, false, 2); }
    else if (BN == 256) { if (bf) ICKA_GEMM(256, kActRuntime, true, 1); else ICKA_GEMM(256, kActRuntime, false, 1); }
    else                { if (bf) ICKA_GEMM(128, kActRuntime, true, 1); else ICKA_GEMM(128, kActRuntime, false, 1); }
  }
  if (pair) {
    if (gelu) { if (bf) ICKA_GEMM(256, ICKA_ACT_GELU_ERF, true, 2); else ICKA_GEMM(256, ICKA_ACT_GELU_ERF, false, 2); }
    else      { if (bf) ICKA_GEMM(256, ICKA_ACT_NONE, true, 2);     else ICKA_GEMM(256, ICKA_ACT_NONE, false, 2); }
  } else if (BN == 256) {
    if (gelu) { if (bf) ICKA_GEMM(256, ICKA_ACT_GELU_ERF, true, 1); else ICKA_GEMM(256, ICKA_ACT_GELU_ERF, false, 1); }
    else      { if (bf) ICKA_GEMM(256, ICKA_ACT_NONE, true, 1);     else ICKA_GEMM(256, ICKA_ACT_NONE, false, 1); }
  } else {
    if (gelu) { if (bf) ICKA_GEMM(128, ICKA_ACT_GELU_ERF, true, 1); else ICKA_GEMM(128, ICKA_ACT_GELU_ERF, false, 1); }
    else      { if (bf) ICKA_GEMM(128, ICKA_ACT_NONE, true, 1);     else ICKA_GEMM(128, ICKA_ACT_NONE, false, 1); }
  }
#undef ICKA_GEMM
}

// out32 / out16 = LayerNorm(A . W^T + bias + residual) with the normalisation fused into the epilogue (LNF).
int icka_gemm_bf16_ln_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                             const float* residual, const float* gamma, const float* beta, float eps, float* out32,
                             void* out16, int M, int N, int K, cudaStream_t st) {
  ICKA_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && N % 8 == 0, "linear_ln(bf16): lda, ldw, N must be multiples of 8");
  ICKA_REQUIRE(icka_aligned(A, 16) && icka_aligned(W, 16) && icka_aligned(out32, 16) && icka_aligned(out16, 16) &&
                   icka_aligned(residual, 16) && icka_aligned(gamma, 8) && icka_aligned(beta, 8),
               "linear_ln(bf16): pointers must be 16-byte aligned");
  CUtensorMap ta, tb;
  int rc = icka_make_tmap_bf16(h, &ta, A, M, K, lda, kBM);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tb, W, N, K, ldw, 256);
  if (rc) return rc;
  GemmArgs args{bias, residual, out32, (int64_t)N, M, N, K, 0, nullptr, nullptr, (int64_t)N, 1, (K + kBK - 1) / kBK, 0,
                gamma, beta, eps, static_cast<__nv_bfloat16*>(out16)};
  return launch_gemm<256, ICKA_ACT_NONE, false, 1, 0, true>(h, ta, tb, args, st);
}

// dgrad  dX[M,K] = (dY[M,N] . W[N,K]) (* gelu'(gelu_pre)) (+ residual): in kernel terms an M x K output contracted
// over N, with W consumed in its stored [N,K] layout as the MN-major B operand.
int icka_gemm_bf16_dgrad_launch(icka_handle* h, const void* dY, int64_t ldd, const void* W, int64_t ldw,
                                const float* residual, const void* gelu_pre, int64_t ldg, void* dX, int64_t ldo,
                                int out_dtype, int M, int N, int K, cudaStream_t st) {
  ICKA_REQUIRE(ldd % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0, "dgrad(bf16): pitches must be multiples of 8");
  ICKA_REQUIRE(K % 64 == 0, "dgrad(bf16): in_features %d must be a multiple of 64", K);
  ICKA_REQUIRE(icka_aligned(dY, 16) && icka_aligned(W, 16) && icka_aligned(dX, 16), "dgrad(bf16): 16-byte alignment");
  ICKA_REQUIRE(!residual || icka_aligned(residual, 16), "dgrad(bf16): residual must be 16-byte aligned");
  ICKA_REQUIRE(!gelu_pre || (icka_aligned(gelu_pre, 4) && ldg % 2 == 0), "dgrad(bf16): gelu_pre alignment");
  const int BN = (K > 128) ? 256 : 128;
  CUtensorMap ta, tb;
  int rc = icka_make_tmap_bf16(h, &ta, dY, M, N, ldd, kBM);
  if (rc) return rc;
  rc = icka_make_tmap_bf16_mn(h, &tb, W, N, K, ldw, BN / 64);
  if (rc) return rc;
  GemmArgs args{nullptr, residual, dX, ldo, M, K, N, 0, static_cast<const __nv_bfloat16*>(gelu_pre), nullptr, ldg, 1,
                (N + kBK - 1) / kBK, 0};
  const bool bf = out_dtype == ICKA_BF16;
#define ICKA_GEMM(BN_, ACT_, BF_) return launch_gemm<BN_, ACT_, BF_, 1, 1>(h, ta, tb, args, st)
  if (BN == 256) {
    if (gelu_pre) { if (bf) ICKA_GEMM(256, ICKA_ACT_GELU_ERF_BWD, true); else ICKA_GEMM(256, ICKA_ACT_GELU_ERF_BWD, false); }
    else          { if (bf) ICKA_GEMM(256, ICKA_ACT_NONE, true);         else ICKA_GEMM(256, ICKA_ACT_NONE, false); }
  } else {
    if (gelu_pre) { if (bf) ICKA_GEMM(128, ICKA_ACT_GELU_ERF_BWD, true); else ICKA_GEMM(128, ICKA_ACT_GELU_ERF_BWD, false); }
    else          { if (bf) ICKA_GEMM(128, ICKA_ACT_NONE, true);         else ICKA_GEMM(128, ICKA_ACT_NONE, false); }
  }
#undef ICKA_GEMM
}

// wgrad  dW[N,K] += dY[M,N]^T . X[M,K]: an N x K output contracted over the M tokens, both operands MN-major,
// token range split over CTAs.  dW must already hold the value to accumulate onto (zeros for a fresh gradient).
int icka_gemm_bf16_wgrad_launch(icka_handle* h, const void* dY, int64_t ldd, const void* X, int64_t ldx, float* dW,
                                int64_t ldo, int M, int N, int K, cudaStream_t st) {
  ICKA_REQUIRE(ldd % 8 == 0 && ldx % 8 == 0 && ldo % 2 == 0, "wgrad(bf16): pitches must be multiples of 8 (dW: 2)");
  ICKA_REQUIRE(N % 64 == 0 && K % 64 == 0, "wgrad(bf16): out_features %d and in_features %d must be multiples of 64",
               N, K);
  ICKA_REQUIRE(icka_aligned(dY, 16) && icka_aligned(X, 16) && icka_aligned(dW, 8), "wgrad(bf16): alignment");
  const int BN = (K > 128) ? 256 : 128;
  CUtensorMap ta, tb;
  int rc = icka_make_tmap_bf16_mn(h, &ta, dY, M, N, ldd, kBM / 64);
  if (rc) return rc;
  rc = icka_make_tmap_bf16_mn(h, &tb, X, M, K, ldx, BN / 64);
  if (rc) return rc;
  const int num_kb = (M + kBK - 1) / kBK;
  const int mn_tiles = ((N + kBM - 1) / kBM) * ((K + BN - 1) / BN);
  // enough token ranges for ~2 tiles per SM, each at least 8 k-blocks (512 tokens) long
  int splits = (2 * h->sm_count + mn_tiles - 1) / mn_tiles;
  if (splits > (num_kb + 7) / 8) splits = (num_kb + 7) / 8;
  if (splits < 1) splits = 1;
  const int kbps = (num_kb + splits - 1) / splits;
  splits = (num_kb + kbps - 1) / kbps;
  GemmArgs args{nullptr, nullptr, dW, ldo, N, K, M, 0, nullptr, nullptr, 0, splits, kbps, 0};
  if (BN == 256) return launch_gemm<256, ICKA_ACT_NONE, false, 1, 2>(h, ta, tb, args, st);
  return launch_gemm<128, ICKA_ACT_NONE, false, 1, 2>(h, ta, tb, args, st);
}
