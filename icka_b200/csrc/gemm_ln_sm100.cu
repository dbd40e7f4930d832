// Dense + bias + residual + BertLayerNorm in ONE tcgen05 kernel (BertSelfOutput CMIM:561-565, BertOutput CMIM:532-536):
//
//   out32 / out16 [M, N] = LayerNorm_{gamma, beta, eps}( A[M,K] . W[N,K]^T + bias + residual[M,N] )      N = 256 * C
//
// The row statistics of LayerNorm span all N columns, but an fp32 accumulator row of N = 768 does not fit the 512 TMEM
// columns of one SM.  A thread-block CLUSTER of C = N / 256 CTAs (3 for H = 768, 4 for H = 1024) owns a 128-row block:
// CTA r computes columns [256 r, 256 r + 256) with the same pipeline as gemm_sm100.cu (TMA producer warp, single-thread
// tcgen05.mma issuer, two 256-column accumulators in TMEM, 8 epilogue warps) and the CTAs exchange per-row partial sums
// through distributed shared memory:
//
//   pass 1   thread = accumulator row (TMEM lane): x = acc + bias + residual, accumulate sum(x), sum(x^2) in two registers,
//            park x in TMEM (tcgen05.st over the accumulator) -- no second read of the residual, no global round trip
//   exchange every epilogue thread writes its (sum, sumsq) into the stats table of EVERY CTA of the cluster
//            (st.shared::cluster), one release-arrive per warp on each CTA's mbarrier; wait (acquire.cluster) on the own one
//   pass 2   x back from TMEM, y = gamma (x - mean) rstd + beta, stored as fp32 and as the bf16 operand copy of the next GEMM
//
// Global traffic with thread = row: every lane reads / writes whole 32-byte sectors of its own row with 256-bit accesses
// (ld / st.global.v8), so no byte moves twice although a warp instruction touches 32 different lines; the residual is
// prefetched one 32-column chunk ahead in registers and one tile ahead into L2.  The eight epilogue warps form two groups
// that alternate tiles (group g owns TMEM accumulator g): the latency of the statistics exchange, of the residual loads and
// of the accumulator hand-over of one group is filled by the other group's passes.  (A first version staged 32 x 32 tiles
// in shared memory for TMA stores with all eight warps on one tile: every latency sat on the critical path -- 360 us for
// 131072 x 768 x 768 against 370 us for GEMM + layernorm_kernel; probes in tools/ln_probe.sh.)
// HBM traffic per 128 x 256 tile: residual 128 KB in, 128 + 64 KB out, A once per cluster from HBM (the other CTAs hit L2):
// the standalone GEMM + layernorm_kernel pair moved 128 KB more per tile (pre-LayerNorm tensor out and in again).
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <stdlib.h>

namespace {

using namespace sm100;

constexpr int kBM = 128, kBN = 256, kBK = 64, kUmmaK = 16;
constexpr int kStages = 3;
constexpr int kTileBytes = 32 * 128;            // 32 rows x 32 fp32: a warp's residual-in / result-out staging tile
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kABytes = kBM * kBK * 2;          // 16 KB
constexpr int kBBytes = kBN * kBK * 2;          // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kMaxC = 4;

struct LnArgs {
  const float* bias;       // [N]
  const float* residual;   // [M, N] fp32 (pitch N) or null
  const float* gamma;      // [N]
  const float* beta;       // [N]
  float eps;
  int M, N, K;
  float* out32;            // [M, N] fp32
  __nv_bfloat16* out16;    // [M, N] bf16 or null
  int debug;   // developer probes (ICKA_LN_DEBUG; results are WRONG): 1 = no residual loads, 2 = no staging / TMA stores,
               // 4 = no statistics exchange, 8 = no pass 2 at all
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// one full 32-byte sector per lane
__device__ __forceinline__ void ldg_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void stg_v8_f32(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void stg_v8_b32(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int32_t c0, int32_t c1) {   // global -> L2 only
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];\n" ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};\n" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;\n" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (((++spins) & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      printf("icka_b200: cluster mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}
template <int C>
struct LnSmem {
  static constexpr int kStatsBytes = 2 * C * kBM * 8;                // [accumulator / group][source CTA][row] float2
  static constexpr int kVecBytes = 3 * kBN * 4;                      // bias | gamma | beta of this CTA's columns
  static constexpr size_t kBytes = (size_t)kStages * kStageBytes + 2 * kEpiWarps * kTileBytes + kStatsBytes + kVecBytes +
                                   512 /*barriers*/ + 1024 /*align slack*/;
};

template <int C>
__global__ void __launch_bounds__(kThreads, 1)
gemm_ln_cluster_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_o32,
                       const LnArgs args) {
  static_assert(C >= 2 && C <= kMaxC, "cluster of 2..4 CTAs");
  using SM = LnSmem<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (size_t)kStages * kABytes;
  uint8_t* stage_s = smem + (size_t)kStages * kStageBytes;                  // 16 x 4 KB, 1024-aligned
  float2* stats = reinterpret_cast<float2*>(stage_s + 2 * kEpiWarps * kTileBytes);
  float* vec_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stats) + SM::kStatsBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(vec_s) + SM::kVecBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* stats_bar = tmem_empty_bar + 2;
  uint64_t* res_bars = stats_bar + 2;                                       // [warp][buffer]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_bars + 2 * kEpiWarps);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = (int)blockIdx.x / C, num_clusters = (int)gridDim.x / C;
  const int M = args.M, N = args.N, K = args.K;
  const int m_tiles = (M + kBM - 1) / kBM;
  const int num_kb = (K + kBK - 1) / kBK;
  const int col0 = (int)rank * kBN;                     // this CTA's first output column

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_res);
    tma_prefetch_desc(&tmap_o32);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kEpiWarps / 2);    // accumulator a belongs to epilogue group a (4 warps)
      mbar_init(&stats_bar[a], (kEpiWarps / 2) * C);    // one arrive per warp of group a of every CTA in the cluster
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bars[i], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, 2 * kBN);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kBN; i += kThreads) {
    vec_s[i] = args.bias ? args.bias[col0 + i] : 0.0f;
    vec_s[kBN + i] = args.gamma[col0 + i];
    vec_s[2 * kBN + i] = args.beta[col0 + i];
  }
  tc_fence_before();
  cluster_sync_all();                                   // peers' barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int mb = cluster_id; mb < m_tiles; mb += num_clusters) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(smem_a + (size_t)stage * kABytes, &tmap_a, &full_bar[stage], kb * kBK, mb * kBM);
          tma_load_2d(smem_b + (size_t)stage * kBBytes, &tmap_b, &full_bar[stage], kb * kBK, col0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int mb = cluster_id; mb < m_tiles; mb += num_clusters, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kBN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * kABytes);
          const uint32_t b_addr = smem_u32(smem_b + (size_t)stage * kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_bf16(tmem_d, make_kmajor_sw128_desc(a_addr + k * (kUmmaK * 2)),
                      make_kmajor_sw128_desc(b_addr + k * (kUmmaK * 2)), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps (one per TMEM lane quadrant) alternate tiles =====================
    // Group g owns accumulator g (tiles it = g, g + 2, ...): while it waits for the cluster's row statistics, for its
    // residual tiles or for the next accumulator, the other group's passes keep the SM busy.  Each warp owns two 4 KB
    // shared-memory tiles (32 rows x 32 fp32, 128-byte swizzle): in pass 1 they receive the residual by TMA (chunk c + 2
    // is in flight while chunk c is consumed), in pass 2 they stage the fp32 result for TMA stores.
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int grp = ew >> 2;
    const int row_in_tile = quad * 32 + lane;
    uint8_t* buf[2] = {stage_s + (size_t)(2 * ew) * kTileBytes, stage_s + (size_t)(2 * ew + 1) * kTileBytes};
    uint64_t* res_bar = res_bars + 2 * ew;                // one per buffer
    uint32_t res_phase[2] = {0u, 0u};
    const float* bias_s = vec_s;
    const float* gamma_s = vec_s + kBN;
    const float* beta_s = vec_s + 2 * kBN;
    uint32_t stats_remote[kMaxC], bar_remote[kMaxC];
#pragma unroll
    for (int p = 0; p < C; ++p) {
      stats_remote[p] = map_to_cta(stats, (uint32_t)p);
      bar_remote[p] = map_to_cta(&stats_bar[grp], (uint32_t)p);
    }
    const float inv_n = 1.0f / (float)N;
    const int acc = grp;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kBN);
    const bool use_res = args.residual != nullptr && !(args.debug & 1);
    const int sw = lane & 7;                              // 128-byte swizzle: 16-byte slot j of row r lives at slot j ^ (r & 7)
    int it = grp;
    for (int mb = cluster_id + grp * num_clusters; mb < m_tiles; mb += 2 * num_clusters, it += 2) {
      const uint32_t ph = (it >> 1) & 1;
      const int row0 = mb * kBM + quad * 32;              // first row of this warp's 32-row block
      const int row = row0 + lane;
      const bool row_ok = row < M;
      if (use_res && lane == 0) {
        bulk_wait_read0();                                // the previous tile's TMA stores have left both buffers
        // this group's NEXT tile starts its way from HBM to L2 now (two tile times ahead)
        if (row0 + 2 * num_clusters * kBM < M) {
#pragma unroll
          for (int c = 0; c < kBN / 32; ++c) tma_prefetch_2d(&tmap_res, col0 + c * 32, row0 + 2 * num_clusters * kBM);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          mbar_arrive_expect_tx(&res_bar[c], kTileBytes);
          tma_load_2d(buf[c], &tmap_res, &res_bar[c], col0 + c * 32, row0);
        }
      }
      mbar_wait(&tmem_full_bar[acc], ph);
      tc_fence_after();

      // ---- pass 1: x = acc + bias + residual -> TMEM, row partial sums ----
      float sum = 0.0f, sq = 0.0f;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
        const int b = c & 1;
        if (use_res) {
          mbar_wait(&res_bar[b], res_phase[b]);
          res_phase[b] ^= 1u;
        }
        tmem_ld_wait();
        const uint8_t* rrow = buf[b] + lane * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bs = *reinterpret_cast<const float4*>(bias_s + c * 32 + q * 4);
          float4 rs = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (use_res) rs = *reinterpret_cast<const float4*>(rrow + ((q ^ sw) << 4));
          const float x0 = __uint_as_float(r[4 * q]) + bs.x + rs.x;
          const float x1 = __uint_as_float(r[4 * q + 1]) + bs.y + rs.y;
          const float x2 = __uint_as_float(r[4 * q + 2]) + bs.z + rs.z;
          const float x3 = __uint_as_float(r[4 * q + 3]) + bs.w + rs.w;
          sum += (x0 + x1) + (x2 + x3);
          sq = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, sq))));
          r[4 * q] = __float_as_uint(x0);
          r[4 * q + 1] = __float_as_uint(x1);
          r[4 * q + 2] = __float_as_uint(x2);
          r[4 * q + 3] = __float_as_uint(x3);
        }
        tmem_st_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
        if (use_res && c + 2 < kBN / 32) {
          __syncwarp();                                   // every lane has read buffer b
          if (lane == 0) {
            mbar_arrive_expect_tx(&res_bar[b], kTileBytes);
            tma_load_2d(buf[b], &tmap_res, &res_bar[b], col0 + (c + 2) * 32, row0);
          }
        }
      }
      tmem_st_wait();

      // ---- exchange the row partials over distributed shared memory ----
      float tsum = sum, tsq = sq;
      if (!(args.debug & 4)) {
        const uint32_t slot = (uint32_t)(((acc * C + (int)rank) * kBM + row_in_tile) * 8);
#pragma unroll
        for (int p = 0; p < C; ++p) st_cluster_f2(stats_remote[p] + slot, sum, sq);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int p = 0; p < C; ++p) mbar_arrive_cluster(bar_remote[p]);      // release.cluster, cumulative over the warp's stores
        }
        mbar_wait_cluster(&stats_bar[acc], ph);
        tsum = 0.0f;
        tsq = 0.0f;
#pragma unroll
        for (int p = 0; p < C; ++p) {
          const float2 v = stats[(acc * C + p) * kBM + row_in_tile];
          tsum += v.x;
          tsq += v.y;
        }
      }
      const float mean = tsum * inv_n;
      const float rstd = rsqrtf(fmaxf(tsq * inv_n - mean * mean, 0.0f) + args.eps);

      // ---- pass 2: normalise; fp32 through the staging tiles + TMA, bf16 as whole 32-byte sectors per lane ----
      if (args.debug & 8) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        continue;
      }
      __nv_bfloat16* o16 = (args.out16 && row_ok) ? args.out16 + (size_t)row * N + col0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
        const int b = c & 1;
        if (lane == 0) bulk_wait_read1();                 // the store that used buffer b two chunks ago has been read
        __syncwarp();
        tmem_ld_wait();
        if (c + 1 == kBN / 32) {          // last TMEM read of this accumulator by this warp
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (args.debug & 2) continue;
        uint8_t* yrow = buf[b] + lane * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 g = *reinterpret_cast<const float4*>(gamma_s + c * 32 + q * 4);
          const float4 bt = *reinterpret_cast<const float4*>(beta_s + c * 32 + q * 4);
          float4 y;
          y.x = fmaf(g.x, (__uint_as_float(r[4 * q]) - mean) * rstd, bt.x);
          y.y = fmaf(g.y, (__uint_as_float(r[4 * q + 1]) - mean) * rstd, bt.y);
          y.z = fmaf(g.z, (__uint_as_float(r[4 * q + 2]) - mean) * rstd, bt.z);
          y.w = fmaf(g.w, (__uint_as_float(r[4 * q + 3]) - mean) * rstd, bt.w);
          *reinterpret_cast<float4*>(yrow + ((q ^ sw) << 4)) = y;
          r[2 * q] = pack_bf16x2(y.x, y.y);
          r[2 * q + 1] = pack_bf16x2(y.z, y.w);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_o32, buf[b], col0 + c * 32, row0);
          bulk_commit();
        }
        if (o16) {
          stg_v8_b32(o16 + c * 32, r);
          stg_v8_b32(o16 + c * 32 + 16, r + 8);
        }
      }
    }
    if (lane == 0) bulk_wait_all();      // global writes performed before the kernel ends
  }

  // ---- teardown: peers may still write into / arrive on this CTA's shared memory until every CTA is here ----
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kBN);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 [rows, cols] tensor, box {32 cols, 32 rows} = 128-byte rows, SWIZZLE_128B (residual loads and result stores)
int make_f32_tile_tmap(icka_handle* h, CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(h->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    ICKA_FAIL(ICKA_ERR_CUDA, "cuTensorMapEncodeTiled(fp32 tile) failed (%d) rows=%lld cols=%lld", (int)r, (long long)rows,
              (long long)cols);
  return ICKA_OK;
}

template <int C>
int launch_ln(icka_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tres, const CUtensorMap& to32,
              const LnArgs& args, cudaStream_t st) {
  auto kern = gemm_ln_cluster_kernel<C>;
  constexpr size_t smem = LnSmem<C>::kBytes;
  ICKA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (C > 8) ICKA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent: as many clusters as the device can keep resident at once (one CTA per SM, clusters inside a GPC)
  static int cached_clusters[kMaxC + 1][16] = {};
  int dev = h->device < 16 ? h->device : 0;
  int max_clusters = cached_clusters[C][dev];
  if (max_clusters == 0) {
    cfg.gridDim = dim3(C * (h->sm_count / C));
    ICKA_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
    if (max_clusters < 1) ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "linear_ln: no cluster of %d CTAs fits this device", C);
    cached_clusters[C][dev] = max_clusters;
    if (getenv("ICKA_LN_DEBUG")) fprintf(stderr, "icka_b200: linear_ln cluster kernel: %d clusters of %d CTAs resident\n", max_clusters, C);
  }
  const int m_tiles = (args.M + kBM - 1) / kBM;
  const int clusters = m_tiles < max_clusters ? m_tiles : max_clusters;
  cfg.gridDim = dim3(C * clusters);
  ICKA_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tres, to32, args));
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

}  // namespace

// true when the cluster kernel serves this problem (bf16 operands, N = 768 or 1024, 32-byte aligned residual)
bool icka_gemm_ln_cluster_supported(int N, int K, const void* residual) {
  return (N == 512 || N == 768 || N == 1024) && K % 8 == 0 && (residual == nullptr || icka_aligned(residual, 32));
}

int icka_gemm_bf16_ln_cluster_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                                     const float* bias, const float* residual, const float* gamma, const float* beta,
                                     float eps, float* out32, void* out16, int M, int N, int K, cudaStream_t st) {
  ICKA_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "linear_ln(cluster): lda, ldw must be multiples of 8");
  ICKA_REQUIRE(icka_aligned(A, 16) && icka_aligned(W, 16) && icka_aligned(out32, 32) && icka_aligned(out16, 32) &&
                   icka_aligned(residual, 32) && icka_aligned(gamma, 4) && icka_aligned(beta, 4),
               "linear_ln(cluster): pointer alignment");
  ICKA_REQUIRE(h->smem_optin >= LnSmem<kMaxC>::kBytes, "linear_ln(cluster): device offers too little shared memory");
  CUtensorMap ta, tb, tres, to32;
  int rc = icka_make_tmap_bf16(h, &ta, A, M, K, lda, kBM);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &tb, W, N, K, ldw, kBN);
  if (rc) return rc;
  rc = make_f32_tile_tmap(h, &tres, residual ? residual : out32, M, N);
  if (rc) return rc;
  rc = make_f32_tile_tmap(h, &to32, out32, M, N);
  if (rc) return rc;
  const char* dbg = getenv("ICKA_LN_DEBUG");
  LnArgs args{bias, residual, gamma, beta, eps, M, N, K, out32, static_cast<__nv_bfloat16*>(out16), dbg ? atoi(dbg) : 0};
  switch (N / kBN) {
    case 2: return launch_ln<2>(h, ta, tb, tres, to32, args, st);
    case 3: return launch_ln<3>(h, ta, tb, tres, to32, args, st);
    case 4: return launch_ln<4>(h, ta, tb, tres, to32, args, st);
  }
  ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "linear_ln(cluster): N = %d is not 512, 768 or 1024", N);
}
