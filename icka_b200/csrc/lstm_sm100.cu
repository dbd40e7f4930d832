// Recurrent half of the emission head's BiLSTM (CMIM:905-908 `nn.LSTM(H, H, batch_first, bidirectional)`, call
// CMIM:1042) as ONE persistent, weight-stationary tcgen05 kernel for sm_100a (SURVEY 8f "next" row 1).
//
//   gates_t = Gx[:, t] + h_{t-1} . W_hh^T          Gx = x . W_ih^T + b_ih + b_hh, one big tensor-core GEMM (icka_linear_fwd)
//   i, f, g, o = sigmoid, sigmoid, tanh, sigmoid    (PyTorch gate order)
//   c_t = f * c_{t-1} + i * g ;  h_t = o * tanh(c_t)
//
// The recurrence is S serial steps of a [B, H] x [H, 4H] product per direction.  Launching a GEMM per step would
// stream W_hh (4.7 MB bf16 per direction) from L2 S times per tile and pay a launch + pipeline fill per step; here
//   * weight-stationary CTA PAIRS: the 2 x 3072 x 768 weights are cut into 2 x 16 slices of 48 hidden units (192 gate
//     columns); a slice belongs to a cluster of two CTAs, each keeping 96 of its columns (147 KB bf16) in shared
//     memory for the whole sequence (loaded once by TMA, SWIZZLE_128B, K-major B operand).  The pair issues
//     tcgen05.mma.cta_group::2 (M = 256: two sentence tiles, N = 192): every staged byte of A feeds twice the tensor
//     work of a single-CTA N = 96 tile, which is what the kernel was short of -- with 5 x 16 KB of A in flight the
//     N = 96 version kept only ~1000 cycles of MMA work queued against a ~2800-cycle issue -> commit -> refill loop
//     (ncu: tensor pipe 32 % active);
//   * a work item is (step t, PAIR of 128-sentence tiles): each CTA streams h_{t-1} of ITS tile (128 x 768 bf16) as
//     the A operand through a 4-stage TMA ring straight out of the time-major output sequence y[t-1] (h is written
//     exactly once; an out-of-bounds row coordinate delivers the zero state at t = 0), the leader CTA issues 48
//     MMAs into one of 2 TMEM accumulator slots (128 lanes x 192 columns in each CTA), and 8 epilogue warps per CTA
//     (thread = sentence = TMEM lane; two warp halves x two sequential blocks of 12 units) add Gx (prefetched one
//     block ahead), apply the cell update with the cell state in registers (first two tiles of a CTA) or in the spare
//     TMEM columns next to the accumulators (tiles 3 and 4 at 2048 sentences per launch: four independent step chains
//     per CTA hide the ~13 us chain latency of a tile), and write h_t (bf16) into a shared-memory tile that the
//     publisher warp stores into y[t] with one TMA instruction;
//   * sentence tiles are independent recurrences: the only cross-CTA dependency is "the 16 CTAs that own tile m in my
//     direction (one per slice pair) have published h_{t-1}" -- a per-(direction, tile) arrival counter in global memory (red.release /
//     ld.acquire + fence.proxy.async before the TMA reads); no grid-wide barrier exists;
//   * a publisher warp turns "all 8 epilogue warps stored their part of the item" into ONE gpu-scope release: a
//     MEMBAR.GPU on a busy SM costs ~4 us and must not stall the warps that do the cell arithmetic.
// All CTAs must be co-resident (they wait on each other): the kernel is launched cooperatively (clusters of 2).
//
// Column order inside a slice (chosen on the host when the weights are permuted once):
//   column c = blk * 48 + jg * 16 + gate * 4 + jj   <->   hidden unit  slice * 48 + blk * 12 + jg * 4 + jj,
//   blk in 0..3 (= 2 * warp half + sequential block), gate in (i, f, g, o), jg in 0..2, jj in 0..3;
//   columns 0..95 live in the pair's CTA 0, 96..191 in CTA 1.
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <stdlib.h>

namespace {

using namespace sm100;

constexpr int kH = 768;                       // hidden size this kernel is built for
constexpr int kU = 48;                        // hidden units per CTA pair
constexpr int kNS = kH / kU;                  // 16 slices per direction
constexpr int kN = 4 * kU;                    // 192 gate columns per pair (UMMA N)
constexpr int kNHalf = kN / 2;                // 96 of them in each CTA's shared memory
constexpr int kKB = kH / 64;                  // 12 k-chunks of 64 bf16 = 128 B
constexpr int kWChunkBytes = kNHalf * 128;    // 12,288
constexpr int kWBytes = kKB * kWChunkBytes;   // 147,456
constexpr int kABytes = 128 * 128;            // one 128-row x 64-k A tile
constexpr int kStages = 4;                    // (4 or 5 stages measure the same in pair mode; the 16 KB go to the h staging tile)
constexpr int kStageTileBytes = 128 * kNHalf; // 128 rows x 48 units bf16 = 12,288: h_t of one item, stored by TMA
constexpr int kSlots = 2;                     // TMEM accumulator slots
constexpr int kSlotCols = 256;                // column stride between slots (192 used)
constexpr int kEpiWarps = 8;
constexpr int kPubWarp = 2 + kEpiWarps;       // warp 10: publishes finished items
constexpr int kThreads = 32 * (kPubWarp + 1);
constexpr int kCtasPerGroup = 2 * kNS * 2;    // 2 directions x 16 slices x 2 CTAs = 64
constexpr size_t kSmemBytes =
    (size_t)kWBytes + (size_t)kStages * kABytes + kStageTileBytes + 1024 /*align*/ + 256 /*barriers*/;

struct LstmArgs {
  const __nv_bfloat16* gx;   // [S*Bn, 2*4H] bf16 TIME-MAJOR (row = t * Bn + sentence), columns ordered
                             // [dir][slice][blk][jg][gate][4]
  int* cnt;                  // [2][MT] arrival counters (zeroed before launch)
  __nv_bfloat16* y;          // [S, Bn, 2H] bf16 TIME-MAJOR: forward states in [:H], backward in [H:]
  float* h_n;                // [2, Bn, H] fp32 or null (already offset to this launch's first sentence)
  float* c_n;                // [2, Bn, H] fp32 or null
  int Bn;                    // sentences of the whole call (pitch of the time planes of gx / y and of the direction
                             // planes of h_n / c_n)
  int B, S, b0, MT, TPG;     // sentences of this launch, steps, first sentence of this launch, 128-row tiles, tiles
                             // per CTA group (a pair walks them two at a time)
  unsigned long long* trace; // developer timeline (ICKA_LSTM_TRACE): [item][8] globaltimer stamps of CTA 0, or null
  int debug;                 // developer probes (ICKA_LSTM_DEBUG): 1 = no dependency wait, 2 = no cell arithmetic /
                             // state stores, 4 = publish without the gpu-scope release, 16 = no MMAs, 32 = no state stores (results are WRONG)
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Developer timeline (tools/lstm_trace.py), slots per item: 0 dependency satisfied (producer), 1 last A chunk issued
// (producer), 2 last MMA committed (MMA warp), 3 accumulator seen (epilogue warp 2), 4 tile written (epilogue warp 2),
// 5 TMA store complete (publisher), 6 counter released (publisher)
#define LSTM_TRACE(slot_, it_)                                                                      \
  do {                                                                                              \
    if (args.trace != nullptr && blockIdx.x == 0 && (it_) < 4096) args.trace[(size_t)(it_) * 8 + (slot_)] = globaltimer_ns(); \
  } while (0)

// TMA store of a {48 units, 128 sentences, 1 step} box of the output sequence from shared memory (bulk async group)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // writes performed, not only the source read
}

// 256-bit store (sm_100: STG.E.ENL2.256), 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e,
                                             uint32_t f, uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d),
               "r"(e), "r"(f), "r"(g), "r"(h)
               : "memory");
}

// Activations: MUFU.TANH, sigmoid(x) = 0.5 tanh(x / 2) + 0.5 -- 5 MUFU per unit and sentence.  (ex2 + rcp forms, 10 MUFU, and
// tanh.approx.f16x2, which splits into two MUFU.TANH.F16, both measured slower or equal.)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// Cell epilogue of one CTA: NP items (one sentence tile of this CTA each) per step, walked in (step, item) order.
// Thread = sentence row (TMEM lane); per item it handles two blocks of 12 hidden units one after the other; the cell
// state of its NP rows lives in registers for the whole sequence.
template <int NP>
__device__ __forceinline__ void lstm_epilogue(const LstmArgs& args, uint64_t* acc_full, uint32_t acc_empty_leader0,
                                              uint64_t* pub_bar, uint64_t* pub_free, uint8_t* h_tile, uint32_t lane_taddr,
                                              int warp, int lane, int dir, int slice, int tile_first) {
  const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
  const int half = (warp - 2) >> 2;          // which half of the slice's 48 units
  const int S = args.S;
  const int row0 = tile_first * 128 + quad * 32 + lane;   // item p: row0 + p * 256
  // cell state: items 0 and 1 of a step in registers; items 2 and 3 (2048 sentences per launch) in the 2 x 64 TMEM
  // columns the two 192-column accumulator slots leave free (thread = lane owns its own 12 columns per block)
  constexpr int NPR = NP < 2 ? NP : 2;
  float c[NPR][2][12];
#pragma unroll
  for (int p = 0; p < NPR; ++p)
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int q = 0; q < 12; ++q) c[p][hh][q] = 0.0f;
  // Gx does not depend on the recurrence: it is fetched ONE BLOCK AHEAD, off the step-to-step critical path
  uint32_t gw_nxt[24];
  auto fetch = [&](int t, int p, int hh, uint32_t (&gw)[24]) {
    const int pos = dir ? (S - 1 - t) : t;
    const int row = row0 + p * 256;
    if (row < args.B) {
      const size_t col = (size_t)((dir * kNS + slice) * 4 + half * 2 + hh) * 48;
      const uint4* gp = reinterpret_cast<const uint4*>(args.gx + ((size_t)pos * args.Bn + row) * (8 * kH) + col);
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const uint4 v = __ldg(gp + q);
        gw[4 * q] = v.x;
        gw[4 * q + 1] = v.y;
        gw[4 * q + 2] = v.z;
        gw[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 24; ++q) gw[q] = 0u;
    }
  };
  fetch(0, 0, 0, gw_nxt);
  const bool probe = (args.debug & 2) != 0;      // probe: no cell arithmetic, no state stores
  int it = 0;
  for (int t = 0; t < S; ++t) {
    const int pos = dir ? (S - 1 - t) : t;
#pragma unroll
    for (int p = 0; p < NP; ++p, ++it) {
      const int slot = it % kSlots;
      const uint32_t slot_phase = (it / kSlots) & 1;
      const int row = row0 + p * 256;
      const bool valid = row < args.B;
      uint32_t hw0[6], hw1[6];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int blk = half * 2 + hh;
        uint32_t gw[24];
#pragma unroll
        for (int q = 0; q < 24; ++q) gw[q] = gw_nxt[q];
        if (hh == 0) fetch(t, p, 1, gw_nxt);
        else if (p + 1 < NP) fetch(t, p + 1, 0, gw_nxt);
        else if (t + 1 < S) fetch(t + 1, 0, 0, gw_nxt);

        if (hh == 0) {
          mbar_wait(&acc_full[slot], slot_phase);
          tc_fence_after();
          if (warp == 2 && lane == 0) LSTM_TRACE(3, it);
        }
        // 48 accumulator columns = 3 groups of 4 units x (i, f, g, o): 16 columns are live at a time, the next
        // group's tcgen05.ld is in flight while this one is being computed
        const uint32_t taddr = lane_taddr + (uint32_t)(slot * kSlotCols + blk * 48);
        uint32_t ra[16], rb[16];
        float hv[12];
        // this block's cell state: registers (items 0, 1) or the spare TMEM columns of slot p - 2 (items 2, 3)
        constexpr int kSpareCol = 192;
        const int pr = p < NPR ? p : 0;
        const uint32_t c_taddr = lane_taddr + (uint32_t)((p >= 2 ? p - 2 : 0) * kSlotCols + kSpareCol + blk * 12);
        float cc[12];
        if (p < 2) {
#pragma unroll
          for (int q = 0; q < 12; ++q) cc[q] = c[pr][hh][q];
        } else if (t > 0) {
          uint32_t cw[12];
          tmem_ld8(c_taddr, &cw[0]);
          tmem_ld4(c_taddr + 8, &cw[8]);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 12; ++q) cc[q] = __uint_as_float(cw[q]);
        } else {
#pragma unroll
          for (int q = 0; q < 12; ++q) cc[q] = 0.0f;
        }
        auto cell4 = [&](const uint32_t (&r)[16], int jg) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int e = jg * 16 + g * 4 + jj;
              const uint32_t w = gw[e >> 1];
              const float gxv = __uint_as_float((e & 1) ? (w & 0xffff0000u) : (w << 16));
              pre[g] = __uint_as_float(r[g * 4 + jj]) + gxv;
            }
            const float ig = sigmoid_fast(pre[0]), fg = sigmoid_fast(pre[1]), gg = tanh_fast(pre[2]),
                        og = sigmoid_fast(pre[3]);
            const int j = jg * 4 + jj;
            cc[j] = probe ? cc[j] : fmaf(fg, cc[j], ig * gg);
            hv[j] = og * tanh_fast(cc[j]);
          }
        };
        tmem_ld16(taddr, ra);
        tmem_ld_wait();
        tmem_ld16(taddr + 16, rb);
        cell4(ra, 0);
        tmem_ld_wait();
        tmem_ld16(taddr + 32, ra);
        cell4(rb, 1);
        tmem_ld_wait();
        if (hh == 1) {   // this warp's last read of the slot: hand it back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_leader0 + (uint32_t)slot * 8u);
        }
        cell4(ra, 2);
        if (p < 2) {
#pragma unroll
          for (int q = 0; q < 12; ++q) c[pr][hh][q] = cc[q];
        } else {
          uint32_t cw[12];
#pragma unroll
          for (int q = 0; q < 12; ++q) cw[q] = __float_as_uint(cc[q]);
          tmem_st8(c_taddr, &cw[0]);
          tmem_st4(c_taddr + 8, &cw[8]);
          tmem_st_wait();
        }

        uint32_t hw[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) hw[q] = pack_bf16x2(hv[2 * q], hv[2 * q + 1]);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          if (hh == 0) hw0[q] = hw[q];
          else hw1[q] = hw[q];
        }
        if (valid && !probe && !(args.debug & 32)) {   // probe 32: cell arithmetic, no state stores
          const int unit0 = slice * kU + blk * 12;
          if (t == S - 1) {
            if (args.h_n) {
              float4* o = reinterpret_cast<float4*>(args.h_n + ((size_t)dir * args.Bn + row) * kH + unit0);
#pragma unroll
              for (int q = 0; q < 3; ++q) o[q] = make_float4(hv[4 * q], hv[4 * q + 1], hv[4 * q + 2], hv[4 * q + 3]);
            }
            if (args.c_n) {
              float4* o = reinterpret_cast<float4*>(args.c_n + ((size_t)dir * args.Bn + row) * kH + unit0);
#pragma unroll
              for (int q = 0; q < 3; ++q)
                o[q] = make_float4(cc[4 * q], cc[4 * q + 1], cc[4 * q + 2], cc[4 * q + 3]);
            }
          }
        }
      }
      // h_t of the item goes through a [128 rows x 48 units] shared-memory tile that the publisher warp stores with ONE
      // TMA instruction: with thread = row every global store instruction would be 32 separate L2 write transactions
      // (rows are 3 KB apart) and the LSU, not the tensor pipe, would bound the kernel (measured: 3.8-4.9 us of a
      // 12.6-13.4 us step).  The tile is free again once the previous item's store has read it (pub_free).
      if (lane == 0) mbar_wait(&pub_free[0], (it & 1) ^ 1);
      __syncwarp();
      if (!probe && !(args.debug & 32)) {
        uint4* dst = reinterpret_cast<uint4*>(h_tile + (quad * 32 + lane) * kNHalf + half * 48);   // 96 B per row
        dst[0] = make_uint4(hw0[0], hw0[1], hw0[2], hw0[3]);
        dst[1] = make_uint4(hw0[4], hw0[5], hw1[0], hw1[1]);
        dst[2] = make_uint4(hw1[2], hw1[3], hw1[4], hw1[5]);
      }
      fence_proxy_async();                    // generic-proxy writes of the tile -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        if (warp == 2) LSTM_TRACE(4, it);
        mbar_arrive(&pub_bar[0]);
      }
    }
  }
}

// MAXNP: most items per step a CTA of this instantiation walks (2: <= 1024 sentences per launch, cell state entirely in
// registers; 4: <= 2048, two more tiles' state in TMEM) -- two kernels so the common case keeps its register allocation
template <int MAXNP>
__global__ void __launch_bounds__(kThreads, 1)
lstm_rec_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                        const __grid_constant__ CUtensorMap tmap_y, const LstmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kWBytes;
  uint8_t* h_tile = smem_a + (size_t)kStages * kABytes;   // [128 rows][48 units] bf16, dense (the TMA store's box)
  uint64_t* bars = reinterpret_cast<uint64_t*>(h_tile + kStageTileBytes);
  uint64_t* full_bar = bars;                 // leader's are used: both CTAs' TMA bytes are credited there
  uint64_t* empty_bar = bars + kStages;      // own: the multicast commit arrives in both CTAs
  uint64_t* acc_full = bars + 2 * kStages;   // own (multicast commit)
  uint64_t* acc_empty = acc_full + kSlots;   // leader's: 2 x 8 epilogue warps arrive
  uint64_t* pub_bar = acc_empty + kSlots;
  uint64_t* pub_free = pub_bar + kSlots;
  uint64_t* w_bar = pub_free + kSlots;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int slice = pair % kNS;
  const int dir = (pair / kNS) & 1;
  const int group = pair / (2 * kNS);
  const int tile0 = group * args.TPG;
  const int tile1 = min(args.MT, tile0 + args.TPG);
  const int npairs = (tile1 - tile0 + 1) / 2;      // items per step; item p: tiles tile0 + 2p (CTA 0), + 2p + 1 (CTA 1)
  const int S = args.S;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_y);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kSlots; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 2 * kEpiWarps);
      mbar_init(&pub_bar[a], kEpiWarps);
      mbar_init(&pub_free[a], 1);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_base_slot, kSlots * kSlotCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();                          // the peer's barriers are initialised before anyone uses them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // the weight half of this CTA: resident for the whole sequence
  if (warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(w_bar, kWBytes);
    const int wrow = (dir * kNS + slice) * kN + (int)rank * kNHalf;
    for (int kb = 0; kb < kKB; ++kb) tma_load_2d(smem_w + (size_t)kb * kWChunkBytes, &tmap_w, w_bar, kb * 64, wrow);
  }
  if (warp == 1 && lane == 0) mbar_wait(w_bar, 0);
  __syncwarp();
  cluster_sync_all();                          // the leader's MMAs read BOTH halves: both have landed past this point

  if (npairs > 0) {
    if (warp == 0) {
      // ===================== TMA producer (one per CTA) =====================
      if (lane == 0) {
        const uint32_t full_leader0 = map_to_cta(&full_bar[0], 0);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < S; ++t) {
          for (int p = 0; p < npairs; ++p) {
            const int m = tile0 + 2 * p + (int)rank;
            const bool tile_ok = m < tile1;
            if (t > 0 && tile_ok && !(args.debug & 1)) {
              // h_{t-1} of this tile is complete once the 16 CTAs that own it (one per slice pair of this direction)
              // have arrived t times
              const int* c = args.cnt + dir * args.MT + m;
              const int need = kNS * t;
              if (ld_acquire_gpu(c) < need) {
                const long long t0 = clock64();
                while (ld_acquire_gpu(c) < need) {
                  __nanosleep(40);
                  if (clock64() - t0 > 4000000000LL) {
                    printf("icka_b200: lstm step wait timed out (block %d t %d tile %d have %d need %d)\n",
                           (int)blockIdx.x, t, m, ld_acquire_gpu(c), need);
                    __trap();
                  }
                }
              }
              fence_proxy_async_all();   // generic-proxy writes of the other CTAs -> visible to the TMA reads below
            }
            // A operand = h_{t-1} of this tile = the rows the previous step wrote into the (time-major) output
            // sequence; at t = 0 (and for the missing tile of an odd count) an out-of-bounds row makes TMA deliver zeros
            LSTM_TRACE(0, t * npairs + p);
            const int prev = dir ? (S - t) : (t - 1);
            const int arow = (t == 0 || !tile_ok) ? S * args.Bn : prev * args.Bn + args.b0 + m * 128;
            for (int kb = 0; kb < kKB; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kABytes);   // armed for the pair
              tma_load_2d_pair(smem_a + (size_t)stage * kABytes, &tmap_h, full_leader0 + (uint32_t)stage * 8u,
                               dir * kH + kb * 64, arow);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            LSTM_TRACE(1, t * npairs + p);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (one thread of the leader CTA) =====================
      if (lane == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(256, kN);
        // One thread issues ~50 MMAs per item: the loop must cost less than the MMAs themselves, so the
        // shared-memory descriptors are formed once and stepped by adding to their 14-bit address field.
        const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(smem_a));
        const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(smem_w));
        const bool no_mma = (args.debug & 16) != 0;       // probe
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = 0; t < S; ++t) {
          for (int p = 0; p < npairs; ++p, ++it) {
            const int slot = it % kSlots;
            const uint32_t slot_phase = (it / kSlots) & 1;
            mbar_wait(&acc_empty[slot], slot_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(slot * kSlotCols);
            uint64_t b_desc = b_desc0;
#pragma unroll 1
            for (int kb = 0; kb < kKB; ++kb, b_desc += (kWChunkBytes >> 4)) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (kABytes >> 4));
              if (!no_mma) {
                umma_bf16_pair(tmem_d, a_desc, b_desc, idesc, kb > 0 ? 1u : 0u);
                umma_bf16_pair(tmem_d, a_desc + 2, b_desc + 2, idesc, 1u);
                umma_bf16_pair(tmem_d, a_desc + 4, b_desc + 4, idesc, 1u);
                umma_bf16_pair(tmem_d, a_desc + 6, b_desc + 6, idesc, 1u);
              }
              umma_commit_pair(&empty_bar[stage], 3);     // the slot is free in both CTAs once these retire
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_pair(&acc_full[slot], 3);         // wake the epilogue warps of both CTAs
            LSTM_TRACE(2, it);
          }
        }
      }
    } else if (warp == kPubWarp) {
      // ===================== publisher: one gpu-scope release per item, off the epilogue warps' path =====================
      // The epilogue warps arrive on pub_bar[slot] (release.cta) after their h stores; this thread acquires the
      // barrier and performs the ONE gpu-scope release of the CTA (cumulative over everything that happened-before it).
      if (lane == 0) {
        const int items = S * npairs;
        for (int it = 0; it < items; ++it) {
          const int t = it / npairs;
          const int m = tile0 + 2 * (it % npairs) + (int)rank;
          const int pos = dir ? (S - 1 - t) : t;
          mbar_wait(&pub_bar[0], it & 1);                 // all 8 epilogue warps have written the tile
          if (m < tile1) {
            if (!(args.debug & 34))
              tma_store_3d(&tmap_y, h_tile, dir * kH + slice * kU, args.b0 + m * 128, pos);   // rows >= B are clipped
            bulk_commit_and_wait();
            LSTM_TRACE(5, it);
            if (args.debug & 4) {
              atomicAdd(args.cnt + dir * args.MT + m, 1);
            } else {
              red_release_gpu_add(args.cnt + dir * args.MT + m, 1);   // the consumer's fence.proxy.async orders its TMA reads
            }
            LSTM_TRACE(6, it);
          }
          mbar_arrive(&pub_free[0]);                      // the tile may be overwritten
        }
      }
    } else {
      // ===================== cell epilogue: thread = sentence, 2 x 12 hidden units per item =====================
      const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      const uint32_t acc_empty_leader0 = map_to_cta(&acc_empty[0], 0);
      const int tile_first = tile0 + (int)rank;
      if (npairs == 1) {
        lstm_epilogue<1>(args, acc_full, acc_empty_leader0, pub_bar, pub_free, h_tile, lane_taddr, warp, lane, dir, slice, tile_first);
      } else if (npairs == 2) {
        lstm_epilogue<2>(args, acc_full, acc_empty_leader0, pub_bar, pub_free, h_tile, lane_taddr, warp, lane, dir, slice, tile_first);
      } else if constexpr (MAXNP >= 4) {
        if (npairs == 3)
          lstm_epilogue<3>(args, acc_full, acc_empty_leader0, pub_bar, pub_free, h_tile, lane_taddr, warp, lane, dir, slice, tile_first);
        else
          lstm_epilogue<4>(args, acc_full, acc_empty_leader0, pub_bar, pub_free, h_tile, lane_taddr, warp, lane, dir, slice, tile_first);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                          // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kSlots * kSlotCols);
  }
}

}  // namespace

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// output sequence as {2H columns, B sentences, S steps}; box {48 units, 128 sentences, 1 step}, no swizzle (dense tile)
static int make_tmap_y_store(icka_handle* h, CUtensorMap* tm, void* y, int B, int S) {
  const cuuint64_t gdim[3] = {(cuuint64_t)(2 * kH), (cuuint64_t)B, (cuuint64_t)S};
  const cuuint64_t gstride[2] = {(cuuint64_t)(2 * kH) * 2, (cuuint64_t)B * (2 * kH) * 2};
  const cuuint32_t box[3] = {(cuuint32_t)kU, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn3>(h->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, y, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) ICKA_FAIL(ICKA_ERR_CUDA, "cuTensorMapEncodeTiled(y store) failed (%d) B=%d S=%d", (int)r, B, S);
  return ICKA_OK;
}

// sentences per launch: a CTA walks at most 4 tiles of 128 per step (cell state of two in registers, of two more in the
// spare TMEM columns) = 8 per pair
static int lstm_groups(int sm_count) { return sm_count / kCtasPerGroup > 0 ? sm_count / kCtasPerGroup : 1; }
static int lstm_chunk(int sm_count) { return 8 * 128 * lstm_groups(sm_count); }

extern "C" int64_t icka_lstm_rec_workspace_bytes(int B, int H) {
  if (B < 0 || H != kH) return -1;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    sms = kCtasPerGroup;
  const size_t chunk = (size_t)lstm_chunk(sms);
  const size_t Bc = (size_t)B < chunk ? (size_t)B : chunk;
  const size_t MT = (Bc + 127) / 128;
  return (int64_t)align_up(2 * MT * sizeof(int), 1024);
}

int icka_lstm_rec1_launch(icka_handle* h, const void* gx, const void* w_hh_perm, void* workspace,
                          int64_t workspace_bytes, void* y, float* h_n, float* c_n, int B, int S, int H, void* stream);

// Which kernel serves a batch of B sentences: 1 = single CTAs, 24-unit slices (lstm_sm100_single.cu: lower step
// latency, one or two sentence tiles), 2 = CTA pairs, 48-unit slices (this file: twice the tensor work per staged
// byte).  The two differ in the column order of gx / w_hh_perm, so the caller asks BEFORE preparing its operands.
extern "C" int icka_lstm_rec_variant(int B) {
  const char* force = getenv("ICKA_LSTM_VARIANT");
  if (force && (atoi(force) == 1 || atoi(force) == 2)) return atoi(force);
  return B <= 256 ? 1 : 2;
}

static int lstm_rec2_launch(icka_handle* h, const void* gx, const void* w_hh_perm, void* workspace,
                            int64_t workspace_bytes, void* y, float* h_n, float* c_n, int B, int S, int H,
                            void* stream);

extern "C" int icka_lstm_rec_fwd(icka_handle* h, const void* gx, const void* w_hh_perm, void* workspace,
                                 int64_t workspace_bytes, void* y, float* h_n, float* c_n, int B, int S, int H,
                                 int variant, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(variant == 1 || variant == 2, "lstm_rec: variant %d (ask icka_lstm_rec_variant)", variant);
  if (variant == 1) return icka_lstm_rec1_launch(h, gx, w_hh_perm, workspace, workspace_bytes, y, h_n, c_n, B, S, H, stream);
  return lstm_rec2_launch(h, gx, w_hh_perm, workspace, workspace_bytes, y, h_n, c_n, B, S, H, stream);
}

static int lstm_rec2_launch(icka_handle* h, const void* gx, const void* w_hh_perm, void* workspace,
                            int64_t workspace_bytes, void* y, float* h_n, float* c_n, int B, int S, int H,
                            void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(H == kH, "lstm_rec: hidden size %d not supported by the persistent kernel (built for %d)", H, kH);
  ICKA_REQUIRE(B >= 0 && S >= 1, "lstm_rec: bad shape B=%d S=%d", B, S);
  ICKA_REQUIRE(gx && w_hh_perm && workspace && y, "lstm_rec: null pointer");
  ICKA_REQUIRE(icka_aligned(gx, 16) && icka_aligned(w_hh_perm, 16) && icka_aligned(workspace, 1024) &&
                   icka_aligned(y, 16),
               "lstm_rec: pointers must be 16-byte aligned (workspace: 1024)");
  ICKA_REQUIRE(!h_n || icka_aligned(h_n, 16), "lstm_rec: h_n must be 16-byte aligned");
  ICKA_REQUIRE(!c_n || icka_aligned(c_n, 16), "lstm_rec: c_n must be 16-byte aligned");
  if (B == 0) return ICKA_OK;
  ICKA_REQUIRE(h->sm_count >= kCtasPerGroup, "lstm_rec: needs %d co-resident CTAs, device has %d SMs", kCtasPerGroup,
               h->sm_count);
  const int chunk = lstm_chunk(h->sm_count);
  const int Bc_max = B < chunk ? B : chunk;
  const size_t MT_max = ((size_t)Bc_max + 127) / 128;
  const size_t cnt_bytes = align_up(2 * MT_max * sizeof(int), 1024);
  ICKA_REQUIRE((size_t)workspace_bytes >= cnt_bytes, "lstm_rec: workspace of %lld B, need %lld", (long long)workspace_bytes,
               (long long)cnt_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  CUtensorMap tw, th, ty;
  int rc = icka_make_tmap_bf16(h, &tw, w_hh_perm, 2 * 4 * kH, kH, kH, kNHalf);
  if (rc) return rc;
  // the A operand of step t is read straight out of the output sequence: [S*B rows, 2H columns], box 128 x 64
  rc = icka_make_tmap_bf16(h, &th, y, (int64_t)S * B, 2 * kH, 2 * kH, 128);
  if (rc) return rc;
  rc = make_tmap_y_store(h, &ty, y, B, S);
  if (rc) return rc;
  ICKA_CUDA(cudaFuncSetAttribute(lstm_rec_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  ICKA_CUDA(cudaFuncSetAttribute(lstm_rec_tcgen05_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const char* dbg = getenv("ICKA_LSTM_DEBUG");

  // sentences are independent recurrences: launches of <= `chunk` sentences, one after the other on the stream
  for (int b0 = 0; b0 < B; b0 += chunk) {
    LstmArgs args;
    args.B = (B - b0 < chunk) ? B - b0 : chunk;
    args.S = S;
    args.Bn = B;
    args.b0 = b0;
    args.MT = (args.B + 127) / 128;
    int groups = lstm_groups(h->sm_count);
    const int pair_tiles = (args.MT + 1) / 2;            // a pair walks two tiles at a time
    if (groups > pair_tiles) groups = pair_tiles;
    args.TPG = 2 * ((pair_tiles + groups - 1) / groups);
    groups = (args.MT + args.TPG - 1) / args.TPG;
    args.cnt = reinterpret_cast<int*>(ws);
    args.gx = static_cast<const __nv_bfloat16*>(gx) + (size_t)b0 * (8 * kH);     // time-major: row = t * B + b
    args.y = static_cast<__nv_bfloat16*>(y) + (size_t)b0 * (2 * kH);
    args.h_n = h_n ? h_n + (size_t)b0 * kH : nullptr;
    args.c_n = c_n ? c_n + (size_t)b0 * kH : nullptr;
    args.debug = dbg ? atoi(dbg) : 0;
    {
      const char* tr = getenv("ICKA_LSTM_TRACE");   // developer timeline: device address of a >= 256 KB buffer
      args.trace = tr ? reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 0)) : nullptr;
    }
    ICKA_CUDA(cudaMemsetAsync(ws, 0, cnt_bytes, st));   // arrival counters
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * kCtasPerGroup);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    if (args.TPG <= 4) ICKA_CUDA(cudaLaunchKernelEx(&cfg, lstm_rec_tcgen05_kernel<2>, tw, th, ty, args));
    else ICKA_CUDA(cudaLaunchKernelEx(&cfg, lstm_rec_tcgen05_kernel<4>, tw, th, ty, args));
    ICKA_LAUNCHED(h);
  }
  return ICKA_OK;
}
