// SINGLE-CTA variant of the persistent recurrent kernel (variant 1 of icka_lstm_rec_fwd; the CTA-pair variant and the
// design notes shared by both live in lstm_sm100.cu).  Used for small batches (<= 256 sentences: one or two sentence
// tiles, where the step-to-step latency matters more than tensor throughput -- 8.2 vs 10.7 us per step) and wherever a
// cooperative launch of 2-CTA clusters is not available.
//
// Recurrent half of the emission head's BiLSTM (CMIM:905-908 `nn.LSTM(H, H, batch_first, bidirectional)`, call
// CMIM:1042) as ONE persistent, weight-stationary tcgen05 kernel for sm_100a (SURVEY 8f "next" row 1).
//
//   gates_t = Gx[:, t] + h_{t-1} . W_hh^T          Gx = x . W_ih^T + b_ih + b_hh, one big tensor-core GEMM (icka_linear_fwd)
//   i, f, g, o = sigmoid, sigmoid, tanh, sigmoid    (PyTorch gate order)
//   c_t = f * c_{t-1} + i * g ;  h_t = o * tanh(c_t)
//
// The recurrence is S serial steps of a [B, H] x [H, 4H] product per direction.  Launching a GEMM per step would
// stream W_hh (4.7 MB bf16 per direction) from L2 S times per tile and pay a launch + pipeline fill per step; here
//   * the 2 x 3072 x 768 weights are split into 2 x 32 slices of 24 hidden units (96 gate columns, 147 KB bf16) and
//     each slice stays in the shared memory of ONE CTA for the whole sequence (loaded once by TMA, SWIZZLE_128B,
//     K-major B operand of tcgen05.mma);
//   * a work item is (step t, 128-sentence tile m): the CTA streams h_{t-1}[tile m] (128 x 768 bf16, written by the
//     32 slice CTAs of its direction) from L2 through a 5-stage TMA ring as the A operand -- straight out of the
//     time-major output sequence y[t-1], so h is written exactly once -- accumulates the 128 x 96 gate
//     pre-activations in TMEM (4 accumulator slots), and 8 epilogue warps (thread = sentence, two column halves)
//     add Gx (prefetched one item ahead), apply the cell update out of TMEM with the cell state in registers and
//     write h_t (bf16) into y[t];
//   * a publisher warp turns "all 8 epilogue warps stored their part of item (t, m)" into ONE gpu-scope release on
//     the tile's arrival counter, so the ~4 us a MEMBAR.GPU takes on a busy SM never stalls the cell arithmetic;
//   * sentence tiles are independent recurrences, so a CTA walks items in (t, m) order and only waits for
//     "all 32 slices have published h_{t-1} of tile m" -- a per-(direction, tile) arrival counter in global memory
//     (red.release / ld.acquire + fence.proxy.async before the TMA reads).  With several tiles per CTA the wait
//     for tile m overlaps the work on the other tiles; no grid-wide barrier exists.
// All CTAs must be co-resident (they wait on each other): the kernel is launched cooperatively.
//
// Column order inside a slice (chosen on the host when the weights are permuted once):
//   column c = half * 48 + jg * 16 + gate * 4 + jj   <->   hidden unit  slice * 24 + half * 12 + jg * 4 + jj,
//   gate in (i, f, g, o), jg in 0..2, jj in 0..3
// so each epilogue thread reads one contiguous block of 48 TMEM columns (three 16-column groups, each holding the
// four gates of four units) and 96 contiguous bytes of Gx.
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <stdlib.h>

namespace {

using namespace sm100;

constexpr int kH = 768;                       // hidden size this kernel is built for
constexpr int kU = 24;                        // hidden units per CTA
constexpr int kNS = kH / kU;                  // 32 slices per direction
constexpr int kN = 4 * kU;                    // 96 gate columns per CTA (UMMA N)
constexpr int kKB = kH / 64;                  // 12 k-chunks of 64 bf16 = 128 B
constexpr int kWChunkBytes = kN * 128;        // 12,288
constexpr int kWBytes = kKB * kWChunkBytes;   // 147,456
constexpr int kABytes = 128 * 128;            // one 128-row x 64-k A tile
constexpr int kStages = 5;
constexpr int kSlots = 4;                     // TMEM accumulator slots
constexpr int kSlotCols = 128;                // column stride between slots (96 used)
constexpr int kEpiWarps = 8;
constexpr int kPubWarp = 2 + kEpiWarps;       // warp 10: publishes finished items
constexpr int kThreads = 32 * (kPubWarp + 1);
constexpr size_t kSmemBytes = (size_t)kWBytes + (size_t)kStages * kABytes + 1024 /*align*/ + 256 /*barriers*/;

struct LstmArgs {
  const __nv_bfloat16* gx;   // [S*Bn, 2*4H] bf16 TIME-MAJOR (row = t * Bn + sentence), columns ordered [dir][slice][half][jg][gate][4]
  int* cnt;                  // [2][MT] arrival counters (zeroed before launch)
  __nv_bfloat16* y;          // [S, Bn, 2H] bf16 TIME-MAJOR: forward states in [:H], backward in [H:]
  float* h_n;                // [2, Bn, H] fp32 or null (already offset to this launch's first sentence)
  float* c_n;                // [2, Bn, H] fp32 or null
  int Bn;                    // sentences of the whole call (pitch of the time planes of gx / y and of the direction
                             // planes of h_n / c_n)
  int B, S, b0, MT, TPG;     // sentences of this launch, steps, first sentence of this launch, 128-row tiles, tiles
                             // per CTA group
  int debug;                 // developer probes (ICKA_LSTM_DEBUG): 1 = no dependency wait, 2 = no cell arithmetic /
                             // state stores, 4 = publish without the gpu-scope release, 8 = no A loads, 16 = no MMAs
                             // (results are WRONG)
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

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// Cell epilogue of one CTA: NT sentence tiles, walked in (step, tile) order.  Thread = sentence row (TMEM lane) and 12
// of the slice's 24 hidden units; the cell state of its NT rows lives in registers for the whole sequence.
template <int NT>
__device__ __forceinline__ void lstm_epilogue(const LstmArgs& args, uint64_t* acc_full, uint64_t* acc_empty,
                                              uint64_t* pub_bar, uint64_t* pub_free, uint32_t lane_taddr, int warp,
                                              int lane, int dir, int slice, int tile0) {
  const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
  const int half = (warp - 2) >> 2;          // which 12 of the slice's 24 units
  const int unit0 = slice * kU + half * 12;
  const size_t gx_col = (size_t)((dir * kNS + slice) * 2 + half) * 48;
  const int S = args.S;
  const int row0 = tile0 * 128 + quad * 32 + lane;
  float c[NT][12];
#pragma unroll
  for (int ti = 0; ti < NT; ++ti)
#pragma unroll
    for (int q = 0; q < 12; ++q) c[ti][q] = 0.0f;
  // Gx does not depend on the recurrence: it is fetched ONE ITEM AHEAD, off the step-to-step critical path
  uint32_t gw_nxt[24];
  auto fetch = [&](int t, int ti, uint32_t (&gw)[24]) {
    const int pos = dir ? (S - 1 - t) : t;
    const int row = row0 + ti * 128;
    if (row < args.B) {
      const uint4* gp = reinterpret_cast<const uint4*>(args.gx + ((size_t)pos * args.Bn + row) * (8 * kH) + gx_col);
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
  fetch(0, 0, gw_nxt);
  int it = 0;
  for (int t = 0; t < S; ++t) {
    const int pos = dir ? (S - 1 - t) : t;
#pragma unroll
    for (int ti = 0; ti < NT; ++ti, ++it) {
      const int slot = it % kSlots;
      const uint32_t slot_phase = (it / kSlots) & 1;
      const int row = row0 + ti * 128;
      const bool valid = row < args.B;
      uint32_t gw[24];
#pragma unroll
      for (int q = 0; q < 24; ++q) gw[q] = gw_nxt[q];
      if (ti + 1 < NT) fetch(t, ti + 1, gw_nxt);
      else if (t + 1 < S) fetch(t + 1, 0, gw_nxt);

      mbar_wait(&acc_full[slot], slot_phase);
      tc_fence_after();
      // 48 accumulator columns = 3 groups of 4 units x (i, f, g, o): 16 columns are live at a time, the next group's
      // tcgen05.ld is in flight while this one is being computed
      const uint32_t taddr = lane_taddr + (uint32_t)(slot * kSlotCols + half * 48);
      const bool probe = (args.debug & 2) != 0;      // probe: no cell arithmetic, no state stores
      uint32_t ra[16], rb[16];
      float hv[12];
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
          c[ti][j] = probe ? c[ti][j] : fmaf(fg, c[ti][j], ig * gg);
          hv[j] = og * tanh_fast(c[ti][j]);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[slot]);   // the MMA warp may reuse the slot
      cell4(ra, 2);
      if (valid && !(args.debug & 2)) {
        uint32_t hw[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) hw[q] = pack_bf16x2(hv[2 * q], hv[2 * q + 1]);
        // 24 bytes per row at byte offset 48 * slice + 24 * half: one 16-byte and one 8-byte store, ordered so the
        // 16-byte one is aligned (half 0: 16 + 8, half 1: 8 + 16)
        uint8_t* yp = reinterpret_cast<uint8_t*>(args.y + ((size_t)pos * args.Bn + row) * (2 * kH) + dir * kH + unit0);
        if (half == 0) {
          *reinterpret_cast<uint4*>(yp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint2*>(yp + 16) = make_uint2(hw[4], hw[5]);
        } else {
          *reinterpret_cast<uint2*>(yp) = make_uint2(hw[0], hw[1]);
          *reinterpret_cast<uint4*>(yp + 8) = make_uint4(hw[2], hw[3], hw[4], hw[5]);
        }
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
              o[q] = make_float4(c[ti][4 * q], c[ti][4 * q + 1], c[ti][4 * q + 2], c[ti][4 * q + 3]);
          }
        }
      }
      // hand the item to the publisher warp: this warp's h stores are ordered before the arrival (release.cta)
      __syncwarp();
      if (lane == 0) {
        mbar_wait(&pub_free[slot], ((it / kSlots) & 1) ^ 1);   // the publisher is done with this slot's previous item
        mbar_arrive(&pub_bar[slot]);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
lstm_rec1_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
                        const LstmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + (size_t)kStages * kABytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;
  uint64_t* acc_empty = acc_full + kSlots;
  uint64_t* pub_bar = acc_empty + kSlots;
  uint64_t* pub_free = pub_bar + kSlots;
  uint64_t* w_bar = pub_free + kSlots;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int slice = blockIdx.x % kNS;
  const int dir = (blockIdx.x / kNS) & 1;
  const int group = blockIdx.x / (2 * kNS);
  const int tile0 = group * args.TPG;
  const int tile1 = min(args.MT, tile0 + args.TPG);
  const int S = args.S;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_h);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kSlots; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], kEpiWarps);
      mbar_init(&pub_bar[a], kEpiWarps);
      mbar_init(&pub_free[a], 1);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_slot, kSlots * kSlotCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (tile0 < tile1) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        // the weight slice: resident for the whole sequence
        mbar_arrive_expect_tx(w_bar, kWBytes);
        const int wrow = (dir * kNS + slice) * kN;
        for (int kb = 0; kb < kKB; ++kb) tma_load_2d(smem_w + (size_t)kb * kWChunkBytes, &tmap_w, w_bar, kb * 64, wrow);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < S; ++t) {
          for (int m = tile0; m < tile1; ++m) {
            if (t > 0 && !(args.debug & 1)) {
              // h_{t-1} of this tile is complete once all slices of this direction have arrived t times
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
            // sequence; at t = 0 an out-of-bounds row makes TMA deliver zeros (h_{-1} = 0)
            const int prev = dir ? (S - t) : (t - 1);
            const int arow = (t == 0) ? S * args.Bn : prev * args.Bn + args.b0 + m * 128;
            for (int kb = 0; kb < kKB; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (args.debug & 8) {
                mbar_arrive(&full_bar[stage]);           // probe: no A traffic
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], kABytes);
                tma_load_2d(smem_a + (size_t)stage * kABytes, &tmap_h, &full_bar[stage], dir * kH + kb * 64, arow);
              }
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16_f32(128, kN);
        mbar_wait(w_bar, 0);
        tc_fence_after();
        // One thread issues ~50 short (N = 96) MMAs per item: the loop must cost less than the MMAs themselves, so
        // the shared-memory descriptors are formed once and stepped by adding to their 14-bit address field.
        const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(smem_a));
        const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(smem_w));
        const bool no_mma = (args.debug & 16) != 0;       // probe
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = 0; t < S; ++t) {
          for (int m = tile0; m < tile1; ++m, ++it) {
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
                umma_bf16(tmem_d, a_desc, b_desc, idesc, kb > 0 ? 1u : 0u);
                umma_bf16(tmem_d, a_desc + 2, b_desc + 2, idesc, 1u);
                umma_bf16(tmem_d, a_desc + 4, b_desc + 4, idesc, 1u);
                umma_bf16(tmem_d, a_desc + 6, b_desc + 6, idesc, 1u);
              }
              umma_commit(&empty_bar[stage]);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&acc_full[slot]);
          }
        }
      }
    } else if (warp == kPubWarp) {
      // ===================== publisher: one gpu-scope release per item, off the epilogue warps' path =====================
      // The epilogue warps arrive on pub_bar[slot] (release.cta) after their h stores; this thread acquires the
      // barrier and performs the ONE gpu-scope release of the CTA (cumulative over everything that happened-before
      // it) -- so the ~1 us a MEMBAR.GPU takes never stalls the warps that do the cell arithmetic.
      if (lane == 0) {
        const int ntiles = tile1 - tile0;
        const int items = S * ntiles;
        for (int it = 0; it < items; ++it) {
          const int m = tile0 + it % ntiles;
          const int slot = it % kSlots;
          mbar_wait(&pub_bar[slot], (it / kSlots) & 1);
          if (args.debug & 4) {
            atomicAdd(args.cnt + dir * args.MT + m, 1);
          } else {
            red_release_gpu_add(args.cnt + dir * args.MT + m, 1);   // the consumer's fence.proxy.async orders its TMA reads
          }
          mbar_arrive(&pub_free[slot]);
        }
      }
    } else {
      // ===================== cell epilogue: thread = sentence, 12 hidden units =====================
      const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      switch (tile1 - tile0) {
        case 1: lstm_epilogue<1>(args, acc_full, acc_empty, pub_bar, pub_free, lane_taddr, warp, lane, dir, slice, tile0); break;
        case 2: lstm_epilogue<2>(args, acc_full, acc_empty, pub_bar, pub_free, lane_taddr, warp, lane, dir, slice, tile0); break;
        case 3: lstm_epilogue<3>(args, acc_full, acc_empty, pub_bar, pub_free, lane_taddr, warp, lane, dir, slice, tile0); break;
        default: lstm_epilogue<4>(args, acc_full, acc_empty, pub_bar, pub_free, lane_taddr, warp, lane, dir, slice, tile0); break;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kSlots * kSlotCols);
  }
}

}  // namespace

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// sentences per launch: the cell state of a CTA's tiles lives in registers, at most 4 tiles of 128 per CTA
static int lstm_chunk(int sm_count) { return 4 * 128 * (sm_count / (2 * kNS) > 0 ? sm_count / (2 * kNS) : 1); }

// variant 1 launcher (called by icka_lstm_rec_fwd in lstm_sm100.cu, which has validated the arguments)
int icka_lstm_rec1_launch(icka_handle* h, const void* gx, const void* w_hh_perm, void* workspace,
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
  ICKA_REQUIRE(h->sm_count >= 2 * kNS, "lstm_rec: needs %d co-resident CTAs, device has %d SMs", 2 * kNS, h->sm_count);
  const int chunk = lstm_chunk(h->sm_count);
  const int Bc_max = B < chunk ? B : chunk;
  const size_t MT_max = ((size_t)Bc_max + 127) / 128;
  const size_t cnt_bytes = align_up(2 * MT_max * sizeof(int), 1024);
  const size_t need = cnt_bytes;
  ICKA_REQUIRE((size_t)workspace_bytes >= need, "lstm_rec: workspace of %lld B, need %lld", (long long)workspace_bytes,
               (long long)need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  CUtensorMap tw, th;
  int rc = icka_make_tmap_bf16(h, &tw, w_hh_perm, 2 * 4 * kH, kH, kH, kN);
  if (rc) return rc;
  // the A operand of step t is read straight out of the output sequence: [S*B rows, 2H columns], box 128 x 64
  rc = icka_make_tmap_bf16(h, &th, y, (int64_t)S * B, 2 * kH, 2 * kH, 128);
  if (rc) return rc;
  ICKA_CUDA(cudaFuncSetAttribute(lstm_rec1_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const char* dbg = getenv("ICKA_LSTM_DEBUG");

  // sentences are independent recurrences: launches of <= `chunk` sentences, one after the other on the stream
  for (int b0 = 0; b0 < B; b0 += chunk) {
    LstmArgs args;
    args.B = (B - b0 < chunk) ? B - b0 : chunk;
    args.S = S;
    args.Bn = B;
    args.MT = (args.B + 127) / 128;
    args.b0 = b0;
    int groups = h->sm_count / (2 * kNS);
    if (groups > args.MT) groups = args.MT;
    args.TPG = (args.MT + groups - 1) / groups;
    groups = (args.MT + args.TPG - 1) / args.TPG;
    args.cnt = reinterpret_cast<int*>(ws);
    args.gx = static_cast<const __nv_bfloat16*>(gx) + (size_t)b0 * (8 * kH);     // time-major: row = t * B + b
    args.y = static_cast<__nv_bfloat16*>(y) + (size_t)b0 * (2 * kH);
    args.h_n = h_n ? h_n + (size_t)b0 * kH : nullptr;
    args.c_n = c_n ? c_n + (size_t)b0 * kH : nullptr;
    args.debug = dbg ? atoi(dbg) : 0;
    ICKA_CUDA(cudaMemsetAsync(ws, 0, cnt_bytes, st));   // arrival counters
    void* kargs[3] = {&tw, &th, &args};
    ICKA_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_rec1_tcgen05_kernel), dim3(groups * 2 * kNS),
                                          dim3(kThreads), kargs, kSmemBytes, st));
    ICKA_LAUNCHED(h);
  }
  return ICKA_OK;
}
