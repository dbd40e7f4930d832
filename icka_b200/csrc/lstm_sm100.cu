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
//     32 slice CTAs of its direction) from L2 through a 5-stage TMA ring as the A operand, accumulates the
//     128 x 96 gate pre-activations in TMEM (4 accumulator slots), and 8 epilogue warps (thread = sentence, two
//     column halves) add Gx, apply the cell update out of TMEM and write h_t (bf16, next step's operand), c_t
//     (fp32) and the output sequence;
//   * sentence tiles are independent recurrences, so a CTA walks items in (t, m) order and only waits for
//     "all 32 slices have published h_{t-1} of tile m" -- a per-(direction, tile) arrival counter in global memory
//     (red.release / ld.acquire + fence.proxy.async before the TMA reads).  With several tiles per CTA the wait
//     for tile m overlaps the work on the other tiles; no grid-wide barrier exists.
// All CTAs must be co-resident (they wait on each other): the kernel is launched cooperatively.
//
// Column order inside a slice (chosen on the host when the weights are permuted once):
//   column c = half * 48 + gate * 12 + j   <->   hidden unit  slice * 24 + half * 12 + j,  gate in (i, f, g, o)
// so each epilogue thread reads one contiguous block of 48 TMEM columns and 96 contiguous bytes of Gx.
#include "common.cuh"
#include "sm100_ptx.cuh"

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
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr size_t kSmemBytes = (size_t)kWBytes + (size_t)kStages * kABytes + 1024 /*align*/ + 256 /*barriers*/;

struct LstmArgs {
  const __nv_bfloat16* gx;   // [B*S, 2*4H] bf16, columns ordered [dir][slice][half][gate][12]
  int* cnt;                  // [2][MT] arrival counters (zeroed before launch)
  __nv_bfloat16* hbuf;       // [2 parity][2 dir][Bp][H] bf16, parity 0 zeroed (h_{-1} = 0)
  float* cbuf;               // [2 dir][Bp][H] fp32, zeroed (c_{-1} = 0)
  __nv_bfloat16* y;          // [B, S, 2H] bf16: forward states in [:H], backward in [H:]
  float* h_n;                // [2, B, H] fp32 or null
  float* c_n;                // [2, B, H] fp32 or null
  int B, S, Bp, MT, TPG;     // sentences, steps, padded sentences (MT * 128), 128-row tiles, tiles per group
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

__global__ void __launch_bounds__(kThreads, 1)
lstm_rec_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h,
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
  uint64_t* w_bar = acc_empty + kSlots;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int slice = blockIdx.x % kNS;
  const int dir = (blockIdx.x / kNS) & 1;
  const int group = blockIdx.x / (2 * kNS);
  const int tile0 = group * args.TPG;
  const int tile1 = min(args.MT, tile0 + args.TPG);
  const int S = args.S, Bp = args.Bp;

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
            if (t > 0) {
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
            const int arow = ((t & 1) * 2 + dir) * Bp + m * 128;
            for (int kb = 0; kb < kKB; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], kABytes);
              tma_load_2d(smem_a + (size_t)stage * kABytes, &tmap_h, &full_bar[stage], kb * 64, arow);
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
        const uint32_t w_addr = smem_u32(smem_w);
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
            for (int kb = 0; kb < kKB; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * kABytes);
              const uint32_t b_addr = w_addr + (uint32_t)kb * kWChunkBytes;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_d, make_kmajor_sw128_desc(a_addr + k * 32), make_kmajor_sw128_desc(b_addr + k * 32), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit(&empty_bar[stage]);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&acc_full[slot]);
          }
        }
      }
    } else {
      // ===================== cell epilogue: thread = sentence, 12 hidden units =====================
      const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
      const int half = (warp - 2) >> 2;          // which 12 of the slice's 24 units
      const int unit0 = slice * kU + half * 12;
      const size_t gx_col = (size_t)((dir * kNS + slice) * 2 + half) * 48;
      int it = 0;
      for (int t = 0; t < S; ++t) {
        const int pos = dir ? (S - 1 - t) : t;
        for (int m = tile0; m < tile1; ++m, ++it) {
          const int slot = it % kSlots;
          const uint32_t slot_phase = (it / kSlots) & 1;
          const int row = m * 128 + quad * 32 + lane;
          const bool valid = row < args.B;
          // everything that does not depend on the recurrence is fetched before the accumulator is waited for
          uint32_t gw[24];
          float c[12];
          float* cp = args.cbuf + ((size_t)dir * Bp + row) * kH + unit0;
          if (valid) {
            const uint4* gp = reinterpret_cast<const uint4*>(args.gx + ((size_t)row * S + pos) * (8 * kH) + gx_col);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              const uint4 v = __ldg(gp + q);
              gw[4 * q] = v.x;
              gw[4 * q + 1] = v.y;
              gw[4 * q + 2] = v.z;
              gw[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(cp + 4 * q);
              c[4 * q] = v.x;
              c[4 * q + 1] = v.y;
              c[4 * q + 2] = v.z;
              c[4 * q + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 24; ++q) gw[q] = 0u;
#pragma unroll
            for (int q = 0; q < 12; ++q) c[q] = 0.0f;
          }
          mbar_wait(&acc_full[slot], slot_phase);
          tc_fence_after();
          uint32_t r[48];
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * kSlotCols + half * 48);
          tmem_ld16(taddr, &r[0]);
          tmem_ld16(taddr + 16, &r[16]);
          tmem_ld16(taddr + 32, &r[32]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[slot]);   // the MMA warp may reuse the slot

          float hv[12];
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            float pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int e = g * 12 + j;
              const uint32_t w = gw[e >> 1];
              const float gxv = __uint_as_float((e & 1) ? (w & 0xffff0000u) : (w << 16));
              pre[g] = __uint_as_float(r[e]) + gxv;
            }
            const float ig = sigmoid_fast(pre[0]), fg = sigmoid_fast(pre[1]), gg = tanh_fast(pre[2]),
                        og = sigmoid_fast(pre[3]);
            c[j] = fmaf(fg, c[j], ig * gg);
            hv[j] = og * tanh_fast(c[j]);
          }
          if (valid) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
              *reinterpret_cast<float4*>(cp + 4 * q) = make_float4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
            uint2 hw[3];
#pragma unroll
            for (int q = 0; q < 3; ++q)
              hw[q] = make_uint2(pack_bf16x2(hv[4 * q], hv[4 * q + 1]), pack_bf16x2(hv[4 * q + 2], hv[4 * q + 3]));
            uint2* hp = reinterpret_cast<uint2*>(args.hbuf + ((size_t)(((t + 1) & 1) * 2 + dir) * Bp + row) * kH + unit0);
            uint2* yp = reinterpret_cast<uint2*>(args.y + ((size_t)row * S + pos) * (2 * kH) + dir * kH + unit0);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              hp[q] = hw[q];
              yp[q] = hw[q];
            }
            if (t == S - 1) {
              if (args.h_n) {
                float4* o = reinterpret_cast<float4*>(args.h_n + ((size_t)dir * args.B + row) * kH + unit0);
#pragma unroll
                for (int q = 0; q < 3; ++q) o[q] = make_float4(hv[4 * q], hv[4 * q + 1], hv[4 * q + 2], hv[4 * q + 3]);
              }
              if (args.c_n) {
                float4* o = reinterpret_cast<float4*>(args.c_n + ((size_t)dir * args.B + row) * kH + unit0);
#pragma unroll
                for (int q = 0; q < 3; ++q) o[q] = make_float4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
              }
            }
          }
          // publish: every epilogue thread's h stores are ordered before the one arrival of this CTA
          __threadfence();
          asm volatile("bar.sync 1, %0;\n" ::"n"(32 * kEpiWarps) : "memory");
          if (threadIdx.x == 64) {
            fence_proxy_async_all();
            red_release_gpu_add(args.cnt + dir * args.MT + m, 1);
          }
        }
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

extern "C" int64_t icka_lstm_rec_workspace_bytes(int B, int H) {
  if (B < 0 || H != kH) return -1;
  const size_t MT = ((size_t)B + 127) / 128, Bp = MT * 128;
  return (int64_t)(align_up(2 * MT * sizeof(int), 1024) + 4 * Bp * kH * 2 + 2 * Bp * kH * 4);
}

extern "C" int icka_lstm_rec_fwd(icka_handle* h, const void* gx, const void* w_hh_perm, void* workspace,
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
  ICKA_REQUIRE(workspace_bytes >= icka_lstm_rec_workspace_bytes(B, H), "lstm_rec: workspace of %lld B, need %lld",
               (long long)workspace_bytes, (long long)icka_lstm_rec_workspace_bytes(B, H));
  ICKA_REQUIRE(h->sm_count >= 2 * kNS, "lstm_rec: needs %d co-resident CTAs, device has %d SMs", 2 * kNS, h->sm_count);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  LstmArgs args;
  args.B = B;
  args.S = S;
  args.MT = (B + 127) / 128;
  args.Bp = args.MT * 128;
  int groups = h->sm_count / (2 * kNS);
  if (groups > args.MT) groups = args.MT;
  args.TPG = (args.MT + groups - 1) / groups;
  groups = (args.MT + args.TPG - 1) / args.TPG;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t cnt_bytes = align_up(2 * (size_t)args.MT * sizeof(int), 1024);
  const size_t h_bytes = 4 * (size_t)args.Bp * kH * 2, c_bytes = 2 * (size_t)args.Bp * kH * 4;
  args.cnt = reinterpret_cast<int*>(ws);
  args.hbuf = reinterpret_cast<__nv_bfloat16*>(ws + cnt_bytes);
  args.cbuf = reinterpret_cast<float*>(ws + cnt_bytes + h_bytes);
  args.gx = static_cast<const __nv_bfloat16*>(gx);
  args.y = static_cast<__nv_bfloat16*>(y);
  args.h_n = h_n;
  args.c_n = c_n;
  ICKA_CUDA(cudaMemsetAsync(ws, 0, cnt_bytes + h_bytes + c_bytes, st));   // counters, h_{-1} = 0, c_{-1} = 0

  CUtensorMap tw, th;
  int rc = icka_make_tmap_bf16(h, &tw, w_hh_perm, 2 * 4 * kH, kH, kH, kN);
  if (rc) return rc;
  rc = icka_make_tmap_bf16(h, &th, args.hbuf, 4 * (int64_t)args.Bp, kH, kH, 128);
  if (rc) return rc;

  ICKA_CUDA(cudaFuncSetAttribute(lstm_rec_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  void* kargs[3] = {&tw, &th, &args};
  ICKA_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_rec_tcgen05_kernel), dim3(groups * 2 * kNS),
                                        dim3(kThreads), kargs, kSmemBytes, st));
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
