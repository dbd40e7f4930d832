// Training (BPTT) of the emission head's BiLSTM (CMIM:905-908, 1042) for bf16 operands: ONE launch per time step that does
// the recurrent product AND the cell arithmetic of BOTH directions.
//
// The first training path issued, per direction and step, a tcgen05 GEMM launch (h_{t-1} . W_hh^T, 128-row tiles for a
// batch of 32) and a cell kernel, on two streams: ~21 us per step, 5 of the 6.3 ms of a B = 32 training step with the
// reference's head.  A step is a [B x 768] x [768 x 3072] product per direction -- 0.15 GFLOP at B = 32 -- followed by
// element-wise work on the result, so the whole step fits one small kernel:
//   forward   CTA (slice, dir) owns 16 hidden units = 64 gate rows of W_hh (98 KB bf16, staged by cp.async, L2-resident
//             across steps): gates = h_{t-1} . W_slice^T on mma.sync m16n8k16 with the n-tiles ordered so that a thread ends
//             up holding i, f, g, o of ITS units, then the cell update in registers: c_t, h_t (written straight into the
//             output sequence, which is also the next step's A operand), and the saved gate activations.
//   backward  the same kernel shape on the transposed weights: dh_rec[:, slice] = dpre_{t+1} . W_hh[:, slice]
//             (K = 3072; the A operand streams through a 6-stage ring of 256-wide chunks), then the cell backward of the slice:
//             dpre_t goes into the time-major [S, B, 8H] gradient buffer that the next launch reads as its A operand
//             and that the big weight / input gradient GEMMs consume afterwards.
// Dependencies between steps are kernel boundaries; 2 x 128 launches replace ~1000, and a CUDA graph replays them.
// (The persistent tcgen05 kernel of lstm_sm100.cu serves inference, where thousands of sentences give each CTA four
// independent step chains to overlap; a training batch of 32-128 sentences is one chain, bound by step latency.)
#include "common.cuh"

namespace {

constexpr int kTH = 768;            // hidden size (the reference's H; other sizes keep the per-step GEMM path)
constexpr int kUnits = 16;          // hidden units per CTA
constexpr int kSlices = kTH / kUnits;
constexpr int kKC = 256;            // contraction chunk of the A operand
constexpr int kStages = 6;          // A chunks in flight: the step is bound by L2 latency, not bandwidth
constexpr int kRows = 32;           // sentence rows per pass
constexpr int kPitchA = kKC + 8;    // bf16 elements: +16 B per row keeps ldmatrix conflict-free
constexpr int kStepThreads = 256;           // 8 warps = 2 contraction halves x 2 m-tiles x 2 unit halves
constexpr size_t kWBytesMax = (size_t)64 * (kTH + 8) * 2 > (size_t)kUnits * (4 * kTH + 8) * 2
                                  ? (size_t)64 * (kTH + 8) * 2 : (size_t)kUnits * (4 * kTH + 8) * 2;
constexpr size_t kABytes = (size_t)kRows * kPitchA * 2;
constexpr size_t kXchBytes = (size_t)8 * 32 * 8 * sizeof(float);   // partial accumulators exchanged between the two halves
constexpr size_t kStepSmem = kWBytesMax + kStages * kABytes + kXchBytes;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1])
               : "r"((uint32_t)__cvta_generic_to_shared(smem)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

struct StepArgs {
  // forward
  const __nv_bfloat16* gx;     // [S, B, 8H] bf16, TIME-MAJOR (position order): x . W_ih^T + b for both directions
  const __nv_bfloat16* w[2];   // forward: W_hh [4H, H] per direction; backward: W_hh^T [H, 4H] per direction
  __nv_bfloat16* y_op;         // [S, B, 2H] bf16: the output sequence = h of every step (A operand of the next step)
  float* y32;                  // [S, B, 2H] fp32 copy (the autograd output)
  float* acts;                 // [2, S, B, 4H] fp32: i, f, g, o after their nonlinearities, indexed by STEP
  float* c_all;                // [2, S, B, H] fp32 cell states, indexed by step
  // backward
  const float* dy;             // [S, B, 2H] fp32
  __nv_bfloat16* dg;           // [S, B, 8H] bf16 gate pre-activation gradients, time-major
  float* dc;                   // [2, B, H] fp32 running cell gradient (zero before the first launch)
  int B, S, t;
};

// stage `rows` rows of `cols` bf16 (cols % 8 == 0) from global (row pitch ld) into shared memory (row pitch pitch)
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, int pitch, const __nv_bfloat16* src, int64_t ld, int rows,
                                           int cols, int rows_valid) {
  const int per_row = cols / 8;
  for (int i = threadIdx.x; i < rows * per_row; i += kStepThreads) {
    const int r = i / per_row, c = (i % per_row) * 8;
    if (r < rows_valid) cp_async16(dst + (size_t)r * pitch + c, src + (size_t)r * ld + c);
    else *reinterpret_cast<uint4*>(dst + (size_t)r * pitch + c) = make_uint4(0, 0, 0, 0);
  }
}

// BWD = false: N = 64 gate rows (n-tile 2 gate + unit half), K = 768.  BWD = true: N = 16 units, K = 3072.
template <bool BWD>
__global__ void __launch_bounds__(kStepThreads) lstm_step_kernel(const StepArgs args) {
  constexpr int K = BWD ? 4 * kTH : kTH;
  constexpr int NROWS = BWD ? kUnits : 64;
  constexpr int NI = BWD ? 1 : 4;             // n-tiles per warp
  constexpr int NKC = K / kKC;
  constexpr int kPitchW = K + 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem_raw + kWBytesMax);

  const int slice = blockIdx.x, dir = blockIdx.y;
  const int B = args.B, S = args.S, t = args.t;
  const int u0 = slice * kUnits;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 1, ng = (warp >> 1) & 1, kh = warp >> 2;   // kh: which half of every 768-wide contraction chunk
  const int g = lane >> 2, q = lane & 3;
  float* xch = reinterpret_cast<float*>(smem_raw + kWBytesMax + kStages * kABytes);
  const int pos = dir ? S - 1 - t : t;
  // the step whose result feeds this one: forward reads h_{t-1}, backward reads dpre_{t+1}
  const bool has_rec = BWD ? (t < S - 1) : (t > 0);
  const int pos_src = BWD ? (dir ? S - 2 - t : t + 1) : (dir ? S - t : t - 1);
  const int nrb = (B + kRows - 1) / kRows;
  const int nchunks = has_rec ? nrb * NKC : 0;

  // every sequence tensor is TIME-MAJOR ([S, B, .]): a step touches one contiguous block per tensor.  (Batch-major rows are
  // S x 12 KB apart -- one DRAM page and one TLB entry per sentence and tensor: the first version of this kernel spent
  // ~4 us per 32 sentences and step on exactly that.)
  const __nv_bfloat16* a_base = BWD ? args.dg + (size_t)pos_src * B * 8 * kTH + (size_t)dir * 4 * kTH
                                    : args.y_op + (size_t)pos_src * B * 2 * kTH + (size_t)dir * kTH;
  const int64_t a_ld = BWD ? (int64_t)8 * kTH : (int64_t)2 * kTH;
  auto issue_chunk = [&](int c) {
    const int rb = c / NKC, kc = c % NKC;
    const int rows_valid = min(kRows, B - rb * kRows);
    stage_rows(As + (size_t)(c % kStages) * kRows * kPitchA, kPitchA, a_base + (size_t)rb * kRows * a_ld + kc * kKC, a_ld, kRows, kKC,
               rows_valid);
  };
  // Programmatic dependent launch: the next step's grid may start now -- all it does before ITS griddepcontrol.wait is
  // stage weights, which no step writes -- and this grid waits for the previous step to complete (and flush) only here,
  // after its own weight loads are in flight.
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  if (has_rec) {
    // weights: forward rows gate * H + u0 + j (j < 16) of W_hh; backward rows u0 + j of W_hh^T
    const __nv_bfloat16* wsrc = args.w[dir];
    for (int i = threadIdx.x; i < NROWS * (K / 8); i += kStepThreads) {
      const int r = i / (K / 8), c = (i % (K / 8)) * 8;
      const int grow = BWD ? u0 + r : (r >> 4) * kTH + u0 + (r & 15);
      cp_async16(Ws + (size_t)r * kPitchW + c, wsrc + (size_t)grow * K + c);
    }
  }
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  if (has_rec) {
#pragma unroll
    for (int c = 0; c < kStages - 1; ++c) {       // the weights travel in the first group
      if (c < nchunks) issue_chunk(c);
      cp_async_commit();
    }
  }

  const int unit = u0 + ng * 8 + 2 * q;          // this thread's units: unit, unit + 1
  for (int rb = 0; rb < nrb; ++rb) {
    // ---- operands of the cell arithmetic: fetched BEFORE the product so that their latency (gx, the saved activations and
    //      the gradients stream from HBM) hides behind it.  Warp kh finishes row group kh: row g + 8 kh of m-tile mt. ----
    const int row = rb * kRows + mt * 16 + g + 8 * kh;
    const bool live = row < B;
    const size_t st_row = ((size_t)dir * S + t) * B + (live ? row : 0);      // [dir][step][row]
    uint32_t gxw[4] = {0u, 0u, 0u, 0u};
    float2 pf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pf[i] = make_float2(0.0f, 0.0f);
    if (live) {
      if constexpr (!BWD) {
        const __nv_bfloat16* gxp = args.gx + ((size_t)pos * B + row) * 8 * kTH + (size_t)dir * 4 * kTH + unit;
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) gxw[gate] = __ldg(reinterpret_cast<const uint32_t*>(gxp + gate * kTH));
        if (t > 0) pf[0] = *reinterpret_cast<const float2*>(args.c_all + (st_row - B) * kTH + unit);
      } else {
        const float* ap = args.acts + st_row * 4 * kTH + unit;
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) pf[gate] = *reinterpret_cast<const float2*>(ap + gate * kTH);
        pf[4] = *reinterpret_cast<const float2*>(args.c_all + st_row * kTH + unit);
        if (t > 0) pf[5] = *reinterpret_cast<const float2*>(args.c_all + (st_row - B) * kTH + unit);
        pf[6] = *reinterpret_cast<const float2*>(args.dy + ((size_t)pos * B + row) * 2 * kTH + (size_t)dir * kTH + unit);
        pf[7] = *reinterpret_cast<const float2*>(args.dc + ((size_t)dir * B + row) * kTH + unit);
      }
    }
    constexpr int NACC = BWD ? 4 : NI;             // backward: one n-tile, four interleaved accumulators (no dependent chain)
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    if (has_rec) {
      for (int kc = 0; kc < NKC; ++kc) {
        const int c = rb * NKC + kc;
        cp_async_wait<kStages - 2>();              // chunk c has landed (one group is committed per iteration)
        __syncthreads();                           // ... for every thread, and chunk c - 1 has been consumed by all
        if (c + kStages - 1 < nchunks) issue_chunk(c + kStages - 1);
        cp_async_commit();
        const __nv_bfloat16* Ab = As + (size_t)(c % kStages) * kRows * kPitchA;
        const __nv_bfloat16* a_ptr = Ab + (size_t)(mt * 16 + (lane & 15)) * kPitchA + (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < kKC / 32; ++ks) {
          const int k0 = kh * (kKC / 2) + ks * 16;
          uint32_t a[4];
          ldsm_x4(a, a_ptr + k0);
          const int kw = kc * kKC + k0;
          if constexpr (BWD) {
            uint32_t b[2];
            ldsm_x2(b, Ws + (size_t)(ng * 8 + (lane & 7)) * kPitchW + kw + ((lane >> 3) & 1) * 8);
            mma16816(acc[ks & 3], a, b[0], b[1]);
          } else {
            // n-tiles ng, ng + 2 (gates i, f of the warp's 8 units) and ng + 4, ng + 6 (g, o)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              uint32_t b[4];
              const int tile = ng + 2 * (2 * h2 + (lane >> 4));
              ldsm_x4(b, Ws + (size_t)(tile * 8 + (lane & 7)) * kPitchW + kw + ((lane >> 3) & 1) * 8);
              mma16816(acc[2 * h2], a, b[0], b[1]);
              mma16816(acc[2 * h2 + 1], a, b[2], b[3]);
            }
          }
        }
      }
      if constexpr (BWD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][j] = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
      }
      // ---- the two contraction halves meet: each warp hands the partner the partial sums of the partner's row group ----
      float* mine = xch + ((size_t)warp * 32 + lane) * 8;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        mine[2 * i] = acc[i][2 * (1 - kh)];
        mine[2 * i + 1] = acc[i][2 * (1 - kh) + 1];
      }
      __syncthreads();
      const float* theirs = xch + ((size_t)(warp ^ 4) * 32 + lane) * 8;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        acc[i][2 * kh] += theirs[2 * i];
        acc[i][2 * kh + 1] += theirs[2 * i + 1];
      }
    }
    if (!live) continue;
    // ---- cell arithmetic.  MUFU forms (tanh.approx, sigmoid through it), as in the inference kernel: the states are rounded
    //      to bf16 for the next step's product anyway, and backward differentiates the SAVED activations ----
    if constexpr (!BWD) {
      float ig[2], fg[2], gg[2], og[2], cn[2], hn[2];
      const float cpv[2] = {pf[0].x, pf[0].y};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float pre[4];
#pragma unroll
        for (int gate = 0; gate < 4; ++gate)
          pre[gate] = acc[gate][2 * kh + e] + __uint_as_float(e ? (gxw[gate] & 0xffff0000u) : (gxw[gate] << 16));
        ig[e] = sigmoid_fast(pre[0]);
        fg[e] = sigmoid_fast(pre[1]);
        gg[e] = tanh_fast(pre[2]);
        og[e] = sigmoid_fast(pre[3]);
        cn[e] = fg[e] * cpv[e] + ig[e] * gg[e];
        hn[e] = og[e] * tanh_fast(cn[e]);
      }
      *reinterpret_cast<float2*>(args.c_all + st_row * kTH + unit) = make_float2(cn[0], cn[1]);
      float* ap = args.acts + st_row * 4 * kTH + unit;
      *reinterpret_cast<float2*>(ap) = make_float2(ig[0], ig[1]);
      *reinterpret_cast<float2*>(ap + kTH) = make_float2(fg[0], fg[1]);
      *reinterpret_cast<float2*>(ap + 2 * kTH) = make_float2(gg[0], gg[1]);
      *reinterpret_cast<float2*>(ap + 3 * kTH) = make_float2(og[0], og[1]);
      const size_t yo = ((size_t)pos * B + row) * 2 * kTH + (size_t)dir * kTH + unit;
      *reinterpret_cast<uint32_t*>(args.y_op + yo) = pack_bf16x2(hn[0], hn[1]);
      *reinterpret_cast<float2*>(args.y32 + yo) = make_float2(hn[0], hn[1]);
    } else {
      const float igv[2] = {pf[0].x, pf[0].y}, fgv[2] = {pf[1].x, pf[1].y}, ggv[2] = {pf[2].x, pf[2].y},
                  ogv[2] = {pf[3].x, pf[3].y}, cnv[2] = {pf[4].x, pf[4].y}, cpv[2] = {pf[5].x, pf[5].y},
                  dyy[2] = {pf[6].x, pf[6].y}, dcc[2] = {pf[7].x, pf[7].y};
      float d_i[2], d_f[2], d_g[2], d_o[2], dcn[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float tc = tanh_fast(cnv[e]);
        const float dh = dyy[e] + acc[0][2 * kh + e];
        const float dct = dcc[e] + dh * ogv[e] * (1.0f - tc * tc);
        d_i[e] = dct * ggv[e] * igv[e] * (1.0f - igv[e]);
        d_f[e] = dct * cpv[e] * fgv[e] * (1.0f - fgv[e]);
        d_g[e] = dct * igv[e] * (1.0f - ggv[e] * ggv[e]);
        d_o[e] = dh * tc * ogv[e] * (1.0f - ogv[e]);
        dcn[e] = dct * fgv[e];
      }
      *reinterpret_cast<float2*>(args.dc + ((size_t)dir * B + row) * kTH + unit) = make_float2(dcn[0], dcn[1]);
      __nv_bfloat16* dp = args.dg + ((size_t)pos * B + row) * 8 * kTH + (size_t)dir * 4 * kTH + unit;
      *reinterpret_cast<uint32_t*>(dp) = pack_bf16x2(d_i[0], d_i[1]);
      *reinterpret_cast<uint32_t*>(dp + kTH) = pack_bf16x2(d_f[0], d_f[1]);
      *reinterpret_cast<uint32_t*>(dp + 2 * kTH) = pack_bf16x2(d_g[0], d_g[1]);
      *reinterpret_cast<uint32_t*>(dp + 3 * kTH) = pack_bf16x2(d_o[0], d_o[1]);
    }
  }
}

// one step of both directions, launched so that it may overlap the tail of the previous launch (see the kernel)
template <bool BWD>
cudaError_t launch_step(const StepArgs& a, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kSlices, 2);
  cfg.blockDim = dim3(kStepThreads);
  cfg.dynamicSmemBytes = kStepSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lstm_step_kernel<BWD>, a);
}

}  // namespace

extern "C" int icka_lstm_bidir_fwd_save(icka_handle* h, const void* gx, const void* w_hh_fwd, const void* w_hh_bwd, void* y_op,
                                        float* y32, float* acts, float* c_all, int B, int S, int H, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(H == kTH, "lstm_bidir_fwd_save: built for hidden size %d, got %d", kTH, H);
  ICKA_REQUIRE(B >= 0 && S >= 1, "lstm_bidir_fwd_save: bad shape B=%d S=%d", B, S);
  ICKA_REQUIRE(gx && w_hh_fwd && w_hh_bwd && y_op && y32 && acts && c_all, "lstm_bidir_fwd_save: null pointer");
  ICKA_REQUIRE(icka_aligned(gx, 16) && icka_aligned(w_hh_fwd, 16) && icka_aligned(w_hh_bwd, 16) && icka_aligned(y_op, 16) &&
                   icka_aligned(y32, 16) && icka_aligned(acts, 16) && icka_aligned(c_all, 16),
               "lstm_bidir_fwd_save: pointers must be 16-byte aligned");
  ICKA_REQUIRE(h->smem_optin >= kStepSmem, "lstm_bidir_fwd_save: device offers too little shared memory");
  if (B == 0) return ICKA_OK;
  ICKA_CUDA(cudaFuncSetAttribute(lstm_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStepSmem));
  StepArgs a{};
  a.gx = static_cast<const __nv_bfloat16*>(gx);
  a.w[0] = static_cast<const __nv_bfloat16*>(w_hh_fwd);
  a.w[1] = static_cast<const __nv_bfloat16*>(w_hh_bwd);
  a.y_op = static_cast<__nv_bfloat16*>(y_op);
  a.y32 = y32;
  a.acts = acts;
  a.c_all = c_all;
  a.B = B;
  a.S = S;
  for (int t = 0; t < S; ++t) {
    a.t = t;
    ICKA_CUDA(launch_step<false>(a, static_cast<cudaStream_t>(stream)));
    ICKA_LAUNCHED(h);
  }
  return ICKA_OK;
}

extern "C" int icka_lstm_bidir_bwd(icka_handle* h, const float* dy, const void* w_hh_t_fwd, const void* w_hh_t_bwd,
                                   const float* acts, const float* c_all, void* dg, float* dc_scratch, int B, int S, int H,
                                   void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(H == kTH, "lstm_bidir_bwd: built for hidden size %d, got %d", kTH, H);
  ICKA_REQUIRE(B >= 0 && S >= 1, "lstm_bidir_bwd: bad shape B=%d S=%d", B, S);
  ICKA_REQUIRE(dy && w_hh_t_fwd && w_hh_t_bwd && acts && c_all && dg && dc_scratch, "lstm_bidir_bwd: null pointer");
  ICKA_REQUIRE(icka_aligned(dy, 16) && icka_aligned(w_hh_t_fwd, 16) && icka_aligned(w_hh_t_bwd, 16) && icka_aligned(dg, 16) &&
                   icka_aligned(acts, 16) && icka_aligned(c_all, 16) && icka_aligned(dc_scratch, 16),
               "lstm_bidir_bwd: pointers must be 16-byte aligned");
  ICKA_REQUIRE(h->smem_optin >= kStepSmem, "lstm_bidir_bwd: device offers too little shared memory");
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ICKA_CUDA(cudaFuncSetAttribute(lstm_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStepSmem));
  ICKA_CUDA(cudaMemsetAsync(dc_scratch, 0, (size_t)2 * B * H * sizeof(float), st));
  StepArgs a{};
  a.w[0] = static_cast<const __nv_bfloat16*>(w_hh_t_fwd);
  a.w[1] = static_cast<const __nv_bfloat16*>(w_hh_t_bwd);
  a.acts = const_cast<float*>(acts);
  a.c_all = const_cast<float*>(c_all);
  a.dy = dy;
  a.dg = static_cast<__nv_bfloat16*>(dg);
  a.dc = dc_scratch;
  a.B = B;
  a.S = S;
  for (int t = S - 1; t >= 0; --t) {
    a.t = t;
    ICKA_CUDA(launch_step<true>(a, st));
    ICKA_LAUNCHED(h);
  }
  return ICKA_OK;
}
