// Emission head around the recurrent kernel (SURVEY 8f "next" row 1; CMIM:905-910, 1042-1043):
//   * lstm_cell_kernel      one LSTM step on CUDA cores -- the per-step path: fp32 parity mode (FFMA GEMM h.W_hh^T
//                           + this kernel) and shapes the persistent tcgen05 kernel (lstm_sm100.cu) is not built for
//   * emission_head_kernel  `self.classifier = nn.Linear(2H, T)` (CMIM:910, 1043): [rows, 2H] x [T, 2H]^T with
//                           T = 15 -- far too narrow for a 128 x 256 tensor-core tile, HBM-bound on the state read
//   * add_f32_kernel        b_ih + b_hh (weight preparation)
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

// gates_h [B,4H] fp32 = h_{t-1} . W_hh^T (gate order i,f,g,o; NULL at t = 0), gx = x_t . W_ih^T + b_ih + b_hh
template <typename T>
__global__ void lstm_cell_kernel(const float* __restrict__ gates_h, const T* __restrict__ gx, int64_t ldgx,
                                 float* __restrict__ c, T* __restrict__ h_out, T* __restrict__ y, int64_t ldy,
                                 float* __restrict__ h_f32, int B, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  float pre[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    pre[g] = to_f32<T>(gx[(int64_t)b * ldgx + g * H + j]);
    if (gates_h) pre[g] = gates_h[(int64_t)b * 4 * H + g * H + j] + pre[g];
  }
  const float ig = sigmoid_exact(pre[0]), fg = sigmoid_exact(pre[1]), gg = tanhf(pre[2]), og = sigmoid_exact(pre[3]);
  const float cn = fg * c[idx] + ig * gg;
  const float hn = og * tanhf(cn);
  c[idx] = cn;
  h_out[idx] = from_f32<T>(hn);
  if (y) y[(int64_t)b * ldy + j] = from_f32<T>(hn);
  if (h_f32) h_f32[idx] = hn;
}

// ---- training (BPTT) of the per-step path ------------------------------------------------------------------------
// Forward step that keeps what the backward pass needs: the gate activations (i, f, g, o after sigmoid / tanh) and
// the new cell state, both fp32.
template <typename T>
__global__ void lstm_cell_fwd_save_kernel(const float* __restrict__ gates_h, const T* __restrict__ gx, int64_t ldgx,
                                          const float* __restrict__ c_prev, float* __restrict__ c_out,
                                          float* __restrict__ acts, T* __restrict__ h_out, T* __restrict__ y_op,
                                          int64_t ldyo, float* __restrict__ y32, int64_t ldy32, int B, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  float pre[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    pre[g] = to_f32<T>(gx[(int64_t)b * ldgx + g * H + j]);
    if (gates_h) pre[g] = gates_h[(int64_t)b * 4 * H + g * H + j] + pre[g];
  }
  const float ig = sigmoid_exact(pre[0]), fg = sigmoid_exact(pre[1]), gg = tanhf(pre[2]), og = sigmoid_exact(pre[3]);
  const float cn = fg * (c_prev ? c_prev[idx] : 0.0f) + ig * gg;
  const float hn = og * tanhf(cn);
  c_out[idx] = cn;
  float* a = acts + (int64_t)b * 4 * H + j;
  a[0] = ig;
  a[H] = fg;
  a[2 * H] = gg;
  a[3 * H] = og;
  h_out[idx] = from_f32<T>(hn);
  y_op[(int64_t)b * ldyo + j] = from_f32<T>(hn);
  y32[(int64_t)b * ldy32 + j] = hn;
}

// One BPTT step: dh = dy (gradient of the output slot) + dh_rec (from step t + 1 through W_hh); dc carries the cell
// gradient from step t + 1 and leaves as the gradient for step t - 1; dpre = gradient of the gate pre-activations.
template <typename T>
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ dh_rec,
                                     float* __restrict__ dc, const float* __restrict__ acts,
                                     const float* __restrict__ c_prev, const float* __restrict__ c_new,
                                     T* __restrict__ dpre, int64_t lddp, int B, int H) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx % H);
  const float* a = acts + (int64_t)b * 4 * H + j;
  const float ig = a[0], fg = a[H], gg = a[2 * H], og = a[3 * H];
  const float tc = tanhf(c_new[idx]);
  const float dh = dy[(int64_t)b * lddy + j] + (dh_rec ? dh_rec[idx] : 0.0f);
  const float dct = dc[idx] + dh * og * (1.0f - tc * tc);
  T* d = dpre + (int64_t)b * lddp + j;
  d[0] = from_f32<T>(dct * gg * ig * (1.0f - ig));
  d[H] = from_f32<T>(dct * (c_prev ? c_prev[idx] : 0.0f) * fg * (1.0f - fg));
  d[2 * H] = from_f32<T>(dct * ig * (1.0f - gg * gg));
  d[3 * H] = from_f32<T>(dh * tc * og * (1.0f - og));
  dc[idx] = dct * fg;
}

// input rows may be time-major (row = t * B + b, tm_S = S > 0): the emissions are always written batch-major
__device__ __forceinline__ int64_t out_row(int64_t r, int64_t M, int tm_S) {
  if (tm_S <= 0) return r;
  const int64_t B = M / tm_S;
  return (r % B) * tm_S + r / B;
}

constexpr int kHeadThreads = 256;
constexpr int kHeadRows = 4;      // rows per warp pass: every weight fetched from shared memory is used 4 times
constexpr int kHeadMaxT = 16;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// [B, S, H] (fp32 or bf16) -> [S, B, H] bf16: the recurrent kernel reads Gx and writes the states one TIME STEP at a
// time, so time-major tensors make every step touch one contiguous block (DRAM-page and TLB friendly) instead of B
// pieces scattered S rows apart.
template <typename T>
__global__ void cast_time_major_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int S, int H) {
  const int64_t n8 = (int64_t)B * S * H / 8;
  const int h8 = H / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % h8);
    const int64_t r = i / h8;                 // output row t * B + b
    const int b = (int)(r % B), t = (int)(r / B);
    float v[8];
    load8<T>(x + ((int64_t)b * S + t) * H + c * 8, v);
    *reinterpret_cast<uint4*>(y + r * H + c * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// out[r, t] = bias[t] + sum_k x[r, k] W[t, k].  Warp = 4 rows at a time, lane = 8 consecutive k per 256-wide sweep;
// W (fp32) lives in shared memory for the life of the block; fp32 accumulation, fixed reduction order.
template <typename T, int NT>
__global__ void __launch_bounds__(kHeadThreads)
emission_head_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ W, const float* __restrict__ bias,
                     float* __restrict__ out, int64_t M, int K, int tm_S) {
  extern __shared__ float w_s[];   // [NT][K]
  for (int i = threadIdx.x * 4; i < NT * K; i += kHeadThreads * 4)
    *reinterpret_cast<float4*>(w_s + i) = __ldg(reinterpret_cast<const float4*>(W + i));
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t warps_total = (int64_t)gridDim.x * (kHeadThreads / 32);
  for (int64_t r0 = ((int64_t)blockIdx.x * (kHeadThreads / 32) + warp) * kHeadRows; r0 < M; r0 += warps_total * kHeadRows) {
    float acc[kHeadRows][NT];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[r][t] = 0.0f;
    for (int k0 = lane * 8; k0 < K; k0 += 256) {
      float xv[kHeadRows][8];
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        if (r0 + r < M) {
          load8<T>(x + (r0 + r) * ldx + k0, xv[r]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) xv[r][i] = 0.0f;
        }
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const float4 wa = *reinterpret_cast<const float4*>(w_s + t * K + k0);
        const float4 wb = *reinterpret_cast<const float4*>(w_s + t * K + k0 + 4);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[r][t] = fmaf(xv[r][i], wv[i], acc[r][t]);
      }
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float v = acc[r][t];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][t] = v;
      }
    // lane t of the warp writes tag t of each row: NT consecutive floats per row
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      float mine = 0.0f;
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (lane == t) mine = acc[r][t];
      if (lane < NT && r0 + r < M) out[out_row(r0 + r, M, tm_S) * NT + lane] = mine + __ldg(bias + lane);
    }
  }
}

// bf16 states: the same product on mma.sync tensor cores.  The FFMA kernel above is bound by FMA issue + shared-memory
// reads of W (measured 1.3 TB/s at T = 15); here the T x K weights are the M = 16 operand of m16n8k16 (tags padded to
// 16), 8 state rows are the N operand, and W is split into hi + lo bf16 halves (two MMAs) so the weights keep ~16
// mantissa bits -- the result matches the fp32-weight product to ~1e-6 while the kernel becomes a pure stream of the
// states.  The contraction index is permuted so that every lane loads 16 contiguous bytes of its row (8 rows x 64 B per
// warp request) and 16 contiguous bytes of two weight rows: logical k-pairs {2q, 2q+1} / {2q+8, 2q+9} of the two
// MMAs of a 32-wide block are physical elements q*8 + {0,1} / {2,3} and q*8 + {4,5} / {6,7} for A and B alike.
constexpr int kMmaHeadRows = 32;          // rows per warp pass (4 groups of 8)
constexpr int kMmaHeadPad = 32;           // row pitch K + 32 bf16: the two rows of a quarter-warp hit disjoint banks

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kHeadThreads)
emission_head_mma_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ W,
                         const float* __restrict__ bias, float* __restrict__ out, int64_t M, int K, int T, int tm_S) {
  extern __shared__ __align__(16) unsigned char head_smem[];
  const int ldw = K + kMmaHeadPad;
  __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(head_smem);   // [16][ldw]
  __nv_bfloat16* w_lo = w_hi + 16 * ldw;
  for (int i = threadIdx.x; i < 16 * K; i += kHeadThreads) {
    const int t = i / K, k = i % K;
    const float w = t < T ? __ldg(W + (size_t)t * K + k) : 0.0f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    w_hi[t * ldw + k] = hi;
    w_lo[t * ldw + k] = __float2bfloat16_rn(w - __bfloat162float(hi));
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int g = lane >> 2, q = lane & 3;
  const float bias_g = g < T ? __ldg(bias + g) : 0.0f, bias_g8 = g + 8 < T ? __ldg(bias + g + 8) : 0.0f;
  const __nv_bfloat16* wh0 = w_hi + g * ldw + q * 8;
  const __nv_bfloat16* wh1 = w_hi + (g + 8) * ldw + q * 8;
  const __nv_bfloat16* wl0 = w_lo + g * ldw + q * 8;
  const __nv_bfloat16* wl1 = w_lo + (g + 8) * ldw + q * 8;
  const int64_t warps_total = (int64_t)gridDim.x * (kHeadThreads / 32);
  for (int64_t r0 = ((int64_t)blockIdx.x * (kHeadThreads / 32) + warp) * kMmaHeadRows; r0 < M;
       r0 += warps_total * kMmaHeadRows) {
    float acc[4][4];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[gi][i] = 0.0f;
    const __nv_bfloat16* xp[4];
    bool ok[4];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const int64_t row = r0 + gi * 8 + g;
      ok[gi] = row < M;
      xp[gi] = x + (ok[gi] ? row : 0) * ldx + q * 8;
    }
#pragma unroll 4
    for (int kb = 0; kb < K; kb += 32) {
      uint4 xv[4];
#pragma unroll
      for (int gi = 0; gi < 4; ++gi)
        xv[gi] = ok[gi] ? __ldg(reinterpret_cast<const uint4*>(xp[gi] + kb)) : make_uint4(0u, 0u, 0u, 0u);
      const uint4 h0 = *reinterpret_cast<const uint4*>(wh0 + kb), h1 = *reinterpret_cast<const uint4*>(wh1 + kb);
      const uint4 l0 = *reinterpret_cast<const uint4*>(wl0 + kb), l1 = *reinterpret_cast<const uint4*>(wl1 + kb);
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) {
        mma_bf16_16816(acc[gi], h0.x, h1.x, h0.y, h1.y, xv[gi].x, xv[gi].y);
        mma_bf16_16816(acc[gi], h0.z, h1.z, h0.w, h1.w, xv[gi].z, xv[gi].w);
        mma_bf16_16816(acc[gi], l0.x, l1.x, l0.y, l1.y, xv[gi].x, xv[gi].y);
        mma_bf16_16816(acc[gi], l0.z, l1.z, l0.w, l1.w, xv[gi].z, xv[gi].w);
      }
    }
    // d0, d1: tag g of rows 2q, 2q+1 of the group; d2, d3: tag g + 8
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t row = r0 + gi * 8 + 2 * q + i;
        if (row < M) {
          const int64_t orow = out_row(row, M, tm_S);
          if (g < T) out[orow * T + g] = acc[gi][i] + bias_g;
          if (g + 8 < T) out[orow * T + g + 8] = acc[gi][2 + i] + bias_g8;
        }
      }
    }
  }
}

// Backward of the classifier (CMIM:910, 1043) in one pass over the states:  dx[r, k] = sum_t dout[r, t] W[t, k]  and
// dW[t, k] += sum_r dout[r, t] x[r, k].  The label count is far below a tensor-core tile (T <= 16), and the generic fp32
// wgrad spent 3.4 ms on the 16384 x 1536 states of a B = 128 step.  A thread owns four consecutive k: its W[:, 4] and
// dW[:, 4] live in registers (tags padded to 16 with zeros), the dout rows of a slab are broadcast from shared memory, and
// x / dx stream through once (HBM-bound); the per-block dW partials are combined with fp32 atomics.
constexpr int kHeadBwdRows = 32;     // dout rows staged per slab
template <typename T>
__global__ void __launch_bounds__(384)
emission_head_bwd_kernel(const float* __restrict__ dout, const T* __restrict__ x, int64_t ldx, const float* __restrict__ W,
                         float* __restrict__ dx, float* __restrict__ dW, int64_t M, int K, int NT, int tm_S) {
  __shared__ __align__(16) float d_s[kHeadBwdRows][kHeadMaxT];
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const bool live = c < K;
  float w[kHeadMaxT][4], acc[kHeadMaxT][4];
#pragma unroll
  for (int t = 0; t < kHeadMaxT; ++t) {
    const float4 v = (live && t < NT) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)t * K + c)) : make_float4(0, 0, 0, 0);
    w[t][0] = v.x; w[t][1] = v.y; w[t][2] = v.z; w[t][3] = v.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.0f;
  }
  for (int64_t r0 = (int64_t)blockIdx.x * kHeadBwdRows; r0 < M; r0 += (int64_t)gridDim.x * kHeadBwdRows) {
    __syncthreads();
    for (int i = threadIdx.x; i < kHeadBwdRows * kHeadMaxT; i += blockDim.x) {
      const int r = i / kHeadMaxT, t = i % kHeadMaxT;
      d_s[r][t] = (r0 + r < M && t < NT) ? __ldg(dout + out_row(r0 + r, M, tm_S) * NT + t) : 0.0f;
    }
    __syncthreads();
    if (!live) continue;
    const int rows = (int)(M - r0 < kHeadBwdRows ? M - r0 : kHeadBwdRows);
    for (int rb = 0; rb < rows; rb += 4) {
      float xv[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (rb + u < rows) {
          const T* xp = x + (r0 + rb + u) * ldx + c;
          if constexpr (sizeof(T) == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xp));
            xv[u][0] = v.x; xv[u][1] = v.y; xv[u][2] = v.z; xv[u][3] = v.w;
          } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(xp));
            xv[u][0] = __uint_as_float(v.x << 16); xv[u][1] = __uint_as_float(v.x & 0xffff0000u);
            xv[u][2] = __uint_as_float(v.y << 16); xv[u][3] = __uint_as_float(v.y & 0xffff0000u);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) xv[u][j] = 0.0f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float g[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int t4 = 0; t4 < kHeadMaxT; t4 += 4) {
          const float4 dv = *reinterpret_cast<const float4*>(&d_s[(rb + u) & (kHeadBwdRows - 1)][t4]);
          const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              g[j] = fmaf(d4[i], w[t4 + i][j], g[j]);
              acc[t4 + i][j] = fmaf(d4[i], xv[u][j], acc[t4 + i][j]);
            }
        }
        if (dx != nullptr && rb + u < rows)
          *reinterpret_cast<float4*>(dx + (r0 + rb + u) * (int64_t)K + c) = make_float4(g[0], g[1], g[2], g[3]);
      }
    }
  }
  if (live && dW != nullptr) {
#pragma unroll
    for (int t = 0; t < kHeadMaxT; ++t)
      if (t < NT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dW + (size_t)t * K + c + j, acc[t][j]);
      }
  }
}

__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                               int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

template <typename T, int NT>
int launch_head(icka_handle* h, const void* x, int64_t ldx, const float* W, const float* bias, float* out, int64_t M,
                int K, int tm_S, cudaStream_t st) {
  const size_t smem = (size_t)NT * K * sizeof(float);
  if (smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "emission_head: T=%d K=%d needs %zu B shared memory (max %zu)", NT, K, smem,
              h->smem_optin);
  auto kern = emission_head_kernel<T, NT>;
  ICKA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  const int64_t want = (M + (kHeadThreads / 32) * kHeadRows - 1) / ((kHeadThreads / 32) * kHeadRows);
  const int grid = (int)(want < (int64_t)h->sm_count * per_sm ? want : (int64_t)h->sm_count * per_sm);
  kern<<<grid, kHeadThreads, smem, st>>>(static_cast<const T*>(x), ldx, W, bias, out, M, K, tm_S);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

template <typename T>
int dispatch_head(icka_handle* h, const void* x, int64_t ldx, const float* W, const float* bias, float* out, int64_t M,
                  int K, int T_, int tm_S, cudaStream_t st) {
  switch (T_) {
#define ICKA_HEAD_CASE(n) \
  case n:                 \
    return launch_head<T, n>(h, x, ldx, W, bias, out, M, K, tm_S, st);
    ICKA_HEAD_CASE(1) ICKA_HEAD_CASE(2) ICKA_HEAD_CASE(3) ICKA_HEAD_CASE(4) ICKA_HEAD_CASE(5) ICKA_HEAD_CASE(6)
    ICKA_HEAD_CASE(7) ICKA_HEAD_CASE(8) ICKA_HEAD_CASE(9) ICKA_HEAD_CASE(10) ICKA_HEAD_CASE(11) ICKA_HEAD_CASE(12)
    ICKA_HEAD_CASE(13) ICKA_HEAD_CASE(14) ICKA_HEAD_CASE(15) ICKA_HEAD_CASE(16)
#undef ICKA_HEAD_CASE
  }
  ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "emission_head: num_labels %d not in 1..%d", T_, kHeadMaxT);
}

}  // namespace

extern "C" int icka_lstm_cell_fwd(icka_handle* h, const float* gates_h, const void* gx, int64_t ldgx, float* c,
                                  void* h_out, void* y, int64_t ldy, float* h_f32, int dtype, int B, int H,
                                  void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && H >= 1, "lstm_cell: bad shape B=%d H=%d", B, H);
  ICKA_REQUIRE(gx && c && h_out, "lstm_cell: null pointer");
  ICKA_REQUIRE(ldgx >= 4 * (int64_t)H && (!y || ldy >= H), "lstm_cell: pitches smaller than the logical extents");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "lstm_cell: bad dtype %d", dtype);
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = (int64_t)B * H;
  const int grid = (int)((n + 255) / 256);
  if (dtype == ICKA_F32)
    lstm_cell_kernel<float><<<grid, 256, 0, st>>>(gates_h, static_cast<const float*>(gx), ldgx, c,
                                                  static_cast<float*>(h_out), static_cast<float*>(y), ldy, h_f32, B, H);
  else
    lstm_cell_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(gates_h, static_cast<const __nv_bfloat16*>(gx), ldgx, c,
                                                          static_cast<__nv_bfloat16*>(h_out),
                                                          static_cast<__nv_bfloat16*>(y), ldy, h_f32, B, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_emission_head_fwd(icka_handle* h, const void* x, int64_t ldx, const float* W, const float* bias,
                                      float* out, int dtype, int64_t M, int K, int T, int time_major_S, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(M >= 0 && K >= 8 && T >= 1, "emission_head: bad shape M=%lld K=%d T=%d", (long long)M, K, T);
  ICKA_REQUIRE(K % 8 == 0 && ldx >= K && ldx % 8 == 0, "emission_head: K and the row pitch must be multiples of 8");
  ICKA_REQUIRE(x && W && bias && out, "emission_head: null pointer");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(W, 16), "emission_head: x and W must be 16-byte aligned");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "emission_head: bad dtype %d", dtype);
  ICKA_REQUIRE(time_major_S >= 0 && (time_major_S == 0 || M % time_major_S == 0),
               "emission_head: M=%lld is not a multiple of the time-major S=%d", (long long)M, time_major_S);
  if (M == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tm_S = time_major_S;
  if (dtype == ICKA_F32) return dispatch_head<float>(h, x, ldx, W, bias, out, M, K, T, tm_S, st);
  const size_t smem = (size_t)2 * 16 * (K + kMmaHeadPad) * sizeof(__nv_bfloat16);
  if (T <= 16 && K % 32 == 0 && smem <= h->smem_optin) {   // tensor-core stream (hi/lo split weights)
    ICKA_CUDA(cudaFuncSetAttribute(emission_head_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = smem > 100 * 1024 ? 1 : 2;
    const int64_t want = (M + (kHeadThreads / 32) * kMmaHeadRows - 1) / ((kHeadThreads / 32) * kMmaHeadRows);
    const int grid = (int)(want < (int64_t)h->sm_count * per_sm ? want : (int64_t)h->sm_count * per_sm);
    emission_head_mma_kernel<<<grid, kHeadThreads, smem, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, W, bias, out,
                                                               M, K, T, tm_S);
    ICKA_LAUNCHED(h);
    return ICKA_OK;
  }
  return dispatch_head<__nv_bfloat16>(h, x, ldx, W, bias, out, M, K, T, tm_S, st);
}

extern "C" int icka_add_f32(icka_handle* h, const float* a, const float* b, float* out, int64_t n, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(n >= 0 && a && b && out, "add_f32: bad arguments");
  if (n == 0) return ICKA_OK;
  add_f32_kernel<<<(int)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, n);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_cast_bf16_time_major(icka_handle* h, const void* x, void* y_bf16, int in_dtype, int B, int S, int H,
                                         void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && H >= 8 && H % 8 == 0, "cast_bf16_time_major: bad shape B=%d S=%d H=%d", B, S, H);
  ICKA_REQUIRE(x && y_bf16 && icka_aligned(x, 16) && icka_aligned(y_bf16, 16), "cast_bf16_time_major: bad pointers");
  ICKA_REQUIRE(in_dtype == ICKA_F32 || in_dtype == ICKA_BF16, "cast_bf16_time_major: bad dtype %d", in_dtype);
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n8 = (int64_t)B * S * H / 8;
  const int64_t want = (n8 + 255) / 256;
  const int grid = (int)(want < (int64_t)h->sm_count * 16 ? want : (int64_t)h->sm_count * 16);
  if (in_dtype == ICKA_F32)
    cast_time_major_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x),
                                                         static_cast<__nv_bfloat16*>(y_bf16), B, S, H);
  else
    cast_time_major_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                                 static_cast<__nv_bfloat16*>(y_bf16), B, S, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_lstm_cell_fwd_save(icka_handle* h, const float* gates_h, const void* gx, int64_t ldgx, const float* c_prev,
                                       float* c_out, float* acts, void* h_out, void* y_op, int64_t ldyo, float* y32,
                                       int64_t ldy32, int dtype, int B, int H, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && H >= 1, "lstm_cell_fwd_save: bad shape B=%d H=%d", B, H);
  ICKA_REQUIRE(gx && c_out && acts && h_out && y_op && y32, "lstm_cell_fwd_save: null pointer");
  ICKA_REQUIRE(ldgx >= 4 * (int64_t)H && ldyo >= H && ldy32 >= H, "lstm_cell_fwd_save: pitches smaller than the extents");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "lstm_cell_fwd_save: bad dtype %d", dtype);
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)(((int64_t)B * H + 255) / 256);
  if (dtype == ICKA_F32)
    lstm_cell_fwd_save_kernel<float><<<grid, 256, 0, st>>>(gates_h, static_cast<const float*>(gx), ldgx, c_prev, c_out, acts,
                                                           static_cast<float*>(h_out), static_cast<float*>(y_op), ldyo, y32,
                                                           ldy32, B, H);
  else
    lstm_cell_fwd_save_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        gates_h, static_cast<const __nv_bfloat16*>(gx), ldgx, c_prev, c_out, acts, static_cast<__nv_bfloat16*>(h_out),
        static_cast<__nv_bfloat16*>(y_op), ldyo, y32, ldy32, B, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_lstm_cell_bwd(icka_handle* h, const float* dy, int64_t lddy, const float* dh_rec, float* dc,
                                  const float* acts, const float* c_prev, const float* c_new, void* dpre, int64_t lddp,
                                  int dtype, int B, int H, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && H >= 1, "lstm_cell_bwd: bad shape B=%d H=%d", B, H);
  ICKA_REQUIRE(dy && dc && acts && c_new && dpre, "lstm_cell_bwd: null pointer");
  ICKA_REQUIRE(lddy >= H && lddp >= 4 * (int64_t)H, "lstm_cell_bwd: pitches smaller than the extents");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "lstm_cell_bwd: bad dtype %d", dtype);
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)(((int64_t)B * H + 255) / 256);
  if (dtype == ICKA_F32)
    lstm_cell_bwd_kernel<float><<<grid, 256, 0, st>>>(dy, lddy, dh_rec, dc, acts, c_prev, c_new, static_cast<float*>(dpre),
                                                      lddp, B, H);
  else
    lstm_cell_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dy, lddy, dh_rec, dc, acts, c_prev, c_new,
                                                              static_cast<__nv_bfloat16*>(dpre), lddp, B, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

// ---- one direction of the training recurrence as ONE host call ---------------------------------------------------
// The per-step kernels are tiny; issued from Python (ctypes per launch) the training step was bound by the host
// (23 ms for 128 steps x 2 directions x 4 launches).  These loops issue the same launches from C.
extern "C" int icka_lstm_dir_fwd_save(icka_handle* h, const void* gx_dir, int64_t ld_gx_row, int64_t gx_pos_stride,
                                      const void* w_hh, float* acts, float* c_all, void* y_op_dir, float* y32_dir,
                                      int64_t ld_y_row, int64_t y_pos_stride, float* gates_scratch, void* h_scratch,
                                      int dtype, int B, int S, int H, int reverse, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && H >= 1, "lstm_dir_fwd_save: bad shape B=%d S=%d H=%d", B, S, H);
  ICKA_REQUIRE(gx_dir && w_hh && acts && c_all && y_op_dir && y32_dir && gates_scratch && h_scratch,
               "lstm_dir_fwd_save: null pointer");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "lstm_dir_fwd_save: bad dtype %d", dtype);
  if (B == 0) return ICKA_OK;
  const size_t esz = dtype == ICKA_BF16 ? 2 : 4;
  uint8_t* hbuf = static_cast<uint8_t*>(h_scratch);
  const size_t hbytes = (size_t)B * H * esz;
  for (int t = 0; t < S; ++t) {
    const int pos = reverse ? S - 1 - t : t;
    void* h_prev = hbuf + (size_t)((t + 1) & 1) * hbytes;
    void* h_next = hbuf + (size_t)(t & 1) * hbytes;
    if (t > 0) {
      int rc = icka_linear_fwd(h, h_prev, H, w_hh, H, nullptr, nullptr, gates_scratch, 4 * (int64_t)H, dtype, ICKA_F32, B,
                               4 * H, H, ICKA_ACT_NONE, stream);
      if (rc) return rc;
    }
    int rc = icka_lstm_cell_fwd_save(
        h, t > 0 ? gates_scratch : nullptr, static_cast<const uint8_t*>(gx_dir) + (size_t)pos * gx_pos_stride * esz,
        ld_gx_row, t > 0 ? c_all + (size_t)(t - 1) * B * H : nullptr, c_all + (size_t)t * B * H,
        acts + (size_t)t * B * 4 * H, h_next, static_cast<uint8_t*>(y_op_dir) + (size_t)pos * y_pos_stride * esz, ld_y_row,
        y32_dir + (size_t)pos * y_pos_stride, ld_y_row, dtype, B, H, stream);
    if (rc) return rc;
  }
  return ICKA_OK;
}

extern "C" int icka_lstm_dir_bwd(icka_handle* h, const float* dy_dir, int64_t ld_dy_row, int64_t dy_pos_stride,
                                 const void* w_hh, const float* acts, const float* c_all, void* dg_dir, int64_t ld_dg_row,
                                 int64_t dg_pos_stride, float* dc_scratch, float* dh_scratch, int dtype, int B, int S,
                                 int H, int reverse, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && H >= 1, "lstm_dir_bwd: bad shape B=%d S=%d H=%d", B, S, H);
  ICKA_REQUIRE(dy_dir && w_hh && acts && c_all && dg_dir && dc_scratch && dh_scratch, "lstm_dir_bwd: null pointer");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "lstm_dir_bwd: bad dtype %d", dtype);
  if (B == 0) return ICKA_OK;
  const size_t esz = dtype == ICKA_BF16 ? 2 : 4;
  ICKA_CUDA(cudaMemsetAsync(dc_scratch, 0, (size_t)B * H * sizeof(float), static_cast<cudaStream_t>(stream)));
  for (int t = S - 1; t >= 0; --t) {
    const int pos = reverse ? S - 1 - t : t;
    void* dpre = static_cast<uint8_t*>(dg_dir) + (size_t)pos * dg_pos_stride * esz;
    int rc = icka_lstm_cell_bwd(h, dy_dir + (size_t)pos * dy_pos_stride, ld_dy_row, t < S - 1 ? dh_scratch : nullptr,
                                dc_scratch, acts + (size_t)t * B * 4 * H, t > 0 ? c_all + (size_t)(t - 1) * B * H : nullptr,
                                c_all + (size_t)t * B * H, dpre, ld_dg_row, dtype, B, H, stream);
    if (rc) return rc;
    if (t > 0) {
      rc = icka_linear_dgrad(h, dpre, ld_dg_row, w_hh, H, nullptr, nullptr, 0, dh_scratch, H, dtype, ICKA_F32, B, 4 * H, H,
                             stream);
      if (rc) return rc;
    }
  }
  return ICKA_OK;
}

extern "C" int icka_emission_head_bwd(icka_handle* h, const float* dout, const void* x, int64_t ldx, const float* W, float* dx,
                                      float* dW, int dtype, int64_t M, int K, int T, int accumulate, int time_major_S,
                                      void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(M >= 0 && K >= 4 && T >= 1 && T <= kHeadMaxT, "emission_head_bwd: bad shape M=%lld K=%d T=%d (T <= %d)",
               (long long)M, K, T, kHeadMaxT);
  ICKA_REQUIRE(K % 4 == 0 && ldx >= K && ldx % 4 == 0, "emission_head_bwd: K and the row pitch must be multiples of 4");
  ICKA_REQUIRE(dout && x && W && (dx || dW), "emission_head_bwd: null pointer");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(W, 16) && icka_aligned(dx, 16), "emission_head_bwd: 16-byte alignment");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "emission_head_bwd: bad dtype %d", dtype);
  ICKA_REQUIRE(time_major_S >= 0 && (time_major_S == 0 || M % time_major_S == 0),
               "emission_head_bwd: M=%lld is not a multiple of the time-major S=%d", (long long)M, time_major_S);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dW && !accumulate) ICKA_CUDA(cudaMemsetAsync(dW, 0, (size_t)T * K * sizeof(float), st));
  if (M == 0) return ICKA_OK;
  const int k4 = K / 4;
  int threads = (k4 + 31) / 32 * 32;
  if (threads > 384) threads = 384;
  const int col_blocks = (k4 + threads - 1) / threads;
  int64_t row_blocks = (M + kHeadBwdRows - 1) / kHeadBwdRows;
  const int64_t cap = (h->sm_count + col_blocks - 1) / col_blocks;
  if (row_blocks > cap) row_blocks = cap;
  dim3 grid((unsigned)row_blocks, (unsigned)col_blocks);
  if (dtype == ICKA_F32)
    emission_head_bwd_kernel<float><<<grid, threads, 0, st>>>(dout, static_cast<const float*>(x), ldx, W, dx, dW, M, K, T,
                                                              time_major_S);
  else
    emission_head_bwd_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(dout, static_cast<const __nv_bfloat16*>(x), ldx, W,
                                                                      dx, dW, M, K, T, time_major_S);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
