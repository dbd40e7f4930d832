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

// out[r, t] = bias[t] + sum_k x[r, k] W[t, k].  Warp = 4 rows at a time, lane = 8 consecutive k per 256-wide sweep;
// W (fp32) lives in shared memory for the life of the block; fp32 accumulation, fixed reduction order.
template <typename T, int NT>
__global__ void __launch_bounds__(kHeadThreads)
emission_head_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ W, const float* __restrict__ bias,
                     float* __restrict__ out, int64_t M, int K) {
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
      if (lane < NT && r0 + r < M) out[(r0 + r) * NT + lane] = mine + __ldg(bias + lane);
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
                int K, cudaStream_t st) {
  const size_t smem = (size_t)NT * K * sizeof(float);
  if (smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "emission_head: T=%d K=%d needs %zu B shared memory (max %zu)", NT, K, smem,
              h->smem_optin);
  auto kern = emission_head_kernel<T, NT>;
  ICKA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  const int64_t want = (M + (kHeadThreads / 32) * kHeadRows - 1) / ((kHeadThreads / 32) * kHeadRows);
  const int grid = (int)(want < (int64_t)h->sm_count * per_sm ? want : (int64_t)h->sm_count * per_sm);
  kern<<<grid, kHeadThreads, smem, st>>>(static_cast<const T*>(x), ldx, W, bias, out, M, K);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

template <typename T>
int dispatch_head(icka_handle* h, const void* x, int64_t ldx, const float* W, const float* bias, float* out, int64_t M,
                  int K, int T_, cudaStream_t st) {
  switch (T_) {
#define ICKA_HEAD_CASE(n) \
  case n:                 \
    return launch_head<T, n>(h, x, ldx, W, bias, out, M, K, st);
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
                                      float* out, int dtype, int64_t M, int K, int T, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(M >= 0 && K >= 8 && T >= 1, "emission_head: bad shape M=%lld K=%d T=%d", (long long)M, K, T);
  ICKA_REQUIRE(K % 8 == 0 && ldx >= K && ldx % 8 == 0, "emission_head: K and the row pitch must be multiples of 8");
  ICKA_REQUIRE(x && W && bias && out, "emission_head: null pointer");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(W, 16), "emission_head: x and W must be 16-byte aligned");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "emission_head: bad dtype %d", dtype);
  if (M == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ICKA_F32) return dispatch_head<float>(h, x, ldx, W, bias, out, M, K, T, st);
  return dispatch_head<__nv_bfloat16>(h, x, ldx, W, bias, out, M, K, T, st);
}

extern "C" int icka_add_f32(icka_handle* h, const float* a, const float* b, float* out, int64_t n, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(n >= 0 && a && b && out, "add_f32: bad arguments");
  if (n == 0) return ICKA_OK;
  add_f32_kernel<<<(int)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, n);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
