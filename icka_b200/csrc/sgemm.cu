// fp32 CUDA-core path of icka_linear_fwd: the 1e-5 parity vehicle (north_star "fp32 path").
//
// kind::tf32 tensor-core MMA keeps a 10-bit mantissa (~1e-3), far outside the fp32 gate, so the fp32
// precision mode runs on FFMA: a classic 128x128x16 register-tiled kernel (8x8 outputs per thread),
// operands read as float4 along K and transposed into shared memory.  The throughput target of the
// project applies to the bf16 tcgen05 kernel in gemm_sm100.cu, not to this one.
//
//   out[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual      (all fp32)
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int LDS_ = BM + 4;
constexpr int kThreads = 256;

__device__ __forceinline__ float4 ldg_row(const float* __restrict__ base, int64_t ld, int row, int rows, int k, int K) {
  if (row < rows && k < K) return *reinterpret_cast<const float4*>(base + (size_t)row * ld + k);
  return make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int ACT, bool OUT_BF16>
__global__ void __launch_bounds__(kThreads) sgemm_tn_kernel(
    const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
    const float* __restrict__ bias, const float* __restrict__ residual, void* __restrict__ out, int64_t ldo,
    float* __restrict__ pre_act_out, int M, int N, int K) {
  __shared__ __align__(16) float As[BK][LDS_];
  __shared__ __align__(16) float Ws[BK][LDS_];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid / 4;          // 0..63
  const int lk = (tid % 4) * 4;      // 0,4,8,12
  const int ty = tid / 16, tx = tid % 16;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float4 pa[2], pw[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      pa[i] = ldg_row(A, lda, m0 + lrow + 64 * i, M, k0 + lk, K);
      pw[i] = ldg_row(W, ldw, n0 + lrow + 64 * i, N, k0 + lk, K);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lrow + 64 * i;
      As[lk + 0][r] = pa[i].x; As[lk + 1][r] = pa[i].y; As[lk + 2][r] = pa[i].z; As[lk + 3][r] = pa[i].w;
      Ws[lk + 0][r] = pw[i].x; Ws[lk + 1][r] = pw[i].y; Ws[lk + 2][r] = pw[i].z; Ws[lk + 3][r] = pw[i].w;
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    stage();
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ws[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int nb = n0 + jh * 64 + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        if (n >= N) continue;
        float v = acc[i][jh * 4 + j];
        if (bias) v += bias[n];
        if (ACT != ICKA_ACT_NONE && ACT != ICKA_ACT_TANH && pre_act_out) pre_act_out[(size_t)m * N + n] = v;
        if (ACT == ICKA_ACT_GELU_ERF) v = gelu_erf(v);
        if (ACT == ICKA_ACT_TANH) v = tanhf(v);
        if (ACT == ICKA_ACT_RELU) v = act_relu(v);
        if (ACT == ICKA_ACT_SWISH) v = act_swish(v);
        if (residual) v += residual[(size_t)m * N + n];
        if (OUT_BF16)
          static_cast<__nv_bfloat16*>(out)[(size_t)m * ldo + n] = __float2bfloat16_rn(v);
        else
          static_cast<float*>(out)[(size_t)m * ldo + n] = v;
      }
    }
  }
}

// Backward GEMMs of the fp32 parity path, any operand orientation through element strides:
//   out[m,n] (+)= (sum_k A(m,k) B(k,n)) * gelu'(gelu_pre[m,n]) + residual[m,n]
// with A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs].  64x64 tile, 4x4 outputs per thread.
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256) sgemm_strided_kernel(
    const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const float* __restrict__ B, int64_t b_rs, int64_t b_cs,
    const float* __restrict__ residual, const float* __restrict__ gelu_pre, int64_t ldg, float* __restrict__ out,
    int64_t ldo, int M, int N, int K, int accumulate) {
  __shared__ float As[GK][GT + 1];
  __shared__ float Bs[GK][GT + 1];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    for (int i = tid; i < GT * GK; i += 256) {
      // consecutive threads walk the operand's contiguous direction
      const int am = (a_cs == 1) ? i / GK : i % GT, ak = (a_cs == 1) ? i % GK : i / GT;
      As[ak][am] = (m0 + am < M && k0 + ak < K) ? A[(size_t)(m0 + am) * a_rs + (size_t)(k0 + ak) * a_cs] : 0.0f;
      const int bn = (b_rs == 1) ? i / GK : i % GT, bk = (b_rs == 1) ? i % GK : i / GT;
      Bs[bk][bn] = (n0 + bn < N && k0 + bk < K) ? B[(size_t)(k0 + bk) * b_rs + (size_t)(n0 + bn) * b_cs] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (gelu_pre) v *= gelu_erf_grad(gelu_pre[(size_t)m * ldg + n]);
      if (residual) v += residual[(size_t)m * N + n];
      float* o = out + (size_t)m * ldo + n;
      *o = accumulate ? *o + v : v;
    }
  }
}

}  // namespace

int icka_sgemm_strided_launch(icka_handle* h, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs,
                              int64_t b_cs, const float* residual, const float* gelu_pre, int64_t ldg, float* out,
                              int64_t ldo, int M, int N, int K, int accumulate, cudaStream_t st) {
  dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT);
  ICKA_REQUIRE(grid.y <= 65535, "sgemm_strided: M=%d too large for one launch", M);
  sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, residual, gelu_pre, ldg, out, ldo, M, N, K,
                                             accumulate);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

int icka_sgemm_launch(icka_handle* h, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                      const float* residual, void* out, int64_t ldo, int out_dtype, int M, int N, int K, int act,
                      float* pre_act_out, cudaStream_t st) {
  ICKA_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "linear(fp32): K, lda, ldw must be multiples of 4");
  ICKA_REQUIRE(icka_aligned(A, 16) && icka_aligned(W, 16), "linear(fp32): A and W must be 16-byte aligned");
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  ICKA_REQUIRE(grid.y <= 65535, "linear(fp32): M=%d too large for one launch", M);
#define ICKA_SGEMM(ACT_, BF_) \
  sgemm_tn_kernel<ACT_, BF_><<<grid, kThreads, 0, st>>>(A, lda, W, ldw, bias, residual, out, ldo, pre_act_out, M, N, K)
  const bool bf = out_dtype == ICKA_BF16;
  if (act == ICKA_ACT_TANH) {
    if (bf) ICKA_SGEMM(ICKA_ACT_TANH, true); else ICKA_SGEMM(ICKA_ACT_TANH, false);
  } else if (act == ICKA_ACT_GELU_ERF) {
    if (bf) ICKA_SGEMM(ICKA_ACT_GELU_ERF, true); else ICKA_SGEMM(ICKA_ACT_GELU_ERF, false);
  } else if (act == ICKA_ACT_RELU) {
    if (bf) ICKA_SGEMM(ICKA_ACT_RELU, true); else ICKA_SGEMM(ICKA_ACT_RELU, false);
  } else if (act == ICKA_ACT_SWISH) {
    if (bf) ICKA_SGEMM(ICKA_ACT_SWISH, true); else ICKA_SGEMM(ICKA_ACT_SWISH, false);
  } else {
    if (bf) ICKA_SGEMM(ICKA_ACT_NONE, true); else ICKA_SGEMM(ICKA_ACT_NONE, false);
  }
#undef ICKA_SGEMM
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
