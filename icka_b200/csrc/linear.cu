// nn.Linear forward (+ fused epilogue) and its two backward GEMMs, dispatched on the operand precision mode
// (bf16 -> tcgen05 tensor-core kernel, fp32 -> FFMA parity kernels).  See include/icka_b200.h.
#include "common.cuh"

int icka_sgemm_launch(icka_handle* h, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                      const float* residual, void* out, int64_t ldo, int out_dtype, int M, int N, int K, int act,
                      float* pre_act_out, cudaStream_t st);
int icka_sgemm_strided_launch(icka_handle* h, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs,
                              int64_t b_cs, const float* residual, const float* gelu_pre, int64_t ldg, float* out,
                              int64_t ldo, int M, int N, int K, int accumulate, cudaStream_t st);
int icka_gemm_bf16_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                          const float* residual, void* out, int64_t ldo, int out_dtype, int M, int N, int K, int act,
                          void* aux_out, cudaStream_t st);
int icka_gemm_bf16_dgrad_launch(icka_handle* h, const void* dY, int64_t ldd, const void* W, int64_t ldw,
                                const float* residual, const void* gelu_pre, int64_t ldg, void* dX, int64_t ldo,
                                int out_dtype, int M, int N, int K, cudaStream_t st);
int icka_gemm_bf16_wgrad_launch(icka_handle* h, const void* dY, int64_t ldd, const void* X, int64_t ldx, float* dW,
                                int64_t ldo, int M, int N, int K, cudaStream_t st);

int icka_gemm_bf16_ln_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                             const float* residual, const float* gamma, const float* beta, float eps, float* out32,
                             void* out16, int M, int N, int K, cudaStream_t st);

int icka_gemm_bf16_splitk_ln_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                                    const float* residual, const float* gamma, const float* beta, float eps, float* out32,
                                    void* out16, int M, int N, int K, cudaStream_t st);
bool icka_gemm_ln_cluster_supported(int N, int K, const void* residual);
int icka_gemm_bf16_ln_cluster_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                                     const float* bias, const float* residual, const float* gamma, const float* beta,
                                     float eps, float* out32, void* out16, int M, int N, int K, cudaStream_t st);

// icka_set_ln_mode: 0 = choose per shape (default), 1 = single-CTA kernel (rows re-read from L2), 2 = cluster kernel
static int g_ln_mode = 0;
extern "C" int icka_set_ln_mode(int mode) {
  ICKA_REQUIRE(mode >= 0 && mode <= 2, "ln mode %d not in 0..2", mode);
  g_ln_mode = mode;
  return ICKA_OK;
}

extern "C" int icka_linear_ln_fwd(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                                  const float* bias, const float* residual, const float* gamma, const float* beta,
                                  float eps, float* out_f32, void* out_bf16, int in_dtype, int M, int N, int K,
                                  void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(A && W && gamma && beta && out_f32, "linear_ln: null pointer");
  ICKA_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear_ln: bad shape M=%d N=%d K=%d", M, N, K);
  ICKA_REQUIRE(lda >= K && ldw >= K, "linear_ln: pitches smaller than the logical extents");
  ICKA_REQUIRE(in_dtype == ICKA_F32 || in_dtype == ICKA_BF16, "linear_ln: bad in_dtype %d", in_dtype);
  if (M == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == ICKA_BF16) {
    // N = 768 / 1024 (and 512): a cluster of N / 256 CTAs shares each 128-row block and exchanges the row statistics
    // over distributed shared memory (csrc/gemm_ln_sm100.cu); other widths: one CTA walks the n-tiles of its rows
    if (g_ln_mode == 0) {
      // skinny problems (the single-query encoders): split-K partials + one reduce-and-normalise pass
      int rc = icka_gemm_bf16_splitk_ln_launch(h, A, lda, W, ldw, bias, residual, gamma, beta, eps, out_f32, out_bf16, M, N, K,
                                               st);
      if (rc <= 0) return rc;
      if (M < 2048) {   // too few row blocks for a CTA (cluster) per block: plain GEMM, then the row kernel in place
        rc = icka_gemm_bf16_launch(h, A, lda, W, ldw, bias, residual, out_f32, N, ICKA_F32, M, N, K, ICKA_ACT_NONE, nullptr,
                                   st);
        if (rc) return rc;
        return icka_layernorm_fwd(h, out_f32, gamma, beta, eps, out_f32, out_bf16, M, N, stream);
      }
    }
    const bool cluster_ok = icka_gemm_ln_cluster_supported(N, K, residual);
    if (g_ln_mode == 2 && !cluster_ok) ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "linear_ln: cluster kernel does not serve N=%d K=%d", N, K);
    if (cluster_ok && g_ln_mode != 1)
      return icka_gemm_bf16_ln_cluster_launch(h, A, lda, W, ldw, bias, residual, gamma, beta, eps, out_f32, out_bf16, M,
                                              N, K, st);
    return icka_gemm_bf16_ln_launch(h, A, lda, W, ldw, bias, residual, gamma, beta, eps, out_f32, out_bf16, M, N, K, st);
  }
  // fp32 parity path: the FFMA GEMM writes the pre-LayerNorm rows, the row kernel normalises them in place
  int rc = icka_sgemm_launch(h, static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, bias, residual,
                             out_f32, N, ICKA_F32, M, N, K, ICKA_ACT_NONE, nullptr, st);
  if (rc) return rc;
  return icka_layernorm_fwd(h, out_f32, gamma, beta, eps, out_f32, out_bf16, M, N, stream);
}

extern "C" int icka_linear_fwd_ex(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                                  const float* bias, const float* residual, void* out, int64_t ldo, void* pre_act_out,
                                  int in_dtype, int out_dtype, int M, int N, int K, int act, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(A && W && out, "linear: null pointer");
  ICKA_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear: bad shape M=%d N=%d K=%d", M, N, K);
  ICKA_REQUIRE(lda >= K && ldw >= K && ldo >= N, "linear: pitches smaller than the logical extents");
  ICKA_REQUIRE(in_dtype == ICKA_F32 || in_dtype == ICKA_BF16, "linear: bad in_dtype %d", in_dtype);
  ICKA_REQUIRE(out_dtype == ICKA_F32 || out_dtype == ICKA_BF16, "linear: bad out_dtype %d", out_dtype);
  ICKA_REQUIRE(act == ICKA_ACT_NONE || act == ICKA_ACT_GELU_ERF || act == ICKA_ACT_TANH || act == ICKA_ACT_RELU ||
                   act == ICKA_ACT_SWISH, "linear: bad activation %d", act);
  ICKA_REQUIRE(!pre_act_out || act == ICKA_ACT_GELU_ERF || act == ICKA_ACT_RELU || act == ICKA_ACT_SWISH,
               "linear: pre_act_out is only defined for the FFN activations (gelu / relu / swish)");
  if (M == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == ICKA_BF16)
    return icka_gemm_bf16_launch(h, A, lda, W, ldw, bias, residual, out, ldo, out_dtype, M, N, K, act, pre_act_out, st);
  return icka_sgemm_launch(h, static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, bias, residual,
                           out, ldo, out_dtype, M, N, K, act, static_cast<float*>(pre_act_out), st);
}

extern "C" int icka_linear_fwd(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                               const float* bias, const float* residual, void* out, int64_t ldo, int in_dtype,
                               int out_dtype, int M, int N, int K, int act, void* stream) {
  return icka_linear_fwd_ex(h, A, lda, W, ldw, bias, residual, out, ldo, nullptr, in_dtype, out_dtype, M, N, K, act,
                            stream);
}

extern "C" int icka_linear_dgrad(icka_handle* h, const void* dY, int64_t ldd, const void* W, int64_t ldw,
                                 const float* residual, const void* gelu_pre, int64_t ldg, void* dX, int64_t ldo,
                                 int in_dtype, int out_dtype, int M, int N, int K, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(dY && W && dX, "linear_dgrad: null pointer");
  ICKA_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear_dgrad: bad shape M=%d N=%d K=%d", M, N, K);
  ICKA_REQUIRE(ldd >= N && ldw >= K && ldo >= K, "linear_dgrad: pitches smaller than the logical extents");
  ICKA_REQUIRE(!gelu_pre || ldg >= K, "linear_dgrad: gelu_pre pitch smaller than K");
  ICKA_REQUIRE(in_dtype == ICKA_F32 || in_dtype == ICKA_BF16, "linear_dgrad: bad in_dtype %d", in_dtype);
  ICKA_REQUIRE(out_dtype == ICKA_F32 || out_dtype == ICKA_BF16, "linear_dgrad: bad out_dtype %d", out_dtype);
  if (M == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == ICKA_BF16)
    return icka_gemm_bf16_dgrad_launch(h, dY, ldd, W, ldw, residual, gelu_pre, ldg, dX, ldo, out_dtype, M, N, K, st);
  ICKA_REQUIRE(out_dtype == ICKA_F32, "linear_dgrad(fp32): output must be fp32");
  return icka_sgemm_strided_launch(h, static_cast<const float*>(dY), ldd, 1, static_cast<const float*>(W), ldw, 1,
                                   residual, static_cast<const float*>(gelu_pre), ldg, static_cast<float*>(dX), ldo,
                                   M, K, N, 0, st);
}

extern "C" int icka_linear_wgrad(icka_handle* h, const void* dY, int64_t ldd, const void* X, int64_t ldx, float* dW,
                                 int in_dtype, int M, int N, int K, int accumulate, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(dY && X && dW, "linear_wgrad: null pointer");
  ICKA_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear_wgrad: bad shape M=%d N=%d K=%d", M, N, K);
  ICKA_REQUIRE(ldd >= N && ldx >= K, "linear_wgrad: pitches smaller than the logical extents");
  ICKA_REQUIRE(in_dtype == ICKA_F32 || in_dtype == ICKA_BF16, "linear_wgrad: bad in_dtype %d", in_dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == ICKA_BF16) {
    if (!accumulate) ICKA_CUDA(cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), st));
    if (M == 0) return ICKA_OK;
    return icka_gemm_bf16_wgrad_launch(h, dY, ldd, X, ldx, dW, K, M, N, K, st);
  }
  if (M == 0) {
    if (!accumulate) ICKA_CUDA(cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), st));
    return ICKA_OK;
  }
  // dW[n,k] = sum_m dY[m,n] X[m,k]: A(n,m) = dY[m*ldd + n], B(m,k) = X[m*ldx + k]
  return icka_sgemm_strided_launch(h, static_cast<const float*>(dY), 1, ldd, static_cast<const float*>(X), ldx, 1,
                                   nullptr, nullptr, 0, dW, K, N, K, M, accumulate, st);
}
