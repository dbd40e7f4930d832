// icka_linear_fwd: nn.Linear + fused epilogue, dispatched on the operand precision mode
// (bf16 -> tcgen05 tensor-core kernel, fp32 -> FFMA parity kernel).  See include/icka_b200.h.
#include "common.cuh"

int icka_sgemm_launch(icka_handle* h, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                      const float* residual, void* out, int64_t ldo, int out_dtype, int M, int N, int K, int act,
                      cudaStream_t st);
int icka_gemm_bf16_launch(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                          const float* residual, void* out, int64_t ldo, int out_dtype, int M, int N, int K, int act,
                          cudaStream_t st);

extern "C" int icka_linear_fwd(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                               const float* bias, const float* residual, void* out, int64_t ldo, int in_dtype,
                               int out_dtype, int M, int N, int K, int act, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(A && W && out, "linear: null pointer");
  ICKA_REQUIRE(M >= 0 && N >= 1 && K >= 1, "linear: bad shape M=%d N=%d K=%d", M, N, K);
  ICKA_REQUIRE(lda >= K && ldw >= K && ldo >= N, "linear: pitches smaller than the logical extents");
  ICKA_REQUIRE(in_dtype == ICKA_F32 || in_dtype == ICKA_BF16, "linear: bad in_dtype %d", in_dtype);
  ICKA_REQUIRE(out_dtype == ICKA_F32 || out_dtype == ICKA_BF16, "linear: bad out_dtype %d", out_dtype);
  ICKA_REQUIRE(act == ICKA_ACT_NONE || act == ICKA_ACT_GELU_ERF, "linear: bad activation %d", act);
  if (M == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == ICKA_BF16)
    return icka_gemm_bf16_launch(h, A, lda, W, ldw, bias, residual, out, ldo, out_dtype, M, N, K, act, st);
  return icka_sgemm_launch(h, static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, bias, residual,
                           out, ldo, out_dtype, M, N, K, act, st);
}
