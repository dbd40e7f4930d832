// Shared host/device helpers for libicka_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/icka_b200.h"

struct icka_handle {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  size_t smem_optin;
  std::atomic<long long> launches;
  void* encode_tiled;   // PFN cuTensorMapEncodeTiled, resolved through the runtime (no -lcuda)
};

void icka_set_error(const char* fmt, ...);

#define ICKA_FAIL(code, ...)      \
  do {                            \
    icka_set_error(__VA_ARGS__);  \
    return (code);                \
  } while (0)

#define ICKA_REQUIRE(cond, ...)                           \
  do {                                                    \
    if (!(cond)) ICKA_FAIL(ICKA_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define ICKA_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      ICKA_FAIL(ICKA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                __LINE__);                                                                     \
  } while (0)

// After a kernel launch: catch launch-configuration errors without synchronising.
#define ICKA_LAUNCHED(h)                 \
  do {                                   \
    ICKA_CUDA(cudaGetLastError());       \
    (h)->launches.fetch_add(1);          \
  } while (0)

#define ICKA_CHECK_HANDLE(h) ICKA_REQUIRE((h) != nullptr, "null handle")

static inline bool icka_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// erf-GELU exactly as the reference writes it (CMIM:31-37): x * 0.5 * (1 + erf(x / sqrt(2)))
__device__ __forceinline__ float gelu_erf(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f)); }

// Same function through the Abramowitz-Stegun 7.1.26 rational form of erf (|error| <= 1.5e-7 before
// rounding): 2 MUFU + ~12 FMA-pipe instructions instead of erff's ~30.  Used where the result is rounded
// to bf16 (2^-9 relative) straight away, so the epilogue of the FFN-up GEMM keeps pace with the MMA.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * az * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(hx, copysignf(erf_abs, z), hx);
}
