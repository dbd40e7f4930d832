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
  void* workspace;      // device scratch [ICKA_WORKSPACE_BYTES]: split-K partial tiles of skinny forward GEMMs
  const unsigned long long* seed_base;   // device-resident dropout seed base (icka_set_seed_base) or null
};

#define ICKA_WORKSPACE_BYTES ((size_t)32 << 20)

void icka_set_error(const char* fmt, ...);

// 2-D bf16 TMA descriptor: tensor [rows, cols] with row pitch `ld` elements, box {64 cols, box_rows}, SWIZZLE_128B.
int icka_make_tmap_bf16(icka_handle* h, CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows);

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

// Every entry point runs against the HANDLE's device, whatever the caller's current device is (a caller that holds
// tensors on cuda:1 while cuda:0 is current, e.g. one nn.DataParallel worker thread per GPU): set it for the duration of
// the call and restore it, as a PyTorch op's device guard would.
struct icka_device_guard {
  int prev = -1;
  bool switched = false;
  explicit icka_device_guard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~icka_device_guard() {
    if (switched) cudaSetDevice(prev);
  }
  icka_device_guard(const icka_device_guard&) = delete;
  icka_device_guard& operator=(const icka_device_guard&) = delete;
};

#define ICKA_CHECK_HANDLE(h)                       \
  ICKA_REQUIRE((h) != nullptr, "null handle");     \
  icka_device_guard icka_guard_((h)->device)

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

// The other two entries of the reference's ACT2FN table (CMIM:39-43): relu, and swish(x) = x * sigmoid(x).
__device__ __forceinline__ float act_relu(float x) { return fmaxf(x, 0.0f); }
__device__ __forceinline__ float act_swish(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float act_swish_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }   // bf16 outputs
__device__ __forceinline__ float act_relu_grad(float x) { return x > 0.0f ? 1.0f : 0.0f; }
__device__ __forceinline__ float act_swish_grad(float x) {
  const float s = 1.0f / (1.0f + expf(-x));
  return s * fmaf(x, 1.0f - s, 1.0f);
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

// d/dx of the erf-GELU: Phi(x) + x * phi(x)  (exact form for the fp32 path, fast form for bf16 outputs)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
__device__ __forceinline__ float gelu_erf_grad_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;   // exp(-z^2) = exp(-x^2 / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * az * -1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.0f);
  return fmaf(0.5f, copysignf(erf_abs, z), 0.5f) + x * 0.3989422804014327f * e;
}

// Two lanes of the same function on Blackwell's packed fp32x2 FMA pipe (FFMA2 / FMUL2 / FADD2): the
// polynomial, the exponent argument and the final blend take one instruction per PAIR of outputs.
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_splat(float a) { return f2_pack(a, a); }

__device__ __forceinline__ float2 gelu_erf_fast2(float x0, float x1) {
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t z = f2_mul(x, f2_splat(0.70710678118654752440f));
  float z0, z1;
  f2_unpack(z, z0, z1);
  const uint64_t az = f2_pack(fabsf(z0), fabsf(z1));
  float d0, d1;
  f2_unpack(f2_fma(f2_splat(0.3275911f), az, f2_splat(1.0f)), d0, d1);
  float t0, t1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t = f2_pack(t0, t1);
  uint64_t poly = f2_fma(f2_splat(1.061405429f), t, f2_splat(-1.453152027f));
  poly = f2_fma(poly, t, f2_splat(1.421413741f));
  poly = f2_fma(poly, t, f2_splat(-0.284496736f));
  poly = f2_fma(poly, t, f2_splat(0.254829592f));
  poly = f2_mul(poly, t);
  float a0, a1;
  f2_unpack(f2_mul(f2_mul(az, az), f2_splat(-1.4426950408889634f)), a0, a1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  float r0, r1;
  f2_unpack(f2_fma(f2_mul(poly, f2_splat(-1.0f)), f2_pack(e0, e1), f2_splat(1.0f)), r0, r1);
  const uint64_t erfv = f2_pack(copysignf(r0, z0), copysignf(r1, z1));
  const uint64_t hx = f2_mul(x, f2_splat(0.5f));
  float y0, y1;
  f2_unpack(f2_fma(hx, erfv, hx), y0, y1);
  return make_float2(y0, y1);
}

// erf-GELU through a fitted tanh form (used by the FFN-up epilogue, whose bf16 output is the next GEMM's A
// operand):  0.5 x (1 + erf(x / sqrt 2)) = 0.5 x (1 + tanh(x (k1 + k2 x^2 + k3 x^4)))  with max |difference| 2.5e-5
// over the real line (coefficients fitted against the erf form), evaluated with MUFU.TANH (relative error
// 2^-11, i.e. <= 4e-4 absolute at |x| ~ 1.5: a few percent of one bf16 ulp of the result).  6 packed fp32x2
// instructions + 2 MUFU per PAIR of outputs, against 14 + 4 for the rational erf above: the epilogue of the
// K = 768 FFN-up GEMM was MUFU / issue bound (ncu: 606 us vs 464 us for the bare mainloop).
__device__ __forceinline__ float2 gelu_erf_tanh2(float x0, float x1) {
  const uint64_t x = f2_pack(x0, x1);
  const uint64_t t = f2_mul(x, x);
  uint64_t p = f2_fma(t, f2_splat(-3.51516789e-04f), f2_splat(3.70056460e-02f));
  p = f2_fma(t, p, f2_splat(7.97507884e-01f));
  float u0, u1;
  f2_unpack(f2_mul(p, x), u0, u1);
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  const uint64_t hx = f2_mul(x, f2_splat(0.5f));
  float y0, y1;
  f2_unpack(f2_fma(hx, f2_pack(t0, t1), hx), y0, y1);
  return make_float2(y0, y1);
}
