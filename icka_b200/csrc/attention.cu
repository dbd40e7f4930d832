// Cross-attention core (BertCoAttention, CMIM:598-623): scores -> +mask -> softmax -> P.V -> merge heads.
//
// The key side of ICKA's cross-attention is tiny (49 regions; 196 for the 448-px variant; 128 text
// tokens for image->text), so the whole K_h / V_h of one (sentence, head) stays resident in shared
// memory and each thread owns one query row: the score row, the softmax statistics and the 64-wide
// context accumulator never leave registers and the merged-head layout is written directly.
// The kernel is bound by the HBM traffic of Q, K|V and ctx (543,744 B / sentence / layer
// in bf16), not by its 19 MFLOP / sentence.
//
// Arithmetic: fp32 throughout; scores = (q.k) * 0.125 + mask  (the reference divides by sqrt(64)
// AFTER the dot product, CMIM:605-607; x0.125 is exact); softmax is evaluated online (running max)
// which reassociates the reference's exp(s - max)/sum within fp32 rounding.
#include "common.cuh"

namespace {

constexpr int kD = 64;
constexpr int kRows = 128;   // query rows per block == threads per block

template <typename T>
__device__ __forceinline__ void load8(const T* p, float* out);

template <>
__device__ __forceinline__ void load8<float>(const float* p, float* out) {
  const float4 a = reinterpret_cast<const float4*>(p)[0];
  const float4 b = reinterpret_cast<const float4*>(p)[1];
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
  out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* out) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float(w[i] << 16);
    out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <typename T>
__device__ __forceinline__ void store8(T* p, const float* v);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float* v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// grid = (ceil(Sq / kRows), nh, B)
template <typename T>
__global__ void __launch_bounds__(kRows) cross_attn_kernel(
    const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, const T* __restrict__ v, int64_t ldkv,
    const float* __restrict__ mask_add, T* __restrict__ ctx, int64_t ldc, int Sq, int Skv) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                      // [Skv][64]
  float* Vs = Ks + (size_t)Skv * kD;     // [Skv][64]
  float* Ms = Vs + (size_t)Skv * kD;     // [Skv]
  const int b = blockIdx.z, h = blockIdx.y;
  const int row = blockIdx.x * kRows + threadIdx.x;

  // cooperative, coalesced load of this head's K and V slices (8 elements per thread per step)
  const T* kb = k + (size_t)b * Skv * ldkv + (size_t)h * kD;
  const T* vb = v + (size_t)b * Skv * ldkv + (size_t)h * kD;
  for (int i = threadIdx.x; i < Skv * (kD / 8); i += kRows) {
    const int r = i / (kD / 8), c = (i % (kD / 8)) * 8;
    float t[8];
    load8<T>(kb + (size_t)r * ldkv + c, t);
    reinterpret_cast<float4*>(Ks + r * kD + c)[0] = make_float4(t[0], t[1], t[2], t[3]);
    reinterpret_cast<float4*>(Ks + r * kD + c)[1] = make_float4(t[4], t[5], t[6], t[7]);
    load8<T>(vb + (size_t)r * ldkv + c, t);
    reinterpret_cast<float4*>(Vs + r * kD + c)[0] = make_float4(t[0], t[1], t[2], t[3]);
    reinterpret_cast<float4*>(Vs + r * kD + c)[1] = make_float4(t[4], t[5], t[6], t[7]);
  }
  for (int i = threadIdx.x; i < Skv; i += kRows) Ms[i] = mask_add ? mask_add[(size_t)b * Skv + i] : 0.0f;
  __syncthreads();
  if (row >= Sq) return;

  float qr[kD];
  const T* qp = q + ((size_t)b * Sq + row) * ldq + (size_t)h * kD;
#pragma unroll
  for (int c = 0; c < kD; c += 8) load8<T>(qp + c, qr + c);

  float acc[kD];
#pragma unroll
  for (int c = 0; c < kD; ++c) acc[c] = 0.0f;
  float m = -INFINITY, l = 0.0f;

  for (int r = 0; r < Skv; ++r) {
    const float4* kr = reinterpret_cast<const float4*>(Ks + r * kD);
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) {
      const float4 kk = kr[c];   // warp-uniform address -> broadcast
      s0 = fmaf(qr[4 * c + 0], kk.x, s0);
      s1 = fmaf(qr[4 * c + 1], kk.y, s1);
      s2 = fmaf(qr[4 * c + 2], kk.z, s2);
      s3 = fmaf(qr[4 * c + 3], kk.w, s3);
    }
    const float s = ((s0 + s1) + (s2 + s3)) * 0.125f + Ms[r];
    if (s > m) {   // running max moved: rescale what has been accumulated so far
      const float scale = expf(m - s);   // m = -inf on the first key -> 0
      l *= scale;
#pragma unroll
      for (int c = 0; c < kD; ++c) acc[c] *= scale;
      m = s;
    }
    const float p = expf(s - m);
    l += p;
    const float4* vr = reinterpret_cast<const float4*>(Vs + r * kD);
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) {
      const float4 vv = vr[c];
      acc[4 * c + 0] = fmaf(p, vv.x, acc[4 * c + 0]);
      acc[4 * c + 1] = fmaf(p, vv.y, acc[4 * c + 1]);
      acc[4 * c + 2] = fmaf(p, vv.z, acc[4 * c + 2]);
      acc[4 * c + 3] = fmaf(p, vv.w, acc[4 * c + 3]);
    }
  }
  const float inv = 1.0f / l;
  T* op = ctx + ((size_t)b * Sq + row) * ldc + (size_t)h * kD;
#pragma unroll
  for (int c = 0; c < kD; c += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = acc[c + j] * inv;
    store8<T>(op + c, t);
  }
}

}  // namespace

extern "C" int icka_cross_attn_core_fwd(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                                        int64_t ldkv, const float* mask_add, void* ctx, int64_t ldc, int dtype,
                                        int B, int Sq, int Skv, int nh, int d, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(q && k && v && ctx, "cross_attn: null pointer");
  ICKA_REQUIRE(B >= 0 && Sq >= 1 && Skv >= 1 && nh >= 1, "cross_attn: bad shape");
  ICKA_REQUIRE(d == kD, "cross_attn: head dim %d != 64 (ICKA uses 768/12 and 1024/16)", d);
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "cross_attn: bad dtype %d", dtype);
  ICKA_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldc % 8 == 0, "cross_attn: pitches must be multiples of 8 elements");
  ICKA_REQUIRE(icka_aligned(q, 16) && icka_aligned(k, 16) && icka_aligned(v, 16) && icka_aligned(ctx, 16),
               "cross_attn: pointers must be 16-byte aligned");
  ICKA_REQUIRE(B <= 65535 && nh <= 65535, "cross_attn: B or nh exceeds grid limits; shard the batch");
  if (B == 0) return ICKA_OK;
  const size_t smem = ((size_t)2 * Skv * kD + Skv) * sizeof(float);
  if (smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "cross_attn: Skv=%d needs %zu B shared memory (max %zu)", Skv, smem, h->smem_optin);
  dim3 grid((Sq + kRows - 1) / kRows, nh, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ICKA_BF16) {
    ICKA_CUDA(cudaFuncSetAttribute(cross_attn_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cross_attn_kernel<__nv_bfloat16><<<grid, kRows, smem, st>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k),
        static_cast<const __nv_bfloat16*>(v), ldkv, mask_add, static_cast<__nv_bfloat16*>(ctx), ldc, Sq, Skv);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(cross_attn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cross_attn_kernel<float><<<grid, kRows, smem, st>>>(static_cast<const float*>(q), ldq, static_cast<const float*>(k),
                                                        static_cast<const float*>(v), ldkv, mask_add,
                                                        static_cast<float*>(ctx), ldc, Sq, Skv);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
