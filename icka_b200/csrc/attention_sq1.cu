// Single-query cross-attention core, forward and backward: the image->text encoders (CMIM:981-989) attend ONE CLIP token
// per sentence to the S text states.  In fp32 mode and on the training path they run the reference formulation
// (BertCoAttention, CMIM:598-623) instead of the folded inference form, and the general kernels -- built around 64/128-row
// query tiles -- spent 212 us per backward call on tiles with one valid row (two calls per training step).
//
// One warp per (sentence, head).  K and V of the head ([Skv, 64]) are staged once in shared memory with coalesced 16-byte
// loads (the tile pitch is padded by 16 B so that row-wise 16-byte reads are bank-conflict-free); the rest is a few hundred
// FMAs per lane:
//   scores, dP   lane = key: dot products of a tile row with q (and dO), softmax statistics by warp shuffles
//   ctx, dq      lane = two head dims: a loop over the keys, the probability broadcast from shared memory
//   dK, dV       rank-1 rows  ds[s] q / 8  and  p'[s] dO  written straight to global memory, one row per warp store
// Dropout on the probabilities (CMIM:616) uses the Philox indexing every attention kernel of the library shares
// (philox.cuh attn_group), so forward, backward and the test-side mask replay agree.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr int kD = 64;
constexpr int kMaxWarps = 8;

struct Sq1Drop {
  uint32_t thresh;   // 0 = off
  float scale;
  uint64_t seed;
  const unsigned long long* base;
};

template <typename T>
struct Tile {
  static constexpr int kVec = 16 / (int)sizeof(T);              // elements per 16-byte piece
  static constexpr int kPitch = kD * (int)sizeof(T) + 16;       // bytes per key row
};

template <typename T>
__device__ __forceinline__ void loadv(const uint8_t* p, float* v);
template <>
__device__ __forceinline__ void loadv<float>(const uint8_t* p, float* v) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void loadv<__nv_bfloat16>(const uint8_t* p, float* v) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
  v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ float2 load2(const uint8_t* p);
template <>
__device__ __forceinline__ float2 load2<float>(const uint8_t* p) { return *reinterpret_cast<const float2*>(p); }
template <>
__device__ __forceinline__ float2 load2<__nv_bfloat16>(const uint8_t* p) {
  const uint32_t t = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(t << 16), __uint_as_float(t & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void store2(T* p, float a, float b);
template <>
__device__ __forceinline__ void store2<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <>
__device__ __forceinline__ void store2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b);
}
template <typename T>
__device__ __forceinline__ float exp_of(float x);                 // fp32 keeps expf (1e-5 parity), bf16 the MUFU form
template <>
__device__ __forceinline__ float exp_of<float>(float x) { return expf(x); }
template <>
__device__ __forceinline__ float exp_of<__nv_bfloat16>(float x) { return exp2f(x * 1.4426950408889634f); }

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stage one [Skv, 64] head tile (global row pitch ld elements) into the padded shared-memory tile: cp.async, so that the
// whole tile is in flight at once (a warp is alone with its 32 KB of K and V, five warps per SM: with loads staged through
// registers, four per lane in flight, the kernel ran at 1.5 TB/s -- 45 us per forward launch at 128 sentences)
template <typename T>
__device__ __forceinline__ void stage_tile(const T* __restrict__ g, int64_t ld, int Skv, uint8_t* s, int lane) {
  constexpr int kVecPerRow = kD / Tile<T>::kVec;                // 8 (bf16) or 16 (fp32) pieces per row
  const int n = Skv * kVecPerRow;
  for (int i = lane; i < n; i += 32) {
    const int r = i / kVecPerRow, c = i % kVecPerRow;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s + (size_t)r * Tile<T>::kPitch + c * 16);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(reinterpret_cast<const uint4*>(g + (size_t)r * ld) + c));
  }
}

__host__ __device__ inline int round4(int n) { return (n + 3) / 4 * 4; }

// per-warp shared memory: K tile | V tile | q[64] | dO[64] | a[Skv] | b[Skv] | c[Skv]   (floats after the tiles)
template <typename T>
__host__ __device__ inline size_t warp_smem_bytes(int Skv) {
  return (size_t)2 * Skv * Tile<T>::kPitch + (size_t)(2 * kD + 3 * round4(Skv)) * sizeof(float);
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(32 * kMaxWarps)
attn_sq1_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, const T* __restrict__ v, int64_t ldkv,
                const float* __restrict__ mask_add, T* __restrict__ ctx, int64_t ldc,                        // forward
                const T* __restrict__ dctx, int64_t lddc, T* __restrict__ dq, int64_t lddq, T* __restrict__ dk,
                T* __restrict__ dv, int64_t lddkv,                                                              // backward
                int items, int Skv, int nh, Sq1Drop drop) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int kVec = Tile<T>::kVec;
  constexpr int kPitch = Tile<T>::kPitch;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + warp;
  if (item >= items) return;                                        // warps are independent: no block-wide barrier below
  const int b = item / nh, h = item % nh;
  uint8_t* ks = smem_raw + (size_t)warp * warp_smem_bytes<T>(Skv);
  uint8_t* vs = ks + (size_t)Skv * kPitch;
  float* qs = reinterpret_cast<float*>(vs + (size_t)Skv * kPitch);
  float* dos = qs + kD;
  float* pa = dos + kD;                  // scores -> exponentials -> dropped probabilities p'
  float* pb = pa + round4(Skv);          // dP -> ds
  float* pc = pb + round4(Skv);          // undropped probabilities p (backward)

  stage_tile<T>(k + (size_t)b * Skv * ldkv + (size_t)h * kD, ldkv, Skv, ks, lane);
  stage_tile<T>(v + (size_t)b * Skv * ldkv + (size_t)h * kD, ldkv, Skv, vs, lane);
  {
    const T* qg = q + (size_t)b * ldq + (size_t)h * kD;
    const float2 t = load2<T>(reinterpret_cast<const uint8_t*>(qg + 2 * lane));
    qs[2 * lane] = t.x;
    qs[2 * lane + 1] = t.y;
    if (BWD) {
      const T* dg = dctx + (size_t)b * lddc + (size_t)h * kD;
      const float2 u = load2<T>(reinterpret_cast<const uint8_t*>(dg + 2 * lane));
      dos[2 * lane] = u.x;
      dos[2 * lane + 1] = u.y;
    }
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncwarp();

  // ---- lane = key: score = q . K[s] / 8 + mask,  dP'[s] = dO . V[s] ----
  float mx = -INFINITY;
  {
    float qr[kD], dr[BWD ? kD : 1];
#pragma unroll
    for (int c = 0; c < kD; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qs + c);
      qr[c] = t.x; qr[c + 1] = t.y; qr[c + 2] = t.z; qr[c + 3] = t.w;
      if (BWD) {
        const float4 u = *reinterpret_cast<const float4*>(dos + c);
        dr[c] = u.x; dr[c + 1] = u.y; dr[c + 2] = u.z; dr[c + 3] = u.w;
      }
    }
    for (int s = lane; s < Skv; s += 32) {
      const uint8_t* kr = ks + (size_t)s * kPitch;
      const uint8_t* vr = vs + (size_t)s * kPitch;
      float a0 = 0.0f, a1 = 0.0f, d0 = 0.0f, d1 = 0.0f;
#pragma unroll
      for (int c = 0; c < kD; c += kVec) {
        float kk[kVec];
        loadv<T>(kr + c * sizeof(T), kk);
#pragma unroll
        for (int j = 0; j < kVec; j += 2) {
          a0 = fmaf(kk[j], qr[c + j], a0);
          a1 = fmaf(kk[j + 1], qr[c + j + 1], a1);
        }
        if (BWD) {
          float vv[kVec];
          loadv<T>(vr + c * sizeof(T), vv);
#pragma unroll
          for (int j = 0; j < kVec; j += 2) {
            d0 = fmaf(vv[j], dr[c + j], d0);
            d1 = fmaf(vv[j + 1], dr[c + j + 1], d1);
          }
        }
      }
      const float sc = (a0 + a1) * 0.125f + (mask_add ? __ldg(mask_add + (size_t)b * Skv + s) : 0.0f);
      pa[s] = sc;
      if (BWD) pb[s] = d0 + d1;
      mx = fmaxf(mx, sc);
    }
  }
  mx = wmax(mx);
  float l = 0.0f;
  for (int s = lane; s < Skv; s += 32) {     // each lane revisits only the entries it wrote
    const float e = exp_of<T>(pa[s] - mx);
    pa[s] = e;
    l += e;
  }
  l = wsum(l);
  const float inv_l = 1.0f / l;
  const uint64_t seed = icka_rng::effective_seed(drop.seed, drop.base);
  const uint64_t drow = (uint64_t)b * nh + h;                      // row index ((b nh + h) Sq + 0) with Sq = 1
  float delta = 0.0f;
  for (int s = lane; s < Skv; s += 32) {
    const float p = pa[s] * inv_l;
    float keep = 1.0f;
    if (drop.thresh) {
      const uint32_t bits =
          icka_rng::keep_bits4(seed, icka_rng::kSiteAttention, icka_rng::attn_group(drow, Skv, s), drop.thresh);
      keep = ((bits >> (s & 3)) & 1u) ? drop.scale : 0.0f;
    }
    pa[s] = p * keep;                         // P' = P keep / (1 - p_drop): what multiplies V
    if (BWD) {
      const float dp = pb[s] * keep;          // dP = dP' keep / (1 - p_drop)
      pb[s] = dp;
      pc[s] = p;
      delta = fmaf(p, dp, delta);
    }
  }
  if (BWD) {
    delta = wsum(delta);
    for (int s = lane; s < Skv; s += 32) pb[s] = pc[s] * (pb[s] - delta) * 0.125f;     // dS / 8: the 1/sqrt(d) folded in
  }
  __syncwarp();

  // ---- lane = head dims 2 lane, 2 lane + 1 ----
  const int dcol = 2 * lane;
  if (!BWD) {
    float c0 = 0.0f, c1 = 0.0f, e0 = 0.0f, e1 = 0.0f;
    int s = 0;
    for (; s + 1 < Skv; s += 2) {
      const float2 va = load2<T>(vs + (size_t)s * kPitch + dcol * sizeof(T));
      const float2 vb = load2<T>(vs + (size_t)(s + 1) * kPitch + dcol * sizeof(T));
      const float p0 = pa[s], p1 = pa[s + 1];
      c0 = fmaf(p0, va.x, c0); c1 = fmaf(p0, va.y, c1);
      e0 = fmaf(p1, vb.x, e0); e1 = fmaf(p1, vb.y, e1);
    }
    if (s < Skv) {
      const float2 va = load2<T>(vs + (size_t)s * kPitch + dcol * sizeof(T));
      c0 = fmaf(pa[s], va.x, c0); c1 = fmaf(pa[s], va.y, c1);
    }
    store2<T>(ctx + (size_t)b * ldc + (size_t)h * kD + dcol, c0 + e0, c1 + e1);
    return;
  }
  {
    const float q0 = qs[dcol], q1 = qs[dcol + 1], o0 = dos[dcol], o1 = dos[dcol + 1];
    float c0 = 0.0f, c1 = 0.0f;
    T* dkp = dk + (size_t)b * Skv * lddkv + (size_t)h * kD + dcol;
    T* dvp = dv + (size_t)b * Skv * lddkv + (size_t)h * kD + dcol;
#pragma unroll 4
    for (int s = 0; s < Skv; ++s) {
      const float ds = pb[s], pd = pa[s];
      const float2 kk = load2<T>(ks + (size_t)s * kPitch + dcol * sizeof(T));
      c0 = fmaf(ds, kk.x, c0);
      c1 = fmaf(ds, kk.y, c1);
      store2<T>(dkp + (size_t)s * lddkv, ds * q0, ds * q1);
      store2<T>(dvp + (size_t)s * lddkv, pd * o0, pd * o1);
    }
    store2<T>(dq + (size_t)b * lddq + (size_t)h * kD + dcol, c0, c1);
  }
}

template <typename T, bool BWD>
int launch_sq1(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const float* mask_add,
               void* ctx, int64_t ldc, const void* dctx, int64_t lddc, void* dq, int64_t lddq, void* dk, void* dv,
               int64_t lddkv, int B, int Skv, int nh, Sq1Drop drop, cudaStream_t st) {
  const size_t wbytes = warp_smem_bytes<T>(Skv);
  int warps = (int)(h->smem_optin / wbytes);
  if (warps < 1) return 1;                                         // tile does not fit: the caller's general kernel takes it
  if (warps > kMaxWarps) warps = kMaxWarps;
  const int items = B * nh;
  // spread the items over the SMs before filling the blocks: a block is resident alone when its tiles are large
  const int per_sm = (items + h->sm_count - 1) / h->sm_count;
  if (warps > per_sm) warps = per_sm < 1 ? 1 : per_sm;
  const size_t smem = wbytes * warps;
  ICKA_CUDA(cudaFuncSetAttribute(attn_sq1_kernel<T, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
  attn_sq1_kernel<T, BWD><<<(items + warps - 1) / warps, 32 * warps, smem, st>>>(
      static_cast<const T*>(q), ldq, static_cast<const T*>(k), static_cast<const T*>(v), ldkv, mask_add,
      static_cast<T*>(ctx), ldc, static_cast<const T*>(dctx), lddc, static_cast<T*>(dq), lddq, static_cast<T*>(dk),
      static_cast<T*>(dv), lddkv, items, Skv, nh, drop);
  ICKA_LAUNCHED(h);
  return 0;
}

}  // namespace

// Both return 0 = launched, < 0 = error, > 0 = shape outside this kernel's envelope (the caller falls through).
int icka_attn_sq1_fwd_launch(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                             const float* mask_add, void* ctx, int64_t ldc, int dtype, int B, int Skv, int nh,
                             uint32_t thresh, float scale, uint64_t seed, const unsigned long long* base, cudaStream_t st) {
  const Sq1Drop drop{thresh, scale, seed, base};
  if ((int64_t)B * nh > INT32_MAX) return 1;
  if (dtype == ICKA_BF16)
    return launch_sq1<__nv_bfloat16, false>(h, q, ldq, k, v, ldkv, mask_add, ctx, ldc, nullptr, 0, nullptr, 0, nullptr,
                                            nullptr, 0, B, Skv, nh, drop, st);
  return launch_sq1<float, false>(h, q, ldq, k, v, ldkv, mask_add, ctx, ldc, nullptr, 0, nullptr, 0, nullptr, nullptr, 0, B,
                                  Skv, nh, drop, st);
}

int icka_attn_sq1_bwd_launch(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                             const float* mask_add, const void* dctx, int64_t lddc, void* dq, int64_t lddq, void* dk,
                             void* dv, int64_t lddkv, int dtype, int B, int Skv, int nh, uint32_t thresh, float scale,
                             uint64_t seed, const unsigned long long* base, cudaStream_t st) {
  const Sq1Drop drop{thresh, scale, seed, base};
  if ((int64_t)B * nh > INT32_MAX) return 1;
  if (dtype == ICKA_BF16)
    return launch_sq1<__nv_bfloat16, true>(h, q, ldq, k, v, ldkv, mask_add, nullptr, 0, dctx, lddc, dq, lddq, dk, dv, lddkv,
                                           B, Skv, nh, drop, st);
  return launch_sq1<float, true>(h, q, ldq, k, v, ldkv, mask_add, nullptr, 0, dctx, lddc, dq, lddq, dk, dv, lddkv, B, Skv,
                                 nh, drop, st);
}
