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
#include "philox.cuh"

namespace {

constexpr int kD = 64;
constexpr int kRows = 128;   // query rows per block == threads per block

// Dropout on the attention probabilities (CMIM:616), training only: P' = P * keep / (1 - p) goes into P.V while the
// softmax normaliser keeps the undropped sum.  thresh == 0 means "off".
struct DropArgs {
  uint32_t thresh;
  float scale;
  uint64_t seed;
  const unsigned long long* base;   // device-resident seed base (icka_set_seed_base) or null
};

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
    const float* __restrict__ mask_add, T* __restrict__ ctx, int64_t ldc, int Sq, int Skv, DropArgs drop) {
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
  const uint64_t drow = ((uint64_t)b * gridDim.y + h) * (uint64_t)Sq + (uint64_t)row;
  uint32_t keep = 0xfu;

  for (int r = 0; r < Skv; ++r) {
    if (drop.thresh && (r & 3) == 0)
      keep = icka_rng::keep_bits4(icka_rng::effective_seed(drop.seed, drop.base), icka_rng::kSiteAttention, icka_rng::attn_group(drow, Skv, r), drop.thresh);
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
    float p = expf(s - m);
    l += p;
    if (drop.thresh) p = (keep >> (r & 3) & 1u) ? p * drop.scale : 0.0f;
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


// ------------------------------------------------------------------------------------------------
// bf16 path: same contract, scores and P.V on the tensor cores (mma.sync m16n8k16, fp32 accumulate).
// The attention core is 1.2 % of the layer's FLOPs and is bound by the HBM traffic of Q, K|V and ctx, so
// the legacy warp-level MMA is enough to take the arithmetic off the critical path; the tcgen05 budget
// goes to the projections around it.  One block = one (sentence, head, 128-query tile): 8 warps x 16
// query rows.  Q (128x64), and K / V in 64-key blocks, are staged in shared memory with 16-byte cp.async
// (row pitch 144 B -> conflict-free ldmatrix); softmax runs online over key blocks in registers (quad
// shuffles), P is rounded to bf16 for the second MMA, and the context tile is written back through the
// Q buffer so that global stores are 16 bytes per lane and row-contiguous.
// ------------------------------------------------------------------------------------------------
constexpr int kMmaThreads = 256;
constexpr int kPitch = 72;        // bf16 elements per smem row (64 + 8 pad)
constexpr int kKeyBlk = 64;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(sa));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(sa));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// grid = (ceil(Sq / (16 * WARPS)), nh, B); WARPS x 16 query rows per block.  The kernel is a load -> compute ->
// store sequence per block, so HBM stays busy only through the overlap of independent blocks: 64-row blocks
// (4 per SM at 128 registers) interleave better than 128-row ones (2 per SM) at the price of staging K / V twice
// from L2.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) cross_attn_mma_kernel(
    const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
    const __nv_bfloat16* __restrict__ v, int64_t ldkv, const float* __restrict__ mask_add,
    __nv_bfloat16* __restrict__ ctx, int64_t ldc, int Sq, int Skv, DropArgs drop) {
  constexpr int kRowsT = WARPS * 16, kThreadsT = WARPS * 32;
  __shared__ __align__(16) __nv_bfloat16 Qs[kRowsT * kPitch];
  __shared__ __align__(16) __nv_bfloat16 Ks[kKeyBlk * kPitch];
  __shared__ __align__(16) __nv_bfloat16 Vs[kKeyBlk * kPitch];
  __shared__ float Ms[kKeyBlk];
  const int b = blockIdx.z, h = blockIdx.y;
  const int row0 = blockIdx.x * kRowsT;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int g = lane >> 2, t = lane & 3;
  constexpr float kLog2e = 1.4426950408889634f;

  // ---- stage the Q tile (rows beyond Sq are zero-filled) ----
  const __nv_bfloat16* qb = q + ((size_t)b * Sq + row0) * ldq + (size_t)h * kD;
  for (int i = tid; i < kRowsT * 8; i += kThreadsT) {
    const int r = i >> 3, c = (i & 7) * 8;
    if (row0 + r < Sq) cp_async16(Qs + r * kPitch + c, qb + (size_t)r * ldq + c);
    else *reinterpret_cast<uint4*>(Qs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
  }

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;   // rows g and g+8 of this warp's 16
  uint32_t qf[4][4];
  const bool warp_active = row0 + warp * 16 < Sq;

  const __nv_bfloat16* kb = k + (size_t)b * Skv * ldkv + (size_t)h * kD;
  const __nv_bfloat16* vb = v + (size_t)b * Skv * ldkv + (size_t)h * kD;
  for (int key0 = 0; key0 < Skv; key0 += kKeyBlk) {
    if (key0 > 0) __syncthreads();   // previous block fully consumed
    for (int i = tid; i < kKeyBlk * 8; i += kThreadsT) {
      const int r = i >> 3, c = (i & 7) * 8;
      if (key0 + r < Skv) {
        cp_async16(Ks + r * kPitch + c, kb + (size_t)(key0 + r) * ldkv + c);
        cp_async16(Vs + r * kPitch + c, vb + (size_t)(key0 + r) * ldkv + c);
      } else {
        *reinterpret_cast<uint4*>(Ks + r * kPitch + c) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Vs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
      }
    }
    for (int i = tid; i < kKeyBlk; i += kThreadsT) {
      const int key = key0 + i;
      Ms[i] = (key < Skv) ? (mask_add ? mask_add[(size_t)b * Skv + key] * kLog2e : 0.0f) : -INFINITY;
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    if (!warp_active) continue;

    if (key0 == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ldmatrix_x4(qf[kk], Qs + (warp * 16 + (lane & 15)) * kPitch + kk * 16 + (lane >> 4) * 8);
    }
    const int nkt = min(8, (Skv - key0 + 7) / 8);   // 8-key tiles that hold at least one real key

    // ---- S = Q K^T for this key block ----
    float sacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
      if (j < nkt) {
        uint32_t kf[4];
        ldmatrix_x4(kf, Ks + (j * 8 + (lane & 7)) * kPitch + (lane >> 3) * 8);          // d = 0..31
        mma_bf16_16816(sacc[j], qf[0], kf[0], kf[1]);
        mma_bf16_16816(sacc[j], qf[1], kf[2], kf[3]);
        ldmatrix_x4(kf, Ks + (j * 8 + (lane & 7)) * kPitch + 32 + (lane >> 3) * 8);     // d = 32..63
        mma_bf16_16816(sacc[j], qf[2], kf[0], kf[1]);
        mma_bf16_16816(sacc[j], qf[3], kf[2], kf[3]);
      }
    }
    // ---- scale, mask, online softmax (log2 domain) ----
    float bm0 = -INFINITY, bm1 = -INFINITY;
    constexpr float kScale = 0.125f * kLog2e;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mk0 = Ms[j * 8 + 2 * t], mk1 = Ms[j * 8 + 2 * t + 1];
      sacc[j][0] = fmaf(sacc[j][0], kScale, mk0);
      sacc[j][1] = fmaf(sacc[j][1], kScale, mk1);
      sacc[j][2] = fmaf(sacc[j][2], kScale, mk0);
      sacc[j][3] = fmaf(sacc[j][3], kScale, mk1);
      bm0 = fmaxf(bm0, fmaxf(sacc[j][0], sacc[j][1]));
      bm1 = fmaxf(bm1, fmaxf(sacc[j][2], sacc[j][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    const float sc0 = fast_exp2(m0 - mn0), sc1 = fast_exp2(m1 - mn1);   // first block: exp2(-inf) = 0
    m0 = mn0;
    m1 = mn1;
    float ps0 = 0.0f, ps1 = 0.0f;
    uint32_t pf[8][2];
    const uint64_t drow0 = ((uint64_t)b * gridDim.y + h) * (uint64_t)Sq + (uint64_t)(row0 + warp * 16 + g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p0 = fast_exp2(sacc[j][0] - mn0), p1 = fast_exp2(sacc[j][1] - mn0);
      float p2 = fast_exp2(sacc[j][2] - mn1), p3 = fast_exp2(sacc[j][3] - mn1);
      ps0 += p0 + p1;
      ps1 += p2 + p3;
      if (drop.thresh) {   // this thread's two keys of the tile share one Philox group of four
        const int key = key0 + j * 8 + 2 * t;
        const uint32_t k0 = icka_rng::keep_bits4(icka_rng::effective_seed(drop.seed, drop.base), icka_rng::kSiteAttention, icka_rng::attn_group(drow0, Skv, key), drop.thresh) >> (key & 3);
        const uint32_t k1 = icka_rng::keep_bits4(icka_rng::effective_seed(drop.seed, drop.base), icka_rng::kSiteAttention, icka_rng::attn_group(drow0 + 8, Skv, key), drop.thresh) >> (key & 3);
        p0 = (k0 & 1u) ? p0 * drop.scale : 0.0f;
        p1 = (k0 & 2u) ? p1 * drop.scale : 0.0f;
        p2 = (k1 & 1u) ? p2 * drop.scale : 0.0f;
        p3 = (k1 & 2u) ? p3 * drop.scale : 0.0f;
      }
      pf[j][0] = pack_bf16x2(p0, p1);
      pf[j][1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * sc0 + ps0;
    l1 = l1 * sc1 + ps1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= sc0; o[j][1] *= sc0; o[j][2] *= sc1; o[j][3] *= sc1;
    }
    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks * 2 < nkt) {
        const uint32_t a[4] = {pf[2 * ks][0], pf[2 * ks][1], pf[2 * ks + 1][0], pf[2 * ks + 1][1]};
#pragma unroll
        for (int jn = 0; jn < 8; jn += 2) {
          uint32_t vf[4];
          ldmatrix_x4_trans(vf, Vs + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + jn * 8 + (lane >> 4) * 8);
          mma_bf16_16816(o[jn], a, vf[0], vf[1]);
          mma_bf16_16816(o[jn + 1], a, vf[2], vf[3]);
        }
      }
    }
  }

  // ---- normalise, stage through the Q buffer (each warp owns its 16 rows), coalesced write-out ----
  if (warp_active) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __nv_bfloat16* ow = Qs + (warp * 16) * kPitch;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(ow + g * kPitch + j * 8 + 2 * t) = pack_bf16x2(o[j][0] * i0, o[j][1] * i0);
      *reinterpret_cast<uint32_t*>(ow + (g + 8) * kPitch + j * 8 + 2 * t) = pack_bf16x2(o[j][2] * i1, o[j][3] * i1);
    }
    __syncwarp();
    __nv_bfloat16* ob = ctx + ((size_t)b * Sq + row0 + warp * 16) * ldc + (size_t)h * kD;
#pragma unroll
    for (int i = lane; i < 16 * 8; i += 32) {
      const int r = i >> 3, c = (i & 7) * 8;
      if (row0 + warp * 16 + r < Sq)
        *reinterpret_cast<uint4*>(ob + (size_t)r * ldc + c) = *reinterpret_cast<const uint4*>(ow + r * kPitch + c);
    }
  }
}

}  // namespace

int icka_attn_sq1_fwd_launch(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                             const float* mask_add, void* ctx, int64_t ldc, int dtype, int B, int Skv, int nh,
                             uint32_t thresh, float scale, uint64_t seed, const unsigned long long* base, cudaStream_t st);
int icka_attn_tcgen05_launch(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                             const float* mask_add, void* ctx, int64_t ldc, int B, int Sq, int Skv, int nh,
                             uint32_t drop_thresh, float drop_scale, uint64_t seed, const unsigned long long* seed_base,
                             cudaStream_t st);

// 0 = pick per shape (tcgen05 kernels for Skv <= 224, mma.sync beyond), 1 = always the mma.sync kernel,
// 2 = the first wide tcgen05 variant for 64 < Skv <= 224 (one softmax group, P through shared memory),
// 3 = same as 0 (the second wide variant -- two softmax groups, P in tensor memory -- is the default)
int g_attn_mode = 0;   // shared with i2t_pool.cu and attention_sm100.cu
extern "C" int icka_set_attn_mode(int mode) {
  if (mode < 0 || mode > 3) ICKA_FAIL(ICKA_ERR_INVALID, "attention mode %d not in 0..3", mode);
  g_attn_mode = mode;
  return ICKA_OK;
}

extern "C" int icka_cross_attn_core_fwd(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                                        int64_t ldkv, const float* mask_add, void* ctx, int64_t ldc, int dtype,
                                        int B, int Sq, int Skv, int nh, int d, void* stream) {
  return icka_cross_attn_core_fwd_drop(h, q, ldq, k, v, ldkv, mask_add, ctx, ldc, dtype, B, Sq, Skv, nh, d, 0.0f, 0, stream);
}

extern "C" int icka_cross_attn_core_fwd_drop(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                                             int64_t ldkv, const float* mask_add, void* ctx, int64_t ldc, int dtype,
                                             int B, int Sq, int Skv, int nh, int d, float p_drop, uint64_t seed,
                                             void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "cross_attn: dropout p=%f outside [0, 1)", (double)p_drop);
  const DropArgs drop{p_drop > 0.0f ? icka_rng::keep_threshold(p_drop) : 0u, 1.0f / (1.0f - p_drop), seed, h->seed_base};
  ICKA_REQUIRE(q && k && v && ctx, "cross_attn: null pointer");
  ICKA_REQUIRE(B >= 0 && Sq >= 1 && Skv >= 1 && nh >= 1, "cross_attn: bad shape");
  ICKA_REQUIRE(d == kD, "cross_attn: head dim %d != 64 (ICKA uses 768/12 and 1024/16)", d);
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "cross_attn: bad dtype %d", dtype);
  ICKA_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldc % 8 == 0, "cross_attn: pitches must be multiples of 8 elements");
  ICKA_REQUIRE(icka_aligned(q, 16) && icka_aligned(k, 16) && icka_aligned(v, 16) && icka_aligned(ctx, 16),
               "cross_attn: pointers must be 16-byte aligned");
  ICKA_REQUIRE(B <= 65535 && nh <= 65535, "cross_attn: B or nh exceeds grid limits; shard the batch");
  if (B == 0) return ICKA_OK;
  if (Sq == 1 && g_attn_mode != 1) {   // image->text encoders: one query per sentence (attention_sq1.cu)
    const int rc = icka_attn_sq1_fwd_launch(h, q, ldq, k, v, ldkv, mask_add, ctx, ldc, dtype, B, Skv, nh, drop.thresh,
                                            drop.scale, drop.seed, drop.base, static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  const size_t smem = ((size_t)2 * Skv * kD + Skv) * sizeof(float);
  if (smem > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "cross_attn: Skv=%d needs %zu B shared memory (max %zu)", Skv, smem, h->smem_optin);
  dim3 grid((Sq + kRows - 1) / kRows, nh, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ICKA_BF16 && g_attn_mode != 1) {
    const int rc = icka_attn_tcgen05_launch(h, q, ldq, k, v, ldkv, mask_add, ctx, ldc, B, Sq, Skv, nh, drop.thresh,
                                            drop.scale, drop.seed, drop.base, st);
    if (rc <= 0) return rc;      // launched (0) or failed (< 0); > 0: shape outside that kernel's envelope
  }
  if (dtype == ICKA_BF16) {
    constexpr int kWarpsPerBlock = 4;
    dim3 grid_mma((Sq + 16 * kWarpsPerBlock - 1) / (16 * kWarpsPerBlock), nh, B);
    ICKA_CUDA(cudaFuncSetAttribute(cross_attn_mma_kernel<kWarpsPerBlock>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
    cross_attn_mma_kernel<kWarpsPerBlock><<<grid_mma, 32 * kWarpsPerBlock, 0, st>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k),
        static_cast<const __nv_bfloat16*>(v), ldkv, mask_add, static_cast<__nv_bfloat16*>(ctx), ldc, Sq, Skv, drop);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(cross_attn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cross_attn_kernel<float><<<grid, kRows, smem, st>>>(static_cast<const float*>(q), ldq, static_cast<const float*>(k),
                                                        static_cast<const float*>(v), ldkv, mask_add,
                                                        static_cast<float*>(ctx), ldc, Sq, Skv, drop);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
