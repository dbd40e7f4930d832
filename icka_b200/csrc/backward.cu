// Backward kernels of the fusion path that are not GEMMs (training, SURVEY 8a20 / 8e: the reference
// trains this stack through torch autograd, My_cross_attention.py:814-844):
//   icka_colsum            bias gradients (column sums of an upstream gradient)
//   icka_layernorm_bwd     backward of BertLayerNorm (CMIM:518-522) fused with the bias-gradient column sums
//   icka_cross_attn_core_bwd  backward of the attention core (CMIM:598-623): dQ, dK, dV from d(ctx)
//   icka_gate_blend_bwd    backward of the gate + blend (CMIM:1029-1036)
// All of them are HBM / L2 bound streaming passes except the attention backward, which recomputes the
// probabilities from Q, K, V (nothing but Q and K|V is saved by the forward).
#include "common.cuh"
#include "philox.cuh"

namespace {

struct DropArgs {   // attention-probability dropout (CMIM:616); thresh == 0: off
  uint32_t thresh;
  float scale;
  uint64_t seed;
  const unsigned long long* base;   // device-resident seed base (icka_set_seed_base) or null
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Column sums: out[n] (+)= sum_m x[m, n].  Block = 32 x 8 threads owning 64 columns (2 per thread, so
// bf16 reads are 4 bytes); row slabs are spread over blockIdx.y and combined with fp32 atomics.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int64_t ld, float* __restrict__ out,
                                                     int M, int N, int rows_per_block) {
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + 2 * tx;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s0 = 0.0f, s1 = 0.0f;
  if (c < N) {
    for (int m = m0 + ty; m < m1; m += 8) {
      const T* p = x + (size_t)m * ld + c;
      if (c + 1 < N) {
        if constexpr (sizeof(T) == 2) {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
          s0 += __uint_as_float(u << 16);
          s1 += __uint_as_float(u & 0xffff0000u);
        } else {
          const float2 v = *reinterpret_cast<const float2*>(p);
          s0 += v.x;
          s1 += v.y;
        }
      } else {
        s0 += to_f32<T>(p[0]);
      }
    }
  }
  red[ty][2 * tx] = s0;
  red[ty][2 * tx + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int col = blockIdx.x * 64 + threadIdx.x;
    if (col < N) atomicAdd(out + col, s);
  }
}

// Wide form for 16-byte-aligned rows: a thread owns 16 bytes of columns (8 bf16 / 4 fp32), a warp 512 contiguous bytes of a
// row, and four rows are in flight per thread -- the 4-byte form above is latency-bound (54 us for the 100 MB intermediate
// gradient of a B=128 training step; this one runs at HBM speed).
template <typename T>
__global__ void __launch_bounds__(256) colsum_wide_kernel(const T* __restrict__ x, int64_t ld, float* __restrict__ out,
                                                          int M, int N, int rows_per_block) {
  constexpr int VE = 16 / (int)sizeof(T);
  __shared__ float red[8][32 * VE + 4];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * VE;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc[VE];
#pragma unroll
  for (int j = 0; j < VE; ++j) acc[j] = 0.0f;
  auto add = [&](const uint4& u) {
    if constexpr (sizeof(T) == 2) {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += __uint_as_float(w[j] << 16);
        acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
      }
    } else {
      acc[0] += __uint_as_float(u.x); acc[1] += __uint_as_float(u.y);
      acc[2] += __uint_as_float(u.z); acc[3] += __uint_as_float(u.w);
    }
  };
  if (c < N) {
    const T* base = x + c;
    int m = m0 + ty;
    for (; m + 24 < m1; m += 32) {
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)m * ld));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(m + 8) * ld));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(m + 16) * ld));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(m + 24) * ld));
      add(u0); add(u1); add(u2); add(u3);
    }
    for (; m < m1; m += 8) add(__ldg(reinterpret_cast<const uint4*>(base + (size_t)m * ld)));
  }
#pragma unroll
  for (int j = 0; j < VE; ++j) red[ty][tx * VE + j] = acc[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VE; i += 256) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][i];
    const int col = blockIdx.x * 32 * VE + i;
    if (col < N) atomicAdd(out + col, t);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  y = w * xhat + b with xhat = (x - mean) * rstd:
//   g = dy * w;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dw += dy * xhat;  db += dy
// One warp per row (the row lives in registers), each block walks a strided set of rows keeping its
// column partial sums of dw / db / dx in registers, then folds them across warps and adds them to the
// global vectors.  dx is written as fp32 (the residual-stream gradient) and/or bf16 (the next GEMM's
// operand); colsum(dx) is the bias gradient of the dense layer that produced x.
// ------------------------------------------------------------------------------------------------
constexpr int kLnVec = 8;   // N <= 1024

__device__ __forceinline__ void ln_fold_columns(const float4 (&acc)[kLnVec], float* __restrict__ out,
                                                float4 (*red)[32 * kLnVec / 2], int warp, int lane, int nvec) {
  if (out == nullptr) return;   // block-uniform
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kLnVec / 2; ++i) red[warp][lane + 32 * i] = acc[half * (kLnVec / 2) + i];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * (kLnVec / 2); e += 256) {
      const int c = half * 32 * (kLnVec / 2) + e;   // float4 column index
      if (c < nvec) {
        float4 s = red[0][e];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
          const float4 t = red[w][e];
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        float* o = out + 4 * c;
        atomicAdd(o + 0, s.x);
        atomicAdd(o + 1, s.y);
        atomicAdd(o + 2, s.z);
        atomicAdd(o + 3, s.w);
      }
    }
  }
}

__global__ void __launch_bounds__(256) layernorm_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma, float eps,
    float* __restrict__ dx32, __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ dbias, int M, int N) {
  __shared__ float4 red[8][32 * kLnVec / 2];   // one quantity at a time, half the vectors per pass
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nvec = N / 4;
  float4 ag[kLnVec], ab[kLnVec], ax[kLnVec];
#pragma unroll
  for (int i = 0; i < kLnVec; ++i) ag[i] = ab[i] = ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gm[kLnVec];
#pragma unroll
  for (int i = 0; i < kLnVec; ++i) {
    const int c = lane + 32 * i;
    gm[i] = (c < nvec) ? reinterpret_cast<const float4*>(gamma)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invN = 1.0f / (float)N;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * N);
    const float4* dr = reinterpret_cast<const float4*>(dy + (size_t)row * N);
    float4 v[kLnVec], d[kLnVec];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        v[i] = xr[c];
        d[i] = dr[c];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      } else {
        v[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = warp_sum(sum) * invN;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * invN + eps);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      // v <- xhat, accumulate dw / db, d <- g = dy * w
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;
      ag[i].x += d[i].x * v[i].x; ag[i].y += d[i].y * v[i].y; ag[i].z += d[i].z * v[i].z; ag[i].w += d[i].w * v[i].w;
      ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
      d[i].x *= gm[i].x; d[i].y *= gm[i].y; d[i].z *= gm[i].z; d[i].w *= gm[i].w;
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
    const float c1 = warp_sum(s1) * invN, c2 = warp_sum(s2) * invN;
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float4 o;
        o.x = rstd * (d[i].x - c1 - v[i].x * c2);
        o.y = rstd * (d[i].y - c1 - v[i].y * c2);
        o.z = rstd * (d[i].z - c1 - v[i].z * c2);
        o.w = rstd * (d[i].w - c1 - v[i].w * c2);
        ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
        if (dx32) reinterpret_cast<float4*>(dx32 + (size_t)row * N)[c] = o;
        if (dx16) {
          uint2 p;
          p.x = pack_bf16x2(o.x, o.y);
          p.y = pack_bf16x2(o.z, o.w);
          reinterpret_cast<uint2*>(dx16 + (size_t)row * N)[c] = p;
        }
      }
    }
  }
  // fold the per-warp column partials and add them to the global vectors
  ln_fold_columns(ag, dgamma, red, warp, lane, nvec);
  ln_fold_columns(ab, dbeta, red, warp, lane, nvec);
  ln_fold_columns(ax, dbias, red, warp, lane, nvec);
}

// ------------------------------------------------------------------------------------------------
// Attention-core backward.  One block per (head, sentence); probabilities are recomputed from Q, K, V:
//   S = Q K^T / 8 + mask;  P = softmax(S);  dV = P^T dO;  dP = dO V^T;  dS = P * (dP - rowsum(P * dP));
//   dQ = dS K / 8;  dK = dS^T Q / 8.
// Phase 1: one thread per query row of the current query tile (scores, softmax, dS rows in shared
// memory, dQ row written out).  Phase 2: the block turns P / dS (TQ x Skv) and dO / Q (TQ x 64) into
// dV / dK contributions, each thread owning a fixed set of (key, 4 dims) outputs accumulated in shared
// memory across query tiles (no atomics: a block owns its head's dK / dV completely).
// fp32 arithmetic throughout; T is the storage type of q, k, v, d(ctx), dq, dk, dv.
// ------------------------------------------------------------------------------------------------
constexpr int kD = 64;
constexpr int kAttnBwdThreads = 128;

template <typename T>
__device__ __forceinline__ void load_row64(const T* p, float* out) {
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int c = 0; c < kD; c += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(p + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        out[c + 2 * i] = __uint_as_float(w[i] << 16);
        out[c + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < kD; c += 4) {
      const float4 a = *reinterpret_cast<const float4*>(p + c);
      out[c] = a.x; out[c + 1] = a.y; out[c + 2] = a.z; out[c + 3] = a.w;
    }
  }
}

template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d) {
  if constexpr (sizeof(T) == 2) {
    uint2 u;
    u.x = pack_bf16x2(a, b);
    u.y = pack_bf16x2(c, d);
    *reinterpret_cast<uint2*>(p) = u;
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  }
}

template <typename T>
__global__ void __launch_bounds__(kAttnBwdThreads) cross_attn_bwd_kernel(
    const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, const T* __restrict__ v, int64_t ldkv,
    const float* __restrict__ mask_add, const T* __restrict__ dctx, int64_t ldc, T* __restrict__ dq, int64_t lddq,
    T* __restrict__ dk, T* __restrict__ dv, int64_t lddkv, int Sq, int Skv, int TQ, DropArgs drop) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                              // [Skv][64]
  float* Vs = Ks + (size_t)Skv * kD;           // [Skv][64]
  float* dKs = Vs + (size_t)Skv * kD;          // [Skv][64]
  float* dVs = dKs + (size_t)Skv * kD;         // [Skv][64]
  float* Qs = dVs + (size_t)Skv * kD;          // [TQ][64]
  float* dOs = Qs + (size_t)TQ * kD;           // [TQ][64]
  float* Ps = dOs + (size_t)TQ * kD;           // [TQ][Skv + 1]
  float* dSs = Ps + (size_t)TQ * (Skv + 1);    // [TQ][Skv + 1]   (already scaled by 1/8)
  float* Ms = dSs + (size_t)TQ * (Skv + 1);    // [Skv]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const int pitch = Skv + 1;

  const T* kb = k + (size_t)b * Skv * ldkv + (size_t)h * kD;
  const T* vb = v + (size_t)b * Skv * ldkv + (size_t)h * kD;
  for (int i = tid; i < Skv * (kD / 8); i += kAttnBwdThreads) {
    const int r = i / (kD / 8), c = (i % (kD / 8)) * 8;
    if constexpr (sizeof(T) == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(kb + (size_t)r * ldkv + c);
      const uint4 w = *reinterpret_cast<const uint4*>(vb + (size_t)r * ldkv + c);
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Ks[r * kD + c + 2 * j] = __uint_as_float(uu[j] << 16);
        Ks[r * kD + c + 2 * j + 1] = __uint_as_float(uu[j] & 0xffff0000u);
        Vs[r * kD + c + 2 * j] = __uint_as_float(ww[j] << 16);
        Vs[r * kD + c + 2 * j + 1] = __uint_as_float(ww[j] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        Ks[r * kD + c + j] = static_cast<float>(kb[(size_t)r * ldkv + c + j]);
        Vs[r * kD + c + j] = static_cast<float>(vb[(size_t)r * ldkv + c + j]);
      }
    }
  }
  for (int i = tid; i < Skv * kD; i += kAttnBwdThreads) dKs[i] = dVs[i] = 0.0f;
  for (int i = tid; i < Skv; i += kAttnBwdThreads) Ms[i] = mask_add ? mask_add[(size_t)b * Skv + i] : 0.0f;
  __syncthreads();

  for (int q0 = 0; q0 < Sq; q0 += TQ) {
    const int nq = min(TQ, Sq - q0);
    // ---- phase 1: thread = query row ----
    if (tid < nq) {
      const int row = q0 + tid;
      float qr[kD], dor[kD];
      load_row64<T>(q + ((size_t)b * Sq + row) * ldq + (size_t)h * kD, qr);
      load_row64<T>(dctx + ((size_t)b * Sq + row) * ldc + (size_t)h * kD, dor);
#pragma unroll
      for (int c = 0; c < kD; ++c) {
        Qs[tid * kD + c] = qr[c];
        dOs[tid * kD + c] = dor[c];
      }
      float* prow = Ps + tid * pitch;
      float* dsrow = dSs + tid * pitch;
      float mx = -INFINITY;
      for (int j = 0; j < Skv; ++j) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + j * kD);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < kD / 4; ++c) {
          const float4 kk = kr[c];
          s0 = fmaf(qr[4 * c], kk.x, s0);
          s1 = fmaf(qr[4 * c + 1], kk.y, s1);
          s2 = fmaf(qr[4 * c + 2], kk.z, s2);
          s3 = fmaf(qr[4 * c + 3], kk.w, s3);
        }
        const float s = ((s0 + s1) + (s2 + s3)) * 0.125f + Ms[j];
        prow[j] = s;
        mx = fmaxf(mx, s);
      }
      float l = 0.0f;
      for (int j = 0; j < Skv; ++j) {
        const float p = expf(prow[j] - mx);
        prow[j] = p;
        l += p;
      }
      const float inv = 1.0f / l;
      float delta = 0.0f;
      const uint64_t drow = ((uint64_t)b * gridDim.x + h) * (uint64_t)Sq + (uint64_t)row;
      uint32_t keep = 0xfu;
      for (int j = 0; j < Skv; ++j) {
        if (drop.thresh && (j & 3) == 0)
          keep = icka_rng::keep_bits4(icka_rng::effective_seed(drop.seed, drop.base), icka_rng::kSiteAttention, icka_rng::attn_group(drow, Skv, j), drop.thresh);
        const float4* vr = reinterpret_cast<const float4*>(Vs + j * kD);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int c = 0; c < kD / 4; ++c) {
          const float4 vv = vr[c];
          s0 = fmaf(dor[4 * c], vv.x, s0);
          s1 = fmaf(dor[4 * c + 1], vv.y, s1);
          s2 = fmaf(dor[4 * c + 2], vv.z, s2);
          s3 = fmaf(dor[4 * c + 3], vv.w, s3);
        }
        // with dropout the forward used P' = P keep/(1-p): dP = dP' keep/(1-p), and delta = sum P' dP' = sum P dP
        const float ks = drop.thresh ? ((keep >> (j & 3) & 1u) ? drop.scale : 0.0f) : 1.0f;
        const float dp = ((s0 + s1) + (s2 + s3)) * ks;
        const float p = prow[j] * inv;
        prow[j] = p;
        dsrow[j] = dp;
        delta = fmaf(p, dp, delta);
      }
      float dqr[kD];
#pragma unroll
      for (int c = 0; c < kD; ++c) dqr[c] = 0.0f;
      for (int j = 0; j < Skv; ++j) {
        const float ds = prow[j] * (dsrow[j] - delta) * 0.125f;
        dsrow[j] = ds;
        if (drop.thresh) {   // phase 2 forms dV from the DROPPED probabilities
          if ((j & 3) == 0)
            keep = icka_rng::keep_bits4(icka_rng::effective_seed(drop.seed, drop.base), icka_rng::kSiteAttention, icka_rng::attn_group(drow, Skv, j), drop.thresh);
          prow[j] = (keep >> (j & 3) & 1u) ? prow[j] * drop.scale : 0.0f;
        }
        const float4* kr = reinterpret_cast<const float4*>(Ks + j * kD);
#pragma unroll
        for (int c = 0; c < kD / 4; ++c) {
          const float4 kk = kr[c];
          dqr[4 * c] = fmaf(ds, kk.x, dqr[4 * c]);
          dqr[4 * c + 1] = fmaf(ds, kk.y, dqr[4 * c + 1]);
          dqr[4 * c + 2] = fmaf(ds, kk.z, dqr[4 * c + 2]);
          dqr[4 * c + 3] = fmaf(ds, kk.w, dqr[4 * c + 3]);
        }
      }
      T* dqp = dq + ((size_t)b * Sq + row) * lddq + (size_t)h * kD;
#pragma unroll
      for (int c = 0; c < kD; c += 4) store4<T>(dqp + c, dqr[c], dqr[c + 1], dqr[c + 2], dqr[c + 3]);
    }
    __syncthreads();
    // ---- phase 2: thread = (key group, 4 dims); dV += P^T dO, dK += dS^T Q ----
    {
      const int dq4 = (tid & 15) * 4;   // dims dq4 .. dq4+3
      for (int j = tid >> 4; j < Skv; j += kAttnBwdThreads / 16) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), ak = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < nq; ++i) {
          const float p = Ps[i * pitch + j], ds = dSs[i * pitch + j];
          const float4 o = *reinterpret_cast<const float4*>(dOs + i * kD + dq4);
          const float4 qq = *reinterpret_cast<const float4*>(Qs + i * kD + dq4);
          av.x = fmaf(p, o.x, av.x); av.y = fmaf(p, o.y, av.y); av.z = fmaf(p, o.z, av.z); av.w = fmaf(p, o.w, av.w);
          ak.x = fmaf(ds, qq.x, ak.x); ak.y = fmaf(ds, qq.y, ak.y); ak.z = fmaf(ds, qq.z, ak.z); ak.w = fmaf(ds, qq.w, ak.w);
        }
        float4* pv = reinterpret_cast<float4*>(dVs + j * kD + dq4);
        float4* pk = reinterpret_cast<float4*>(dKs + j * kD + dq4);
        float4 cv = *pv, ck = *pk;
        cv.x += av.x; cv.y += av.y; cv.z += av.z; cv.w += av.w;
        ck.x += ak.x; ck.y += ak.y; ck.z += ak.z; ck.w += ak.w;
        *pv = cv;
        *pk = ck;
      }
    }
    __syncthreads();
  }
  // ---- write dK, dV ----
  T* dkb = dk + (size_t)b * Skv * lddkv + (size_t)h * kD;
  T* dvb = dv + (size_t)b * Skv * lddkv + (size_t)h * kD;
  for (int i = tid; i < Skv * (kD / 4); i += kAttnBwdThreads) {
    const int r = i / (kD / 4), c = (i % (kD / 4)) * 4;
    const float4 a = *reinterpret_cast<const float4*>(dKs + r * kD + c);
    const float4 g = *reinterpret_cast<const float4*>(dVs + r * kD + c);
    store4<T>(dkb + (size_t)r * lddkv + c, a.x, a.y, a.z, a.w);
    store4<T>(dvb + (size_t)r * lddkv + c, g.x, g.y, g.z, g.w);
  }
}

// ------------------------------------------------------------------------------------------------
// bf16 attention backward on the tensor cores (mma.sync m16n8k16, fp32 accumulate) for Sq <= 128: the same
// tiling as the forward kernel (attention.cu): one block = one (sentence, head), 8 warps x 16 query rows,
// keys in blocks of 64.  The softmax statistics are rebuilt in a first sweep over the key blocks; the
// second sweep forms P and dS = P * (dP - delta) / 8 per key block with delta = rowsum(dO * O) taken from
// the saved forward output, accumulates dQ += dS K in registers, parks P and dS (bf16) in shared memory,
// and then the 8 warps split dV = P^T dO and dK = dS^T Q of that key block (operands fetched with
// ldmatrix.trans straight from the [row][key] / [row][dim] tiles).
// ------------------------------------------------------------------------------------------------
constexpr int kMmaThreads = 256;
constexpr int kPitch = 72;        // bf16 elements per smem row (64 + 8 pad): conflict-free ldmatrix
constexpr int kKeyBlk = 64;
constexpr int kRowsQ = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t sa = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_fence_wait() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
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

// grid = (nh, B); Sq <= 128
__global__ void __launch_bounds__(kMmaThreads) cross_attn_bwd_mma_kernel(
    const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
    const __nv_bfloat16* __restrict__ v, int64_t ldkv, const float* __restrict__ mask_add,
    const __nv_bfloat16* __restrict__ o, int64_t ldo, const __nv_bfloat16* __restrict__ dctx, int64_t ldc,
    __nv_bfloat16* __restrict__ dq, int64_t lddq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
    int64_t lddkv, int Sq, int Skv, DropArgs drop) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_dyn);   // [128][72]
  __nv_bfloat16* dOs = Qs + kRowsQ * kPitch;                          // [128][72]
  __nv_bfloat16* Ps = dOs + kRowsQ * kPitch;                          // [128][72]  P   of the current key block
  __nv_bfloat16* dSs = Ps + kRowsQ * kPitch;                          // [128][72]  dS  of the current key block
  __nv_bfloat16* Ks = dSs + kRowsQ * kPitch;                          // [64][72]
  __nv_bfloat16* Vs = Ks + kKeyBlk * kPitch;                          // [64][72]
  float* Ms = reinterpret_cast<float*>(Vs + kKeyBlk * kPitch);        // [64]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int g = lane >> 2, t = lane & 3;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float kScale = 0.125f * kLog2e;

  // ---- stage Q and dO (rows beyond Sq are zero-filled) ----
  const __nv_bfloat16* qb = q + (size_t)b * Sq * ldq + (size_t)h * kD;
  const __nv_bfloat16* dob = dctx + (size_t)b * Sq * ldc + (size_t)h * kD;
  for (int i = tid; i < kRowsQ * 8; i += kMmaThreads) {
    const int r = i >> 3, c = (i & 7) * 8;
    if (r < Sq) {
      cp_async16(Qs + r * kPitch + c, qb + (size_t)r * ldq + c);
      cp_async16(dOs + r * kPitch + c, dob + (size_t)r * ldc + c);
    } else {
      *reinterpret_cast<uint4*>(Qs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dOs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_fence_wait();
  __syncthreads();

  // ---- delta[row] = sum_d dO[row][d] * O[row][d]; two lanes per row, then fetched for rows g and g+8 ----
  float delta0, delta1;
  {
    const int r = warp * 16 + (lane >> 1), c0 = (lane & 1) * 32;
    float acc = 0.0f;
    if (r < Sq) {
      const __nv_bfloat16* op = o + ((size_t)b * Sq + r) * ldo + (size_t)h * kD + c0;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        const uint4 uo = *reinterpret_cast<const uint4*>(op + c);
        const uint4 ud = *reinterpret_cast<const uint4*>(dOs + r * kPitch + c0 + c);
        const uint32_t wo[4] = {uo.x, uo.y, uo.z, uo.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc = fmaf(__uint_as_float(wo[j] << 16), __uint_as_float(wd[j] << 16), acc);
          acc = fmaf(__uint_as_float(wo[j] & 0xffff0000u), __uint_as_float(wd[j] & 0xffff0000u), acc);
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    delta0 = __shfl_sync(0xffffffffu, acc, 2 * g);
    delta1 = __shfl_sync(0xffffffffu, acc, 2 * (g + 8));
  }

  uint32_t qf[4][4], dof[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    ldmatrix_x4(qf[kk], Qs + (warp * 16 + (lane & 15)) * kPitch + kk * 16 + (lane >> 4) * 8);
    ldmatrix_x4(dof[kk], dOs + (warp * 16 + (lane & 15)) * kPitch + kk * 16 + (lane >> 4) * 8);
  }

  const __nv_bfloat16* kb = k + (size_t)b * Skv * ldkv + (size_t)h * kD;
  const __nv_bfloat16* vb = v + (size_t)b * Skv * ldkv + (size_t)h * kD;

  auto load_keys = [&](int key0, bool with_v) {
    for (int i = tid; i < kKeyBlk * 8; i += kMmaThreads) {
      const int r = i >> 3, c = (i & 7) * 8;
      if (key0 + r < Skv) {
        cp_async16(Ks + r * kPitch + c, kb + (size_t)(key0 + r) * ldkv + c);
        if (with_v) cp_async16(Vs + r * kPitch + c, vb + (size_t)(key0 + r) * ldkv + c);
      } else {
        *reinterpret_cast<uint4*>(Ks + r * kPitch + c) = make_uint4(0, 0, 0, 0);
        if (with_v) *reinterpret_cast<uint4*>(Vs + r * kPitch + c) = make_uint4(0, 0, 0, 0);
      }
    }
    if (tid < kKeyBlk) {
      const int key = key0 + tid;
      Ms[tid] = (key < Skv) ? (mask_add ? mask_add[(size_t)b * Skv + key] * kLog2e : 0.0f) : -INFINITY;
    }
    cp_async_fence_wait();
    __syncthreads();
  };
  // scores of this warp's 16 rows against the staged key block, log2 domain (scale and mask applied)
  auto scores = [&](float (&sacc)[8][4]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.0f;
      uint32_t kf[4];
      ldmatrix_x4(kf, Ks + (j * 8 + (lane & 7)) * kPitch + (lane >> 3) * 8);
      mma_bf16_16816(sacc[j], qf[0], kf[0], kf[1]);
      mma_bf16_16816(sacc[j], qf[1], kf[2], kf[3]);
      ldmatrix_x4(kf, Ks + (j * 8 + (lane & 7)) * kPitch + 32 + (lane >> 3) * 8);
      mma_bf16_16816(sacc[j], qf[2], kf[0], kf[1]);
      mma_bf16_16816(sacc[j], qf[3], kf[2], kf[3]);
      const float mk0 = Ms[j * 8 + 2 * t], mk1 = Ms[j * 8 + 2 * t + 1];
      sacc[j][0] = fmaf(sacc[j][0], kScale, mk0);
      sacc[j][1] = fmaf(sacc[j][1], kScale, mk1);
      sacc[j][2] = fmaf(sacc[j][2], kScale, mk0);
      sacc[j][3] = fmaf(sacc[j][3], kScale, mk1);
    }
  };

  // ---- sweep 1: softmax statistics (running max m, sum l) over all key blocks ----
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
  for (int key0 = 0; key0 < Skv; key0 += kKeyBlk) {
    if (key0 > 0) __syncthreads();
    load_keys(key0, false);
    float sacc[8][4];
    scores(sacc);
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bm0 = fmaxf(bm0, fmaxf(sacc[j][0], sacc[j][1]));
      bm1 = fmaxf(bm1, fmaxf(sacc[j][2], sacc[j][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    float ps0 = 0.0f, ps1 = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ps0 += fast_exp2(sacc[j][0] - mn0) + fast_exp2(sacc[j][1] - mn0);
      ps1 += fast_exp2(sacc[j][2] - mn1) + fast_exp2(sacc[j][3] - mn1);
    }
    l0 = l0 * fast_exp2(m0 - mn0) + ps0;
    l1 = l1 * fast_exp2(m1 - mn1) + ps1;
    m0 = mn0;
    m1 = mn1;
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float il0 = 1.0f / l0, il1 = 1.0f / l1;

  // ---- sweep 2: P, dS, dQ per key block, then dK / dV of that block ----
  float dqa[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) dqa[j][0] = dqa[j][1] = dqa[j][2] = dqa[j][3] = 0.0f;
  for (int key0 = 0; key0 < Skv; key0 += kKeyBlk) {
    __syncthreads();   // previous users of Ks / Vs / Ps / dSs are done
    load_keys(key0, true);
    float sacc[8][4];
    scores(sacc);
    uint32_t dsf[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float dp[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // dP = dO V^T for this 8-key tile
      uint32_t vf[4];
      ldmatrix_x4(vf, Vs + (j * 8 + (lane & 7)) * kPitch + (lane >> 3) * 8);
      mma_bf16_16816(dp, dof[0], vf[0], vf[1]);
      mma_bf16_16816(dp, dof[1], vf[2], vf[3]);
      ldmatrix_x4(vf, Vs + (j * 8 + (lane & 7)) * kPitch + 32 + (lane >> 3) * 8);
      mma_bf16_16816(dp, dof[2], vf[0], vf[1]);
      mma_bf16_16816(dp, dof[3], vf[2], vf[3]);
      float p0 = fast_exp2(sacc[j][0] - m0) * il0, p1 = fast_exp2(sacc[j][1] - m0) * il0;
      float p2 = fast_exp2(sacc[j][2] - m1) * il1, p3 = fast_exp2(sacc[j][3] - m1) * il1;
      float k0 = 1.0f, k1 = 1.0f, k2 = 1.0f, k3 = 1.0f;   // keep / (1 - p) of this thread's four probabilities
      if (drop.thresh) {
        const int key = key0 + j * 8 + 2 * t;
        const uint64_t drow = ((uint64_t)b * gridDim.x + h) * (uint64_t)Sq + (uint64_t)(warp * 16 + g);
        const uint32_t b0 = icka_rng::keep_bits4(icka_rng::effective_seed(drop.seed, drop.base), icka_rng::kSiteAttention, icka_rng::attn_group(drow, Skv, key), drop.thresh) >> (key & 3);
        const uint32_t b1 = icka_rng::keep_bits4(icka_rng::effective_seed(drop.seed, drop.base), icka_rng::kSiteAttention, icka_rng::attn_group(drow + 8, Skv, key), drop.thresh) >> (key & 3);
        k0 = (b0 & 1u) ? drop.scale : 0.0f;
        k1 = (b0 & 2u) ? drop.scale : 0.0f;
        k2 = (b1 & 1u) ? drop.scale : 0.0f;
        k3 = (b1 & 2u) ? drop.scale : 0.0f;
      }
      // dP = dP' keep/(1-p) (dp[] is the gradient w.r.t. the dropped probabilities); dV below uses P' = P keep/(1-p)
      const float d0 = p0 * (dp[0] * k0 - delta0) * 0.125f, d1 = p1 * (dp[1] * k1 - delta0) * 0.125f;
      const float d2 = p2 * (dp[2] * k2 - delta1) * 0.125f, d3 = p3 * (dp[3] * k3 - delta1) * 0.125f;
      p0 *= k0; p1 *= k1; p2 *= k2; p3 *= k3;
      dsf[j][0] = pack_bf16x2(d0, d1);
      dsf[j][1] = pack_bf16x2(d2, d3);
      const int r0 = warp * 16 + g, col = j * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(Ps + r0 * kPitch + col) = pack_bf16x2(p0, p1);
      *reinterpret_cast<uint32_t*>(Ps + (r0 + 8) * kPitch + col) = pack_bf16x2(p2, p3);
      *reinterpret_cast<uint32_t*>(dSs + r0 * kPitch + col) = dsf[j][0];
      *reinterpret_cast<uint32_t*>(dSs + (r0 + 8) * kPitch + col) = dsf[j][1];
    }
    // dQ += dS K   (A = dS fragments from registers, B = K via ldmatrix.trans, as P.V in the forward)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a[4] = {dsf[2 * ks][0], dsf[2 * ks][1], dsf[2 * ks + 1][0], dsf[2 * ks + 1][1]};
#pragma unroll
      for (int jn = 0; jn < 8; jn += 2) {
        uint32_t kf[4];
        ldmatrix_x4_trans(kf, Ks + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + jn * 8 + (lane >> 4) * 8);
        mma_bf16_16816(dqa[jn], a, kf[0], kf[1]);
        mma_bf16_16816(dqa[jn + 1], a, kf[2], kf[3]);
      }
    }
    __syncthreads();   // P and dS of every warp are in shared memory
    // dV = P^T dO, dK = dS^T Q for this key block: warp -> 16 keys (mt) x 32 dims (nh2)
    {
      const int mt = warp & 3, nh2 = warp >> 2;
      float dka[4][4], dva[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dka[j][0] = dka[j][1] = dka[j][2] = dka[j][3] = 0.0f;
        dva[j][0] = dva[j][1] = dva[j][2] = dva[j][3] = 0.0f;
      }
      const int mat = lane >> 3;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {   // 16 query rows per step
        // A[m = key][k = row] = X[row][key]: transposed 8x8 blocks of the [row][key] tiles
        const int arow = ks * 16 + (lane & 7) + ((mat >> 1) & 1) * 8;
        const int acol = mt * 16 + (mat & 1) * 8;
        uint32_t ap[4], ads[4];
        ldmatrix_x4_trans(ap, Ps + arow * kPitch + acol);
        ldmatrix_x4_trans(ads, dSs + arow * kPitch + acol);
#pragma unroll
        for (int jn = 0; jn < 4; jn += 2) {
          const int brow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int bcol = nh2 * 32 + jn * 8 + (lane >> 4) * 8;
          uint32_t bo[4], bq[4];
          ldmatrix_x4_trans(bo, dOs + brow * kPitch + bcol);
          ldmatrix_x4_trans(bq, Qs + brow * kPitch + bcol);
          mma_bf16_16816(dva[jn], ap, bo[0], bo[1]);
          mma_bf16_16816(dva[jn + 1], ap, bo[2], bo[3]);
          mma_bf16_16816(dka[jn], ads, bq[0], bq[1]);
          mma_bf16_16816(dka[jn + 1], ads, bq[2], bq[3]);
        }
      }
      const int key_a = key0 + mt * 16 + g, key_b = key_a + 8;
      __nv_bfloat16* dkp = dk + (size_t)b * Skv * lddkv + (size_t)h * kD + nh2 * 32 + 2 * t;
      __nv_bfloat16* dvp = dv + (size_t)b * Skv * lddkv + (size_t)h * kD + nh2 * 32 + 2 * t;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (key_a < Skv) {
          *reinterpret_cast<uint32_t*>(dkp + (size_t)key_a * lddkv + j * 8) = pack_bf16x2(dka[j][0], dka[j][1]);
          *reinterpret_cast<uint32_t*>(dvp + (size_t)key_a * lddkv + j * 8) = pack_bf16x2(dva[j][0], dva[j][1]);
        }
        if (key_b < Skv) {
          *reinterpret_cast<uint32_t*>(dkp + (size_t)key_b * lddkv + j * 8) = pack_bf16x2(dka[j][2], dka[j][3]);
          *reinterpret_cast<uint32_t*>(dvp + (size_t)key_b * lddkv + j * 8) = pack_bf16x2(dva[j][2], dva[j][3]);
        }
      }
    }
  }

  // ---- dQ: stage through the P tile (each warp owns its 16 rows), coalesced write-out ----
  __syncthreads();
  {
    __nv_bfloat16* ow = Ps + (warp * 16) * kPitch;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(ow + g * kPitch + j * 8 + 2 * t) = pack_bf16x2(dqa[j][0], dqa[j][1]);
      *reinterpret_cast<uint32_t*>(ow + (g + 8) * kPitch + j * 8 + 2 * t) = pack_bf16x2(dqa[j][2], dqa[j][3]);
    }
    __syncwarp();
    __nv_bfloat16* ob = dq + ((size_t)b * Sq + warp * 16) * lddq + (size_t)h * kD;
#pragma unroll
    for (int i = lane; i < 16 * 8; i += 32) {
      const int r = i >> 3, c = (i & 7) * 8;
      if (warp * 16 + r < Sq)
        *reinterpret_cast<uint4*>(ob + (size_t)r * lddq + c) = *reinterpret_cast<const uint4*>(ow + r * kPitch + c);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gate + blend backward (forward: elementwise.cu gate_blend_kernel).  One block per sentence.
//   out = g tok + (1-g) fused,  g = sigmoid(w_fold . n + c_fold),  n = LN(fused[0] + tok[0]) * ln_w + ln_b
// ------------------------------------------------------------------------------------------------
constexpr int kGateThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < kGateThreads / 32) ? red[l] : 0.0f;
  return warp_sum(t);
}

__global__ void __launch_bounds__(kGateThreads) gate_blend_bwd_kernel(
    const float* __restrict__ dout, const float* __restrict__ fused, const float* __restrict__ tok,
    const float* __restrict__ gate, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float ln_eps,
    const float* __restrict__ w_fold, float* __restrict__ dfused, float* __restrict__ dtok, float* __restrict__ d_ln_w,
    float* __restrict__ d_ln_b, float* __restrict__ d_w_fold, float* __restrict__ d_c_fold, int S, int H) {
  __shared__ float red[kGateThreads / 32];
  const int b = blockIdx.x;
  const float* f = fused + (size_t)b * S * H;
  const float* t = tok + (size_t)b * S * H;
  const float* d = dout + (size_t)b * S * H;
  float* df = dfused + (size_t)b * S * H;
  float* dt = dtok ? dtok + (size_t)b * S * H : nullptr;
  const float g = gate[b], og = 1.0f - g;

  // d g = sum dout * (tok - fused)
  const int nvec = S * H / 4;
  const float4* f4 = reinterpret_cast<const float4*>(f);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  const float4* d4 = reinterpret_cast<const float4*>(d);
  float acc = 0.0f;
  for (int v = threadIdx.x; v < nvec; v += kGateThreads) {
    const float4 a = t4[v], c = f4[v], e = d4[v];
    acc += (e.x * (a.x - c.x) + e.y * (a.y - c.y)) + (e.z * (a.z - c.z) + e.w * (a.w - c.w));
  }
  const float dg = block_sum(acc, red);
  const float dlogit = dg * g * og;
  if (threadIdx.x == 0) atomicAdd(d_c_fold, dlogit);

  // LayerNorm statistics of the [CLS] rows
  float sum = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) sum += f[k] + t[k];
  const float mean = block_sum(sum, red) / (float)H;
  float sq = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) {
    const float x = (f[k] + t[k]) - mean;
    sq += x * x;
  }
  const float rstd = 1.0f / sqrtf(block_sum(sq, red) / (float)H + ln_eps);
  float s1 = 0.0f, s2 = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) {
    const float xh = ((f[k] + t[k]) - mean) * rstd;
    const float dn = dlogit * w_fold[k];
    atomicAdd(d_w_fold + k, dlogit * (xh * ln_w[k] + ln_b[k]));
    atomicAdd(d_ln_w + k, dn * xh);
    atomicAdd(d_ln_b + k, dn);
    const float gx = dn * ln_w[k];
    s1 += gx;
    s2 += gx * xh;
  }
  const float c1 = block_sum(s1, red) / (float)H, c2 = block_sum(s2, red) / (float)H;

  // streaming part: d fused = (1-g) dout, d tok = g dout; row 0 also carries the gate path
  float4* df4 = reinterpret_cast<float4*>(df);
  float4* dt4 = reinterpret_cast<float4*>(dt);
  const int hvec = H / 4;
  for (int v = threadIdx.x; v < nvec; v += kGateThreads) {
    const float4 e = d4[v];
    float4 a = make_float4(og * e.x, og * e.y, og * e.z, og * e.w);
    float4 c = make_float4(g * e.x, g * e.y, g * e.z, g * e.w);
    if (v < hvec) {
      float add[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = 4 * v + j;
        const float xh = ((f[k] + t[k]) - mean) * rstd;
        const float gx = dlogit * w_fold[k] * ln_w[k];
        add[j] = rstd * (gx - c1 - xh * c2);
      }
      a.x += add[0]; a.y += add[1]; a.z += add[2]; a.w += add[3];
      c.x += add[0]; c.y += add[1]; c.z += add[2]; c.w += add[3];
    }
    df4[v] = a;
    if (dt) dt4[v] = c;
  }
}

// The same backward split in two, so that the streaming part runs on every SM: with one block per sentence a training batch
// of 32-128 sentences left most of the machine idle (99 us at B = 128).
//   stream kernel  block = (sentence, 16 rows):  dfused = (1-g) dout, dtok = g dout, partial dg -> atomicAdd(dg[b])
//   cls kernel     block = sentence: the gate path through the [CLS] row (LayerNorm backward, parameter gradients), added
//                  onto row 0 of dfused / dtok
constexpr int kGateRows = 16;
__global__ void __launch_bounds__(kGateThreads) gate_blend_bwd_stream_kernel(
    const float* __restrict__ dout, const float* __restrict__ fused, const float* __restrict__ tok,
    const float* __restrict__ gate, float* __restrict__ dfused, float* __restrict__ dtok, float* __restrict__ dg_out, int S,
    int H) {
  __shared__ float red[kGateThreads / 32];
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * kGateRows, r1 = min(S, r0 + kGateRows);
  const size_t base = ((size_t)b * S + r0) * H;
  const int nvec = (r1 - r0) * H / 4;
  const float4* f4 = reinterpret_cast<const float4*>(fused + base);
  const float4* t4 = reinterpret_cast<const float4*>(tok + base);
  const float4* d4 = reinterpret_cast<const float4*>(dout + base);
  float4* df4 = reinterpret_cast<float4*>(dfused + base);
  float4* dt4 = dtok ? reinterpret_cast<float4*>(dtok + base) : nullptr;
  const float g = gate[b], og = 1.0f - g;
  float acc = 0.0f;
  for (int v = threadIdx.x; v < nvec; v += kGateThreads) {
    const float4 a = __ldcs(t4 + v), c = __ldcs(f4 + v), e = __ldcs(d4 + v);
    acc += (e.x * (a.x - c.x) + e.y * (a.y - c.y)) + (e.z * (a.z - c.z) + e.w * (a.w - c.w));
    df4[v] = make_float4(og * e.x, og * e.y, og * e.z, og * e.w);
    if (dt4) dt4[v] = make_float4(g * e.x, g * e.y, g * e.z, g * e.w);
  }
  const float part = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(dg_out + b, part);
}

__global__ void __launch_bounds__(kGateThreads) gate_blend_bwd_cls_kernel(
    const float* __restrict__ dg_in, const float* __restrict__ fused, const float* __restrict__ tok,
    const float* __restrict__ gate, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float ln_eps,
    const float* __restrict__ w_fold, float* __restrict__ dfused, float* __restrict__ dtok, float* __restrict__ d_ln_w,
    float* __restrict__ d_ln_b, float* __restrict__ d_w_fold, float* __restrict__ d_c_fold, int S, int H) {
  __shared__ float red[kGateThreads / 32];
  const int b = blockIdx.x;
  const float* f = fused + (size_t)b * S * H;
  const float* t = tok + (size_t)b * S * H;
  float* df = dfused + (size_t)b * S * H;
  float* dt = dtok ? dtok + (size_t)b * S * H : nullptr;
  const float g = gate[b], og = 1.0f - g;
  const float dlogit = dg_in[b] * g * og;
  if (threadIdx.x == 0) atomicAdd(d_c_fold, dlogit);
  float sum = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) sum += f[k] + t[k];
  const float mean = block_sum(sum, red) / (float)H;
  float sq = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) {
    const float x = (f[k] + t[k]) - mean;
    sq += x * x;
  }
  const float rstd = 1.0f / sqrtf(block_sum(sq, red) / (float)H + ln_eps);
  float s1 = 0.0f, s2 = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) {
    const float xh = ((f[k] + t[k]) - mean) * rstd;
    const float dn = dlogit * w_fold[k];
    atomicAdd(d_w_fold + k, dlogit * (xh * ln_w[k] + ln_b[k]));
    atomicAdd(d_ln_w + k, dn * xh);
    atomicAdd(d_ln_b + k, dn);
    const float gx = dn * ln_w[k];
    s1 += gx;
    s2 += gx * xh;
  }
  const float c1 = block_sum(s1, red) / (float)H, c2 = block_sum(s2, red) / (float)H;
  for (int k = threadIdx.x; k < H; k += kGateThreads) {
    const float xh = ((f[k] + t[k]) - mean) * rstd;
    const float gx = dlogit * w_fold[k] * ln_w[k];
    const float add = rstd * (gx - c1 - xh * c2);
    df[k] += add;
    if (dt) dt[k] += add;
  }
}

}  // namespace

extern "C" int icka_colsum(icka_handle* h, const void* x, int64_t ld, int dtype, float* out, int M, int N,
                           int accumulate, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(x && out && M >= 0 && N >= 1 && ld >= N, "colsum: bad arguments");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "colsum: bad dtype %d", dtype);
  ICKA_REQUIRE(ld % 2 == 0 && icka_aligned(x, 8), "colsum: pitch must be even and x 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!accumulate) ICKA_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st));
  if (M == 0) return ICKA_OK;
  const int ve = dtype == ICKA_BF16 ? 8 : 4;
  if (N % ve == 0 && ld % ve == 0 && icka_aligned(x, 16) && M >= 256) {
    const int cb = (N + 32 * ve - 1) / (32 * ve);
    int rb = (8 * h->sm_count + cb - 1) / cb;
    if (rb > (M + 31) / 32) rb = (M + 31) / 32;
    const int rows = (M + rb - 1) / rb;
    dim3 grid_w(cb, (M + rows - 1) / rows);
    if (dtype == ICKA_BF16)
      colsum_wide_kernel<__nv_bfloat16><<<grid_w, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ld, out, M, N, rows);
    else
      colsum_wide_kernel<float><<<grid_w, 256, 0, st>>>(static_cast<const float*>(x), ld, out, M, N, rows);
    ICKA_LAUNCHED(h);
    return ICKA_OK;
  }
  const int col_blocks = (N + 63) / 64;
  int row_blocks = (4 * h->sm_count + col_blocks - 1) / col_blocks;
  if (row_blocks > (M + 63) / 64) row_blocks = (M + 63) / 64;
  if (row_blocks < 1) row_blocks = 1;
  const int rpb = (M + row_blocks - 1) / row_blocks;
  dim3 grid(col_blocks, (M + rpb - 1) / rpb);
  if (dtype == ICKA_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ld, out, M, N, rpb);
  else
    colsum_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), ld, out, M, N, rpb);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_layernorm_bwd(icka_handle* h, const float* dy, const float* x, const float* gamma, float eps,
                                  float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dbias, int M, int N,
                                  void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(M >= 0 && N >= 4 && dy && x && gamma, "layernorm_bwd: bad arguments");
  ICKA_REQUIRE(N % 4 == 0 && N <= 128 * kLnVec, "layernorm_bwd: N=%d must be a multiple of 4 and <= %d", N, 128 * kLnVec);
  ICKA_REQUIRE(dx_f32 || dx_bf16, "layernorm_bwd: no dx output");
  ICKA_REQUIRE(icka_aligned(dy, 16) && icka_aligned(x, 16) && icka_aligned(gamma, 16) && icka_aligned(dx_f32, 16) &&
                   icka_aligned(dx_bf16, 8),
               "layernorm_bwd: pointers must be 16-byte aligned");
  if (M == 0) return ICKA_OK;
  int blocks = (M + 7) / 8;
  if (blocks > 2 * h->sm_count) blocks = 2 * h->sm_count;
  layernorm_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, x, gamma, eps, dx_f32, static_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dbias, M, N);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

int icka_attn_sq1_bwd_launch(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                             const float* mask_add, const void* dctx, int64_t lddc, void* dq, int64_t lddq, void* dk,
                             void* dv, int64_t lddkv, int dtype, int B, int Skv, int nh, uint32_t thresh, float scale,
                             uint64_t seed, const unsigned long long* base, cudaStream_t st);

extern "C" int icka_cross_attn_core_bwd(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                                        int64_t ldkv, const float* mask_add, const void* ctx, int64_t ldctx,
                                        const void* dctx, int64_t ldc, void* dq, int64_t lddq, void* dk, void* dv,
                                        int64_t lddkv, int dtype, int B, int Sq, int Skv, int nh, int d, void* stream) {
  return icka_cross_attn_core_bwd_drop(h, q, ldq, k, v, ldkv, mask_add, ctx, ldctx, dctx, ldc, dq, lddq, dk, dv, lddkv,
                                       dtype, B, Sq, Skv, nh, d, 0.0f, 0, stream);
}

extern "C" int icka_cross_attn_core_bwd_drop(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                                             int64_t ldkv, const float* mask_add, const void* ctx, int64_t ldctx,
                                             const void* dctx, int64_t ldc, void* dq, int64_t lddq, void* dk, void* dv,
                                             int64_t lddkv, int dtype, int B, int Sq, int Skv, int nh, int d,
                                             float p_drop, uint64_t seed, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "cross_attn_bwd: dropout p=%f outside [0, 1)", (double)p_drop);
  const DropArgs drop{p_drop > 0.0f ? icka_rng::keep_threshold(p_drop) : 0u, 1.0f / (1.0f - p_drop), seed, h->seed_base};
  ICKA_REQUIRE(q && k && v && dctx && dq && dk && dv, "cross_attn_bwd: null pointer");
  ICKA_REQUIRE(B >= 0 && Sq >= 1 && Skv >= 1 && nh >= 1, "cross_attn_bwd: bad shape");
  ICKA_REQUIRE(d == kD, "cross_attn_bwd: head dim %d != 64", d);
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "cross_attn_bwd: bad dtype %d", dtype);
  ICKA_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldc % 8 == 0 && lddq % 8 == 0 && lddkv % 8 == 0,
               "cross_attn_bwd: pitches must be multiples of 8 elements");
  ICKA_REQUIRE(icka_aligned(q, 16) && icka_aligned(k, 16) && icka_aligned(v, 16) && icka_aligned(dctx, 16) &&
                   icka_aligned(dq, 16) && icka_aligned(dk, 16) && icka_aligned(dv, 16),
               "cross_attn_bwd: pointers must be 16-byte aligned");
  ICKA_REQUIRE(B <= 65535, "cross_attn_bwd: B exceeds grid limits; shard the batch");
  if (B == 0) return ICKA_OK;
  if (Sq == 1) {   // image->text encoders: one query per sentence (attention_sq1.cu)
    const int rc = icka_attn_sq1_bwd_launch(h, q, ldq, k, v, ldkv, mask_add, dctx, ldc, dq, lddq, dk, dv, lddkv, dtype, B,
                                            Skv, nh, drop.thresh, drop.scale, drop.seed, drop.base,
                                            static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;
  }
  if (dtype == ICKA_BF16 && ctx != nullptr && Sq <= kRowsQ && ldctx % 8 == 0 && icka_aligned(ctx, 16)) {
    // tensor-core path
    const size_t smem_mma = (size_t)(4 * kRowsQ + 2 * kKeyBlk) * kPitch * sizeof(__nv_bfloat16) + kKeyBlk * sizeof(float);
    using T = __nv_bfloat16;
    ICKA_CUDA(cudaFuncSetAttribute(cross_attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
    cross_attn_bwd_mma_kernel<<<dim3(nh, B), kMmaThreads, smem_mma, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const T*>(q), ldq, static_cast<const T*>(k), static_cast<const T*>(v), ldkv, mask_add,
        static_cast<const T*>(ctx), ldctx, static_cast<const T*>(dctx), ldc, static_cast<T*>(dq), lddq,
        static_cast<T*>(dk), static_cast<T*>(dv), lddkv, Sq, Skv, drop);
    ICKA_LAUNCHED(h);
    return ICKA_OK;
  }
  // query tile: as many rows (<= 128 threads) as shared memory allows next to the four [Skv][64] key-side arrays
  const size_t fixed = ((size_t)4 * Skv * kD + Skv) * sizeof(float);
  const size_t per_row = ((size_t)2 * kD + 2 * (Skv + 1)) * sizeof(float);
  if (fixed + per_row > h->smem_optin)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "cross_attn_bwd: Skv=%d needs more than %zu B of shared memory", Skv, h->smem_optin);
  int TQ = (int)((h->smem_optin - fixed) / per_row);
  if (TQ > kAttnBwdThreads) TQ = kAttnBwdThreads;
  if (TQ > Sq) TQ = Sq;
  const size_t smem = fixed + (size_t)TQ * per_row;
  dim3 grid(nh, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ICKA_BF16) {
    using T = __nv_bfloat16;
    ICKA_CUDA(cudaFuncSetAttribute(cross_attn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cross_attn_bwd_kernel<T><<<grid, kAttnBwdThreads, smem, st>>>(
        static_cast<const T*>(q), ldq, static_cast<const T*>(k), static_cast<const T*>(v), ldkv, mask_add,
        static_cast<const T*>(dctx), ldc, static_cast<T*>(dq), lddq, static_cast<T*>(dk), static_cast<T*>(dv), lddkv, Sq,
        Skv, TQ, drop);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(cross_attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cross_attn_bwd_kernel<float><<<grid, kAttnBwdThreads, smem, st>>>(
        static_cast<const float*>(q), ldq, static_cast<const float*>(k), static_cast<const float*>(v), ldkv, mask_add,
        static_cast<const float*>(dctx), ldc, static_cast<float*>(dq), lddq, static_cast<float*>(dk),
        static_cast<float*>(dv), lddkv, Sq, Skv, TQ, drop);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_gate_blend_bwd(icka_handle* h, const float* dout, const float* fused, const float* tok,
                                   const float* gate, const float* ln_w, const float* ln_b, float ln_eps,
                                   const float* w_fold, float* dfused, float* dtok, float* d_ln_w, float* d_ln_b,
                                   float* d_w_fold, float* d_c_fold, int B, int S, int H, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && H >= 4 && dout && fused && tok && gate && ln_w && ln_b && w_fold && dfused &&
                   d_ln_w && d_ln_b && d_w_fold && d_c_fold,
               "gate_blend_bwd: bad arguments");
  ICKA_REQUIRE(H % 4 == 0, "gate_blend_bwd: H must be a multiple of 4");
  ICKA_REQUIRE(icka_aligned(dout, 16) && icka_aligned(fused, 16) && icka_aligned(tok, 16) && icka_aligned(dfused, 16) &&
                   icka_aligned(dtok, 16),
               "gate_blend_bwd: pointers must be 16-byte aligned");
  if (B == 0) return ICKA_OK;
  if (h->workspace != nullptr && (size_t)B * sizeof(float) <= ICKA_WORKSPACE_BYTES && B <= 65535) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* dg = static_cast<float*>(h->workspace);       // [B] per-sentence gate gradients (stream-ordered scratch)
    ICKA_CUDA(cudaMemsetAsync(dg, 0, (size_t)B * sizeof(float), st));
    gate_blend_bwd_stream_kernel<<<dim3((S + kGateRows - 1) / kGateRows, B), kGateThreads, 0, st>>>(dout, fused, tok, gate,
                                                                                                   dfused, dtok, dg, S, H);
    ICKA_LAUNCHED(h);
    gate_blend_bwd_cls_kernel<<<B, kGateThreads, 0, st>>>(dg, fused, tok, gate, ln_w, ln_b, ln_eps, w_fold, dfused, dtok,
                                                          d_ln_w, d_ln_b, d_w_fold, d_c_fold, S, H);
    ICKA_LAUNCHED(h);
    return ICKA_OK;
  }
  gate_blend_bwd_kernel<<<B, kGateThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      dout, fused, tok, gate, ln_w, ln_b, ln_eps, w_fold, dfused, dtok, d_ln_w, d_ln_b, d_w_fold, d_c_fold, S, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of the gate fold (elementwise.cu gate_fold_kernel):  w_fold = Wp^T wa,  c_fold = wa . bp + ba
//   dWp[j][k] = wa[j] dwf[k];  dwa[j] = Wp[j][:] . dwf + bp[j] dc;  dbp[j] = wa[j] dc;  dba = dc
// One block per row j of Wp.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128) gate_fold_bwd_kernel(const float* __restrict__ Wp, const float* __restrict__ bp,
                                                            const float* __restrict__ wa,
                                                            const float* __restrict__ d_w_fold,
                                                            const float* __restrict__ d_c_fold, float* __restrict__ dWp,
                                                            float* __restrict__ dbp, float* __restrict__ dwa,
                                                            float* __restrict__ dba, int H) {
  __shared__ float red[4];
  const int j = blockIdx.x;
  const float waj = wa[j], dc = d_c_fold[0];
  float dot = 0.0f;
  for (int k = threadIdx.x; k < H; k += 128) {
    const float g = d_w_fold[k];
    dWp[(size_t)j * H + k] = waj * g;
    dot = fmaf(Wp[(size_t)j * H + k], g, dot);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    dwa[j] = (red[0] + red[1]) + (red[2] + red[3]) + bp[j] * dc;
    dbp[j] = waj * dc;
    if (j == 0) dba[0] = dc;
  }
}
}  // namespace

extern "C" int icka_gate_fold_bwd(icka_handle* h, const float* Wp, const float* bp, const float* wa,
                                  const float* d_w_fold, const float* d_c_fold, float* dWp, float* dbp, float* dwa,
                                  float* dba, int H, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(H >= 1 && Wp && bp && wa && d_w_fold && d_c_fold && dWp && dbp && dwa && dba, "gate_fold_bwd: bad arguments");
  gate_fold_bwd_kernel<<<H, 128, 0, static_cast<cudaStream_t>(stream)>>>(Wp, bp, wa, d_w_fold, d_c_fold, dWp, dbp, dwa,
                                                                        dba, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

// ---- element-wise activation backward: dx = dy * act'(ref) -----------------------------------------------------------
// ref = the PRE-activation for gelu / relu / swish (BertIntermediate, CMIM:549-550 with config.hidden_act, CMIM:43) and the
// OUTPUT y for tanh (tanh' = 1 - y^2: the prompt mapping networks keep only the activated tensor, CMIM:917, :925).
// The default GELU layer never comes here inside a cross layer (its derivative is fused into the FFN-down dgrad epilogue).
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ ref, T* __restrict__ dx,
                                                      int64_t n, int act) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = to_f32<T>(dy[i]), r = to_f32<T>(ref[i]);
    float d;
    if (act == ICKA_ACT_TANH) d = 1.0f - r * r;
    else if (act == ICKA_ACT_RELU) d = act_relu_grad(r);
    else if (act == ICKA_ACT_SWISH) d = act_swish_grad(r);
    else d = gelu_erf_grad(r);
    dx[i] = from_f32<T>(g * d);
  }
}

extern "C" int icka_act_bwd(icka_handle* h, const void* dy, const void* ref, void* dx, int dtype, int64_t n, int act,
                            void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(dy && ref && dx && n >= 0, "act_bwd: bad arguments");
  ICKA_REQUIRE(dtype == ICKA_F32 || dtype == ICKA_BF16, "act_bwd: bad dtype %d", dtype);
  ICKA_REQUIRE(act == ICKA_ACT_GELU_ERF || act == ICKA_ACT_TANH || act == ICKA_ACT_RELU || act == ICKA_ACT_SWISH,
               "act_bwd: bad activation %d", act);
  if (n == 0) return ICKA_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ICKA_F32)
    act_bwd_kernel<float><<<(int)blocks, 256, 0, st>>>(static_cast<const float*>(dy), static_cast<const float*>(ref),
                                                       static_cast<float*>(dx), n, act);
  else
    act_bwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy),
                                                               static_cast<const __nv_bfloat16*>(ref),
                                                               static_cast<__nv_bfloat16*>(dx), n, act);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
