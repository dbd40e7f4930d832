// HBM-bound helpers of the fusion path: operand casts, region-grid relayout, LayerNorm, gate + blend.
// All loops are 128-bit vectorised and coalesced; none of these kernels has data reuse beyond a row.
#include "common.cuh"
#include "philox.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (GEMM A operand copy; the fp32 residual stream is kept as is)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x,
                                                            __nv_bfloat16* __restrict__ y, int64_t n) {
  const int64_t nvec = n / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * v];
    const float4 b = reinterpret_cast<const float4*>(x)[2 * v + 1];
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y);
    o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(y)[v] = o;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nvec * 8 + threadIdx.x; i < n; i += blockDim.x) y[i] = __float2bfloat16_rn(x[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Split-precision operand ("bf16 x 3"): x fp32 [M, K] -> [M, 3K] bf16 with hi = bf16(x), lo = bf16(x - hi), laid out
//   as_weight = 0:  [hi | lo | hi]        as_weight = 1:  [hi | hi | lo]
// so that ONE bf16 tensor-core GEMM over K' = 3K forms a_hi.w_hi + a_lo.w_hi + a_hi.w_lo: the fp32 product to ~2^-16
// relative (the dropped a_lo.w_lo term).  Used for the single-query image->text chain (CMIM:981-989), whose GEMMs have
// M = batch rows only: one token per sentence walks 2 L layers, and at L = 5 plain bf16 operands bring it to the 2e-2 gate.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ x, int64_t ldx,
                                                           __nv_bfloat16* __restrict__ y, int M, int K, int as_weight) {
  const int kv = K / 4;
  const int64_t total = (int64_t)M * kv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int m = (int)(i / kv), c = (int)(i % kv) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)m * ldx + c);
    const float h0 = __bfloat162float(__float2bfloat16_rn(v.x)), h1 = __bfloat162float(__float2bfloat16_rn(v.y));
    const float h2 = __bfloat162float(__float2bfloat16_rn(v.z)), h3 = __bfloat162float(__float2bfloat16_rn(v.w));
    const uint2 hi = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    const uint2 lo = make_uint2(pack_bf16x2(v.x - h0, v.y - h1), pack_bf16x2(v.z - h2, v.w - h3));
    __nv_bfloat16* row = y + (size_t)m * 3 * K + c;
    *reinterpret_cast<uint2*>(row) = hi;
    *reinterpret_cast<uint2*>(row + K) = as_weight ? hi : lo;
    *reinterpret_cast<uint2*>(row + 2 * K) = as_weight ? lo : hi;
  }
}

// ------------------------------------------------------------------------------------------------
// bf16 -> fp32 widening (exact): a caller that already holds bf16 text states / token embeddings (the
// encoders ran in bf16) hands them over as the GEMM operand directly; this builds the fp32 residual stream.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x,
                                                            float* __restrict__ y, int64_t n) {
  const int64_t nvec = n / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const uint4 a = reinterpret_cast<const uint4*>(x)[v];
    float4 lo, hi;      // a bf16 is the upper half of the fp32 with the same value
    lo.x = __uint_as_float(a.x << 16); lo.y = __uint_as_float(a.x & 0xffff0000u);
    lo.z = __uint_as_float(a.y << 16); lo.w = __uint_as_float(a.y & 0xffff0000u);
    hi.x = __uint_as_float(a.z << 16); hi.y = __uint_as_float(a.z & 0xffff0000u);
    hi.z = __uint_as_float(a.w << 16); hi.w = __uint_as_float(a.w & 0xffff0000u);
    reinterpret_cast<float4*>(y)[2 * v] = lo;
    reinterpret_cast<float4*>(y)[2 * v + 1] = hi;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nvec * 8 + threadIdx.x; i < n; i += blockDim.x) y[i] = __bfloat162float(x[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Region relayout: grid [B, C, R] fp32 (R contiguous)  ->  rows [B*R, C] (C contiguous), CMIM:956.
// One block moves a [64 channels x R] slab through shared memory (coalesced on both sides).
// (Tried in round 2: a [256 channels x R] slab with a region-major bf16-pair tile -- 16-byte tile reads and 16-byte global
// stores, 4x fewer instructions per byte -- whose load side has to gather 4-byte elements in 32-byte runs to stay free of
// bank conflicts: 198 us against 135 us for this kernel at B = 1024.  The 16-byte global LOADS matter more.)
// ------------------------------------------------------------------------------------------------
constexpr int kRegC = 64;

template <typename OutT>
__global__ void __launch_bounds__(256) region_rows_kernel(const float* __restrict__ grid, OutT* __restrict__ rows,
                                                          int C, int R, int Rp) {
  extern __shared__ __align__(16) float tile[];   // [kRegC][Rp]; Rp == R (odd R: conflict-free as is) or R + 1
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kRegC;
  const int nc = min(kRegC, C - c0);
  const float* src = grid + ((size_t)b * C + c0) * R;
  const int total = nc * R;
  if (Rp == R && (total & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    // the slab is one contiguous run: straight 16-byte copies, several in flight per thread
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* t4 = reinterpret_cast<float4*>(tile);
    const int nv = total >> 2;
#pragma unroll 4
    for (int i = threadIdx.x; i < nv; i += 256) t4[i] = __ldcs(s4 + i);
  } else if ((R & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    // padded pitch (even R, e.g. the 14 x 14 grid): a warp copies whole channel rows, 16 bytes per lane from global
    // memory, four scalar stores into the odd-pitch tile (no per-element division, coalesced on the global side)
    const int warp_l = threadIdx.x >> 5, lane_l = threadIdx.x & 31;
    const int r4 = R >> 2;
    for (int c = warp_l; c < nc; c += 8) {
      const float4* s4 = reinterpret_cast<const float4*>(src + (size_t)c * R);
      float* trow = tile + c * Rp;
      for (int i = lane_l; i < r4; i += 32) {
        const float4 v = __ldcs(s4 + i);
        trow[4 * i] = v.x;
        trow[4 * i + 1] = v.y;
        trow[4 * i + 2] = v.z;
        trow[4 * i + 3] = v.w;
      }
    }
  } else {
    for (int i = threadIdx.x; i < total; i += 256) {
      const int c = i / R, r = i - c * R;
      tile[c * Rp + r] = src[i];
    }
  }
  __syncthreads();
  OutT* dst = rows + (size_t)b * R * C + c0;
  // A warp writes one 64-channel segment of a region row per pass.  Reading the tile with lane = channel PAIR (stride
  // 2 * Rp words) is a 2-way bank conflict for every odd or even pitch (ncu: 3.4 M conflicts, the kernel was issue /
  // conflict bound at 60 % of HBM); lane = channel (and channel + 32) with the odd pitch is conflict-free, and two
  // shuffles hand each lane the neighbouring channel it stores next to its own.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (nc == kRegC && Rp == R && (R & 31) == 4) {
    // Unpadded even pitch with R = 4 (mod 32) -- the 14 x 14 grid of the 448-px configuration, R = 196: lane = channel
    // reads column r with a 4-way bank conflict (bank = 4 lane + r), but if the lanes of octet o read column r0 + (o + i) % 4
    // instead, the 32 lanes hit 32 banks; after four rounds every lane has its channel at regions r0 .. r0 + 3.  The two
    // source lanes of a destination lane share an octet, so the shuffles of round i deliver region r0 + (o_src + i) % 4;
    // the destination keeps its four packed results and picks, for each output row, the round that produced it.  This is
    // what lets the slab arrive as one contiguous 16-byte copy (the padded-pitch path stores scalars: 59 % of HBM).
    const int src = 2 * (lane & 15);
    const bool upper = lane >= 16;
    const int oct = lane >> 3, oct_src = src >> 3;
    const int c = (upper ? 32 : 0) + src;
    for (int r0 = 4 * warp; r0 < R; r0 += 32) {
      float2 res[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = r0 + ((oct + i) & 3);
        const float lo = tile[lane * R + rr], hi = tile[(lane + 32) * R + rr];
        const float a0 = __shfl_sync(0xffffffffu, lo, src), a1 = __shfl_sync(0xffffffffu, lo, src + 1);
        const float b0 = __shfl_sync(0xffffffffu, hi, src), b1 = __shfl_sync(0xffffffffu, hi, src + 1);
        res[i] = make_float2(upper ? b0 : a0, upper ? b1 : a1);       // region r0 + (oct_src + i) % 4
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = (q - oct_src) & 3;
        const float2 v = i == 0 ? res[0] : i == 1 ? res[1] : i == 2 ? res[2] : res[3];
        if constexpr (sizeof(OutT) == 2) {
          *reinterpret_cast<uint32_t*>(dst + (size_t)(r0 + q) * C + c) = pack_bf16x2(v.x, v.y);
        } else {
          *reinterpret_cast<float2*>(dst + (size_t)(r0 + q) * C + c) = v;
        }
      }
    }
  } else if (nc == kRegC) {
    const int src = 2 * (lane & 15);                 // lanes 0-15 store channels 0-31, lanes 16-31 channels 32-63
    const bool upper = lane >= 16;
#pragma unroll 2
    for (int r = warp; r < R; r += 8) {
      const float lo = tile[lane * Rp + r], hi = tile[(lane + 32) * Rp + r];
      const float a0 = __shfl_sync(0xffffffffu, lo, src), a1 = __shfl_sync(0xffffffffu, lo, src + 1);
      const float b0 = __shfl_sync(0xffffffffu, hi, src), b1 = __shfl_sync(0xffffffffu, hi, src + 1);
      const float v0 = upper ? b0 : a0, v1 = upper ? b1 : a1;
      const int c = (upper ? 32 : 0) + src;
      if constexpr (sizeof(OutT) == 2) {
        *reinterpret_cast<uint32_t*>(dst + (size_t)r * C + c) = pack_bf16x2(v0, v1);
      } else {
        *reinterpret_cast<float2*>(dst + (size_t)r * C + c) = make_float2(v0, v1);
      }
    }
  } else {
    for (int i = threadIdx.x; i < R * nc; i += 256) {
      const int r = i / nc, c = i - r * nc;
      dst[(size_t)r * C + c] = from_f32<OutT>(tile[c * Rp + r]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BertLayerNorm (CMIM:518-522): u = mean(x); s = mean((x-u)^2); y = w * (x-u)/sqrt(s+eps) + b
// One warp per row, the row lives in registers (N <= 32*4*kMaxVec), two-pass statistics like the
// reference.  Emits the fp32 residual stream and/or the bf16 operand copy for the next GEMM.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxVec = 8;   // N <= 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        float* __restrict__ y32, __nv_bfloat16* __restrict__ y16,
                                                        int M, int N) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= M) return;
  const int nvec = N / 4;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * N);
  float4 v[kMaxVec];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      v[i] = xr[c];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(sum) / (float)N;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)N + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[c];
      const float4 bt = reinterpret_cast<const float4*>(beta)[c];
      float4 o;
      o.x = g.x * ((v[i].x - mean) * rstd) + bt.x;
      o.y = g.y * ((v[i].y - mean) * rstd) + bt.y;
      o.z = g.z * ((v[i].z - mean) * rstd) + bt.z;
      o.w = g.w * ((v[i].w - mean) * rstd) + bt.w;
      if (y32) reinterpret_cast<float4*>(y32 + (size_t)row * N)[c] = o;
      if (y16) {
        uint2 p;
        p.x = pack_bf16x2(o.x, o.y);
        p.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(y16 + (size_t)row * N)[c] = p;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gate fold: w_fold[k] = sum_j wa[j] * Wp[j][k];  c_fold = sum_j wa[j] * bp[j] + ba
// ------------------------------------------------------------------------------------------------
// Block = 32 columns k (lane = column: a warp reads 128 contiguous bytes of a row of Wp) x 8 warps that split the rows j;
// eight rows in flight per warp.  (The first version walked all H rows in one dependent chain per thread: 100 us for a
// 768 x 768 matrix-vector product, 3 % of a training step.)
__global__ void __launch_bounds__(256) gate_fold_kernel(const float* __restrict__ Wp, const float* __restrict__ bp,
                                                        const float* __restrict__ wa, const float* __restrict__ ba,
                                                        float* __restrict__ w_fold, float* __restrict__ c_fold, int H) {
  __shared__ float part[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 32 + lane;
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.0f;
  if (k < H) {
    int j = warp;
    for (; j + 56 < H; j += 64) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = fmaf(__ldg(wa + j + 8 * u), __ldg(Wp + (size_t)(j + 8 * u) * H + k), acc[u]);
    }
    for (; j < H; j += 8) acc[0] = fmaf(__ldg(wa + j), __ldg(Wp + (size_t)j * H + k), acc[0]);
  }
  part[warp][lane] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (warp == 0 && k < H) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    w_fold[k] = t;
  }
  if (blockIdx.x == 0 && warp == 1) {
    float a = 0.0f;
    for (int j = lane; j < H; j += 32) a = fmaf(wa[j], bp[j], a);
    a = warp_sum(a);
    if (lane == 0) c_fold[0] = a + ba[0];
  }
}

// ------------------------------------------------------------------------------------------------
// Gate + blend (CMIM:1029-1036).  One block per sentence: LayerNorm of the [CLS] rows, dot with the
// folded gate vector, sigmoid, then a streaming blend of the sentence's S*H elements.
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

constexpr int kGateRowsFwd = 16;
__global__ void __launch_bounds__(kGateThreads) gate_blend_kernel(
    const float* __restrict__ fused, const float* __restrict__ tok, const float* __restrict__ ln_w,
    const float* __restrict__ ln_b, float ln_eps, const float* __restrict__ w_fold, const float* __restrict__ c_fold,
    float* __restrict__ out, float* __restrict__ gate_out, int S, int H, int rpb) {
  __shared__ float red[kGateThreads / 32];
  // block = (sentence, chunk of rpb rows): every block recomputes the sentence's gate from the [CLS] rows (3 KB,
  // L2-resident) and streams its own rows -- with one block per sentence a training batch of 32-128 sentences left most
  // SMs idle
  const int b = blockIdx.x;
  const float* f = fused + (size_t)b * S * H;
  const float* t = tok + (size_t)b * S * H;
  float* o = out + (size_t)b * S * H;

  float sum = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) sum += f[k] + t[k];
  const float mean = block_sum(sum, red) / (float)H;
  float sq = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) {
    const float d = (f[k] + t[k]) - mean;
    sq += d * d;
  }
  const float rstd = 1.0f / sqrtf(block_sum(sq, red) / (float)H + ln_eps);
  float dot = 0.0f;
  for (int k = threadIdx.x; k < H; k += kGateThreads) {
    const float n = ((f[k] + t[k]) - mean) * rstd * ln_w[k] + ln_b[k];
    dot = fmaf(n, w_fold[k], dot);
  }
  const float logit = block_sum(dot, red) + c_fold[0];
  const float g = 1.0f / (1.0f + expf(-logit));
  const float og = 1.0f - g;
  if (threadIdx.x == 0 && blockIdx.y == 0 && gate_out) gate_out[b] = g;

  const int r0 = blockIdx.y * rpb, r1 = min(S, r0 + rpb);
  const int v0 = (int)((size_t)r0 * H / 4), nvec = (int)((size_t)r1 * H / 4);   // H % 4 != 0: rpb == S (one chunk)
  const float4* f4 = reinterpret_cast<const float4*>(f);
  const float4* t4 = reinterpret_cast<const float4*>(t);
  float4* o4 = reinterpret_cast<float4*>(o);
  for (int v = v0 + threadIdx.x; v < nvec; v += kGateThreads) {
    const float4 a = __ldcs(t4 + v), c = __ldcs(f4 + v);
    float4 r;
    r.x = g * a.x + og * c.x;
    r.y = g * a.y + og * c.y;
    r.z = g * a.z + og * c.z;
    r.w = g * a.w + og * c.w;
    __stcs(o4 + v, r);
  }
}

// ------------------------------------------------------------------------------------------------
// Last LayerNorm of the text->image encoder fused with the gate + blend (inference path):
//   fused = LN2(pre)            (BertOutput.LayerNorm, CMIM:535; `pre` = dense(h) + attention_output)
//   g     = sigmoid(w_fold . LN_g(fused[0] + tok[0]) + c_fold)            (CMIM:1029-1035)
//   out   = g tok + (1-g) fused                                            (CMIM:1036)
// Two launches: a tiny one (one warp per sentence) for the gates, then a streaming one (one warp per row, the
// row lives in registers).  Compared with running layernorm_kernel and gate_blend_kernel back to back this
// never writes the fp32 `fused` tensor nor reads it again: per sentence 393 KB (pre) + 393 KB (tok) in, 393 KB (out) + 197 KB (bf16 fused, the key/value
// operand of the image->text encoders) out, instead of 2.2 MB.
// ------------------------------------------------------------------------------------------------
// LayerNorm of one row held in v[] by a warp (two-pass statistics like the reference); gamma / beta are read
// through L1 each time (3 KB each, always resident) instead of living in registers.
__device__ __forceinline__ void warp_ln_row(float4 (&v)[kMaxVec], int lane, int nvec, float invH, float eps,
                                            const float* __restrict__ gamma, const float* __restrict__ beta) {
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (lane + 32 * i < nvec) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(sum) * invH;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (lane + 32 * i < nvec) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) * invH + eps);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c);
      v[i].x = g.x * (v[i].x * rstd) + bt.x;
      v[i].y = g.y * (v[i].y * rstd) + bt.y;
      v[i].z = g.z * (v[i].z * rstd) + bt.z;
      v[i].w = g.w * (v[i].w * rstd) + bt.w;
    }
  }
}

// Step 1 (tiny): the gate of each sentence from its [CLS] rows.  One warp per sentence.
__global__ void __launch_bounds__(128) ln_gate_kernel(
    const float* __restrict__ pre, const float* __restrict__ ln2_w, const float* __restrict__ ln2_b, float ln2_eps,
    const float* __restrict__ tok, const float* __restrict__ lng_w, const float* __restrict__ lng_b, float lng_eps,
    const float* __restrict__ w_fold, const float* __restrict__ c_fold, float* __restrict__ gate_out, int B, int S,
    int H) {
  const int b = blockIdx.x * 4 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (b >= B) return;
  const int nvec = H / 4;
  const float invH = 1.0f / (float)H;
  const float4* p4 = reinterpret_cast<const float4*>(pre + (size_t)b * S * H);
  const float4* t4 = reinterpret_cast<const float4*>(tok + (size_t)b * S * H);
  float4 f[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    f[i] = (c < nvec) ? p4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  warp_ln_row(f, lane, nvec, invH, ln2_eps, ln2_w, ln2_b);          // fused[0]
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 t = t4[c];
      f[i].x += t.x; f[i].y += t.y; f[i].z += t.z; f[i].w += t.w;
    }
  }
  warp_ln_row(f, lane, nvec, invH, lng_eps, lng_w, lng_b);
  float dot = 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 wf = __ldg(reinterpret_cast<const float4*>(w_fold) + c);
      dot += (f[i].x * wf.x + f[i].y * wf.y) + (f[i].z * wf.z + f[i].w * wf.w);
    }
  }
  const float logit = warp_sum(dot) + c_fold[0];
  if (lane == 0) gate_out[b] = 1.0f / (1.0f + expf(-logit));
}

// Step 2 (streaming, HBM-bound): one warp per row -- LayerNorm, fused copies, blend with the sentence's gate.
__global__ void __launch_bounds__(256) ln_blend_kernel(
    const float* __restrict__ pre, const float* __restrict__ ln2_w, const float* __restrict__ ln2_b, float ln2_eps,
    const float* __restrict__ tok, const float* __restrict__ gate, float* __restrict__ out,
    float* __restrict__ fused32, __nv_bfloat16* __restrict__ fused16, int M, int S, int H) {
  const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= M) return;
  const int nvec = H / 4;
  const float4* p4 = reinterpret_cast<const float4*>(pre + (size_t)row * H);
  const float4* t4 = reinterpret_cast<const float4*>(tok + (size_t)row * H);
  float4 f[kMaxVec], t[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      f[i] = __ldcs(p4 + c);
      t[i] = __ldcs(t4 + c);
    }
  }
  const float g = __ldg(gate + row / S), og = 1.0f - g;
  warp_ln_row(f, lane, nvec, 1.0f / (float)H, ln2_eps, ln2_w, ln2_b);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      if (fused32) reinterpret_cast<float4*>(fused32 + (size_t)row * H)[c] = f[i];
      if (fused16) {
        uint2 p;
        p.x = pack_bf16x2(f[i].x, f[i].y);
        p.y = pack_bf16x2(f[i].z, f[i].w);
        reinterpret_cast<uint2*>(fused16 + (size_t)row * H)[c] = p;
      }
      float4 o;
      o.x = g * t[i].x + og * f[i].x;
      o.y = g * t[i].y + og * f[i].y;
      o.z = g * t[i].z + og * f[i].z;
      o.w = g * t[i].w + og * f[i].w;
      __stcs(reinterpret_cast<float4*>(out + (size_t)row * H) + c, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Hidden-state dropout (CMIM:563, 534: `dropout(dense(h)) + input`), forward and backward share one kernel:
//   y = x * keep * 1/(1-p) (+ residual)      keep = Philox(seed, site, element index) < (1-p) 2^32
// x / y in fp32 or bf16 (the forward reads the fp32 dense output and writes the fp32 pre-LayerNorm rows; the
// backward turns the fp32 upstream gradient into the masked GEMM operand).  n must be a multiple of 4.
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) dropout_kernel(const TI* __restrict__ x, const float* __restrict__ residual,
                                                      TO* __restrict__ y, int64_t n4, uint32_t thresh, float scale,
                                                      uint64_t seed_arg, uint32_t site,
                                                      const unsigned long long* seed_base) {
  const uint64_t seed = icka_rng::effective_seed(seed_arg, seed_base);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
    float v[4];
    if constexpr (sizeof(TI) == 2) {
      const uint2 u = reinterpret_cast<const uint2*>(x)[g];
      v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
      v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    } else {
      const float4 f = reinterpret_cast<const float4*>(x)[g];
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    }
    const uint32_t keep = icka_rng::keep_bits4(seed, site, (uint64_t)g, thresh);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (keep >> j & 1u) ? v[j] * scale : 0.0f;
    if (residual) {
      const float4 r = reinterpret_cast<const float4*>(residual)[g];
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    if constexpr (sizeof(TO) == 2) {
      uint2 o;
      o.x = pack_bf16x2(v[0], v[1]);
      o.y = pack_bf16x2(v[2], v[3]);
      reinterpret_cast<uint2*>(y)[g] = o;
    } else {
      reinterpret_cast<float4*>(y)[g] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// Test helper: the keep mask itself (u8 0/1).  kind 0: hidden site over n elements (n % 4 == 0);
// kind 1: attention site, [rows, Skv] with rows = B*nh*Sq.
__global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* __restrict__ m, int64_t rows, int Skv, int kind,
                                                           uint32_t thresh, uint64_t seed_arg,
                                                           const unsigned long long* seed_base) {
  const uint64_t seed = icka_rng::effective_seed(seed_arg, seed_base);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (kind == 0) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < rows / 4; g += stride) {
      const uint32_t keep = icka_rng::keep_bits4(seed, icka_rng::kSiteHidden, (uint64_t)g, thresh);
      for (int j = 0; j < 4; ++j) m[4 * g + j] = (keep >> j) & 1u;
    }
  } else {
    const int gpr = (Skv + 3) / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * gpr; i += stride) {
      const int64_t row = i / gpr;
      const int k4 = (int)(i % gpr);
      const uint32_t keep = icka_rng::keep_bits4(seed, icka_rng::kSiteAttention, icka_rng::attn_group((uint64_t)row, Skv, 4 * k4), thresh);
      for (int j = 0; j < 4; ++j)
        if (4 * k4 + j < Skv) m[row * Skv + 4 * k4 + j] = (keep >> j) & 1u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Additive attention mask (CMIM:962-965, 976-977): out[b][j] = (1 - mask[b][j]) * -10000
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_additive_kernel(const int64_t* __restrict__ mask, int64_t ld,
                                                            float* __restrict__ out, int B, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * n) {
    const int b = i / n, j = i - b * n;
    out[i] = (1.0f - (float)mask[(size_t)b * ld + j]) * -10000.0f;
  }
}

}  // namespace

extern "C" int icka_mask_additive(icka_handle* h, const int64_t* mask, int64_t ld, float* out, int B, int n,
                                  void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(mask && out && B >= 0 && n >= 1 && ld >= n, "mask_additive: bad arguments");
  if (B == 0) return ICKA_OK;
  mask_additive_kernel<<<(B * n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, ld, out, B, n);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_cast_f32_to_bf16(icka_handle* h, const float* x, void* y, int64_t n, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(n >= 0 && x && y, "cast: bad arguments");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(y, 16), "cast: pointers must be 16-byte aligned");
  if (n == 0) return ICKA_OK;
  int64_t blocks = (n / 8 + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), n);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_split_bf16x3(icka_handle* h, const float* x, int64_t ldx, void* y_bf16, int M, int K, int as_weight,
                                 void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(x && y_bf16 && M >= 0 && K >= 4 && K % 4 == 0 && ldx >= K && ldx % 4 == 0,
               "split_bf16x3: bad arguments (K and the pitch must be multiples of 4)");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(y_bf16, 8), "split_bf16x3: pointers must be 16 / 8-byte aligned");
  if (M == 0) return ICKA_OK;
  int64_t blocks = ((int64_t)M * (K / 4) + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  split_bf16x3_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, static_cast<__nv_bfloat16*>(y_bf16), M, K, as_weight ? 1 : 0);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_cast_bf16_to_f32(icka_handle* h, const void* x, float* y, int64_t n, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(n >= 0 && x && y, "cast: bad arguments");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(y, 16), "cast: pointers must be 16-byte aligned");
  if (n == 0) return ICKA_OK;
  int64_t blocks = (n / 8 + 255) / 256;
  const int64_t cap = (int64_t)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_bf16_f32_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), y, n);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_region_rows(icka_handle* h, const float* grid, void* rows, int out_dtype, int B, int C, int R,
                                void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && C >= 1 && R >= 1 && grid && rows, "region_rows: bad arguments");
  ICKA_REQUIRE(C % 2 == 0, "region_rows: channel count %d must be even", C);
  ICKA_REQUIRE(B <= 65535, "region_rows: B=%d exceeds the grid.y limit; shard the batch", B);
  ICKA_REQUIRE(out_dtype == ICKA_F32 || out_dtype == ICKA_BF16, "region_rows: bad dtype %d", out_dtype);
  if (B == 0) return ICKA_OK;
  const int Rp = ((R & 31) == 4 && C % kRegC == 0) ? R : (R | 1);   // R = 4 (mod 32): unpadded, rotated reads (see the kernel)
  const size_t smem = (size_t)kRegC * Rp * sizeof(float);
  ICKA_REQUIRE(smem <= h->smem_optin, "region_rows: R=%d too large", R);
  dim3 g((C + kRegC - 1) / kRegC, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == ICKA_BF16) {
    ICKA_CUDA(cudaFuncSetAttribute(region_rows_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_rows_kernel<__nv_bfloat16><<<g, 256, smem, st>>>(grid, static_cast<__nv_bfloat16*>(rows), C, R, Rp);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(region_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_rows_kernel<float><<<g, 256, smem, st>>>(grid, static_cast<float*>(rows), C, R, Rp);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_layernorm_fwd(icka_handle* h, const float* x, const float* gamma, const float* beta, float eps,
                                  float* y_f32, void* y_bf16, int M, int N, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(M >= 0 && N >= 4 && x && gamma && beta, "layernorm: bad arguments");
  ICKA_REQUIRE(N % 4 == 0 && N <= 128 * kMaxVec, "layernorm: N=%d must be a multiple of 4 and <= %d", N, 128 * kMaxVec);
  ICKA_REQUIRE(y_f32 || y_bf16, "layernorm: no output");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(gamma, 16) && icka_aligned(beta, 16) &&
                   icka_aligned(y_f32, 16) && icka_aligned(y_bf16, 8),
               "layernorm: pointers must be 16-byte aligned");
  if (M == 0) return ICKA_OK;
  const int rows_per_block = 8;
  layernorm_kernel<<<(M + rows_per_block - 1) / rows_per_block, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, gamma, beta, eps, y_f32, static_cast<__nv_bfloat16*>(y_bf16), M, N);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_gate_fold(icka_handle* h, const float* Wp, const float* bp, const float* wa, const float* ba,
                              float* w_fold, float* c_fold, int H, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(H >= 1 && Wp && bp && wa && ba && w_fold && c_fold, "gate_fold: bad arguments");
  gate_fold_kernel<<<(H + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(Wp, bp, wa, ba, w_fold, c_fold, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_gate_blend_fwd(icka_handle* h, const float* fused, const float* tok, const float* ln_w,
                                   const float* ln_b, float ln_eps, const float* w_fold, const float* c_fold,
                                   float* out, float* gate_out, int B, int S, int H, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && H >= 1 && fused && tok && ln_w && ln_b && w_fold && c_fold && out,
               "gate_blend: bad arguments");
  ICKA_REQUIRE(((size_t)S * H) % 4 == 0, "gate_blend: S*H must be a multiple of 4");
  ICKA_REQUIRE(icka_aligned(fused, 16) && icka_aligned(tok, 16) && icka_aligned(out, 16),
               "gate_blend: pointers must be 16-byte aligned");
  if (B == 0) return ICKA_OK;
  const int rpb = (H % 4 == 0 && S <= 65535 * kGateRowsFwd) ? kGateRowsFwd : S;   // H % 4 != 0: chunks would start unaligned
  gate_blend_kernel<<<dim3(B, (S + rpb - 1) / rpb), kGateThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      fused, tok, ln_w, ln_b, ln_eps, w_fold, c_fold, out, gate_out, S, H, rpb);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_ln_gate_blend_fwd(icka_handle* h, const float* pre, const float* ln2_w, const float* ln2_b,
                                      float ln2_eps, const float* tok, const float* lng_w, const float* lng_b,
                                      float lng_eps, const float* w_fold, const float* c_fold, float* out,
                                      float* fused_f32, void* fused_bf16, float* gate_out, int B, int S, int H,
                                      void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && S >= 1 && H >= 4 && pre && ln2_w && ln2_b && tok && lng_w && lng_b && w_fold && c_fold && out,
               "ln_gate_blend: bad arguments");
  ICKA_REQUIRE(H % 4 == 0 && H <= 128 * kMaxVec, "ln_gate_blend: H=%d must be a multiple of 4 and <= %d", H,
               128 * kMaxVec);
  ICKA_REQUIRE(icka_aligned(pre, 16) && icka_aligned(tok, 16) && icka_aligned(out, 16) && icka_aligned(ln2_w, 16) &&
                   icka_aligned(ln2_b, 16) && icka_aligned(lng_w, 16) && icka_aligned(lng_b, 16) &&
                   icka_aligned(w_fold, 16) && icka_aligned(fused_f32, 16) && icka_aligned(fused_bf16, 8),
               "ln_gate_blend: pointers must be 16-byte aligned");
  ICKA_REQUIRE(gate_out != nullptr, "ln_gate_blend: gate_out is required (it carries the gates between the two launches)");
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ln_gate_kernel<<<(B + 3) / 4, 128, 0, st>>>(pre, ln2_w, ln2_b, ln2_eps, tok, lng_w, lng_b, lng_eps, w_fold, c_fold,
                                              gate_out, B, S, H);
  ICKA_LAUNCHED(h);
  const int M = B * S;
  ln_blend_kernel<<<(M + 7) / 8, 256, 0, st>>>(pre, ln2_w, ln2_b, ln2_eps, tok, gate_out, out, fused_f32,
                                               static_cast<__nv_bfloat16*>(fused_bf16), M, S, H);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_dropout_fwd(icka_handle* h, const void* x, int x_dtype, const float* residual, void* y, int y_dtype,
                                int64_t n, float p_drop, uint64_t seed, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(x && y && n >= 0 && n % 4 == 0, "dropout: bad arguments (n must be a multiple of 4)");
  ICKA_REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "dropout: p=%f outside [0, 1)", (double)p_drop);
  ICKA_REQUIRE((x_dtype == ICKA_F32 || x_dtype == ICKA_BF16) && (y_dtype == ICKA_F32 || y_dtype == ICKA_BF16), "dropout: bad dtype");
  ICKA_REQUIRE(icka_aligned(x, 16) && icka_aligned(y, 16) && icka_aligned(residual, 16), "dropout: pointers must be 16-byte aligned");
  if (n == 0) return ICKA_OK;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)h->sm_count * 8) blocks = (int64_t)h->sm_count * 8;
  const uint32_t th = icka_rng::keep_threshold(p_drop);
  const float sc = 1.0f / (1.0f - p_drop);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  using B16 = __nv_bfloat16;
#define ICKA_DROP(TI, TO) dropout_kernel<TI, TO><<<(int)blocks, 256, 0, st>>>(static_cast<const TI*>(x), residual, static_cast<TO*>(y), n4, th, sc, seed, icka_rng::kSiteHidden, h->seed_base)
  if (x_dtype == ICKA_F32) { if (y_dtype == ICKA_F32) ICKA_DROP(float, float); else ICKA_DROP(float, B16); }
  else                     { if (y_dtype == ICKA_F32) ICKA_DROP(B16, float);   else ICKA_DROP(B16, B16); }
#undef ICKA_DROP
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

extern "C" int icka_dropout_mask(icka_handle* h, uint8_t* mask, int64_t rows, int Skv, int kind, float p_drop,
                                 uint64_t seed, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(mask && rows >= 0 && (kind == 0 || (kind == 1 && Skv >= 1)), "dropout_mask: bad arguments");
  ICKA_REQUIRE(kind != 0 || rows % 4 == 0, "dropout_mask: element count must be a multiple of 4");
  if (rows == 0) return ICKA_OK;
  dropout_mask_kernel<<<h->sm_count * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, rows, Skv, kind,
                                                                                      icka_rng::keep_threshold(p_drop), seed,
                                                                                      h->seed_base);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
