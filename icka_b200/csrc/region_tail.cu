// Tail of the ResNet-152 region producer (SURVEY 8f "next" row 3; resnet/resnet_utils.py:36-43):
//   fc  = x.mean(3).mean(2)                                   [B, C]
//   att = F.adaptive_avg_pool2d(x, [att_size, att_size])      [B, C, a, a]
// and, in the SAME pass over the layer4 output x [B, C, g, g], the K-major region rows [B*a*a, C] (bf16 or fp32) that the
// region projection consumes (CMIM:956-958) -- so a deployment that owns the producer never materialises the
// TMA-hostile 196-byte-pitch view, and icka_region_rows (a second read of the grid) disappears.  a == g is the
// reference's 7 x 7 case (pooling is the identity); a != g covers the hi-res variants (14 x 14 map pooled to 7 x 7, ...).
// One block moves a [64 channels x g*g] slab through shared memory: coalesced 16-byte reads, row-contiguous writes.
#include "common.cuh"

namespace {

constexpr int kTailC = 64;

template <typename OutT>
__global__ void __launch_bounds__(256)
region_tail_kernel(const float* __restrict__ x, float* __restrict__ fc, float* __restrict__ att, OutT* __restrict__ rows,
                   int C, int g, int a, int Gp, int Rp) {
  extern __shared__ __align__(16) float tail_smem[];
  const int G = g * g, R = a * a;
  float* tile = tail_smem;                    // [kTailC][Gp]
  float* pooled = (a == g) ? tile : tile + (size_t)kTailC * Gp;   // [kTailC][Rp]
  const int pp = (a == g) ? Gp : Rp;
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kTailC;
  const int nc = min(kTailC, C - c0);
  const float* src = x + ((size_t)b * C + c0) * G;
  const int total = nc * G;
  if (Gp == G && (total & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* t4 = reinterpret_cast<float4*>(tile);
#pragma unroll 4
    for (int i = threadIdx.x; i < (total >> 2); i += 256) t4[i] = __ldcs(s4 + i);
  } else {
    for (int i = threadIdx.x; i < total; i += 256) {
      const int c = i / G, r = i - c * G;
      tile[c * Gp + r] = src[i];
    }
  }
  __syncthreads();
  // fc: mean over the width, then over the height (the reference's order of the two means)
  if (fc != nullptr && threadIdx.x < nc) {
    const float* t = tile + threadIdx.x * Gp;
    const float inv = 1.0f / (float)g;
    float acc = 0.0f;
    for (int i = 0; i < g; ++i) {
      float row = 0.0f;
      for (int j = 0; j < g; ++j) row += t[i * g + j];
      acc += row * inv;
    }
    fc[(size_t)b * C + c0 + threadIdx.x] = acc * inv;
  }
  if (a != g) {
    // adaptive average pooling: bin [floor(i*g/a), ceil((i+1)*g/a))
    for (int i = threadIdx.x; i < nc * R; i += 256) {
      const int c = i / R, r = i - c * R;
      const int oi = r / a, oj = r - oi * a;
      const int h0 = (oi * g) / a, h1 = ((oi + 1) * g + a - 1) / a;
      const int w0 = (oj * g) / a, w1 = ((oj + 1) * g + a - 1) / a;
      const float* t = tile + c * Gp;
      float acc = 0.0f;
      for (int y = h0; y < h1; ++y)
        for (int z = w0; z < w1; ++z) acc += t[y * g + z];
      pooled[c * Rp + r] = acc / (float)((h1 - h0) * (w1 - w0));
    }
    __syncthreads();
  }
  if (att != nullptr) {   // [B, C, a, a]: each channel's R values are contiguous
    float* dst = att + ((size_t)b * C + c0) * R;
    for (int i = threadIdx.x; i < nc * R; i += 256) {
      const int c = i / R, r = i - c * R;
      dst[i] = pooled[c * pp + r];
    }
  }
  if (rows != nullptr) {  // [B*R, C]: two adjacent channels per thread, a warp covers one 64-channel row segment
    OutT* dst = rows + (size_t)b * R * C + c0;
    constexpr int pairs = kTailC / 2;
#pragma unroll 4
    for (int i = threadIdx.x; i < R * pairs; i += 256) {
      const int r = i / pairs, c = (i - r * pairs) * 2;
      if (c + 1 < nc) {
        const float v0 = pooled[c * pp + r], v1 = pooled[(c + 1) * pp + r];
        if constexpr (sizeof(OutT) == 2) {
          *reinterpret_cast<uint32_t*>(dst + (size_t)r * C + c) = pack_bf16x2(v0, v1);
        } else {
          *reinterpret_cast<float2*>(dst + (size_t)r * C + c) = make_float2(v0, v1);
        }
      } else if (c < nc) {
        dst[(size_t)r * C + c] = from_f32<OutT>(pooled[c * pp + r]);
      }
    }
  }
}

}  // namespace

extern "C" int icka_region_tail_fwd(icka_handle* h, const float* x, float* fc, float* att_f32, void* rows, int rows_dtype,
                                    int B, int C, int g, int att_size, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(B >= 0 && C >= 1 && g >= 1 && att_size >= 1, "region_tail: bad shape B=%d C=%d g=%d att_size=%d", B, C, g,
               att_size);
  ICKA_REQUIRE(x && (fc || att_f32 || rows), "region_tail: null input or no output requested");
  ICKA_REQUIRE(!rows || rows_dtype == ICKA_F32 || rows_dtype == ICKA_BF16, "region_tail: bad rows dtype %d", rows_dtype);
  ICKA_REQUIRE(!rows || (C % 2 == 0 && icka_aligned(rows, 8)), "region_tail: rows need an even C and 8-byte alignment");
  if (B == 0) return ICKA_OK;
  const int G = g * g, R = att_size * att_size;
  const int Gp = (G & 1) ? G : G + 1;          // odd pitch: conflict-free column walks
  const int Rp = (R & 1) ? R : R + 1;
  const size_t smem = (size_t)kTailC * (Gp + (att_size == g ? 0 : Rp)) * sizeof(float);
  ICKA_REQUIRE(smem <= h->smem_optin, "region_tail: a %d x %d map needs %zu B of shared memory (max %zu)", g, g, smem,
               h->smem_optin);
  ICKA_REQUIRE(B <= 65535, "region_tail: B=%d too large for one launch", B);
  dim3 grid((C + kTailC - 1) / kTailC, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows_dtype == ICKA_BF16 && rows) {
    ICKA_CUDA(cudaFuncSetAttribute(region_tail_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_tail_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(x, fc, att_f32, static_cast<__nv_bfloat16*>(rows), C, g,
                                                               att_size, Gp, Rp);
  } else {
    ICKA_CUDA(cudaFuncSetAttribute(region_tail_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    region_tail_kernel<float><<<grid, 256, smem, st>>>(x, fc, att_f32, static_cast<float*>(rows), C, g, att_size, Gp, Rp);
  }
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}
