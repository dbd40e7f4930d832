// Single-query (image -> text) attention pool, folded form.
//
// In ICKA's image->text encoders (cls_layer_Y, CMIM:901, 984-989) ONE projected CLIP token attends over
// the 128 fused text states; 96 % of the layer's FLOPs would go to K/V projections of tokens that are
// then reduced to a single row.  With q fixed per sentence the projections fold through the softmax
// (SURVEY 7.3 #6):
//     scores[h][s] = q_h . (Wk_h x_s + bk_h) / 8 = (U_h . x_s) / 8 + const_h,     U_h = Wk_h^T q_h
//     ctx_h        = sum_s p[h][s] (Wv_h x_s + bv_h) = Wv_h xbar_h + bv_h,        xbar_h = sum_s p[h][s] x_s
// (const_h shifts every key of a head equally and cancels in the softmax).  The two weight products
// become ordinary GEMMs around this kernel (icka_b200/modules.py); this kernel is the part that touches
// the text states: per sentence, U [nh x H] against X [S x H], softmax over s with the additive text
// mask, and the probability-weighted pool of X.  It reads X exactly once (HBM-bound: S*H*2 bytes per
// sentence) and is arithmetically an attention with nh "queries" of head-dim H whose K and V are both X.
//
// One CTA per sentence, 8 warps.  Warp w owns the hidden-dim slice [w*H/8, (w+1)*H/8): it computes the
// partial scores of that slice for a 32-key block (mma.sync m16n8k16 bf16, fp32 accumulate), the 8
// partials are summed through shared memory, every warp redoes the (cheap) online-softmax update, and
// accumulates its own slice of xbar with a second MMA.  X blocks are double-buffered with cp.async.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kKeyBlk = 32;
constexpr int kSpPitch = 32;   // floats; 4-float groups XOR-swizzled by the row (sp_at) -> conflict-free 8-byte accesses
                               // without padding, which keeps the CTA under half an SM's shared memory (2 CTAs / SM)

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

// Element (row, col) of a [16][32] fp32 partial-score tile (col even; the pair col, col+1 stays adjacent).
__device__ __forceinline__ float* sp_at(float* tile, int row, int col) {
  return tile + row * kSpPitch + (col ^ ((row & 7) << 2));
}

// Row-major bf16 tile with H elements per row; 16-byte chunk c of row r is stored at chunk c ^ (r & 7).
template <int H>
__device__ __forceinline__ __nv_bfloat16* tile_chunk(__nv_bfloat16* base, int row, int chunk) {
  return base + (size_t)row * H + ((chunk ^ (row & 7)) << 3);
}

template <int H>
__global__ void __launch_bounds__(kThreads, (H <= 768) ? 2 : 1) i2t_pool_kernel(const __nv_bfloat16* __restrict__ U,
                                                            const __nv_bfloat16* __restrict__ X,
                                                            const float* __restrict__ mask_add,
                                                            __nv_bfloat16* __restrict__ xbar, int S, int nh) {
  constexpr int DW = H / kWarps;        // hidden-dim slice per warp
  constexpr int KS = DW / 16;           // k-steps of the score MMA per warp
  constexpr int NT = DW / 8;            // n-tiles of the pool MMA per warp
  constexpr int CPR = H / 8;            // 16-byte chunks per row
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Xs0 = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Xs1 = Xs0 + (size_t)kKeyBlk * H;
  float* Sp = reinterpret_cast<float*>(Xs1 + (size_t)kKeyBlk * H);          // [kWarps][16][kSpPitch]
  float* Ms = Sp + kWarps * 16 * kSpPitch;                                    // [2][kKeyBlk]

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* Xb = X + (size_t)b * S * H;
  const __nv_bfloat16* Ub = U + (size_t)b * nh * H;

  auto load_block = [&](int blk, __nv_bfloat16* dst, float* mdst) {
    const int key0 = blk * kKeyBlk;
    for (int i = tid; i < kKeyBlk * CPR; i += kThreads) {
      const int r = i / CPR, c = i - r * CPR;
      if (key0 + r < S) cp_async16(tile_chunk<H>(dst, r, c), Xb + (size_t)(key0 + r) * H + c * 8);
      else *reinterpret_cast<uint4*>(tile_chunk<H>(dst, r, c)) = make_uint4(0, 0, 0, 0);
    }
    if (tid < kKeyBlk) {
      const int key = key0 + tid;
      mdst[tid] = (key < S) ? (mask_add ? mask_add[(size_t)b * S + key] * kLog2e : 0.0f) : -INFINITY;
    }
  };

  // ---- stage U (rows >= nh are zero) in buffer 1, X block 0 in buffer 0 ----
  for (int i = tid; i < 16 * CPR; i += kThreads) {
    const int r = i / CPR, c = i - r * CPR;
    if (r < nh) cp_async16(tile_chunk<H>(Xs1, r, c), Ub + (size_t)r * H + c * 8);
    else *reinterpret_cast<uint4*>(tile_chunk<H>(Xs1, r, c)) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  load_block(0, Xs0, Ms);
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 1;\n" ::);
  __syncthreads();
  uint32_t ua[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
    ldmatrix_x4(ua[ks], tile_chunk<H>(Xs1, lane & 15, (warp * DW + ks * 16) / 8 + (lane >> 4)));

  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

  const int nblk = (S + kKeyBlk - 1) / kKeyBlk;
  for (int blk = 0; blk < nblk; ++blk) {
    __nv_bfloat16* Xs = (blk & 1) ? Xs1 : Xs0;
    const float* Mb = Ms + (blk & 1) * kKeyBlk;
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();   // block `blk` visible to all; everyone is done with block blk-1 (and with U)
    if (blk + 1 < nblk) load_block(blk + 1, (blk & 1) ? Xs0 : Xs1, Ms + ((blk + 1) & 1) * kKeyBlk);
    asm volatile("cp.async.commit_group;\n" ::);

    // ---- partial scores of this warp's hidden-dim slice: [16 heads x 32 keys] ----
    float sacc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.0f;
#pragma unroll
      for (int k2 = 0; k2 < KS / 2; ++k2) {
        uint32_t kf[4];
        ldmatrix_x4(kf, tile_chunk<H>(Xs, nt * 8 + (lane & 7), warp * (DW / 8) + k2 * 4 + (lane >> 3)));
        mma_bf16_16816(sacc[nt], ua[2 * k2], kf[0], kf[1]);
        mma_bf16_16816(sacc[nt], ua[2 * k2 + 1], kf[2], kf[3]);
      }
    }
    float* spw = Sp + warp * 16 * kSpPitch;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<float2*>(sp_at(spw, g, nt * 8 + 2 * t)) = make_float2(sacc[nt][0], sacc[nt][1]);
      *reinterpret_cast<float2*>(sp_at(spw, g + 8, nt * 8 + 2 * t)) = make_float2(sacc[nt][2], sacc[nt][3]);
    }
    __syncthreads();
    // ---- sum the 8 partials, scale, mask, online softmax (every warp keeps the full statistics) ----
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float2 a = make_float2(0.0f, 0.0f), c = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const float2 p0 = *reinterpret_cast<const float2*>(sp_at(Sp + w * 16 * kSpPitch, g, nt * 8 + 2 * t));
        const float2 p1 = *reinterpret_cast<const float2*>(sp_at(Sp + w * 16 * kSpPitch, g + 8, nt * 8 + 2 * t));
        a.x += p0.x; a.y += p0.y; c.x += p1.x; c.y += p1.y;
      }
      constexpr float kScale = 0.125f * kLog2e;
      const float mk0 = Mb[nt * 8 + 2 * t], mk1 = Mb[nt * 8 + 2 * t + 1];
      sacc[nt][0] = fmaf(a.x, kScale, mk0);
      sacc[nt][1] = fmaf(a.y, kScale, mk1);
      sacc[nt][2] = fmaf(c.x, kScale, mk0);
      sacc[nt][3] = fmaf(c.y, kScale, mk1);
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      bm0 = fmaxf(bm0, fmaxf(sacc[nt][0], sacc[nt][1]));
      bm1 = fmaxf(bm1, fmaxf(sacc[nt][2], sacc[nt][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    const float sc0 = fast_exp2(m0 - mn0), sc1 = fast_exp2(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    float ps0 = 0.0f, ps1 = 0.0f;
    uint32_t pf[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float p0 = fast_exp2(sacc[nt][0] - mn0), p1 = fast_exp2(sacc[nt][1] - mn0);
      const float p2 = fast_exp2(sacc[nt][2] - mn1), p3 = fast_exp2(sacc[nt][3] - mn1);
      ps0 += p0 + p1;
      ps1 += p2 + p3;
      pf[nt][0] = pack_bf16x2(p0, p1);
      pf[nt][1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * sc0 + ps0;
    l1 = l1 * sc1 + ps1;
#pragma unroll
    for (int j = 0; j < NT; ++j) { o[j][0] *= sc0; o[j][1] *= sc0; o[j][2] *= sc1; o[j][3] *= sc1; }
    // ---- xbar slice += P . X[block, slice] ----
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t a[4] = {pf[2 * ks][0], pf[2 * ks][1], pf[2 * ks + 1][0], pf[2 * ks + 1][1]};
#pragma unroll
      for (int jn = 0; jn < NT; jn += 2) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, tile_chunk<H>(Xs, ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8,
                                            warp * (DW / 8) + jn + (lane >> 4)));
        mma_bf16_16816(o[jn], a, vf[0], vf[1]);
        mma_bf16_16816(o[jn + 1], a, vf[2], vf[3]);
      }
    }
  }

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* ob = xbar + (size_t)b * nh * H + warp * DW;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (g < nh) *reinterpret_cast<uint32_t*>(ob + (size_t)g * H + j * 8 + 2 * t) = pack_bf16x2(o[j][0] * i0, o[j][1] * i0);
    if (g + 8 < nh)
      *reinterpret_cast<uint32_t*>(ob + (size_t)(g + 8) * H + j * 8 + 2 * t) = pack_bf16x2(o[j][2] * i1, o[j][3] * i1);
  }
}

template <int H>
int launch_pool(icka_handle* h, const void* U, const void* X, const float* mask_add, void* xbar, int B, int S, int nh,
                cudaStream_t st) {
  const size_t smem = (size_t)2 * kKeyBlk * H * 2 + (size_t)kWarps * 16 * kSpPitch * 4 + 2 * kKeyBlk * 4;
  if (smem > h->smem_optin) ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "i2t_pool: needs %zu B shared memory", smem);
  ICKA_CUDA(cudaFuncSetAttribute(i2t_pool_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // two CTAs per SM only fit with the largest shared-memory carve-out
  ICKA_CUDA(cudaFuncSetAttribute(i2t_pool_kernel<H>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  i2t_pool_kernel<H><<<B, kThreads, smem, st>>>(static_cast<const __nv_bfloat16*>(U),
                                                static_cast<const __nv_bfloat16*>(X), mask_add,
                                                static_cast<__nv_bfloat16*>(xbar), S, nh);
  ICKA_LAUNCHED(h);
  return ICKA_OK;
}

}  // namespace

int icka_i2t_pool_tcgen05_launch(icka_handle* h, const void* U, const void* X, const float* mask_add, void* xbar, int B,
                                 int S, int H, int nh, cudaStream_t st);
extern int g_attn_mode;   // attention.cu: 0 = tcgen05 kernels where they apply, 1 = mma.sync kernels only

extern "C" int icka_i2t_pool_fwd(icka_handle* h, const void* U, const void* X, const float* mask_add, void* xbar,
                                 int B, int S, int H, int nh, void* stream) {
  ICKA_CHECK_HANDLE(h);
  ICKA_REQUIRE(U && X && xbar, "i2t_pool: null pointer");
  ICKA_REQUIRE(B >= 0 && S >= 1 && nh >= 1 && nh <= 16, "i2t_pool: bad shape B=%d S=%d nh=%d (nh <= 16)", B, S, nh);
  ICKA_REQUIRE(icka_aligned(U, 16) && icka_aligned(X, 16) && icka_aligned(xbar, 16), "i2t_pool: pointers must be 16-byte aligned");
  if (B == 0) return ICKA_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_attn_mode != 1) {
    const int rc = icka_i2t_pool_tcgen05_launch(h, U, X, mask_add, xbar, B, S, H, nh, st);
    if (rc <= 0) return rc;      // launched (0) or failed (< 0); > 0: shape outside that kernel's envelope
  }
  if (H == 768) return launch_pool<768>(h, U, X, mask_add, xbar, B, S, nh, st);
  if (H == 1024) return launch_pool<1024>(h, U, X, mask_add, xbar, B, S, nh, st);
  ICKA_FAIL(ICKA_ERR_UNSUPPORTED, "i2t_pool: hidden size %d not instantiated (768 and 1024 are)", H);
}
