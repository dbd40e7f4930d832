/* icka_b200.h -- C ABI of libicka_b200.so: the B200 (sm_100a) drop-in for ICKA's cross-modal
 * fusion + CRF hot path.
 *
 * The reference (buctcurry/ICKA) is pure Python/PyTorch and has no FFI of its own: the boundary it
 * exposes for this path is nn.Module composition inside Cross_Modal_Interaction_Module.py ("CMIM").
 * Each entry point below replaces the ATen call chain of the cited reference statement(s); the Python
 * mirror of the reference modules (icka_b200/modules.py, icka_b200/crf.py) binds them with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; every pointer is a DEVICE pointer on the handle's
 *     device unless stated otherwise; tensors are dense row-major.
 *   - return 0 on success, <0 on error; icka_last_error() returns a thread-local message.
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *     no entry point synchronises with the host or allocates device memory (icka_create does, once).
 *   - re-entrant: one handle per device; no global mutable state (CMIM is driven by one thread per GPU
 *     under nn.DataParallel, My_cross_attention.py:777-779).
 *   - the handle owns one 32 MiB device workspace (split-K partial tiles of skinny icka_linear_fwd calls):
 *     calls on the SAME handle must be issued to one stream at a time (or be ordered by events).
 *   - there is NO CPU fallback: every entry point fails if the device is not sm_100.
 */
#ifndef ICKA_B200_H_
#define ICKA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct icka_handle icka_handle;

enum icka_status {
  ICKA_OK = 0,
  ICKA_ERR_INVALID = -1,      /* bad argument (shape, alignment, dtype) */
  ICKA_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed       */
  ICKA_ERR_UNSUPPORTED = -3   /* device is not sm_100 / shape outside the kernels' envelope */
};

enum icka_dtype { ICKA_F32 = 0, ICKA_BF16 = 1 };
enum icka_act {
  ICKA_ACT_NONE = 0,
  ICKA_ACT_GELU_ERF = 1,     /* CMIM:31-37 */
  ICKA_ACT_GELU_ERF_BWD = 2, /* internal: multiply by gelu'(pre-activation) (backward of CMIM:550) */
  ICKA_ACT_TANH = 3,         /* torch.nn.Tanh of the prompt mapping networks, CMIM:917, :925 */
  ICKA_ACT_RELU = 4,         /* ACT2FN["relu"], CMIM:43 (config.hidden_act) */
  ICKA_ACT_SWISH = 5         /* ACT2FN["swish"] = x * sigmoid(x), CMIM:38-39, :43 */
};

int icka_version(void);
const char* icka_last_error(void);
int icka_create(int device, icka_handle** out);
int icka_destroy(icka_handle* h);
/* Device-resident dropout seed base.  Every dropout site (icka_dropout_fwd, the *_drop attention entry points,
 * icka_dropout_mask) draws its Philox masks from  seed_argument + *seed_base_dev  when a base is set (NULL clears it).  A
 * training step captured into a CUDA graph replays with frozen kernel arguments: the caller advances the 8-byte counter
 * with a kernel of the graph itself, so every replay drops different elements while backward still regenerates the
 * masks of its own forward.  The pointer must stay valid while the handle uses it. */
int icka_set_seed_base(icka_handle* h, const uint64_t* seed_base_dev);

/* Number of kernels this handle has launched so far (bench.py reports it as gpu_launches). */
int64_t icka_launch_count(const icka_handle* h);

/* ---- ingest ---------------------------------------------------------------------------------- */

/* fp32 -> bf16 copy of a GEMM operand (the residual stream itself stays fp32). */
int icka_cast_f32_to_bf16(icka_handle* h, const float* x, void* y_bf16, int64_t n, void* stream);

/* bf16 -> fp32 widening (exact).  For callers whose encoders already produce bf16 states (the "bf16 path" of the
 * north star): the bf16 tensor is used as the GEMM operand as is, this builds the fp32 residual stream
 * (`+ input_tensor` of CMIM:564) / the fp32 `token_embedding` operand of the gate + blend (CMIM:1036). */
int icka_cast_bf16_to_f32(icka_handle* h, const void* x_bf16, float* y, int64_t n, void* stream);

/* Split-precision GEMM operand ("bf16 x 3"): x fp32 [M, K] (row pitch ldx) -> y bf16 [M, 3K] holding hi = bf16(x) and
 * lo = bf16(x - hi) as [hi | lo | hi] (as_weight = 0, the activation side) or [hi | hi | lo] (as_weight = 1, the nn.Linear
 * weight side).  icka_linear_fwd over K' = 3K of two such operands accumulates a_hi.w_hi + a_lo.w_hi + a_hi.w_lo in fp32:
 * the fp32 product to ~2^-16 relative on the tensor cores.  The single-query image->text encoders (CMIM:981-989; M = batch
 * rows) use it for vismapping (CMIM:954) and their FFN so that the CLIP token stays inside the bf16 gate (2e-2) after the
 * 2 x layer_num1 layers it walks (10 at the script default, My_cross_attention.py:603).  K % 4 == 0. */
int icka_split_bf16x3(icka_handle* h, const float* x, int64_t ldx, void* y_bf16, int M, int K, int as_weight, void* stream);

/* CMIM:956  `visual_embeds_att.view(-1, 2048, 49).permute(0, 2, 1)`:
 * grid [B, C, R] fp32 (R contiguous) -> rows [B*R, C] in `out_dtype` (C contiguous, the K-major GEMM
 * operand the region projection wants). */
int icka_region_rows(icka_handle* h, const float* grid, void* rows, int out_dtype,
                     int B, int C, int R, void* stream);

/* CMIM:962-965 / 976-977  `(1.0 - mask) * -10000.0`: 0/1 int64 mask [B, >= n] (row pitch ld) -> additive
 * fp32 mask [B, n] for the attention kernels. */
int icka_mask_additive(icka_handle* h, const int64_t* mask, int64_t ld, float* out, int B, int n, void* stream);

/* ---- dense layers ---------------------------------------------------------------------------- */

/* nn.Linear with fused epilogue:  out[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) (+ residual[M,N]).
 * Replaces CMIM:958 (vismap2text), :954 (vismapping), :592-594 (query/key/value), :562 (+ the `+ input`
 * of :564 via `residual`), :549-550 (dense + erf-GELU), :533 (+ :535 residual).
 *   in_dtype  ICKA_BF16: A and W are bf16 -> tcgen05/TMEM tensor-core kernel fed by TMA, fp32 accumulate.
 *             ICKA_F32 : A and W are fp32 -> CUDA-core FFMA kernel (the 1e-5 parity path).
 *   bias fp32 (may be NULL); residual fp32 [M,N] (may be NULL); out_dtype ICKA_F32 or ICKA_BF16.
 *   lda / ldw / ldo are row pitches in ELEMENTS (lda >= K, ldw >= K, ldo >= N); residual has pitch N.
 *   bf16 path needs 16-byte aligned pointers and lda, ldw multiples of 8. */
int icka_linear_fwd(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                    const float* bias, const float* residual, void* out, int64_t ldo,
                    int in_dtype, int out_dtype, int M, int N, int K, int act, void* stream);

/* Dense layer + residual + BertLayerNorm in one call (BertSelfOutput CMIM:561-565, BertOutput CMIM:532-536):
 *   out = LayerNorm_{gamma,beta,eps}(A[M,K] . W[N,K]^T + bias + residual[M,N])
 * out_f32 [M,N] (pitch N) and optionally out_bf16 [M,N] (the next GEMM's operand; may be NULL).  bf16 operands:
 * the normalisation runs in the epilogue of the tcgen05 GEMM.  N = 768 / 1024 (the model widths): a thread-block cluster
 * of N / 256 CTAs owns each 128-row block, CTA r holds columns [256 r, 256 r + 256) in TMEM, pre-LayerNorm values are
 * parked in TMEM, the row sums travel over distributed shared memory and the normalised rows leave through TMA stores
 * (csrc/gemm_ln_sm100.cu).  Other widths: one CTA owns whole 128-row blocks, keeps row sums while it walks the n-tiles
 * and re-reads its own (L2-resident) fp32 stores once to normalise them in place.
 * fp32 operands: FFMA GEMM followed by the row kernel. */
int icka_linear_ln_fwd(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                       const float* bias, const float* residual, const float* gamma, const float* beta, float eps,
                       float* out_f32, void* out_bf16, int in_dtype, int M, int N, int K, void* stream);

/* Developer / test switch for icka_linear_ln_fwd with bf16 operands: 0 = choose per shape (default), 1 = the single-CTA
 * kernel, 2 = the cluster kernel (N = 512 / 768 / 1024: N / 256 CTAs per 128-row block, statistics over DSMEM). */
int icka_set_ln_mode(int mode);

/* Training-time forward: icka_linear_fwd that can also keep the pre-activation (A.W^T + bias, in the
 * operand dtype, pitch N) of a GELU layer for the backward pass.  pre_act_out may be NULL. */
int icka_linear_fwd_ex(icka_handle* h, const void* A, int64_t lda, const void* W, int64_t ldw,
                       const float* bias, const float* residual, void* out, int64_t ldo, void* pre_act_out,
                       int in_dtype, int out_dtype, int M, int N, int K, int act, void* stream);

/* Backward of nn.Linear w.r.t. its input (autograd of CMIM:533, :549, :562, :592-594 ...):
 *   dX[M,K] = (dY[M,N] . W[N,K]) * gelu'(gelu_pre[M,K]) + residual[M,K]
 * dY and W in `in_dtype` (W exactly as nn.Linear stores it, no transposed copy); gelu_pre (same dtype, pitch
 * ldg) and residual (fp32, pitch K) may be NULL.  bf16 path: K % 64 == 0. */
int icka_linear_dgrad(icka_handle* h, const void* dY, int64_t ldd, const void* W, int64_t ldw,
                      const float* residual, const void* gelu_pre, int64_t ldg, void* dX, int64_t ldo,
                      int in_dtype, int out_dtype, int M, int N, int K, void* stream);

/* Backward of nn.Linear w.r.t. its weight:  dW[N,K] (+)= dY[M,N]^T . X[M,K]  (fp32 dW, pitch K).
 * accumulate = 0 overwrites dW, 1 adds to it (gradient accumulation, My_cross_attention.py:826-844).
 * bf16 path: N % 64 == 0 and K % 64 == 0; token range split over CTAs, partial sums added with red.global. */
int icka_linear_wgrad(icka_handle* h, const void* dY, int64_t ldd, const void* X, int64_t ldx, float* dW,
                      int in_dtype, int M, int N, int K, int accumulate, void* stream);

/* Backward of an element-wise activation (autograd of CMIM:550 with config.hidden_act != 'gelu' -- ACT2FN, CMIM:43 -- and
 * of the torch.nn.Tanh of the prompt mapping networks, CMIM:917, :925):  dx[i] = dy[i] * act'(ref[i]), all three tensors
 * in `dtype`.  ref is the PRE-activation for ICKA_ACT_GELU_ERF / _RELU / _SWISH and the activation's OUTPUT for
 * ICKA_ACT_TANH (1 - y^2).  dx may alias dy. */
int icka_act_bwd(icka_handle* h, const void* dy, const void* ref, void* dx, int dtype, int64_t n, int act, void* stream);

/* Column sums of x[M,N] (pitch ld, fp32 or bf16) -> out[N] fp32: the bias gradient of a dense layer. */
int icka_colsum(icka_handle* h, const void* x, int64_t ld, int dtype, float* out, int M, int N, int accumulate,
                void* stream);

/* Backward of BertLayerNorm (CMIM:518-522) for y = LN(x):  dx from dy and the saved pre-LN x; dgamma[N],
 * dbeta[N] and dbias[N] (= column sums of dx: the bias gradient of the dense layer that produced x) are
 * ADDED to (each may be NULL).  dx_f32 and/or dx_bf16 (either may be NULL). */
int icka_layernorm_bwd(icka_handle* h, const float* dy, const float* x, const float* gamma, float eps,
                       float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dbias, int M, int N,
                       void* stream);

/* Tile-shape policy of the bf16 path (process-wide tuning/testing knob): 0 = choose per shape (default),
 * 1 = always single-CTA 128 x N tiles, 2 = always CTA pairs (cluster of 2, tcgen05 cta_group::2, 256 x 256). */
int icka_set_gemm_mode(int mode);

/* BertLayerNorm CMIM:518-522 (TF style: biased variance, eps inside the sqrt) over rows of x[M,N]
 * (x already holds dense(...) + input).  Writes y_f32 and/or y_bf16 (either may be NULL). */
int icka_layernorm_fwd(icka_handle* h, const float* x, const float* gamma, const float* beta, float eps,
                       float* y_f32, void* y_bf16, int M, int N, void* stream);

/* ---- attention core -------------------------------------------------------------------------- */

/* BertCoAttention CMIM:598-623 after the three projections: per sentence b and head h
 *   P = softmax( Q_h K_h^T / sqrt(d) + mask_add[b] ),  ctx_h = P V_h,  heads merged in place.
 * q [B*Sq, >=nh*d] (pitch ldq), k/v [B*Skv, ...] (pitch ldkv; k and v may alias one [K|V] buffer),
 * mask_add [B, Skv] fp32 additive (0 / -10000, CMIM:965, :977) or NULL, ctx [B*Sq, nh*d] (pitch ldc).
 * `dtype` applies to q, k, v and ctx.  d must be 64. */
int icka_cross_attn_core_fwd(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                             int64_t ldkv, const float* mask_add, void* ctx, int64_t ldc, int dtype,
                             int B, int Sq, int Skv, int nh, int d, void* stream);

/* Training-time variants with dropout on the attention probabilities (CMIM:616): P' = P * keep / (1 - p_drop)
 * enters P.V, the softmax normaliser keeps the undropped sum.  keep = Philox4x32-10(seed, element) -- forward and
 * backward regenerate the same mask from (seed, p_drop); p_drop = 0 is the plain kernel. */
int icka_cross_attn_core_fwd_drop(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                                  int64_t ldkv, const float* mask_add, void* ctx, int64_t ldc, int dtype,
                                  int B, int Sq, int Skv, int nh, int d, float p_drop, uint64_t seed, void* stream);
int icka_cross_attn_core_bwd_drop(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                                  int64_t ldkv, const float* mask_add, const void* ctx, int64_t ldctx,
                                  const void* dctx, int64_t ldc, void* dq, int64_t lddq, void* dk, void* dv,
                                  int64_t lddkv, int dtype, int B, int Sq, int Skv, int nh, int d, float p_drop,
                                  uint64_t seed, void* stream);

/* Hidden-state dropout (CMIM:563, 534), forward and backward alike:  y = x * keep / (1 - p_drop) (+ residual)
 * over n elements (n % 4 == 0); x / y fp32 or bf16, residual fp32 or NULL; keep = Philox(seed, element index). */
int icka_dropout_fwd(icka_handle* h, const void* x, int x_dtype, const float* residual, void* y, int y_dtype,
                     int64_t n, float p_drop, uint64_t seed, void* stream);

/* Test helper: the keep mask (u8 0/1) the kernels above regenerate.  kind 0: hidden site, `rows` = element count;
 * kind 1: attention site, mask [rows = B*nh*Sq, Skv]. */
int icka_dropout_mask(icka_handle* h, uint8_t* mask, int64_t rows, int Skv, int kind, float p_drop, uint64_t seed,
                      void* stream);

/* Kernel choice of the bf16 attention core and of icka_i2t_pool_fwd (process-wide tuning/testing knob): 0 = per
 * shape (tcgen05 / TMEM kernels for Skv <= 64 resp. S <= 128, mma.sync kernels otherwise; default), 1 = always the
 * mma.sync kernels, 2 = like 0 plus the (not faster) wide tcgen05 attention variant for 64 < Skv <= 224. */
int icka_set_attn_mode(int mode);

/* Backward of icka_cross_attn_core_fwd: probabilities are recomputed from q, k, v (same layouts as the
 * forward); dctx [B*Sq, nh*d] -> dq [B*Sq, nh*d] (pitch lddq), dk / dv [B*Skv, nh*d] (pitch lddkv; may be the
 * halves of one [dK|dV] buffer).  ctx (the forward output, pitch ldctx) may be NULL; with it, bf16 and
 * Sq <= 128 the tensor-core kernel runs (any Skv); otherwise the fp32 CUDA-core kernel (Skv <= ~150). */
int icka_cross_attn_core_bwd(icka_handle* h, const void* q, int64_t ldq, const void* k, const void* v,
                             int64_t ldkv, const float* mask_add, const void* ctx, int64_t ldctx, const void* dctx,
                             int64_t ldc, void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv, int dtype,
                             int B, int Sq, int Skv, int nh, int d, void* stream);

/* Single-query attention in folded form (image->text encoders, CMIM:984-989 with Sq = 1).
 * With one query per sentence, scores[h][s] = (U_h . x_s)/sqrt(d) + mask[s] where U_h = Wk_h^T q_h, and
 * ctx_h = Wv_h xbar_h + bv_h where xbar_h = sum_s softmax_s(scores[h])[s] x_s.  This entry computes the
 * part that touches the text states:  U [B, nh*H] bf16, X [B*S, H] bf16, mask_add [B,S] fp32 additive or
 * NULL  ->  xbar [B, nh*H] bf16.  The weight folds around it are plain icka_linear_fwd calls.
 * H in {768, 1024}, nh <= 16. */
int icka_i2t_pool_fwd(icka_handle* h, const void* U, const void* X, const float* mask_add, void* xbar,
                      int B, int S, int H, int nh, void* stream);

/* ---- gated fusion ---------------------------------------------------------------------------- */

/* Folds cls_layer_both.proj (CMIM:877, :882) and aux_head (CMIM:934, :1034) into one H-vector:
 *   w_fold = Wp^T wa,  c_fold = wa . bp + ba   so that  logit = w_fold . LN(feat) + c_fold. */
int icka_gate_fold(icka_handle* h, const float* Wp, const float* bp, const float* wa, const float* ba,
                   float* w_fold, float* c_fold, int H, void* stream);

/* CMIM:1029-1036: feat = LN_{eps}(fused[:,0] + tok[:,0]); g = sigmoid(w_fold.feat + c_fold);
 * out = g*tok + (1-g)*fused.  fused/tok/out [B,S,H] fp32, gate_out [B] fp32 (may be NULL). */
int icka_gate_blend_fwd(icka_handle* h, const float* fused, const float* tok, const float* ln_w,
                        const float* ln_b, float ln_eps, const float* w_fold, const float* c_fold,
                        float* out, float* gate_out, int B, int S, int H, void* stream);

/* Inference fusion of the text->image encoder's last LayerNorm (CMIM:535) with the gate + blend:
 *   fused = LN_{ln2}(pre);  out = icka_gate_blend_fwd(fused, tok, ...)
 * `pre` [B,S,H] fp32 is dense(h) + attention_output of the last layer (what icka_linear_fwd wrote with its
 * residual epilogue).  fused_f32 [B,S,H] and fused_bf16 [B*S,H] (the key/value operand of the image->text
 * encoders) are optional outputs (NULL to skip); gate_out [B] is required. */
int icka_ln_gate_blend_fwd(icka_handle* h, const float* pre, const float* ln2_w, const float* ln2_b, float ln2_eps,
                           const float* tok, const float* lng_w, const float* lng_b, float lng_eps,
                           const float* w_fold, const float* c_fold, float* out, float* fused_f32,
                           void* fused_bf16, float* gate_out, int B, int S, int H, void* stream);

/* Backward of icka_gate_blend_fwd: dout [B,S,H] -> dfused, dtok [B,S,H] (dtok may be NULL); the gradients of
 * the LayerNorm affine (d_ln_w, d_ln_b [H]) and of the folded gate vector (d_w_fold [H], d_c_fold [1]) are
 * ADDED to.  `gate` is the gate_out of the forward. */
int icka_gate_blend_bwd(icka_handle* h, const float* dout, const float* fused, const float* tok, const float* gate,
                        const float* ln_w, const float* ln_b, float ln_eps, const float* w_fold, float* dfused,
                        float* dtok, float* d_ln_w, float* d_ln_b, float* d_w_fold, float* d_c_fold, int B, int S,
                        int H, void* stream);

/* Backward of icka_gate_fold: (d_w_fold [H], d_c_fold [1]) -> dWp [H,H], dbp [H], dwa [H], dba [1] (overwritten). */
int icka_gate_fold_bwd(icka_handle* h, const float* Wp, const float* bp, const float* wa, const float* d_w_fold,
                       const float* d_c_fold, float* dWp, float* dbp, float* dwa, float* dba, int H, void* stream);

/* ---- CRF ------------------------------------------------------------------------------------- */

/* torchcrf.CRF.decode (call sites CMIM:1051, :1056): Viterbi best path per sentence.
 * emissions [B,S,T] fp32, mask [B,S] u8 (NULL = all on), start/end [T], trans [T,T] (from i to j),
 * tags_out [B,S] i32 (positions >= len hold -1), lens_out [B] i32.  T <= 32.  Bit-exact with the
 * reference's fp32 op order ((score+trans)+emission, first index wins ties).  Inputs must be NaN-free. */
int icka_viterbi_decode(icka_handle* h, const float* emissions, const uint8_t* mask, const float* start,
                        const float* end, const float* trans, int32_t* tags_out, int32_t* lens_out,
                        int B, int S, int T, void* stream);

/* torchcrf.CRF.forward with reduction='none' (call sites CMIM:1047-1048, :1052-1053): per-sentence
 * log-likelihood  llh[b] = score(gold path) - log Z.  tags [B,S] i64. */
int icka_crf_llh_fwd(icka_handle* h, const float* emissions, const int64_t* tags, const uint8_t* mask,
                     const float* start, const float* end, const float* trans, float* llh_out,
                     int B, int S, int T, void* stream);

/* Gradient of sum_b w[b] * llh[b] (llh as icka_crf_llh_fwd; autograd of CMIM:1047-1048): forward-backward
 * marginals minus the gold path.  d_emissions [B,S,T] is overwritten (zeros at masked-off steps); d_start [T],
 * d_end [T], d_trans [T,T] are ADDED to.  w [B] fp32 is the upstream gradient per sentence. */
int icka_crf_llh_bwd(icka_handle* h, const float* emissions, const int64_t* tags, const uint8_t* mask,
                     const float* start, const float* end, const float* trans, const float* w,
                     float* d_emissions, float* d_start, float* d_end, float* d_trans, int B, int S, int T,
                     void* stream);

/* ---- tag post-processing + chunk-F1 ----------------------------------------- */

/* Replaces the host loops that consume the decoded tags: My_cross_attention.py:879-903 / 1052-1077 (walk each
 * sentence while its mask is on, keep positions whose GOLD label is not flagged ICKA_NER_SKIP) and
 * ner_evaluate.py:4-48, 64-110 (get_chunks / evaluate) on the kept positions.
 * pred [B,S] i32 (icka_viterbi_decode's tags_out), gold [B,S] i64 label ids, mask [B,S] u8 (NULL = all on).
 * label_info [n_ids] u16 per label id: bits 0-7 = chunk-type id (name.split('-')[-1]), ICKA_NER_SKIP,
 * ICKA_NER_OUTSIDE (the 'O' tag), ICKA_NER_BEGIN (name.split('-')[0] == 'B').
 * totals [6] u64 are ADDED to: kept tokens, kept tokens with pred == gold, |gold & pred chunks|, |pred chunks|,
 * |gold chunks|, positions with an id outside [0, n_ids) (must stay 0).  per_sentence [B,5] i32 (or NULL) gets
 * the first five per sentence.  Exact integer arithmetic. */
#define ICKA_NER_SKIP (1u << 8)
#define ICKA_NER_OUTSIDE (1u << 9)
#define ICKA_NER_BEGIN (1u << 10)
int icka_ner_chunk_counts(icka_handle* h, const int32_t* pred, const int64_t* gold, const uint8_t* mask,
                          const uint16_t* label_info, int n_ids, unsigned long long* totals,
                          int32_t* per_sentence, int B, int S, void* stream);

/* ---- emission head: BiLSTM + classifier -------------------------------------- */

/* Recurrent half of `self.lstm = nn.LSTM(H, H, batch_first=True, bidirectional=True)` (CMIM:905-908, call :1042):
 * one persistent, weight-stationary tcgen05 kernel walks all S steps of both directions (csrc/lstm_sm100.cu).
 *   gx        [S*B, 2*4H] bf16, TIME-MAJOR rows (row = t*B + sentence; x from icka_cast_bf16_time_major)
 *             = x . W_ih^T + b_ih + b_hh for both directions (an icka_linear_fwd call), columns in the
 *             kernel's slice order:  col = ((dir*16 + slice)*4 + blk)*48 + jg*16 + gate*4 + jj  <->  PyTorch row
 *             gate*H + slice*48 + blk*12 + jg*4 + jj of direction `dir` (gate order i, f, g, o; blk < 4, jg < 3, jj < 4)
 *   w_hh_perm [2*4H, H] bf16: weight_hh_l0 / weight_hh_l0_reverse with their rows in the same order
 *   workspace icka_lstm_rec_workspace_bytes(B, H) bytes, 1024-byte aligned (arrival counters; zeroed by the call)
 *   y         [S, B, 2H] bf16 TIME-MAJOR (forward states in [:H], backward in [H:]); h_n, c_n [2, B, H] fp32 or NULL
 * Every step then reads / writes one contiguous block of gx / y.  More than 2048 sentences run as consecutive
 * launches of <= 2048 (a CTA keeps the cell state of <= 4 sentence tiles: two in registers, two in spare TMEM columns).
 * H = 768 only (other sizes: the per-step path, icka_linear_fwd + icka_lstm_cell_fwd).  The launch is cooperative:
 * 64 or 128 co-resident CTAs in clusters of 2 (tcgen05 cta_group::2 pairs sharing a 48-unit weight slice). */
int64_t icka_lstm_rec_workspace_bytes(int B, int H);
/* Two kernels serve the call; they differ in the slice width and therefore in the column order above:
 *   variant 2 (B > 256): CTA pairs, 48-unit slices -- the order documented above;
 *   variant 1 (B <= 256): single CTAs, 24-unit slices (lower step latency):
 *             col = ((dir*32 + slice)*2 + half)*48 + jg*16 + gate*4 + jj  <->  gate*H + slice*24 + half*12 + jg*4 + jj.
 * icka_lstm_rec_variant(B) says which one a batch of B sentences gets (ICKA_LSTM_VARIANT=1|2 overrides); the caller
 * prepares gx / w_hh_perm for it and passes it back. */
int icka_lstm_rec_variant(int B);
int icka_lstm_rec_fwd(icka_handle* h, const void* gx, const void* w_hh_perm, void* workspace, int64_t workspace_bytes,
                      void* y, float* h_n, float* c_n, int B, int S, int H, int variant, void* stream);

/* One LSTM step (per-step path: fp32 parity mode, or shapes icka_lstm_rec_fwd does not cover):
 *   pre = gates_h[B,4H] (fp32, = h_{t-1} . W_hh^T; NULL at the first step) + gx[B, 4H] (`dtype`, row pitch ldgx)
 *   c = sigmoid(pre_f) * c + sigmoid(pre_i) * tanh(pre_g);  h = sigmoid(pre_o) * tanh(c)      (gate order i, f, g, o)
 * c [B,H] fp32 is updated in place; h goes to h_out [B,H] (`dtype`, the next step's GEMM operand), to y (`dtype`, row
 * pitch ldy; may be NULL) and to h_f32 [B,H] (may be NULL). */
int icka_lstm_cell_fwd(icka_handle* h, const float* gates_h, const void* gx, int64_t ldgx, float* c, void* h_out,
                       void* y, int64_t ldy, float* h_f32, int dtype, int B, int H, void* stream);

/* Training (BPTT through the per-step path; autograd of CMIM:1042).  Forward step like icka_lstm_cell_fwd that also keeps
 * the gate activations `acts` [B,4H] fp32 (i, f, g, o after sigmoid / tanh) and writes the new cell state to c_out
 * (c_prev may be NULL = zeros); h goes to h_out [B,H] (`dtype`), y_op (`dtype`, pitch ldyo) and y32 (fp32, pitch ldy32). */
int icka_lstm_cell_fwd_save(icka_handle* h, const float* gates_h, const void* gx, int64_t ldgx, const float* c_prev,
                            float* c_out, float* acts, void* h_out, void* y_op, int64_t ldyo, float* y32, int64_t ldy32,
                            int dtype, int B, int H, void* stream);

/* One backward step: dh = dy[B,H] (pitch lddy, gradient of this step's output) + dh_rec[B,H] (from the next step through
 * W_hh; NULL at the last step); dc [B,H] in: cell gradient from the next step, out: for the previous step;
 * dpre [B,4H] (`dtype`, pitch lddp) = gradient of the gate pre-activations (operand of the dgrad / wgrad GEMMs). */
int icka_lstm_cell_bwd(icka_handle* h, const float* dy, int64_t lddy, const float* dh_rec, float* dc, const float* acts,
                       const float* c_prev, const float* c_new, void* dpre, int64_t lddp, int dtype, int B, int H,
                       void* stream);

/* One direction of the training recurrence as ONE host call (the per-step launches of icka_linear_fwd /
 * icka_lstm_cell_fwd_save, resp. icka_lstm_cell_bwd / icka_linear_dgrad, issued from C: from Python the step was host-bound).
 * Strides in ELEMENTS: row (sentence) pitch and position (time) stride of the [B, S, .] tensors; `*_dir` pointers are
 * already offset to this direction's column block.  acts [S,B,4H], c_all [S,B,H] fp32 are indexed by STEP.  Scratch:
 * gates [B,4H] fp32 and h [2,B,H] (`dtype`) for the forward; dc, dh [B,H] fp32 for the backward. */
int icka_lstm_dir_fwd_save(icka_handle* h, const void* gx_dir, int64_t ld_gx_row, int64_t gx_pos_stride, const void* w_hh,
                           float* acts, float* c_all, void* y_op_dir, float* y32_dir, int64_t ld_y_row, int64_t y_pos_stride,
                           float* gates_scratch, void* h_scratch, int dtype, int B, int S, int H, int reverse, void* stream);
int icka_lstm_dir_bwd(icka_handle* h, const float* dy_dir, int64_t ld_dy_row, int64_t dy_pos_stride, const void* w_hh,
                      const float* acts, const float* c_all, void* dg_dir, int64_t ld_dg_row, int64_t dg_pos_stride,
                      float* dc_scratch, float* dh_scratch, int dtype, int B, int S, int H, int reverse, void* stream);

/* `self.classifier = nn.Linear(2H, num_labels)` (CMIM:910, :1043): out[M,T] fp32 = x[M,K] (`dtype`, pitch ldx) .
 * W[T,K]^T (fp32) + bias[T].  T <= 16, K % 8 == 0.  fp32 accumulation in a fixed order.
 * time_major_S = S > 0: the rows of x are time-major (row = t*B + b, B = M/S, as icka_lstm_rec_fwd writes them) and
 * the emissions are written batch-major (row b*S + t), the layout the CRF entry points take; 0: same row order. */
int icka_emission_head_fwd(icka_handle* h, const void* x, int64_t ldx, const float* W, const float* bias, float* out,
                           int dtype, int64_t M, int K, int T, int time_major_S, void* stream);

/* Training (backpropagation through time) of the bidirectional LSTM (CMIM:905-908, 1042; autograd of nn.LSTM) with bf16
 * operands, H = 768: one launch per time step computes the recurrent product and the cell arithmetic of both directions
 * (csrc/lstm_train.cu).
 * Every sequence tensor is TIME-MAJOR ([S,B,.], row = position * B + sentence): a step touches one contiguous block.
 *   fwd_save: gx [S,B,8H] bf16 = x . [W_ih; W_ih_reverse]^T + biases (gate order i,f,g,o per direction),
 *             w_hh_* [4H,H] bf16 -> y_op [S,B,2H] bf16 and y32 [S,B,2H] fp32 (the output sequence), acts [2,S,B,4H] fp32
 *             (gate activations per STEP) and c_all [2,S,B,H] fp32 (cell states per step) for the backward pass.
 *   bwd:      dy [S,B,2H] fp32, w_hh_t_* [H,4H] bf16 (W_hh transposed), the saved acts / c_all -> dg [S,B,8H] bf16 = the
 *             gradients of the gate pre-activations; dc_scratch [2,B,H] fp32 is working storage.
 *             The weight / bias / input gradients follow from dg as plain GEMMs (icka_linear_wgrad / _dgrad, icka_colsum). */
int icka_lstm_bidir_fwd_save(icka_handle* h, const void* gx, const void* w_hh_fwd, const void* w_hh_bwd, void* y_op,
                             float* y32, float* acts, float* c_all, int B, int S, int H, void* stream);
int icka_lstm_bidir_bwd(icka_handle* h, const float* dy, const void* w_hh_t_fwd, const void* w_hh_t_bwd, const float* acts,
                        const float* c_all, void* dg, float* dc_scratch, int B, int S, int H, void* stream);

/* Backward of the classifier (CMIM:910, 1043; autograd of torch.nn.Linear(2H, num_labels)) in one pass over the states:
 * dx [M,K] fp32 = dout [M,T] . W [T,K]  (dx may be NULL) and dW [T,K] fp32 (+)= dout^T . x  (dW may be NULL; zeroed
 * first unless accumulate != 0).  x [M,K] fp32 or bf16 with row pitch ldx, T <= 16, K % 4 == 0.  time_major_S = S > 0:
 * the rows of x / dx are time-major (t*B + b) while dout is batch-major (b*S + t), as icka_emission_head_fwd pairs them. */
int icka_emission_head_bwd(icka_handle* h, const float* dout, const void* x, int64_t ldx, const float* W, float* dx,
                           float* dW, int dtype, int64_t M, int K, int T, int accumulate, int time_major_S, void* stream);

/* x [B,S,H] (fp32 or bf16) -> y [S,B,H] bf16: the time-major operand of the input projection of icka_lstm_rec_fwd. */
int icka_cast_bf16_time_major(icka_handle* h, const void* x, void* y_bf16, int in_dtype, int B, int S, int H,
                              void* stream);

/* out = a + b (fp32; weight preparation: b_ih + b_hh). */
int icka_add_f32(icka_handle* h, const float* a, const float* b, float* out, int64_t n, void* stream);

/* ---- region producer tail ---------------------------------------------------- */

/* resnet/resnet_utils.py:36-43 on the layer4 output x [B, C, g, g] fp32, in ONE pass:
 *   fc [B, C] = x.mean(3).mean(2);  att_f32 [B, C, a, a] = adaptive_avg_pool2d(x, [a, a]);
 *   rows [B*a*a, C] (`rows_dtype`) = att.view(-1, C, a*a).permute(0, 2, 1) (CMIM:956) -- the K-major operand of the
 *   region projection, which CrossModalFusion accepts directly (icka_region_rows is then skipped).
 * Any of fc / att_f32 / rows may be NULL (at least one output). */
int icka_region_tail_fwd(icka_handle* h, const float* x, float* fc, float* att_f32, void* rows, int rows_dtype,
                         int B, int C, int g, int att_size, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICKA_B200_H_ */
