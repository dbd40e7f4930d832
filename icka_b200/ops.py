"""Tensor-level wrappers over the C ABI (include/icka_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic step is a kernel of
libicka_b200.so launched on torch's current stream.  There is no eager/CPU fallback -- a CPU tensor, a
missing library or a non-sm_100 device raises ``RuntimeError``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, BF16, F32

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _dev(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError('icka_b200 ops need CUDA tensors (there is no CPU fallback)')
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _ctx(t: torch.Tensor):
    d = _dev(t)
    return _lib.load(), _lib.handle(d), torch.cuda.current_stream(d).cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype:
        raise RuntimeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise RuntimeError(f'{name}: tensor must be contiguous')


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _need(x, torch.float32, 'cast_bf16(x)')
    lib, h, st = _ctx(x)
    y = torch.empty_like(x, dtype=torch.bfloat16)
    _lib.check(lib.icka_cast_f32_to_bf16(h, x.data_ptr(), y.data_ptr(), x.numel(), st), 'icka_cast_f32_to_bf16')
    return y


def cast_f32(x: torch.Tensor) -> torch.Tensor:
    """bf16 -> fp32 (exact widening): the fp32 residual stream of a caller-supplied bf16 tensor."""
    _need(x, torch.bfloat16, 'cast_f32(x)')
    lib, h, st = _ctx(x)
    y = torch.empty_like(x, dtype=torch.float32)
    if x.numel() == 0:
        return y
    _lib.check(lib.icka_cast_bf16_to_f32(h, x.data_ptr(), y.data_ptr(), x.numel(), st), 'icka_cast_bf16_to_f32')
    return y


def split3(x: torch.Tensor, as_weight: bool = False) -> torch.Tensor:
    """fp32 [M, K] (unit stride along K) -> bf16 [M, 3K] split-precision operand: [hi | lo | hi] (activations) or
    [hi | hi | lo] (``as_weight``); see icka_split_bf16x3."""
    if x.dim() != 2 or x.dtype != torch.float32 or x.stride(1) != 1:
        raise RuntimeError('split3: need a 2-D fp32 tensor with unit stride along K')
    M, K = x.shape
    lib, h, st = _ctx(x)
    y = torch.empty(M, 3 * K, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.icka_split_bf16x3(h, x.data_ptr(), _ld(x, K), y.data_ptr(), M, K, int(as_weight), st), 'icka_split_bf16x3')
    return y


def region_rows(grid: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    """[B, C, g, g] (or [B, C, R]) fp32 ResNet grid -> [B*R, C] rows (CMIM:956)."""
    _need(grid, torch.float32, 'region_rows(grid)')
    B, C = grid.shape[0], grid.shape[1]
    R = grid.numel() // max(B * C, 1)
    lib, h, st = _ctx(grid)
    rows = torch.empty(B * R, C, dtype=out_dtype, device=grid.device)
    _lib.check(lib.icka_region_rows(h, grid.data_ptr(), rows.data_ptr(), _DT[out_dtype], B, C, R, st),
               'icka_region_rows')
    return rows


def mask_additive(mask01: torch.Tensor, n: int) -> torch.Tensor:
    """0/1 mask [B, >= n] (any integer/bool/float dtype) -> additive fp32 [B, n]: (1 - m) * -10000 (CMIM:962-965)."""
    if mask01.dim() != 2 or mask01.shape[1] < n:
        raise RuntimeError(f'mask_additive: mask of shape {tuple(mask01.shape)} has fewer than {n} columns')
    m = mask01 if mask01.dtype == torch.int64 else mask01.long()
    if m.stride(1) != 1:
        m = m.contiguous()
    B = m.shape[0]
    lib, h, st = _ctx(m)
    out = torch.empty(B, n, dtype=torch.float32, device=m.device)
    _lib.check(lib.icka_mask_additive(h, m.data_ptr(), m.stride(0) if B > 1 else max(n, m.stride(0)), out.data_ptr(),
                                      B, n, st), 'icka_mask_additive')
    return out


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, residual: Optional[torch.Tensor] = None,
           act: int = ACT_NONE, out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None,
           pre_act_out: Optional[torch.Tensor] = None, alg_k: Optional[int] = None) -> torch.Tensor:
    """out[M,N] = act(a[M,K] . w[N,K]^T + bias) (+ residual).  a/w fp32 -> FFMA path, bf16 -> tcgen05 path.

    ``a`` and ``w`` may be row-pitched 2-D views (stride(1) == 1).  ``pre_act_out`` (contiguous [M,N] in the
    operand dtype, GELU layers only) receives a . w^T + bias for the backward pass.  ``alg_k``: the contraction length of
    the ALGORITHM when the operands are split-precision (``split3``: K' = 3K) -- bookkeeping for the profiler only."""
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1]:
        raise RuntimeError(f'linear: bad shapes {tuple(a.shape)} x {tuple(w.shape)}')
    if a.dtype != w.dtype or a.dtype not in _DT:
        raise RuntimeError(f'linear: operand dtypes {a.dtype}/{w.dtype} must both be fp32 or both bf16')
    if a.stride(1) != 1 or w.stride(1) != 1:
        raise RuntimeError('linear: operands must be unit-stride along K')
    M, K = a.shape
    N = w.shape[0]
    lib, h, st = _ctx(a)
    out_dtype = out_dtype or (out.dtype if out is not None else a.dtype)
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    elif out.shape != (M, N) or out.stride(1) != 1 or out.dtype != out_dtype:
        raise RuntimeError('linear: bad `out`')
    if bias is not None:
        _need(bias, torch.float32, 'linear(bias)')
    if residual is not None:
        _need(residual, torch.float32, 'linear(residual)')
        if residual.shape != (M, N):
            raise RuntimeError('linear: residual shape mismatch')
    if pre_act_out is not None:
        _need(pre_act_out, a.dtype, 'linear(pre_act_out)')
        if pre_act_out.shape != (M, N):
            raise RuntimeError('linear: pre_act_out shape mismatch')
    _lib.check(lib.icka_linear_fwd_ex(h, a.data_ptr(), _ld(a, K), w.data_ptr(), _ld(w, K), _p(bias), _p(residual),
                                      out.data_ptr(), _ld(out, N), _p(pre_act_out), _DT[a.dtype], _DT[out_dtype],
                                      M, N, K, act, st), 'icka_linear_fwd_ex')
    return out


def linear_ln(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], residual: Optional[torch.Tensor],
              gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, want_bf16: bool = False):
    """LayerNorm(a . w^T + bias + residual) -> (y32 [M,N], y16 | None): the dense layer with the normalisation
    fused into the GEMM epilogue (bf16 operands) -- BertSelfOutput / BertOutput in one launch."""
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1] or a.dtype != w.dtype or a.dtype not in _DT:
        raise RuntimeError(f'linear_ln: bad operands {tuple(a.shape)} {a.dtype} x {tuple(w.shape)} {w.dtype}')
    if a.stride(1) != 1 or w.stride(1) != 1:
        raise RuntimeError('linear_ln: operands must be unit-stride along K')
    M, K = a.shape
    N = w.shape[0]
    for t, n in ((gamma, 'gamma'), (beta, 'beta')):
        _need(t, torch.float32, f'linear_ln({n})')
    if bias is not None:
        _need(bias, torch.float32, 'linear_ln(bias)')
    if residual is not None:
        _need(residual, torch.float32, 'linear_ln(residual)')
        if residual.shape != (M, N):
            raise RuntimeError('linear_ln: residual shape mismatch')
    lib, h, st = _ctx(a)
    y32 = torch.empty(M, N, dtype=torch.float32, device=a.device)
    y16 = torch.empty(M, N, dtype=torch.bfloat16, device=a.device) if want_bf16 else None
    _lib.check(lib.icka_linear_ln_fwd(h, a.data_ptr(), _ld(a, K), w.data_ptr(), _ld(w, K), _p(bias), _p(residual),
                                      gamma.data_ptr(), beta.data_ptr(), float(eps), y32.data_ptr(), _p(y16),
                                      _DT[a.dtype], M, N, K, st), 'icka_linear_ln_fwd')
    return y32, y16


def _ld(t: torch.Tensor, cols: int) -> int:
    """Row pitch of a 2-D view (a single-row tensor may report any stride)."""
    return t.stride(0) if t.shape[0] > 1 else max(cols, t.stride(0))


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, *, residual: Optional[torch.Tensor] = None,
                 gelu_pre: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """dX[M,K] = (dy[M,N] . w[N,K]) * gelu'(gelu_pre[M,K]) + residual[M,K]; w exactly as nn.Linear stores it."""
    if dy.dim() != 2 or w.dim() != 2 or dy.shape[1] != w.shape[0]:
        raise RuntimeError(f'linear_dgrad: bad shapes {tuple(dy.shape)} x {tuple(w.shape)}')
    if dy.dtype != w.dtype or dy.dtype not in _DT or dy.stride(1) != 1 or w.stride(1) != 1:
        raise RuntimeError('linear_dgrad: operands must share dtype (fp32 or bf16) and be unit-stride')
    M, N = dy.shape
    K = w.shape[1]
    lib, h, st = _ctx(dy)
    out_dtype = out_dtype or dy.dtype
    dx = torch.empty(M, K, dtype=out_dtype, device=dy.device)
    if residual is not None:
        _need(residual, torch.float32, 'linear_dgrad(residual)')
        if residual.shape != (M, K):
            raise RuntimeError('linear_dgrad: residual shape mismatch')
    if gelu_pre is not None:
        if gelu_pre.dtype != dy.dtype or gelu_pre.shape != (M, K) or gelu_pre.stride(1) != 1:
            raise RuntimeError('linear_dgrad: bad gelu_pre')
    _lib.check(lib.icka_linear_dgrad(h, dy.data_ptr(), _ld(dy, N), w.data_ptr(), _ld(w, K), _p(residual),
                                     _p(gelu_pre), _ld(gelu_pre, K) if gelu_pre is not None else 0, dx.data_ptr(), K,
                                     _DT[dy.dtype], _DT[out_dtype], M, N, K, st), 'icka_linear_dgrad')
    return dx


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> torch.Tensor:
    """dW[N,K] (+)= dy[M,N]^T . x[M,K] (fp32)."""
    if dy.dim() != 2 or x.dim() != 2 or dy.shape[0] != x.shape[0]:
        raise RuntimeError(f'linear_wgrad: bad shapes {tuple(dy.shape)} x {tuple(x.shape)}')
    if dy.dtype != x.dtype or dy.dtype not in _DT or dy.stride(1) != 1 or x.stride(1) != 1:
        raise RuntimeError('linear_wgrad: operands must share dtype (fp32 or bf16) and be unit-stride')
    M, N = dy.shape
    K = x.shape[1]
    lib, h, st = _ctx(dy)
    if out is None:
        out = torch.empty(N, K, dtype=torch.float32, device=dy.device)
        accumulate = False
    else:
        _need(out, torch.float32, 'linear_wgrad(out)')
        if out.shape != (N, K):
            raise RuntimeError('linear_wgrad: bad `out`')
    _lib.check(lib.icka_linear_wgrad(h, dy.data_ptr(), _ld(dy, N), x.data_ptr(), _ld(x, K), out.data_ptr(),
                                     _DT[dy.dtype], M, N, K, int(accumulate), st), 'icka_linear_wgrad')
    return out


def act_bwd(dy: torch.Tensor, ref: torch.Tensor, act: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = dy * act'(ref) element-wise; ref = pre-activation (gelu / relu / swish) or the output (tanh).  ``out`` may be dy."""
    if dy.dtype not in _DT or ref.dtype != dy.dtype or dy.shape != ref.shape or not dy.is_contiguous() or not ref.is_contiguous():
        raise RuntimeError('act_bwd: dy / ref must be contiguous tensors of one shape and dtype (fp32 or bf16)')
    if out is None:
        out = torch.empty_like(dy)
    elif out.dtype != dy.dtype or out.shape != dy.shape or not out.is_contiguous():
        raise RuntimeError('act_bwd: bad `out`')
    lib, h, st = _ctx(dy)
    _lib.check(lib.icka_act_bwd(h, dy.data_ptr(), ref.data_ptr(), out.data_ptr(), _DT[dy.dtype], dy.numel(), int(act), st),
               'icka_act_bwd')
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a 2-D fp32/bf16 view -> fp32 [N] (bias gradient)."""
    if x.dim() != 2 or x.dtype not in _DT or x.stride(1) != 1:
        raise RuntimeError('colsum: need a unit-stride 2-D fp32/bf16 tensor')
    M, N = x.shape
    lib, h, st = _ctx(x)
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    _lib.check(lib.icka_colsum(h, x.data_ptr(), _ld(x, N), _DT[x.dtype], out.data_ptr(), M, N, 0, st), 'icka_colsum')
    return out


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, eps: float, *, want_f32: bool = True,
                  want_bf16: bool = False, want_dbias: bool = True):
    """Backward of layernorm(x): returns (dx32 | None, dx16 | None, dgamma, dbeta, colsum(dx) | None)."""
    _need(dy, torch.float32, 'layernorm_bwd(dy)')
    _need(x, torch.float32, 'layernorm_bwd(x)')
    _need(gamma, torch.float32, 'layernorm_bwd(gamma)')
    M, N = x.shape
    lib, h, st = _ctx(x)
    dx32 = torch.empty_like(x) if want_f32 else None
    dx16 = torch.empty_like(x, dtype=torch.bfloat16) if want_bf16 else None
    acc = torch.zeros(3, N, dtype=torch.float32, device=x.device)
    _lib.check(lib.icka_layernorm_bwd(h, dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), float(eps), _p(dx32), _p(dx16),
                                      acc[0].data_ptr(), acc[1].data_ptr(), acc[2].data_ptr() if want_dbias else None,
                                      M, N, st), 'icka_layernorm_bwd')
    return dx32, dx16, acc[0], acc[1], (acc[2] if want_dbias else None)


def cross_attn_core_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask_add: Optional[torch.Tensor],
                        dctx: torch.Tensor, B: int, Sq: int, Skv: int, nh: int, d: int,
                        ctx: Optional[torch.Tensor] = None, p_drop: float = 0.0, seed: int = 0):
    """Returns (dq [B*Sq, nh*d], dkv [B*Skv, 2*nh*d]) in the dtype of q.  ``ctx`` = the forward output
    (enables the tensor-core kernel for bf16)."""
    if q.dtype not in _DT or k.dtype != q.dtype or v.dtype != q.dtype or dctx.dtype != q.dtype:
        raise RuntimeError('cross_attn_core_bwd: q/k/v/dctx must share dtype fp32 or bf16')
    if q.stride(1) != 1 or k.stride(1) != 1 or v.stride(1) != 1 or dctx.stride(1) != 1 or k.stride(0) != v.stride(0):
        raise RuntimeError('cross_attn_core_bwd: bad strides')
    H = nh * d
    lib, h, st = _ctx(q)
    dq = torch.empty(B * Sq, H, dtype=q.dtype, device=q.device)
    dkv = torch.empty(B * Skv, 2 * H, dtype=q.dtype, device=q.device)
    dk, dv = dkv[:, :H], dkv[:, H:]
    if ctx is not None and (ctx.dtype != q.dtype or ctx.stride(1) != 1):
        raise RuntimeError('cross_attn_core_bwd: bad ctx')
    _lib.check(lib.icka_cross_attn_core_bwd_drop(h, q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                            _p(mask_add), _p(ctx), ctx.stride(0) if ctx is not None else 0,
                                            dctx.data_ptr(), dctx.stride(0), dq.data_ptr(), H,
                                            dk.data_ptr(), dv.data_ptr(), 2 * H, _DT[q.dtype], B, Sq, Skv, nh, d,
                                            float(p_drop), int(seed), st),
               'icka_cross_attn_core_bwd')
    return dq, dkv


def gate_blend_bwd(dout, fused, tok, gate, ln_w, ln_b, ln_eps: float, w_fold, want_dtok: bool = True):
    """Returns (dfused, dtok | None, d_ln_w, d_ln_b, d_w_fold, d_c_fold)."""
    for t, n in ((dout, 'dout'), (fused, 'fused'), (tok, 'tok'), (gate, 'gate')):
        _need(t, torch.float32, f'gate_blend_bwd({n})')
    B, S, H = fused.shape
    lib, h, st = _ctx(fused)
    dfused = torch.empty_like(fused)
    dtok = torch.empty_like(tok) if want_dtok else None
    acc = torch.zeros(3 * H + 1, dtype=torch.float32, device=fused.device)
    d_ln_w, d_ln_b, d_w_fold, d_c_fold = acc[:H], acc[H:2 * H], acc[2 * H:3 * H], acc[3 * H:]
    _lib.check(lib.icka_gate_blend_bwd(h, dout.data_ptr(), fused.data_ptr(), tok.data_ptr(), gate.data_ptr(),
                                       ln_w.data_ptr(), ln_b.data_ptr(), float(ln_eps), w_fold.data_ptr(),
                                       dfused.data_ptr(), _p(dtok), d_ln_w.data_ptr(), d_ln_b.data_ptr(),
                                       d_w_fold.data_ptr(), d_c_fold.data_ptr(), B, S, H, st), 'icka_gate_blend_bwd')
    return dfused, dtok, d_ln_w, d_ln_b, d_w_fold, d_c_fold


def gate_fold_bwd(wp, bp, wa, d_w_fold, d_c_fold):
    """Returns (dWp [H,H], dbp [H], dwa [H], dba [1])."""
    H = wp.shape[0]
    lib, h, st = _ctx(wp)
    f32 = dict(dtype=torch.float32, device=wp.device)
    dwp, dbp, dwa, dba = torch.empty(H, H, **f32), torch.empty(H, **f32), torch.empty(H, **f32), torch.empty(1, **f32)
    _lib.check(lib.icka_gate_fold_bwd(h, wp.data_ptr(), bp.data_ptr(), wa.data_ptr(), d_w_fold.data_ptr(),
                                      d_c_fold.data_ptr(), dwp.data_ptr(), dbp.data_ptr(), dwa.data_ptr(),
                                      dba.data_ptr(), H, st), 'icka_gate_fold_bwd')
    return dwp, dbp, dwa, dba


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, want_f32: bool = True,
              want_bf16: bool = False) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    _need(x, torch.float32, 'layernorm(x)')
    _need(gamma, torch.float32, 'layernorm(gamma)')
    _need(beta, torch.float32, 'layernorm(beta)')
    M, N = x.shape
    lib, h, st = _ctx(x)
    y32 = torch.empty_like(x) if want_f32 else None
    y16 = torch.empty_like(x, dtype=torch.bfloat16) if want_bf16 else None
    _lib.check(lib.icka_layernorm_fwd(h, x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps), _p(y32),
                                      _p(y16), M, N, st), 'icka_layernorm_fwd')
    return y32, y16


def cross_attn_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask_add: Optional[torch.Tensor], B: int,
                    Sq: int, Skv: int, nh: int, d: int, p_drop: float = 0.0, seed: int = 0) -> torch.Tensor:
    """q [B*Sq, nh*d], k/v [B*Skv, nh*d] row-pitched views (k, v may be halves of one [K|V] buffer).
    ``p_drop`` > 0 (training): dropout on the probabilities with the Philox mask of ``seed``."""
    if q.dtype not in _DT or k.dtype != q.dtype or v.dtype != q.dtype:
        raise RuntimeError('cross_attn_core: q/k/v must share dtype fp32 or bf16')
    if q.stride(1) != 1 or k.stride(1) != 1 or v.stride(1) != 1 or k.stride(0) != v.stride(0):
        raise RuntimeError('cross_attn_core: bad strides')
    if mask_add is not None:
        _need(mask_add, torch.float32, 'cross_attn_core(mask_add)')
        if mask_add.shape != (B, Skv):
            raise RuntimeError(f'cross_attn_core: mask_add must be [B, Skv], got {tuple(mask_add.shape)}')
    lib, h, st = _ctx(q)
    ctx = torch.empty(B * Sq, nh * d, dtype=q.dtype, device=q.device)
    _lib.check(lib.icka_cross_attn_core_fwd_drop(h, q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                                 _p(mask_add), ctx.data_ptr(), ctx.stride(0), _DT[q.dtype], B, Sq, Skv,
                                                 nh, d, float(p_drop), int(seed), st), 'icka_cross_attn_core_fwd')
    return ctx


def i2t_pool(u: torch.Tensor, x: torch.Tensor, mask_add: Optional[torch.Tensor], B: int, S: int, H: int, nh: int):
    """Folded single-query attention pool: u [B, nh*H] bf16, x [B*S, H] bf16 -> xbar [B, nh*H] bf16."""
    _need(u, torch.bfloat16, 'i2t_pool(u)')
    _need(x, torch.bfloat16, 'i2t_pool(x)')
    if u.shape != (B, nh * H) or x.shape != (B * S, H):
        raise RuntimeError(f'i2t_pool: bad shapes {tuple(u.shape)} {tuple(x.shape)}')
    if mask_add is not None:
        _need(mask_add, torch.float32, 'i2t_pool(mask_add)')
    lib, h, st = _ctx(u)
    xbar = torch.empty_like(u)
    _lib.check(lib.icka_i2t_pool_fwd(h, u.data_ptr(), x.data_ptr(), _p(mask_add), xbar.data_ptr(), B, S, H, nh, st),
               'icka_i2t_pool_fwd')
    return xbar


def gate_fold(wp: torch.Tensor, bp: torch.Tensor, wa: torch.Tensor, ba: torch.Tensor):
    for t, n in ((wp, 'Wp'), (bp, 'bp'), (wa, 'wa'), (ba, 'ba')):
        _need(t, torch.float32, f'gate_fold({n})')
    H = wp.shape[0]
    lib, h, st = _ctx(wp)
    w_fold = torch.empty(H, dtype=torch.float32, device=wp.device)
    c_fold = torch.empty(1, dtype=torch.float32, device=wp.device)
    _lib.check(lib.icka_gate_fold(h, wp.data_ptr(), bp.data_ptr(), wa.data_ptr(), ba.data_ptr(), w_fold.data_ptr(),
                                  c_fold.data_ptr(), H, st), 'icka_gate_fold')
    return w_fold, c_fold


def gate_blend(fused: torch.Tensor, tok: torch.Tensor, ln_w, ln_b, ln_eps: float, w_fold, c_fold):
    _need(fused, torch.float32, 'gate_blend(fused)')
    _need(tok, torch.float32, 'gate_blend(tok)')
    if fused.shape != tok.shape or fused.dim() != 3:
        raise RuntimeError('gate_blend: fused/tok must both be [B, S, H]')
    B, S, H = fused.shape
    lib, h, st = _ctx(fused)
    out = torch.empty_like(fused)
    gate = torch.empty(B, dtype=torch.float32, device=fused.device)
    _lib.check(lib.icka_gate_blend_fwd(h, fused.data_ptr(), tok.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(),
                                       float(ln_eps), w_fold.data_ptr(), c_fold.data_ptr(), out.data_ptr(),
                                       gate.data_ptr(), B, S, H, st), 'icka_gate_blend_fwd')
    return out, gate


def ln_gate_blend(pre, ln2_w, ln2_b, ln2_eps: float, tok, lng_w, lng_b, lng_eps: float, w_fold, c_fold, *,
                  want_fused_f32: bool = False, want_fused_bf16: bool = True):
    """fused = LN(pre); result = gate-blend(fused, tok).  Returns (result, gate, fused32 | None, fused16 | None)."""
    _need(pre, torch.float32, 'ln_gate_blend(pre)')
    _need(tok, torch.float32, 'ln_gate_blend(tok)')
    if pre.shape != tok.shape or pre.dim() != 3:
        raise RuntimeError('ln_gate_blend: pre/tok must both be [B, S, H]')
    B, S, H = pre.shape
    lib, h, st = _ctx(pre)
    out = torch.empty_like(pre)
    gate = torch.empty(B, dtype=torch.float32, device=pre.device)
    f32 = torch.empty_like(pre) if want_fused_f32 else None
    f16 = torch.empty(B * S, H, dtype=torch.bfloat16, device=pre.device) if want_fused_bf16 else None
    _lib.check(lib.icka_ln_gate_blend_fwd(h, pre.data_ptr(), ln2_w.data_ptr(), ln2_b.data_ptr(), float(ln2_eps),
                                          tok.data_ptr(), lng_w.data_ptr(), lng_b.data_ptr(), float(lng_eps),
                                          w_fold.data_ptr(), c_fold.data_ptr(), out.data_ptr(), _p(f32), _p(f16),
                                          gate.data_ptr(), B, S, H, st), 'icka_ln_gate_blend_fwd')
    return out, gate, f32, f16


def viterbi(emissions: torch.Tensor, mask_u8: Optional[torch.Tensor], start, end, trans):
    """Batch-first fp32 emissions [B,S,T]; returns device tensors (tags [B,S] int32 with -1 pad, lens [B] int32)."""
    _need(emissions, torch.float32, 'viterbi(emissions)')
    B, S, T = emissions.shape
    if mask_u8 is not None:
        _need(mask_u8, torch.uint8, 'viterbi(mask)')
    for t, n in ((start, 'start'), (end, 'end'), (trans, 'trans')):
        _need(t, torch.float32, f'viterbi({n})')
    lib, h, st = _ctx(emissions)
    tags = torch.empty(B, S, dtype=torch.int32, device=emissions.device)
    lens = torch.empty(B, dtype=torch.int32, device=emissions.device)
    _lib.check(lib.icka_viterbi_decode(h, emissions.data_ptr(), _p(mask_u8), start.data_ptr(), end.data_ptr(),
                                       trans.data_ptr(), tags.data_ptr(), lens.data_ptr(), B, S, T, st),
               'icka_viterbi_decode')
    return tags, lens


def crf_llh(emissions: torch.Tensor, tags_i64: torch.Tensor, mask_u8: Optional[torch.Tensor], start, end, trans):
    _need(emissions, torch.float32, 'crf_llh(emissions)')
    _need(tags_i64, torch.int64, 'crf_llh(tags)')
    B, S, T = emissions.shape
    if mask_u8 is not None:
        _need(mask_u8, torch.uint8, 'crf_llh(mask)')
    lib, h, st = _ctx(emissions)
    llh = torch.empty(B, dtype=torch.float32, device=emissions.device)
    _lib.check(lib.icka_crf_llh_fwd(h, emissions.data_ptr(), tags_i64.data_ptr(), _p(mask_u8), start.data_ptr(),
                                    end.data_ptr(), trans.data_ptr(), llh.data_ptr(), B, S, T, st),
               'icka_crf_llh_fwd')
    return llh


def crf_llh_bwd(emissions: torch.Tensor, tags_i64: torch.Tensor, mask_u8: Optional[torch.Tensor], start, end, trans,
                w: torch.Tensor):
    """Gradient of sum_b w[b] * llh[b]: returns (d_emissions [B,S,T], d_start [T], d_end [T], d_trans [T,T])."""
    _need(emissions, torch.float32, 'crf_llh_bwd(emissions)')
    _need(tags_i64, torch.int64, 'crf_llh_bwd(tags)')
    _need(w, torch.float32, 'crf_llh_bwd(w)')
    B, S, T = emissions.shape
    if mask_u8 is not None:
        _need(mask_u8, torch.uint8, 'crf_llh_bwd(mask)')
    lib, h, st = _ctx(emissions)
    de = torch.empty_like(emissions)
    acc = torch.zeros(T * T + 2 * T, dtype=torch.float32, device=emissions.device)
    d_start, d_end, d_trans = acc[:T], acc[T:2 * T], acc[2 * T:].view(T, T)
    _lib.check(lib.icka_crf_llh_bwd(h, emissions.data_ptr(), tags_i64.data_ptr(), _p(mask_u8), start.data_ptr(),
                                    end.data_ptr(), trans.data_ptr(), w.data_ptr(), de.data_ptr(), d_start.data_ptr(),
                                    d_end.data_ptr(), d_trans.data_ptr(), B, S, T, st), 'icka_crf_llh_bwd')
    return de, d_start, d_end, d_trans


def dropout(x: torch.Tensor, p_drop: float, seed: int, *, residual: Optional[torch.Tensor] = None,
            out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """y = x * keep / (1 - p) (+ residual); keep = Philox(seed, element index).  x fp32/bf16 contiguous, numel % 4 == 0."""
    if x.dtype not in _DT or not x.is_contiguous():
        raise RuntimeError('dropout: need a contiguous fp32/bf16 tensor')
    out_dtype = out_dtype or x.dtype
    if residual is not None:
        _need(residual, torch.float32, 'dropout(residual)')
        if residual.shape != x.shape:
            raise RuntimeError('dropout: residual shape mismatch')
    lib, h, st = _ctx(x)
    y = torch.empty_like(x, dtype=out_dtype)
    _lib.check(lib.icka_dropout_fwd(h, x.data_ptr(), _DT[x.dtype], _p(residual), y.data_ptr(), _DT[out_dtype], x.numel(),
                                    float(p_drop), int(seed), st), 'icka_dropout_fwd')
    return y


def dropout_mask(shape, p_drop: float, seed: int, device, attention: bool = False) -> torch.Tensor:
    """Test helper: the u8 keep mask the dropout kernels regenerate.  ``attention``: shape = (B, nh, Sq, Skv)."""
    m = torch.empty(shape, dtype=torch.uint8, device=device)
    d = m.device.index if m.device.index is not None else torch.cuda.current_device()
    lib, h, st = _lib.load(), _lib.handle(d), torch.cuda.current_stream(d).cuda_stream
    if attention:
        rows, skv = m.numel() // shape[-1], shape[-1]
        _lib.check(lib.icka_dropout_mask(h, m.data_ptr(), rows, skv, 1, float(p_drop), int(seed), st), 'icka_dropout_mask')
    else:
        _lib.check(lib.icka_dropout_mask(h, m.data_ptr(), m.numel(), 1, 0, float(p_drop), int(seed), st), 'icka_dropout_mask')
    return m


# ---- emission head (BiLSTM + classifier) -------------------------------------------------------

def add_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _need(a, torch.float32, 'add_f32(a)')
    _need(b, torch.float32, 'add_f32(b)')
    if a.shape != b.shape:
        raise RuntimeError('add_f32: shape mismatch')
    lib, h, st = _ctx(a)
    out = torch.empty_like(a)
    _lib.check(lib.icka_add_f32(h, a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), st), 'icka_add_f32')
    return out


def lstm_rec_workspace_bytes(B: int, H: int) -> int:
    n = int(_lib.load().icka_lstm_rec_workspace_bytes(int(B), int(H)))
    if n < 0:
        raise RuntimeError(f'lstm_rec: hidden size {H} is not supported by the persistent kernel')
    return n


def lstm_rec_variant(B: int) -> int:
    """1 = single-CTA kernel (24-unit slices, small batches), 2 = CTA-pair kernel (48-unit slices)."""
    return int(_lib.load().icka_lstm_rec_variant(int(B)))


def lstm_rec(gx: torch.Tensor, w_hh_perm: torch.Tensor, B: int, S: int, H: int, *, variant: int,
             workspace: Optional[torch.Tensor] = None, want_state: bool = False):
    """gx [S*B, 8H] bf16 (time-major rows, slice-ordered columns), w_hh_perm [8H, H] bf16 -> y [S,B,2H] bf16
    time-major (, h_n, c_n [2,B,H] fp32)."""
    _need(gx, torch.bfloat16, 'lstm_rec(gx)')
    _need(w_hh_perm, torch.bfloat16, 'lstm_rec(w_hh_perm)')
    if gx.shape != (B * S, 8 * H) or w_hh_perm.shape != (8 * H, H):
        raise RuntimeError(f'lstm_rec: bad shapes {tuple(gx.shape)} / {tuple(w_hh_perm.shape)} for B={B} S={S} H={H}')
    lib, h, st = _ctx(gx)
    need = lstm_rec_workspace_bytes(B, H)
    if workspace is None:
        workspace = torch.empty(need + 1024, dtype=torch.uint8, device=gx.device)
    off = (-workspace.data_ptr()) % 1024
    if workspace.numel() - off < need:
        raise RuntimeError(f'lstm_rec: workspace of {workspace.numel()} B, need {need} + alignment')
    y = torch.empty(S, B, 2 * H, dtype=torch.bfloat16, device=gx.device)
    h_n = torch.empty(2, B, H, dtype=torch.float32, device=gx.device) if want_state else None
    c_n = torch.empty(2, B, H, dtype=torch.float32, device=gx.device) if want_state else None
    _lib.check(lib.icka_lstm_rec_fwd(h, gx.data_ptr(), w_hh_perm.data_ptr(), workspace.data_ptr() + off, need,
                                     y.data_ptr(), _p(h_n), _p(c_n), B, S, H, int(variant), st),
               'icka_lstm_rec_fwd')
    return (y, h_n, c_n) if want_state else y


def lstm_cell(gates_h: Optional[torch.Tensor], gx: torch.Tensor, c: torch.Tensor, h_out: torch.Tensor,
              y: Optional[torch.Tensor], h_f32: Optional[torch.Tensor] = None) -> None:
    """One step of the per-step path.  gx [B,4H] / y [B,H] may be row-pitched views; c [B,H] fp32 in place."""
    B, H = c.shape
    _need(c, torch.float32, 'lstm_cell(c)')
    if gx.dtype not in _DT or gx.shape != (B, 4 * H) or gx.stride(1) != 1:
        raise RuntimeError('lstm_cell: bad gx')
    _need(h_out, gx.dtype, 'lstm_cell(h_out)')
    if gates_h is not None:
        _need(gates_h, torch.float32, 'lstm_cell(gates_h)')
        if gates_h.shape != (B, 4 * H):
            raise RuntimeError('lstm_cell: bad gates_h')
    if y is not None and (y.dtype != gx.dtype or y.shape != (B, H) or y.stride(1) != 1):
        raise RuntimeError('lstm_cell: bad y')
    if h_f32 is not None:
        _need(h_f32, torch.float32, 'lstm_cell(h_f32)')
    lib, h, st = _ctx(c)
    _lib.check(lib.icka_lstm_cell_fwd(h, _p(gates_h), gx.data_ptr(), _ld(gx, 4 * H), c.data_ptr(), h_out.data_ptr(),
                                      _p(y), 0 if y is None else _ld(y, H), _p(h_f32), _DT[gx.dtype], B, H, st),
               'icka_lstm_cell_fwd')


def cast_bf16_time_major(x: torch.Tensor) -> torch.Tensor:
    """x [B,S,H] fp32 / bf16 contiguous -> [S,B,H] bf16."""
    if x.dim() != 3 or x.dtype not in _DT or not x.is_contiguous():
        raise RuntimeError('cast_bf16_time_major: x must be a contiguous [B,S,H] fp32 / bf16 tensor')
    B, S, H = x.shape
    lib, h, st = _ctx(x)
    y = torch.empty(S, B, H, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.icka_cast_bf16_time_major(h, x.data_ptr(), y.data_ptr(), _DT[x.dtype], B, S, H, st),
               'icka_cast_bf16_time_major')
    return y


def emission_head(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, time_major_S: int = 0) -> torch.Tensor:
    """x [M,K] (fp32 or bf16, unit stride along K) . w[T,K]^T (fp32) + bias -> [M,T] fp32.  ``time_major_S`` = S:
    the rows of x are t*B + b and the output rows b*S + t."""
    if x.dim() != 2 or x.dtype not in _DT or x.stride(1) != 1:
        raise RuntimeError('emission_head: bad x')
    _need(w, torch.float32, 'emission_head(w)')
    _need(bias, torch.float32, 'emission_head(bias)')
    M, K = x.shape
    T = w.shape[0]
    if w.shape != (T, K) or bias.shape != (T,):
        raise RuntimeError('emission_head: bad weight / bias shape')
    lib, h, st = _ctx(x)
    out = torch.empty(M, T, dtype=torch.float32, device=x.device)
    _lib.check(lib.icka_emission_head_fwd(h, x.data_ptr(), _ld(x, K), w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                          _DT[x.dtype], M, K, T, int(time_major_S), st), 'icka_emission_head_fwd')
    return out


def emission_head_bwd(dout: torch.Tensor, x: torch.Tensor, w: torch.Tensor, want_dx: bool = True, want_dw: bool = True,
                      time_major_S: int = 0):
    """Classifier backward: dout [M,T] fp32, x [M,K] fp32 / bf16, w [T,K] fp32 -> (dx [M,K] fp32 | None, dw [T,K] fp32 | None).
    ``time_major_S`` = S: the rows of x / dx are t*B + b while those of dout are b*S + t (as ``emission_head`` pairs them)."""
    _need(dout, torch.float32, 'emission_head_bwd(dout)')
    _need(w, torch.float32, 'emission_head_bwd(w)')
    if x.dim() != 2 or x.dtype not in _DT or x.stride(1) != 1:
        raise RuntimeError('emission_head_bwd: bad x')
    M, K = x.shape
    T = w.shape[0]
    if dout.shape != (M, T) or w.shape != (T, K):
        raise RuntimeError('emission_head_bwd: shape mismatch')
    lib, h, st = _ctx(x)
    dx = torch.empty(M, K, dtype=torch.float32, device=x.device) if want_dx else None
    dw = torch.empty(T, K, dtype=torch.float32, device=x.device) if want_dw else None
    _lib.check(lib.icka_emission_head_bwd(h, dout.data_ptr(), x.data_ptr(), _ld(x, K), w.data_ptr(), _p(dx), _p(dw),
                                          _DT[x.dtype], M, K, T, 0, int(time_major_S), st), 'icka_emission_head_bwd')
    return dx, dw


def region_tail(x: torch.Tensor, att_size: int, *, want_fc: bool = True, want_att: bool = True,
                rows_dtype: Optional[torch.dtype] = None):
    """layer4 output x [B,C,g,g] fp32 -> (fc [B,C] | None, att [B,C,a,a] fp32 | None, rows [B, a*a, C] | None)."""
    _need(x, torch.float32, 'region_tail(x)')
    if x.dim() != 4 or x.shape[2] != x.shape[3]:
        raise RuntimeError(f'region_tail: expected [B,C,g,g], got {tuple(x.shape)}')
    B, C, g, _ = x.shape
    a = int(att_size)
    lib, h, st = _ctx(x)
    fc = torch.empty(B, C, dtype=torch.float32, device=x.device) if want_fc else None
    att = torch.empty(B, C, a, a, dtype=torch.float32, device=x.device) if want_att else None
    rows = torch.empty(B, a * a, C, dtype=rows_dtype, device=x.device) if rows_dtype is not None else None
    _lib.check(lib.icka_region_tail_fwd(h, x.data_ptr(), _p(fc), _p(att), _p(rows),
                                        _DT[rows_dtype] if rows_dtype is not None else F32, B, C, g, a, st),
               'icka_region_tail_fwd')
    return fc, att, rows


def lstm_cell_fwd_save(gates_h, gx, c_prev, c_out, acts, h_out, y_op, y32) -> None:
    """Training forward step (keeps activations); gx [B,4H], y_op / y32 [B,H] may be row-pitched views."""
    B, H = c_out.shape
    lib, h, st = _ctx(c_out)
    _lib.check(lib.icka_lstm_cell_fwd_save(h, _p(gates_h), gx.data_ptr(), _ld(gx, 4 * H), _p(c_prev), c_out.data_ptr(),
                                           acts.data_ptr(), h_out.data_ptr(), y_op.data_ptr(), _ld(y_op, H),
                                           y32.data_ptr(), _ld(y32, H), _DT[gx.dtype], B, H, st),
               'icka_lstm_cell_fwd_save')


def lstm_cell_bwd(dy, dh_rec, dc, acts, c_prev, c_new, dpre) -> None:
    """One BPTT step; dy [B,H] fp32 and dpre [B,4H] may be row-pitched views, dc is updated in place."""
    B, H = dc.shape
    lib, h, st = _ctx(dc)
    _lib.check(lib.icka_lstm_cell_bwd(h, dy.data_ptr(), _ld(dy, H), _p(dh_rec), dc.data_ptr(), acts.data_ptr(), _p(c_prev),
                                      c_new.data_ptr(), dpre.data_ptr(), _ld(dpre, 4 * H), _DT[dpre.dtype], B, H, st),
               'icka_lstm_cell_bwd')
