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
           act: int = ACT_NONE, out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = act(a[M,K] . w[N,K]^T + bias) (+ residual).  a/w fp32 -> FFMA path, bf16 -> tcgen05 path.

    ``a`` and ``w`` may be row-pitched 2-D views (stride(1) == 1)."""
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
    _lib.check(lib.icka_linear_fwd(h, a.data_ptr(), a.stride(0) if M > 1 else max(K, a.stride(0)), w.data_ptr(),
                                   w.stride(0) if N > 1 else max(K, w.stride(0)), _p(bias), _p(residual),
                                   out.data_ptr(), out.stride(0) if M > 1 else max(N, out.stride(0)),
                                   _DT[a.dtype], _DT[out_dtype], M, N, K, act, st), 'icka_linear_fwd')
    return out


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
                    Sq: int, Skv: int, nh: int, d: int) -> torch.Tensor:
    """q [B*Sq, nh*d], k/v [B*Skv, nh*d] row-pitched views (k, v may be halves of one [K|V] buffer)."""
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
    _lib.check(lib.icka_cross_attn_core_fwd(h, q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                            _p(mask_add), ctx.data_ptr(), ctx.stride(0), _DT[q.dtype], B, Sq, Skv,
                                            nh, d, st), 'icka_cross_attn_core_fwd')
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
