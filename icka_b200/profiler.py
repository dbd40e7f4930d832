"""CUDA-event timer around the C-ABI ops (used by bench.py for the per-kernel roofline numbers).

Inside ``with KernelTimer() as kt:`` every call of an ``icka_b200.ops`` function is bracketed by a pair of
CUDA events recorded on torch's current stream -- the stream the kernel is launched on -- and tagged
with its algorithmic work (FLOPs for the GEMMs, bytes for the HBM-bound kernels; formulas in DESIGN.md).
``summary()`` synchronises once and returns per-op launch counts, mean durations and achieved rates.
"""
from __future__ import annotations

from collections import defaultdict

import torch

from . import ops

_ESZ = {torch.float32: 4, torch.bfloat16: 2, torch.uint8: 1, torch.int32: 4, torch.int64: 8}


def _work_linear(a, w, bias, residual=None, act=0, out_dtype=None, out=None, pre_act_out=None, alg_k=None):
    M, K = a.shape
    N = w.shape[0]
    od = out_dtype or (out.dtype if out is not None else a.dtype)
    nbytes = M * K * _ESZ[a.dtype] + N * K * _ESZ[w.dtype] + M * N * _ESZ[od] + (M * N * 4 if residual is not None else 0)
    # split-precision operands (ops.split3): the algorithm's FLOPs are 2 M N K, the kernel executes 3x that
    return 'linear_bf16_tcgen05' if a.dtype == torch.bfloat16 else 'linear_fp32_ffma', 2.0 * M * N * (alg_k or K), nbytes


def _work(name, args, kwargs):
    if name == 'linear':
        return _work_linear(*args, **kwargs)
    if name == 'linear_ln':
        a, w = args[0], args[1]
        M, K = a.shape
        N = w.shape[0]
        nb = M * K * _ESZ[a.dtype] + N * K * _ESZ[w.dtype] + M * N * 8 + (M * N * 2 if kwargs.get('want_bf16') else 0)
        # GEMM + (split-K reduce +) LayerNorm in one op: its own row, so that the GEMM roofline row holds GEMM launches only
        return ('linear_ln_bf16_tcgen05' if a.dtype == torch.bfloat16 else 'linear_ln_fp32_ffma'), 2.0 * M * N * K, nb
    if name == 'cast_bf16':
        return name, 0.0, args[0].numel() * 6
    if name == 'region_rows':
        return name, 0.0, args[0].numel() * (4 + _ESZ[args[1]])
    if name == 'layernorm':
        x = args[0]
        nb = x.numel() * 4 * (1 + (1 if kwargs.get('want_f32', True) else 0)) + (x.numel() * 2 if kwargs.get('want_bf16') else 0)
        return name, 0.0, nb
    if name == 'cross_attn_core':
        q, k, v, mask, B, Sq, Skv, nh, d = args
        es = _ESZ[q.dtype]
        return name, 4.0 * B * nh * Sq * Skv * d, (2 * B * Sq * nh * d + 2 * B * Skv * nh * d) * es
    if name == 'i2t_pool':
        u, x, mask, B, S, H, nh = args
        return name, 4.0 * B * nh * S * H, (x.numel() + 2 * u.numel()) * 2
    if name == 'gate_blend':
        return name, 0.0, args[0].numel() * 12
    if name == 'ln_gate_blend':
        n = args[0].numel()
        return name, 0.0, n * 4 * 3 + (n * 2 if kwargs.get('want_fused_bf16', True) else 0) + (n * 4 if kwargs.get('want_fused_f32') else 0)
    if name == 'gate_fold':
        return name, 0.0, args[0].numel() * 4
    if name == 'viterbi':
        B, S, T = args[0].shape
        return name, 0.0, B * (S * T * 4 + S + S * 4)
    if name == 'crf_llh':
        B, S, T = args[0].shape
        return name, 0.0, B * (S * T * 4 + S + S * 8)
    # ---- backward (training) ----
    if name == 'linear_dgrad':
        dy, w = args[0], args[1]
        M, N = dy.shape
        K = w.shape[1]
        od = kwargs.get('out_dtype') or dy.dtype
        nb = (M * N + N * K) * _ESZ[dy.dtype] + M * K * _ESZ[od]
        nb += M * K * 4 if kwargs.get('residual') is not None else 0
        nb += M * K * _ESZ[dy.dtype] if kwargs.get('gelu_pre') is not None else 0
        return ('dgrad_bf16_tcgen05' if dy.dtype == torch.bfloat16 else 'dgrad_fp32_ffma'), 2.0 * M * N * K, nb
    if name == 'linear_wgrad':
        dy, x = args[0], args[1]
        M, N = dy.shape
        K = x.shape[1]
        return (('wgrad_bf16_tcgen05' if dy.dtype == torch.bfloat16 else 'wgrad_fp32_ffma'), 2.0 * M * N * K,
                (M * N + M * K) * _ESZ[dy.dtype] + N * K * 4)
    if name == 'colsum':
        return name, 0.0, args[0].numel() * _ESZ[args[0].dtype]
    if name == 'layernorm_bwd':
        x = args[1]
        nb = x.numel() * 4 * (2 + (1 if kwargs.get('want_f32', True) else 0)) + (x.numel() * 2 if kwargs.get('want_bf16') else 0)
        return name, 0.0, nb
    if name == 'cross_attn_core_bwd':
        q, k, v, mask, dctx, B, Sq, Skv, nh, d = args[:10]
        es = _ESZ[q.dtype]
        return name, 10.0 * B * nh * Sq * Skv * d, (3 * B * Sq * nh * d + 4 * B * Skv * nh * d) * es
    if name == 'gate_blend_bwd':
        return name, 0.0, args[0].numel() * 4 * (6 if kwargs.get('want_dtok', True) else 5)
    if name == 'dropout':
        x = args[0]
        od = kwargs.get('out_dtype') or x.dtype
        return name, 0.0, x.numel() * (_ESZ[x.dtype] + _ESZ[od] + (4 if kwargs.get('residual') is not None else 0))
    if name == 'crf_llh_bwd':
        B, S, T = args[0].shape
        return name, 0.0, B * (2 * S * T * 4 + S + S * 8)
    return name, 0.0, 0


class KernelTimer:
    OPS = ('mask_additive', 'cast_bf16', 'region_rows', 'linear', 'linear_ln', 'layernorm', 'cross_attn_core', 'i2t_pool', 'gate_fold', 'gate_blend', 'ln_gate_blend',
           'viterbi', 'crf_llh', 'linear_dgrad', 'linear_wgrad', 'colsum', 'layernorm_bwd', 'cross_attn_core_bwd',
           'gate_blend_bwd', 'gate_fold_bwd', 'crf_llh_bwd', 'dropout')

    def __init__(self):
        self.records = []
        self.shapes = []
        self._saved = {}

    def __enter__(self):
        for name in self.OPS:
            fn = getattr(ops, name)
            self._saved[name] = fn
            setattr(ops, name, self._wrap(name, fn))
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(ops, name, fn)
        return False

    def _wrap(self, name, fn):
        def timed(*args, **kwargs):
            tag, flops, nbytes = _work(name, args, kwargs)
            if name in ('linear', 'linear_ln'):
                self.shapes.append((f'{args[0].shape[0]}x{args[1].shape[0]}x{args[0].shape[1]}', len(self.records)))
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*args, **kwargs)
            e.record()
            self.records.append((tag, s, e, flops, nbytes))
            return out
        return timed

    def summary(self):
        torch.cuda.synchronize()
        agg = defaultdict(lambda: dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        for tag, s, e, flops, nbytes in self.records:
            a = agg[tag]
            a['launches'] += 1
            a['ms'] += s.elapsed_time(e)
            a['flops'] += flops
            a['bytes'] += nbytes
        out = {}
        for tag, a in agg.items():
            sec = a['ms'] * 1e-3
            out[tag] = dict(launches=a['launches'], ms_total=a['ms'], ms_per_launch=a['ms'] / a['launches'],
                            tflops=(a['flops'] / sec / 1e12) if sec > 0 else 0.0,
                            gbs=(a['bytes'] / sec / 1e9) if sec > 0 else 0.0,
                            flops_per_launch=a['flops'] / a['launches'], bytes_per_launch=a['bytes'] / a['launches'])
        return out

    def gemm_shapes(self):
        """Per (M x N x K) GEMM shape: launches, mean ms per launch, TFLOP/s (call after summary())."""
        agg = defaultdict(lambda: [0, 0.0, 0.0])
        for shape, idx in self.shapes:
            tag, s, e, flops, nbytes = self.records[idx]
            a = agg[shape]
            a[0] += 1
            a[1] += s.elapsed_time(e)
            a[2] += flops
        return {k: dict(launches=v[0], ms_per_launch=v[1] / v[0], tflops=v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0)
                for k, v in agg.items()}
