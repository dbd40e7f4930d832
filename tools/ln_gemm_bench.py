"""Dense + residual + LayerNorm: the cluster kernel (icka_linear_ln_fwd, ln mode 2) against the tcgen05 GEMM followed by
layernorm_kernel, on the shapes of a 1024-sentence step.  CUDA events, inputs larger than L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icka_b200 import _lib, ops                                    # noqa: E402

DEV = 'cuda:0'
lib = _lib.load()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for M, N, K in ((131072, 768, 768), (131072, 768, 3072), (32768, 768, 768), (131072, 1024, 1024)):
    g = torch.Generator(DEV).manual_seed(1)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).bfloat16()
    bias, gamma, beta = (torch.randn(N, device=DEV, generator=g) for _ in range(3))
    res = torch.randn(M, N, device=DEV, generator=g)

    def split():
        pre = ops.linear(a, w, bias, residual=res, out_dtype=torch.float32)
        return ops.layernorm(pre, gamma, beta, 1e-12, want_f32=True, want_bf16=True)

    def fused():
        return ops.linear_ln(a, w, bias, res, gamma, beta, 1e-12, want_bf16=True)

    ref32, ref16 = split()
    out = {}
    for mode, name in ((2, 'cluster'), (1, 'single_cta')):
        lib.icka_set_ln_mode(mode)
        y32, y16 = fused()
        torch.cuda.synchronize()
        err = float((y32 - ref32).abs().max())
        out[name] = (timed(fused), err)
    lib.icka_set_ln_mode(0)
    t_split = timed(split)
    t_gemm = timed(lambda: ops.linear(a, w, bias, residual=res, out_dtype=torch.float32))
    bytes_fused = M * K * 2 + M * N * (4 + 4 + 2)
    print(f'M={M} N={N} K={K}: gemm+ln {t_split * 1e3:.1f} us (gemm alone {t_gemm * 1e3:.1f}) | '
          + ' | '.join(f'{k} {v[0] * 1e3:.1f} us err {v[1]:.1e}' for k, v in out.items())
          + f' | cluster: {2 * M * N * K / out["cluster"][0] / 1e9:.0f} TF/s, {bytes_fused / out["cluster"][0] / 1e6:.0f} GB/s')
