"""Developer diagnostic for the tcgen05 GEMM: localises descriptor / swizzle / pipeline mistakes.

Runs small cases in increasing complexity and, on mismatch, prints where the errors sit
(row/column/k structure).  Not part of the test-suite."""
import math
import sys

import torch

sys.path.insert(0, '.')
from icka_b200 import ops  # noqa: E402

DEV = 'cuda:0'


def run(M, N, K, pattern='rand'):
    g = torch.Generator().manual_seed(1)
    if pattern == 'rand':
        a = torch.randn(M, K, generator=g).bfloat16()
        w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    elif pattern == 'ones':
        a = torch.ones(M, K).bfloat16()
        w = torch.ones(N, K).bfloat16()
    elif pattern == 'rowid':      # out[m, n] = m * (sum over k of 1/K) -> identifies row mapping
        a = (torch.arange(M).float().view(M, 1) * torch.ones(1, K) / K).bfloat16()
        w = torch.ones(N, K).bfloat16()
    elif pattern == 'colid':
        a = torch.ones(M, K).bfloat16()
        w = (torch.arange(N).float().view(N, 1) * torch.ones(1, K) / K).bfloat16()
    elif pattern == 'kid':        # a one-hot in k per row: out[m, n] = w[n, m % K]
        a = torch.zeros(M, K)
        a[torch.arange(M), torch.arange(M) % K] = 1
        a = a.bfloat16()
        w = torch.randn(N, K, generator=g).bfloat16()
    got = ops.linear(a.to(DEV), w.to(DEV), None, out_dtype=torch.float32)
    torch.cuda.synchronize()
    got = got.cpu()
    want = (a.double() @ w.double().t()).float()
    err = (got - want).abs()
    bad = err > 1e-2 * want.abs().clamp(min=1.0)
    print(f'M={M} N={N} K={K} {pattern}: max err {float(err.max()):.3e}, bad {int(bad.sum())}/{bad.numel()}', flush=True)
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print('   bad rows (first 16):', rows[:16].tolist(), ' count', len(rows))
        print('   bad cols (first 16):', cols[:16].tolist(), ' count', len(cols))
        print('   got[0,:8] ', got[0, :8].tolist())
        print('   want[0,:8]', want[0, :8].tolist())
        print('   got[1,:8] ', got[1, :8].tolist())
        print('   want[1,:8]', want[1, :8].tolist())
    return not bad.any()


if __name__ == '__main__':
    ok = True
    from icka_b200 import _lib
    if len(sys.argv) > 1:
        _lib.load().icka_set_gemm_mode(int(sys.argv[1]))
        print('gemm mode', sys.argv[1])
    for args in [(128, 128, 16, 'ones'), (128, 128, 16, 'rand'), (128, 128, 64, 'ones'), (128, 128, 64, 'rowid'),
                 (128, 128, 64, 'colid'), (128, 128, 64, 'kid'), (128, 128, 64, 'rand'), (128, 256, 64, 'rand'),
                 (128, 256, 128, 'rand'), (128, 256, 768, 'rand'), (256, 256, 64, 'rand'), (1024, 768, 768, 'rand'),
                 (128 * 300, 768, 768, 'rand'), (256, 256, 64, 'rowid'), (256, 256, 64, 'colid'), (512, 512, 128, 'rand'), (1000, 768, 3072, 'rand')]:
        ok = run(*args) and ok
    print('ALL OK' if ok else 'FAILURES')
    sys.exit(0 if ok else 1)
