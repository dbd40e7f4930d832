"""Viterbi decode timings (CUDA events) at the sweep batch sizes (developer tool)."""
import sys

import torch

sys.path.insert(0, '.')
from icka_b200 import ops, synth  # noqa: E402

DEV = 'cuda:0'
PEAK_GBS = 6547.5


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def main():
    sh = synth.STD
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024, 4096, 65536, 262144]
    for Bv in sizes:
        e = torch.randn(Bv, sh.S, sh.T, device=DEV)
        lens = synth.lengths(Bv, sh.S, torch.Generator().manual_seed(1))
        m_short = synth.prefix_mask(lens, sh.S).to(torch.uint8).to(DEV)
        m_full = torch.ones(Bv, sh.S, dtype=torch.uint8, device=DEV)
        st, en, tr = torch.randn(sh.T, device=DEV), torch.randn(sh.T, device=DEV), torch.randn(sh.T, sh.T, device=DEV)
        for nm, m in (('full-length', m_full), ('tweet-length', m_short)):
            t = timeit(lambda: ops.viterbi(e, m, st, en, tr))
            nb = Bv * 8320
            print(f'viterbi B={Bv:6d} {nm:13s} {t*1e6:9.1f} us  {Bv/t/1e6:8.2f} M sent/s  {nb/t/1e9:8.1f} GB/s '
                  f'algorithmic ({nb/t/1e9/PEAK_GBS*100:5.1f}% of HBM peak)')


if __name__ == '__main__':
    main()
