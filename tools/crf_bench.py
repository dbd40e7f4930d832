"""CRF log-likelihood forward / backward timings (developer tool)."""
import sys

import torch

sys.path.insert(0, '.')
from icka_b200 import ops, synth  # noqa: E402

DEV = 'cuda:0'


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
    sizes = [int(a) for a in sys.argv[1:]] or [128, 1024, 8192]
    for B in sizes:
        batch = synth.crf_batch(B, sh, seed=1)
        e = batch['emissions'].to(DEV)
        tags = batch['tags'].to(DEV)
        full = torch.ones(B, sh.S, dtype=torch.uint8, device=DEV)
        short = batch['mask'].to(torch.uint8).to(DEV)
        cp = {k: v.to(DEV) for k, v in synth.crf_params(sh.T, 2).items()}
        w = torch.ones(B, device=DEV)
        for nm, m in (('full-length', full), ('tweet-length', short)):
            tf = timeit(lambda: ops.crf_llh(e, tags, m, cp['start_transitions'], cp['end_transitions'], cp['transitions']))
            tb = timeit(lambda: ops.crf_llh_bwd(e, tags, m, cp['start_transitions'], cp['end_transitions'],
                                                cp['transitions'], w))
            print(f'crf B={B:6d} {nm:13s} llh fwd {tf*1e6:8.1f} us   llh bwd {tb*1e6:8.1f} us')


if __name__ == '__main__':
    main()
