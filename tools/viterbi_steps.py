import sys, torch
sys.path.insert(0, '.')
from icka_b200 import ops
DEV='cuda:0'
def timeit(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
B=1024; T=15
st, en, tr = torch.randn(T, device=DEV), torch.randn(T, device=DEV), torch.randn(T, T, device=DEV)
for S in (4, 16, 32, 64, 128, 256):
    e = torch.randn(B, S, T, device=DEV); m = torch.ones(B, S, dtype=torch.uint8, device=DEV)
    g = torch.cuda.CUDAGraph()
    ops.viterbi(e, m, st, en, tr); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(10): ops.viterbi(e, m, st, en, tr)
    t = timeit(g.replay) / 10
    print(f'S={S:4d} B={B}: {t:7.2f} us per decode (graph of 10 launches)')
