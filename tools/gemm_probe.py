"""Developer probe: GEMM time with the epilogue partly disabled (mode | debug << 4)."""
import math, sys, torch
sys.path.insert(0, '.')
from icka_b200 import ops, _lib
from icka_b200._lib import ACT_GELU_ERF, ACT_NONE
DEV = 'cuda:0'
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3
B, S, H, I = 1024, 128, 768, 3072
x = torch.randn(B * S, H, device=DEV); xb = x.bfloat16(); fb = torch.randn(B * S, I, device=DEV).bfloat16()
w = lambda n, k: (torch.randn(n, k, device=DEV) / math.sqrt(k)).bfloat16()
Wq, Wi, Wd = w(H, H), w(I, H), w(H, I)
bH, bI = torch.randn(H, device=DEV), torch.randn(I, device=DEV)
cases = [('Q proj bf16 out', lambda: ops.linear(xb, Wq, bH, out_dtype=torch.bfloat16), 2.0 * B * S * H * H),
         ('out proj f32+res', lambda: ops.linear(xb, Wq, bH, residual=x, out_dtype=torch.float32), 2.0 * B * S * H * H),
         ('FFN up gelu', lambda: ops.linear(xb, Wi, bI, act=ACT_GELU_ERF, out_dtype=torch.bfloat16), 2.0 * B * S * H * I),
         ('FFN down f32+res', lambda: ops.linear(fb, Wd, bH, residual=x, out_dtype=torch.float32), 2.0 * B * S * H * I)]
for mode in (1, 2):
    for dbg in (0, 2, 1):
        _lib.load().icka_set_gemm_mode(mode | (dbg << 4))
        for name, fn, fl in cases:
            t = timeit(fn)
            print(f'mode {mode} debug {dbg} {name:18s} {t*1e6:8.1f} us {fl/t/1e12:7.1f} TF/s', flush=True)
_lib.load().icka_set_gemm_mode(0)
