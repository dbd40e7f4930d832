"""Per-kernel CUDA-event timings at the bench shapes (developer tool; prints a table)."""
import math
import sys

import torch

sys.path.insert(0, '.')
from icka_b200 import ops, synth  # noqa: E402
from icka_b200._lib import ACT_GELU_ERF, ACT_NONE  # noqa: E402

DEV = 'cuda:0'
PEAK_TF, PEAK_GBS = 1671.2, 6547.5


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


def main(B=1024, mode=0):
    from icka_b200 import _lib
    _lib.load().icka_set_gemm_mode(mode)
    print('gemm mode', mode)
    S, R, H, I = 128, 49, 768, 3072
    bf = torch.bfloat16
    x = torch.randn(B * S, H, device=DEV)
    xb = x.bfloat16()
    rows = torch.randn(B * R, 2048, device=DEV).bfloat16()
    regs = torch.randn(B * R, H, device=DEV).bfloat16()
    fb = torch.randn(B * S, I, device=DEV).bfloat16()
    w = lambda n, k: (torch.randn(n, k, device=DEV) / math.sqrt(k)).bfloat16()
    bias = lambda n: torch.randn(n, device=DEV)
    Wvm, Wq, Wkv, Wo, Wi, Wd = w(H, 2048), w(H, H), w(2 * H, H), w(H, H), w(I, H), w(H, I)
    bH, b2H, bI = bias(H), bias(2 * H), bias(I)
    q = torch.randn(B * S, H, device=DEV).bfloat16()
    kv = torch.randn(B * R, 2 * H, device=DEV).bfloat16()
    grid = torch.randn(B, 2048, R, device=DEV)
    g1 = torch.ones(H, device=DEV)
    tok = torch.randn(B, S, H, device=DEV)
    fused = torch.randn(B, S, H, device=DEV)
    wf, cf = torch.randn(H, device=DEV) * 0.01, torch.zeros(1, device=DEV)

    def gemm(name, a, W, b, res=None, act=ACT_NONE, od=bf):
        M, K = a.shape
        N = W.shape[0]
        t = timeit(lambda: ops.linear(a, W, b, residual=res, act=act, out_dtype=od))
        tf = 2.0 * M * N * K / t / 1e12
        print(f'{name:28s} M={M:7d} N={N:5d} K={K:5d}  {t*1e6:9.1f} us  {tf:7.1f} TF/s  {tf/PEAK_TF*100:5.1f}% of burst peak')
        return t

    def mem(name, fn, nbytes):
        t = timeit(fn)
        print(f'{name:28s} {nbytes/1e6:9.1f} MB  {t*1e6:9.1f} us  {nbytes/t/1e9:8.1f} GB/s  {nbytes/t/1e9/PEAK_GBS*100:5.1f}% of HBM peak')
        return t

    tot = 0.0
    tot += mem('region_rows', lambda: ops.region_rows(grid, bf), B * 2048 * R * 6)
    tot += mem('cast text', lambda: ops.cast_bf16(x), B * S * H * 6)
    tot += gemm('region proj', rows, Wvm, bH)
    tot += gemm('Q proj', xb, Wq, bH)
    tot += gemm('K|V proj', regs, Wkv, b2H)
    tot += mem('attention core', lambda: ops.cross_attn_core(q, kv[:, :H], kv[:, H:], None, B, S, R, 12, 64),
               B * (S * H * 2 * 2 + R * 2 * H * 2))
    tot += gemm('out proj +res (f32 out)', xb, Wo, bH, res=x, od=torch.float32)
    tot += mem('layernorm', lambda: ops.layernorm(x, g1, g1, 1e-12, want_f32=True, want_bf16=True), B * S * H * 10)
    def gemm_ln(name, a, W, b, res):
        M, K = a.shape
        N = W.shape[0]
        t = timeit(lambda: ops.linear_ln(a, W, b, res, g1, g1, 1e-12, want_bf16=True))
        tf = 2.0 * M * N * K / t / 1e12
        print(f'{name:28s} M={M:7d} N={N:5d} K={K:5d}  {t*1e6:9.1f} us  {tf:7.1f} TF/s  (GEMM + LayerNorm in one launch)')
    gemm_ln('out proj +res +LN fused', xb, Wo, bH, x)
    gemm_ln('FFN down +res +LN fused', fb, Wd, bH, x)
    tot += gemm('FFN up + gelu', xb, Wi, bI, act=ACT_GELU_ERF)
    tot += gemm('FFN down +res (f32 out)', fb, Wd, bH, res=x, od=torch.float32)
    tot += mem('layernorm', lambda: ops.layernorm(x, g1, g1, 1e-12, want_f32=True, want_bf16=True), B * S * H * 10)
    tot += mem('gate blend', lambda: ops.gate_blend(fused, tok, g1, g1, 1e-5, wf, cf), B * S * H * 12)
    mem('ln + gate blend (fused)', lambda: ops.ln_gate_blend(fused, g1, g1, 1e-12, tok, g1, g1, 1e-5, wf, cf),
        B * S * H * 14)
    u = (torch.randn(B, 12 * H, device=DEV) / math.sqrt(H)).bfloat16()
    mem('i2t pool', lambda: ops.i2t_pool(u, xb, None, B, S, H, 12), B * (S * H * 2 + 2 * 12 * H * 2))
    print(f'sum of t2i-side kernels: {tot*1e3:.3f} ms for B={B} -> {B/tot:,.0f} sentences/s (excl. i2t)')

    sh = synth.STD
    for Bv in (256, 4096, 65536):
        e = torch.randn(Bv, sh.S, sh.T, device=DEV)
        lens = synth.lengths(Bv, sh.S, torch.Generator().manual_seed(1))
        m_short = synth.prefix_mask(lens, sh.S).to(torch.uint8).to(DEV)
        m_full = torch.ones(Bv, sh.S, dtype=torch.uint8, device=DEV)
        st, en, tr = torch.randn(sh.T, device=DEV), torch.randn(sh.T, device=DEV), torch.randn(sh.T, sh.T, device=DEV)
        for nm, m in (('full-length', m_full), ('tweet-length', m_short)):
            t = timeit(lambda: ops.viterbi(e, m, st, en, tr))
            nb = Bv * 8320
            print(f'viterbi B={Bv:6d} {nm:13s} {t*1e6:9.1f} us  {Bv/t/1e6:8.2f} M sent/s  {nb/t/1e9:8.1f} GB/s algorithmic ({nb/t/1e9/PEAK_GBS*100:5.1f}% of HBM peak)')


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
