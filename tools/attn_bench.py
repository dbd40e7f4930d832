"""Cross-attention core at the hi-res shape (BASELINE configs[3]: 256 queries x 196 regions) and the std shape, every
kernel variant (icka_set_attn_mode): CUDA-event time, GB/s of the algorithmic bytes (Q + K|V in, ctx out), max error."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icka_b200 import _lib, ops                                    # noqa: E402

DEV = 'cuda:0'
lib = _lib.load()
NAMES = {0: 'default', 1: 'mma.sync', 2: 'tcgen05 wide (1 group, P in smem)', 3: 'tcgen05 wide2 (2 groups, P in TMEM)'}


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


for B, Sq, Skv, nh in ((512, 256, 196, 12), (1024, 128, 49, 12), (256, 128, 128, 12), (128, 1, 128, 12), (1024, 1, 128, 12)):
    H = nh * 64
    g = torch.Generator(DEV).manual_seed(1)
    q = torch.randn(B * Sq, H, device=DEV, generator=g).bfloat16()
    kv = torch.randn(B * Skv, 2 * H, device=DEV, generator=g).bfloat16()
    mask = torch.zeros(B, Skv, device=DEV)
    nbytes = (2 * B * Sq * H + 2 * B * Skv * H) * 2
    lib.icka_set_attn_mode(1)
    ref = ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask, B, Sq, Skv, nh, 64).float()
    for mode in (1, 2, 3, 0):
        if mode in (2, 3) and Skv <= 64:
            continue
        _lib.check(lib.icka_set_attn_mode(mode), 'icka_set_attn_mode')
        fn = lambda: ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask, B, Sq, Skv, nh, 64)
        out = fn().float()
        torch.cuda.synchronize()
        t = timed(fn)
        print(f'B={B} Sq={Sq} Skv={Skv}: {NAMES[mode]:38s} {t * 1e3:7.1f} us  {nbytes / t / 1e6:7.0f} GB/s  max|d vs mma.sync| {float((out - ref).abs().max()):.2e}')
    lib.icka_set_attn_mode(0)
