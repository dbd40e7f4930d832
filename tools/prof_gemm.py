"""Launches each GEMM flavour of one t2i layer a few times (B=1024) -- target command for ncu."""
import math
import sys

import torch

sys.path.insert(0, '.')
from icka_b200 import ops  # noqa: E402
from icka_b200._lib import ACT_GELU_ERF, ACT_NONE  # noqa: E402

DEV = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S, R, H, I = 128, 49, 768, 3072
bf = torch.bfloat16
w = lambda n, k: (torch.randn(n, k, device=DEV) / math.sqrt(k)).bfloat16()
x = torch.randn(B * S, H, device=DEV)
xb, fb = x.bfloat16(), torch.randn(B * S, I, device=DEV).bfloat16()
rows = torch.randn(B * R, 2048, device=DEV).bfloat16()
Wvm, Wq, Wi, Wd = w(H, 2048), w(H, H), w(I, H), w(H, I)
bH, bI = torch.randn(H, device=DEV), torch.randn(I, device=DEV)
for rep in range(2):
    ops.linear(rows, Wvm, bH, out_dtype=bf)                                   # region proj
    ops.linear(xb, Wq, bH, out_dtype=bf)                                      # Q proj
    ops.linear(xb, Wq, bH, residual=x, out_dtype=torch.float32)               # out proj + residual
    ops.linear(xb, Wi, bI, act=ACT_GELU_ERF, out_dtype=bf)                    # FFN up + GELU
    ops.linear(fb, Wd, bH, residual=x, out_dtype=torch.float32)               # FFN down + residual
torch.cuda.synchronize()
print('ok')
