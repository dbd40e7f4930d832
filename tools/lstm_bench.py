"""CUDA-event timing of the emission head (BiLSTM + classifier) stages:  python tools/lstm_bench.py [B] [S]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import icka_b200
from icka_b200 import ops
from icka_b200.config import FusionConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
H, T = 768, 15
icka_b200.set_precision('bf16')
torch.manual_seed(0)
head = icka_b200.EmissionHead(FusionConfig(hidden_size=H), num_labels=T).cuda().eval()
x = torch.randn(B, S, H, device='cuda')
xb = ops.cast_bf16_time_major(x).view(S * B, H)
variant = ops.lstm_rec_variant(B)
wi, b, wh = head.lstm._prepared(variant)
ws = torch.empty(ops.lstm_rec_workspace_bytes(B, H) + 1024, dtype=torch.uint8, device='cuda')


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2]


gx = ops.linear(xb, wi, b, out_dtype=torch.bfloat16)
y = ops.lstm_rec(gx, wh, B, S, H, variant=variant, workspace=ws)
t_cast = timed(lambda: ops.cast_bf16_time_major(x))
t_gx = timed(lambda: ops.linear(xb, wi, b, out_dtype=torch.bfloat16))
t_rec = timed(lambda: ops.lstm_rec(gx, wh, B, S, H, variant=variant, workspace=ws))
w32, b32 = head.classifier.weight.detach().contiguous(), head.classifier.bias.detach().contiguous()
t_head = timed(lambda: ops.emission_head(y.view(S * B, 2 * H), w32, b32, time_major_S=S))
with torch.no_grad():
    t_all = timed(lambda: head(x))
ref = torch.nn.LSTM(H, H, batch_first=True, bidirectional=True).cuda().to(torch.bfloat16)
xr = x.to(torch.bfloat16)
with torch.no_grad():
    t_cudnn = timed(lambda: ref(xr))
flops_gx = 2.0 * B * S * 8 * H * H
flops_rec = 2.0 * B * S * 8 * H * H
print(f'B={B} S={S} variant={variant}: cast {t_cast:.3f} ms | gx GEMM {t_gx:.3f} ms ({flops_gx / t_gx / 1e9:.0f} TF/s) | '
      f'recurrent {t_rec:.3f} ms ({flops_rec / t_rec / 1e9:.0f} TF/s, {t_rec / S * 1e3:.2f} us/step) | '
      f'classifier {t_head:.3f} ms ({B * S * 2 * H * 2 / t_head / 1e6:.0f} GB/s) | '
      f'module {t_all:.3f} ms = {B / t_all * 1e3:.0f} sentences/s | torch nn.LSTM (cuDNN, bf16) {t_cudnn:.3f} ms')
