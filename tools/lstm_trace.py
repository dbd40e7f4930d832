"""Developer timeline of the persistent LSTM kernel (CTA 0): python tools/lstm_trace.py [B]
Prints, per phase, the median time in us between consecutive events of an item and from step to step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icka_b200
from icka_b200 import ops
from icka_b200.config import FusionConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S, H = 128, 768
icka_b200.set_precision('bf16')
torch.manual_seed(0)
head = icka_b200.EmissionHead(FusionConfig(hidden_size=H), num_labels=15).cuda().eval()
x = torch.randn(B, S, H, device='cuda')
variant = ops.lstm_rec_variant(B)
wi, b, wh = head.lstm._prepared(variant)
gx = ops.linear(ops.cast_bf16_time_major(x).view(S * B, H), wi, b, out_dtype=torch.bfloat16)
for _ in range(2):
    ops.lstm_rec(gx, wh, B, S, H, variant=variant)
trace = torch.zeros(4096 * 8, dtype=torch.int64, device='cuda')
os.environ['ICKA_LSTM_TRACE'] = str(trace.data_ptr())
ops.lstm_rec(gx, wh, B, S, H, variant=variant)
torch.cuda.synchronize()
del os.environ['ICKA_LSTM_TRACE']
t = trace.view(4096, 8).cpu().double()
items = int((t[:, 0] > 0).sum())
npairs = items // S
t = t[:items]
names = ['dep ready', 'A issued', 'MMA committed', 'acc seen', 'tile written', 'store done', 'released']
print(f'B={B} variant={variant}: {items} items, {npairs} per step')
for p in range(npairs):
    rows = t[p::npairs][8:-4]
    line = []
    for a in range(6):
        line.append(f'{names[a]} -> {names[a + 1]}: {float((rows[:, a + 1] - rows[:, a]).median()) / 1e3:.2f}')
    step = float((rows[1:, 0] - rows[:-1, 0]).median()) / 1e3
    rel_to_dep = float((rows[1:, 0] - rows[:-1, 6]).median()) / 1e3
    print(f' item {p}: ' + ' | '.join(line) + f' | released -> next dep ready: {rel_to_dep:.2f} | step {step:.2f} us')
