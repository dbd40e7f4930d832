"""Per-step cost of the fused BPTT step kernels (csrc/lstm_train.cu): forward-save and backward loops timed apart.
usage: python tools/lstm_train_bench.py [B ...]"""
import sys
import torch
from icka_b200 import _lib

H, S = 768, 128
lib = _lib.load()
dev = torch.device('cuda', 0)
h = _lib.handle(0)
for B in [int(a) for a in sys.argv[1:]] or [32, 64, 128]:
    torch.manual_seed(0)
    gx = (torch.randn(B, S, 8 * H, device=dev) * 0.5).bfloat16()
    w = [(torch.randn(4 * H, H, device=dev) / H ** 0.5).bfloat16() for _ in range(2)]
    wt = [x.t().contiguous() for x in w]
    y_op = torch.zeros(B, S, 2 * H, device=dev, dtype=torch.bfloat16)
    y32 = torch.zeros(B, S, 2 * H, device=dev)
    acts = torch.zeros(2, S, B, 4 * H, device=dev)
    c_all = torch.zeros(2, S, B, H, device=dev)
    dy = torch.randn(B, S, 2 * H, device=dev)
    dg = torch.zeros(B, S, 8 * H, device=dev, dtype=torch.bfloat16)
    dc = torch.zeros(2, B, H, device=dev)

    def fwd():
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.icka_lstm_bidir_fwd_save(h, gx.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), y_op.data_ptr(),
                                                y32.data_ptr(), acts.data_ptr(), c_all.data_ptr(), B, S, H, st), 'fwd')

    def bwd():
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.icka_lstm_bidir_bwd(h, dy.data_ptr(), wt[0].data_ptr(), wt[1].data_ptr(), acts.data_ptr(),
                                           c_all.data_ptr(), dg.data_ptr(), dc.data_ptr(), B, S, H, st), 'bwd')

    out = []
    for name, fn in (('fwd', fwd), ('bwd', bwd)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append(f'{name} {e0.elapsed_time(e1) / 5 / S * 1e3:.2f} us/step')
    print(f'B={B}: ' + ', '.join(out), flush=True)
