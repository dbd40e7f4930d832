"""Two captured steps in flight on two streams (bench.py --inflight 2): each capture has its own device buffers and its
own library handle slot (split-K workspace); concurrent replays must reproduce the eager results bit for bit."""
import pytest
import torch

from icka_b200 import synth
from icka_b200.pipeline import FusionViterbiPipeline

pytestmark = pytest.mark.gpu


def test_concurrent_graph_replays_match_eager():
    shape = synth.Shape(L=1)
    pipe = FusionViterbiPipeline(shape, 'cuda:0', 'bf16', seed=3)
    B = 192
    ds = [pipe.to_device(pipe.make_host_batch(B, shape, seed, pin=False)) for seed in (11, 12)]
    torch.cuda.synchronize()
    eager = [[t.clone() for t in pipe.step_device(d)] for d in ds]
    torch.cuda.synchronize()
    caps = [pipe.capture(ds[0], slot=0), pipe.capture(ds[1], slot=1)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    for _ in range(6):
        for (g, _), st in zip(caps, streams):
            with torch.cuda.stream(st):
                g.replay()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    for (g, outs), want in zip(caps, eager):
        for got, ref in zip(outs, want):
            assert torch.equal(got, ref)
