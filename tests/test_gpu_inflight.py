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


@pytest.mark.parametrize('use_graphs', [False, True])
def test_infer_host_matches_device_step_and_oracle(use_graphs):
    """The end-to-end call bench.py's `e2e` times: pinned host batches in, host tags / lengths / gates out, batch i+1's
    copies overlapping batch i's kernels on two buffer sets.  Tags must be the oracle's, gates the device step's; the
    bf16-resident host batches (half the bytes) must give the same tags and gates within the bf16 gate."""
    from oracle import crf_ref
    shape = synth.Shape(L=1)
    pipe = FusionViterbiPipeline(shape, 'cuda:0', 'bf16', seed=5)
    B = 96
    hosts = [pipe.make_host_batch(B, shape, 40 + i) for i in range(3)]
    results, (start, end) = pipe.infer_host(hosts, use_graphs=use_graphs)
    torch.cuda.synchronize()
    assert start.elapsed_time(end) > 0
    crf = {k: v.detach().cpu() for k, v in pipe.crf.state_dict().items()}
    for host, (tags, lens, gate) in zip(hosts, results):
        want = crf_ref.viterbi_decode(host['emissions'], host['crf_mask'].bool(), crf['start_transitions'],
                                      crf['end_transitions'], crf['transitions'])
        got = [tags[b, :int(lens[b])].tolist() for b in range(B)]
        assert got == want
        _, _, _, _, gate_dev = pipe.step_device(pipe.to_device(host))
        torch.cuda.synchronize()
        assert torch.equal(gate.view(-1), gate_dev.cpu().view(-1))
    hosts16 = [pipe.make_host_batch(B, shape, 40 + i, bf16_states=True) for i in range(3)]
    assert pipe.h2d_bytes(hosts16[0]) < 0.51 * pipe.h2d_bytes(hosts[0])
    results16, _ = pipe.infer_host(hosts16, use_graphs=use_graphs)
    torch.cuda.synchronize()
    for (tags, lens, gate), (tags16, lens16, gate16) in zip(results, results16):
        assert torch.equal(tags, tags16) and torch.equal(lens, lens16)
        assert float((gate.view(-1) - gate16.view(-1)).abs().max()) <= 2e-2
