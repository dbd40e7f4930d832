"""World-size-2 gloo test (CPU) of the N>1 path: contiguous batch shards, ordered gather of ragged tag
lists (decoded here by the CPU oracle, standing in for the per-rank GPU decode), max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icka_b200 import shard, synth
from oracle import crf_ref


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sh = synth.Shape(S=24, T=7)
        batch = synth.crf_batch(11, sh, seed=3, kind='ties', median_len=9)
        cp = synth.crf_params(sh.T, 4, 'normal')
        full = {'emissions': batch['emissions'], 'mask': batch['mask']}
        mine = shard.shard_batch(full, world, rank)
        local = crf_ref.viterbi_decode(mine['emissions'], mine['mask'], cp['start_transitions'],
                                       cp['end_transitions'], cp['transitions'])
        everything = shard.gather_tag_lists(local)
        slowest = shard.max_over_ranks(10.0 + rank)
        if rank == 0:
            want = crf_ref.viterbi_decode(batch['emissions'], batch['mask'], cp['start_transitions'],
                                          cp['end_transitions'], cp['transitions'])
            q.put((everything == want, slowest))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_decode_matches_single_rank():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert slowest == 11.0


# ---- data-parallel gradient averaging (the training collective) ------------------------------------------
def _make_model():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                               torch.nn.Linear(16, 3))


def _dp_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _make_model()
        if rank != 0:                       # ranks start out different; broadcast must fix that
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        shard.broadcast_parameters(model)
        model[4].bias.requires_grad_(True)
        extra = torch.nn.Parameter(torch.ones(5))          # registered but never used: no gradient arrives
        reducer = shard.GradientAllReducer(list(model.parameters()) + [extra], bucket_bytes=1024)
        g = torch.Generator().manual_seed(11)
        x, y = torch.randn(10, 12, generator=g), torch.randn(10, 3, generator=g)
        mine = shard.shard_batch({'x': x, 'y': y}, world, rank)
        grads = []
        for step in range(2):                               # twice: the hook counters must re-arm
            for p in model.parameters():
                p.grad = None
            loss = ((model(mine['x']) - mine['y']) ** 2).mean()
            loss.backward()
            reducer.finish()
            grads.append([p.grad.clone() for p in model.parameters()])
        if rank == 0:
            # by value (numpy): a torch tensor on a multiprocessing queue travels as a file descriptor that the parent
            # must fetch from THIS process while it is still alive -- a race once the worker exits right after the put
            q.put(([[t.numpy() for t in step] for step in grads], len(reducer.buckets), reducer.launched_early,
                   None if extra.grad is None else extra.grad.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads, n_buckets, early, extra_grad = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = _make_model()
    g = torch.Generator().manual_seed(11)
    x, y = torch.randn(10, 12, generator=g), torch.randn(10, 3, generator=g)
    ((model(x) - y) ** 2).mean().backward()                 # equal shards: mean of shard means == full-batch mean
    for step in range(2):
        for got, p in zip(grads[step], model.parameters()):
            assert torch.allclose(torch.from_numpy(got), p.grad, rtol=1e-5, atol=1e-6)
    assert n_buckets >= 3                                   # 1 KiB buckets split this model
    assert early >= 2 * (n_buckets - 1)                     # all but the unused-parameter bucket start inside backward
    assert extra_grad is not None and float(abs(extra_grad).max()) == 0.0


# ---- gradient accumulation: several backward() calls per finish() (My_cross_attention.py:821-831) ---------------------
def _accum_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _make_model()
        reducer = shard.GradientAllReducer(list(model.parameters()), bucket_bytes=1024)
        g = torch.Generator().manual_seed(13)
        x, y = torch.randn(12, 12, generator=g), torch.randn(12, 3, generator=g)
        mine = shard.shard_batch({'x': x, 'y': y}, world, rank)        # 6 rows per rank = 2 micro-batches of 3
        grads = []
        for step in range(2):                                           # second window: p.grad are bucket views by then
            model.zero_grad(set_to_none=(step == 0))                    # both zero_grad flavours
            micro = [(mine['x'][i:i + 3], mine['y'][i:i + 3]) for i in (0, 3)]
            with reducer.no_sync():
                (((model(micro[0][0]) - micro[0][1]) ** 2).mean() / 2).backward()
            (((model(micro[1][0]) - micro[1][1]) ** 2).mean() / 2).backward()
            reducer.finish()
            grads.append([p.grad.clone().numpy() for p in model.parameters()])
        # a second synchronising backward() before finish() must raise, not race
        raised = False
        model.zero_grad(set_to_none=True)
        ((model(mine['x']) - mine['y']) ** 2).mean().backward()
        try:
            ((model(mine['x']) - mine['y']) ** 2).mean().backward()
        except RuntimeError as e:
            raised = 'no_sync' in str(e)
        reducer.finish()
        if rank == 0:
            q.put((grads, raised))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_accumulation_matches_full_batch():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_accum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads, raised = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = _make_model()
    g = torch.Generator().manual_seed(13)
    x, y = torch.randn(12, 12, generator=g), torch.randn(12, 3, generator=g)
    ((model(x) - y) ** 2).mean().backward()        # equal micro-batches: mean of means == full-batch mean
    for step in range(2):
        for got, p in zip(grads[step], model.parameters()):
            assert torch.allclose(torch.from_numpy(got), p.grad, rtol=1e-5, atol=1e-6)
    assert raised


def test_last_bucket_is_small():
    """The bucket that completes last (first-registered parameters) cannot overlap backward: it is kept small."""
    import torch
    from icka_b200 import shard
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (300_000, 300_000, 200_000, 5_000_000, 4_000_000, 10)]
    r = shard.GradientAllReducer(ps, bucket_bytes=25 << 20, last_bucket_bytes=2 << 20)
    sizes = [sum(p.numel() for p in b.params) for b in r.buckets]
    assert sum(sizes) == sum(p.numel() for p in ps)
    assert [id(p) for p in r.buckets[-1].params] == [id(ps[1]), id(ps[0])]          # 2.4 MB >= 2 MB: the first two parameters
    assert [id(p) for b in r.buckets[:-1] for p in b.params] == [id(p) for p in reversed(ps[2:])]   # reverse registration order
    assert len(r.buckets) == 3                                                       # 16 MB + 20 MB do not share a 25 MB bucket
    r.remove_hooks()
    r1 = shard.GradientAllReducer(ps, last_bucket_bytes=0)
    assert [id(p) for b in r1.buckets for p in b.params] == [id(p) for p in reversed(ps)]
    r1.remove_hooks()
