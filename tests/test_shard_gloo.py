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
