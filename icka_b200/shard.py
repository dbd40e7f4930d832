"""Batch sharding across ranks: sentences are independent, so a batch is cut into contiguous
per-rank shards exactly like the reference's ``DistributedSampler`` / ``DataParallel`` scatter
(My_cross_attention.py:707, 777-779).  Inference needs no data-path collective; the helpers below are
the only distributed logic: shard bounds, and gathering ragged per-rank tag lists back in order.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n`` sentences owned by ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, world: int, rank: int) -> dict:
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, world, rank)
    return {k: v[lo:hi] for k, v in batch.items()}


def gather_tag_lists(local: Sequence[Sequence[int]], group=None) -> List[List[int]]:
    """All ranks obtain the full, ordered list of decoded tag sequences (ragged -> all_gather_object)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [list(x) for x in local]
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, [list(x) for x in local], group=group)
    return [seq for part in parts for seq in part]


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing reduction bench.py uses: the slowest rank defines the step time."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------
# Training: the one collective of the path -- data-parallel gradient averaging (the reference
# does it with apex DistributedDataParallel / nn.DataParallel, My_cross_attention.py:768-779).
# ---------------------------------------------------------------------------------------------------------
def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Every rank starts from rank ``src``'s weights (what DDP does at construction)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


class _Bucket:
    __slots__ = ('params', 'offsets', 'numel', 'flat', 'pending', 'work')

    def __init__(self, params):
        self.params = params
        self.offsets, n = [], 0
        for p in params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.flat = None
        self.pending = len(params)
        self.work = None


class GradientAllReducer:
    """Bucketed, backward-overlapped gradient all-reduce for one process per GPU.

    Parameters are grouped into buckets of about ``bucket_bytes`` (the last one to complete: ``last_bucket_bytes``) in
    REVERSE registration order (roughly the order backward produces their gradients: the kernel-backed autograd nodes of ``icka_b200.autograd`` emit
    all 16 parameter gradients of a cross layer at once, so buckets fill layer by layer).  A
    post-accumulate-grad hook counts a bucket's gradients in; when the last one lands, the bucket is packed
    into one flat fp32 buffer and all-reduced asynchronously (NCCL over NVLink 5 / NVSwitch on GPU ranks --
    the collective runs on NCCL's own stream beside the rest of backward; gloo in the CPU tests).
    ``finish()`` (call it after ``loss.backward()``, before clipping / the optimizer step) waits for the
    transfers and averages over the ranks -- ranks hold equal shards, so this is the global-batch mean the
    reference's DDP produces; ``p.grad`` becomes a view of its bucket (no unpack copy, like DDP's
    ``gradient_as_bucket_view``).

    Gradient accumulation (the reference's ``--gradient_accumulation_steps``, My_cross_attention.py:821-831): run every
    micro-batch but the last under ``with reducer.no_sync():`` -- their ``backward()`` only accumulates into ``p.grad``,
    nothing is counted or sent -- and the last one normally; its hooks see the ACCUMULATED gradients, so the buckets
    carry the sum.  A second ``backward()`` outside ``no_sync()`` before ``finish()`` would accumulate into a buffer an
    asynchronous all-reduce is still reading: it raises instead.
    """

    def __init__(self, params, bucket_bytes: int = 25 << 20, group=None, last_bucket_bytes: int = 4 << 20,
                 overlap: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        plist = [p for p in params if p.requires_grad]
        seen, uniq = set(), []
        for p in plist:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.buckets: List[_Bucket] = []
        # The bucket that completes LAST (the first-registered parameters: their gradients are the last ones backward
        # produces) cannot overlap anything -- the optimizer waits for it -- so it is kept small: just enough of the first
        # parameters to reach ``last_bucket_bytes``.
        tail, tail_bytes = [], 0
        if last_bucket_bytes and len(uniq) > 1:
            for p in uniq:
                if tail_bytes >= last_bucket_bytes or len(tail) == len(uniq) - 1:
                    break
                tail.append(p)
                tail_bytes += p.numel() * 4
            if tail_bytes > bucket_bytes:
                tail = []
        head = uniq[len(tail):]
        cur, cur_bytes = [], 0
        for p in reversed(head):
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(_Bucket(cur))
        if tail:
            self.buckets.append(_Bucket(list(reversed(tail))))
        self._where = {}
        self._handles = []
        for b in self.buckets:
            for p in b.params:
                self._where[id(p)] = b
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.launched_early = 0      # buckets whose all-reduce started from inside backward (diagnostics)
        # overlap = False: every bucket is sent from finish(), after backward (developer knob ICKA_ALLREDUCE_OVERLAP=0)
        self.overlap = overlap and os.environ.get('ICKA_ALLREDUCE_OVERLAP', '1') != '0'
        self._sync = True

    @contextlib.contextmanager
    def no_sync(self):
        """Micro-batches whose gradients are only accumulated locally (all but the last of an accumulation window)."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev

    def remove_hooks(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    # -- internals -------------------------------------------------------------------------------------
    def _on_grad(self, p: torch.Tensor) -> None:
        if not self._sync:
            return
        b = self._where[id(p)]
        if b.pending <= 0 or b.work is not None:
            raise RuntimeError('GradientAllReducer: a gradient arrived for a bucket that was already handed to the '
                               'all-reduce -- a second backward() before finish(); wrap all micro-batches but the last '
                               'in `with reducer.no_sync():`')
        b.pending -= 1
        if b.pending == 0 and self.world > 1 and self.overlap:
            self._launch(b)
            self.launched_early += 1

    def _launch(self, b: _Bucket) -> None:
        ref = b.params[0]
        if b.flat is None or b.flat.device != ref.device:
            b.flat = torch.empty(b.numel, dtype=torch.float32, device=ref.device)
        with torch.no_grad():
            views = [b.flat[off:off + p.numel()] for p, off in zip(b.params, b.offsets)]
            lo, hi = b.flat.data_ptr(), b.flat.data_ptr() + b.flat.numel() * 4
            aliased = [p.grad is not None and lo <= p.grad.data_ptr() < hi for p in b.params]
            if not any(aliased) and all(p.grad is not None for p in b.params):
                torch.cat([p.grad.reshape(-1) for p in b.params], out=b.flat)          # one pack kernel per bucket
            else:
                for p, v, a in zip(b.params, views, aliased):
                    if a and p.grad.data_ptr() == v.data_ptr():
                        continue            # accumulated straight into last step's bucket view: already packed
                    if p.grad is None:
                        v.zero_()
                    else:
                        v.copy_(p.grad.reshape(-1))
        op = dist.ReduceOp.SUM
        b_avg = False
        if self.world > 1 and dist.get_backend(self.group) == 'nccl':
            op, b_avg = dist.ReduceOp.AVG, True            # NCCL scales inside the collective
        b.work = (dist.all_reduce(b.flat, op=op, group=self.group, async_op=True), b_avg)

    def finish(self) -> None:
        """Wait for every bucket, average, scatter back; re-arm the hooks' counters for the next step."""
        if self.world > 1:
            for b in self.buckets:
                if b.work is None:               # some gradient of this bucket never arrived (unused parameter)
                    self._launch(b)
            with torch.no_grad():
                for b in self.buckets:
                    work, averaged = b.work
                    work.wait()
                    if not averaged:
                        b.flat.div_(self.world)
                    for p, off in zip(b.params, b.offsets):
                        p.grad = b.flat[off:off + p.numel()].view_as(p)      # gradients become views of the bucket
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None
