"""Whole-step CUDA graphs for the training path.

One training step of the fusion path is ~130 kernel launches from Python (autograd nodes -> ctypes -> cudaLaunchKernelEx):
at 32-128 sentences per GPU the host, not the GPU, sets the step time (3.7 ms at 32 sentences, 3.9 ms at 128).
``CapturedStep`` records forward + backward + gradient all-reduce + optimizer step once and replays them with one driver
call.  Two things a replay cannot take from the host are moved to the device:

  * dropout seeds -- the per-call seeds the modules draw on the host are frozen at capture time; the part that must change
    every step is an 8-byte counter in device memory (``icka_set_seed_base``) that the graph itself advances, so each replay
    draws fresh Philox masks and backward still regenerates the masks of its own forward;
  * input validation that reads values back (``CRF._validate``) runs in the warm-up passes and is skipped while capturing.

Inputs are static device tensors: copy the next batch into them (``tensor.copy_``) before ``replay()``.
"""
from __future__ import annotations

from typing import Callable

import torch

from . import _lib, modules

_GOLDEN = 0x9E3779B97F4A7C15 - (1 << 64)          # 2^64 / phi as a signed 64-bit increment


class CapturedStep:
    def __init__(self, step: Callable[[], torch.Tensor], device, warmup: int = 3):
        """``step()`` runs one whole training step (zero_grad, forward, backward, reducer.finish(), optimizer.step()) on
        static input tensors and returns the loss tensor.  The optimizer must be capturable (``capturable=True``)."""
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.seed_base = torch.zeros(1, dtype=torch.int64, device=self.device)
        _lib.set_seed_base(idx, self.seed_base.data_ptr(), keep_alive=self.seed_base)
        self._idx = idx
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                       # eager warm-up off the default stream (allocator, caches, NCCL)
            for _ in range(warmup):
                self.seed_base.add_(_GOLDEN)
                step()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        # (the warm-up passes ran on a side stream, the capture runs on torch's capture stream: autograd's note about
        # AccumulateGrad nodes seeing a different stream than the one they were created on is expected here)
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count(idx)
        with torch.cuda.graph(self.graph):
            self.seed_base.add_(_GOLDEN)                    # every replay starts from a new seed base
            self.loss = step()
        self.kernels = _lib.launch_count(idx) - n0          # library kernels one replay launches

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        modules.invalidate_operand_caches()     # the replayed optimizer step changed the weights behind Python's back
        return self.loss

    def close(self) -> None:
        _lib.set_seed_base(self._idx, None)
