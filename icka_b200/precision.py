"""Precision mode of the kernel-backed modules: 'bf16' (tcgen05 operands, fp32 everything else) or 'fp32' (parity path).

State is never mutated inside a forward.  Three levels, innermost wins:

  * ``with precision('fp32'):`` -- a THREAD-LOCAL override for the calls made inside the block on this thread.  The
    reference's ``nn.DataParallel`` runs one replica per Python thread (My_cross_attention.py:777-779): overrides of
    different threads never see each other.
  * a module's own ``precision`` attribute (``CrossModalFusion``, ``MTCCMBertForMMTokenClassificationCRF``, the
    pipelines): applied as such an override for the duration of that module's forward; ``nn.DataParallel`` replicas
    inherit it with the rest of the module's attributes.
  * ``set_precision(mode)`` -- the process-wide default, configuration to be set before work is in flight.
"""
from __future__ import annotations

import threading

import torch

_MODES = ('bf16', 'fp32')
_default = 'bf16'
_tls = threading.local()


def _check(mode: str) -> str:
    if mode not in _MODES:
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {mode!r}")
    return mode


def set_precision(mode: str) -> None:
    """Process-wide default (used by every thread that has no override active)."""
    global _default
    _default = _check(mode)


def get_precision() -> str:
    return getattr(_tls, 'mode', None) or _default


class precision:
    """Context manager: thread-local precision override; ``None`` leaves the current mode in place."""

    def __init__(self, mode):
        self.mode = None if mode is None else _check(mode)

    def __enter__(self):
        self.prev = getattr(_tls, 'mode', None)
        if self.mode is not None:
            _tls.mode = self.mode
        return self

    def __exit__(self, *exc):
        _tls.mode = self.prev
        return False


def compute_dtype() -> torch.dtype:
    return torch.bfloat16 if get_precision() == 'bf16' else torch.float32
