"""Drop-in for ``torchcrf.CRF`` as the reference uses it (CMIM:3, 911-912, 1047-1057).

Same constructor (``CRF(num_tags, batch_first=False)``), parameter names (``start_transitions``,
``end_transitions``, ``transitions`` -- initialised U(-0.1, 0.1) like pytorch-crf), ``forward(emissions,
tags, mask=None, reduction='sum')`` and ``decode(emissions, mask=None) -> List[List[int]]``, and the same
``ValueError``s from input validation.  The arithmetic is one CUDA kernel per call
(``icka_viterbi_decode`` / ``icka_crf_llh_fwd``); ``decode`` does a single device->host copy of the
``[B,S]`` int32 tags and ``[B]`` lengths instead of pytorch-crf's ``.item()`` per decoded token.

``forward`` is differentiable: when autograd is recording, the log-likelihood is one
``icka_b200.autograd.CrfLlhFn`` node whose backward is the forward-backward kernel ``icka_crf_llh_bwd``
(gradients for the emissions and the three parameter tensors, as the reference trains them through
``loss = -crf(..., reduction='token_mean')``, CMIM:1047-1048).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import ops
from .autograd import CrfLlhFn


class CRF(nn.Module):
    def __init__(self, num_tags: int, batch_first: bool = False) -> None:
        if num_tags <= 0:
            raise ValueError(f'invalid number of tags: {num_tags}')
        super().__init__()
        self.num_tags = num_tags
        self.batch_first = batch_first
        self.start_transitions = nn.Parameter(torch.empty(num_tags))
        self.end_transitions = nn.Parameter(torch.empty(num_tags))
        self.transitions = nn.Parameter(torch.empty(num_tags, num_tags))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.uniform_(self.start_transitions, -0.1, 0.1)
        nn.init.uniform_(self.end_transitions, -0.1, 0.1)
        nn.init.uniform_(self.transitions, -0.1, 0.1)

    def __repr__(self) -> str:
        return f'{self.__class__.__name__}(num_tags={self.num_tags})'

    # -- validation: same conditions and ValueError as pytorch-crf's _validate ---------------------
    def _validate(self, emissions, tags=None, mask=None) -> None:
        capturing = emissions.is_cuda and torch.cuda.is_current_stream_capturing()
        if emissions.dim() != 3:
            raise ValueError(f'emissions must have dimension of 3, got {emissions.dim()}')
        if emissions.size(2) != self.num_tags:
            raise ValueError(
                f'expected last dimension of emissions is {self.num_tags}, '
                f'got {emissions.size(2)}')
        if tags is not None and emissions.shape[:2] != tags.shape:
            raise ValueError(
                'the first two dimensions of emissions and tags must match, '
                f'got {tuple(emissions.shape[:2])} and {tuple(tags.shape)}')
        if mask is not None:
            if emissions.shape[:2] != mask.shape:
                raise ValueError(
                    'the first two dimensions of emissions and mask must match, '
                    f'got {tuple(emissions.shape[:2])} and {tuple(mask.shape)}')
            first = mask[:, 0] if self.batch_first else mask[0]
            # (the two value checks read device memory back: not possible while a CUDA graph is being captured -- the
            # warm-up passes of icka_b200.graphs.CapturedStep have run them on the same tensors)
            if not capturing and not bool(first.all()):
                raise ValueError('mask of the first timestep must all be on')
        if tags is not None and tags.numel() and not capturing and bool(((tags < 0) | (tags >= self.num_tags)).any()):
            # pytorch-crf indexes start_transitions / transitions / emissions with every tag id (masked positions too):
            # an id outside [0, num_tags) -- e.g. an ignore-index of -100 on padding -- is an IndexError there
            raise IndexError(f'tags must lie in [0, {self.num_tags}); got values in '
                             f'[{int(tags.min())}, {int(tags.max())}]')

    def _batch_first(self, emissions, tags, mask):
        if not self.batch_first:
            emissions = emissions.transpose(0, 1)
            tags = None if tags is None else tags.transpose(0, 1)
            mask = None if mask is None else mask.transpose(0, 1)
        e = emissions.float().contiguous()
        if mask is None:
            m = None
        elif mask.dtype == torch.uint8:
            m = mask.contiguous()
        elif mask.dtype == torch.bool:
            m = mask.contiguous().view(torch.uint8)
        else:
            m = (mask != 0).contiguous().view(torch.uint8)
        y = None if tags is None else tags.long().contiguous()
        return e, y, m

    def _params(self):
        return (self.start_transitions.detach().float().contiguous(),
                self.end_transitions.detach().float().contiguous(),
                self.transitions.detach().float().contiguous())

    def forward(self, emissions: torch.Tensor, tags: torch.LongTensor, mask: Optional[torch.Tensor] = None,
                reduction: str = 'sum') -> torch.Tensor:
        """Log-likelihood of ``tags`` (the reference negates it, CMIM:1047-1048)."""
        self._validate(emissions, tags=tags, mask=mask)
        if reduction not in ('none', 'sum', 'mean', 'token_mean'):
            raise ValueError(f'invalid reduction: {reduction}')
        e, y, m = self._batch_first(emissions, tags, mask)
        params = (self.start_transitions, self.end_transitions, self.transitions)
        if torch.is_grad_enabled() and (e.requires_grad or any(p.requires_grad for p in params)):
            llh = CrfLlhFn.apply(e, *(p.float().contiguous() for p in params), y, m)
        else:
            llh = ops.crf_llh(e, y, m, *self._params())
        if reduction == 'none':
            return llh
        if reduction == 'sum':
            return llh.sum()
        if reduction == 'mean':
            return llh.mean()
        n_tok = e.shape[0] * e.shape[1] if m is None else m.sum()
        return llh.sum() / n_tok

    def decode_tensors(self, emissions: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """Device-side result: (tags [B,S] int32, -1 beyond each length; lens [B] int32). No host sync."""
        e, _, m = self._batch_first(emissions, None, mask)
        return ops.viterbi(e, m, *self._params())

    def decode(self, emissions: torch.Tensor, mask: Optional[torch.Tensor] = None) -> List[List[int]]:
        self._validate(emissions, mask=mask)
        tags, lens = self.decode_tensors(emissions, mask)
        tags_h = tags.cpu()
        lens_h = lens.cpu().tolist()
        rows = tags_h.tolist()
        return [rows[b][:lens_h[b]] for b in range(len(rows))]
