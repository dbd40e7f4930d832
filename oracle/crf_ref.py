"""CPU restatement of the linear-chain CRF the reference uses (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED.  The reference does ``from torchcrf import CRF``
(Cross_Modal_Interaction_Module.py:3) and calls ``CRF(num_tags, batch_first=True)``
(:911-912), ``crf(emissions, tags=, mask=, reduction='token_mean')`` (:1047-1048,
:1052-1053) and ``crf.decode(emissions, mask=)`` (:1051, :1056).  ``torchcrf`` is the
PyPI package ``pytorch-crf`` (kmkurn); it is not vendored under /root/reference, no version
is pinned anywhere in the reference (``reduction='token_mean'`` implies >= 0.7.0; latest
upstream is 0.7.2), it is not installed here and cannot be (no network), and the reference
has no test or golden vector at this boundary.  This file restates the package's published
algorithm -- same fp32 operation order, same tie-breaking -- and is
anchored by: a brute-force path enumerator (``brute_force_*``), score self-consistency
checks, and ``torch.logsumexp``/autograd cross-checks in ``tests/test_oracle_crf.py``.

Conventions (pytorch-crf): ``transitions[i, j]`` is the score of moving from tag i to tag j.
All arithmetic is done in the dtype of the inputs (fp32 in ICKA).
"""
from __future__ import annotations

import itertools
from typing import List, Optional

import torch


def validate(emissions: torch.Tensor, num_tags: int, tags: Optional[torch.Tensor] = None,
             mask: Optional[torch.Tensor] = None, batch_first: bool = True) -> None:
    """Input checks of pytorch-crf's ``_validate`` (ValueError text mirrors upstream)."""
    if emissions.dim() != 3:
        raise ValueError(f'emissions must have dimension of 3, got {emissions.dim()}')
    if emissions.size(2) != num_tags:
        raise ValueError(
            f'expected last dimension of emissions is {num_tags}, '
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
        no_empty_seq = not batch_first and mask[0].all()
        no_empty_seq_bf = batch_first and mask[:, 0].all()
        if not no_empty_seq and not no_empty_seq_bf:
            raise ValueError('mask of the first timestep must all be on')


def viterbi_decode(emissions: torch.Tensor, mask: Optional[torch.Tensor],
                   start: torch.Tensor, end: torch.Tensor, trans: torch.Tensor) -> List[List[int]]:
    """Best tag path per sentence; ``emissions`` is batch-first ``[B, S, T]``, ``mask`` ``[B, S]``.

    Follows pytorch-crf ``_viterbi_decode`` step for step:
      score_0[j]   = start[j] + e_0[j]
      cand[i][j]   = (score[i] + trans[i][j]) + e_t[j]        (two fp32 roundings, this order)
      next, bp     = max / argmax over i (first maximal index wins)
      score        = mask_t ? next : score                     (bp recorded for every t)
      score       += end ;  last = argmax_j score ;  walk bp[len-2 .. 0]
    """
    B, S, T = emissions.shape
    if mask is None:
        mask = torch.ones(B, S, dtype=torch.bool)
    mask = mask.bool()
    e = emissions.transpose(0, 1)          # time-major, as upstream does for batch_first
    m = mask.transpose(0, 1)
    score = start + e[0]                   # [B, T]
    history = []
    for t in range(1, S):
        cand = (score.unsqueeze(2) + trans) + e[t].unsqueeze(1)   # [B, T_from, T_to]
        nxt, idx = cand.max(dim=1)
        score = torch.where(m[t].unsqueeze(1), nxt, score)
        history.append(idx)
    score = score + end
    seq_ends = m.long().sum(dim=0) - 1
    out: List[List[int]] = []
    for b in range(B):
        last = int(score[b].max(dim=0)[1])
        path = [last]
        for bp in reversed(history[:int(seq_ends[b])]):
            last = int(bp[b][path[-1]])
            path.append(last)
        path.reverse()
        out.append(path)
    return out


def path_score(emissions, tags, mask, start, end, trans) -> torch.Tensor:
    """Gold-path score (numerator), pytorch-crf ``_compute_score``; batch-first inputs -> [B]."""
    B, S, T = emissions.shape
    e = emissions.transpose(0, 1)
    y = tags.transpose(0, 1)
    m = mask.transpose(0, 1).to(emissions.dtype)
    ar = torch.arange(B)
    score = start[y[0]] + e[0, ar, y[0]]
    for t in range(1, S):
        score = score + trans[y[t - 1], y[t]] * m[t]
        score = score + e[t, ar, y[t]] * m[t]
    seq_ends = mask.transpose(0, 1).long().sum(dim=0) - 1
    last = y[seq_ends, ar]
    return score + end[last]


def log_partition(emissions, mask, start, end, trans) -> torch.Tensor:
    """log Z (denominator), pytorch-crf ``_compute_normalizer``; batch-first inputs -> [B]."""
    B, S, T = emissions.shape
    e = emissions.transpose(0, 1)
    m = mask.transpose(0, 1).bool()
    score = start + e[0]
    for t in range(1, S):
        nxt = torch.logsumexp((score.unsqueeze(2) + trans) + e[t].unsqueeze(1), dim=1)
        score = torch.where(m[t].unsqueeze(1), nxt, score)
    return torch.logsumexp(score + end, dim=1)


def log_likelihood(emissions, tags, mask, start, end, trans, reduction: str = 'sum') -> torch.Tensor:
    """pytorch-crf ``CRF.forward`` (returns +llh; ICKA negates it, CMIM:1047)."""
    if reduction not in ('none', 'sum', 'mean', 'token_mean'):
        raise ValueError(f'invalid reduction: {reduction}')
    if mask is None:
        mask = torch.ones(emissions.shape[:2], dtype=torch.bool)
    llh = path_score(emissions, tags, mask, start, end, trans) - \
        log_partition(emissions, mask, start, end, trans)
    if reduction == 'none':
        return llh
    if reduction == 'sum':
        return llh.sum()
    if reduction == 'mean':
        return llh.mean()
    return llh.sum() / mask.to(emissions.dtype).sum()


# ---------------------------------------------------------------------------------------------
# brute force (tiny T, S only) -- anchors the restatement on first principles
# ---------------------------------------------------------------------------------------------

def _seq_score(e, path, start, end, trans):
    s = float(start[path[0]]) + float(e[0, path[0]])
    for t in range(1, len(path)):
        s += float(trans[path[t - 1], path[t]]) + float(e[t, path[t]])
    return s + float(end[path[-1]])


def brute_force_best(e: torch.Tensor, length: int, start, end, trans):
    """Enumerate all T**length paths of one sentence in float64; returns (best_score, [paths at max])."""
    T = e.shape[1]
    e = e.double()
    start, end, trans = start.double(), end.double(), trans.double()
    best, arg = None, []
    for path in itertools.product(range(T), repeat=length):
        s = _seq_score(e, path, start, end, trans)
        if best is None or s > best + 1e-12:
            best, arg = s, [list(path)]
        elif abs(s - best) <= 1e-12:
            arg.append(list(path))
    return best, arg


def brute_force_logZ(e: torch.Tensor, length: int, start, end, trans) -> float:
    T = e.shape[1]
    e = e.double()
    start, end, trans = start.double(), end.double(), trans.double()
    scores = [_seq_score(e, p, start, end, trans) for p in itertools.product(range(T), repeat=length)]
    return float(torch.logsumexp(torch.tensor(scores, dtype=torch.float64), dim=0))
