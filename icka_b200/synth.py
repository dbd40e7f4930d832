"""Seeded synthetic Twitter-shaped inputs for the fusion + Viterbi path.

There is no network for datasets or checkpoints, so tests and ``bench.py`` feed tensors of the shapes
the reference's driver builds (My_cross_attention.py:250-472, 798-817): BERT-like text states,
a post-ReLU ResNet-152 grid, a CLIP feature, the second encoder's token states, prefix masks with
tweet-shaped lengths, and emission scores for the CRF.  Base seed 19260817 is the reference's own
default (My_cross_attention.py:577-580); rank r uses ``seed + r``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

BASE_SEED = 19260817


@dataclass
class Shape:
    S: int = 128          # text length (max_seq_length, My_cross_attention.py:375-383)
    R: int = 49           # 7x7 ResNet-152 grid (CMIM:956)
    H: int = 768
    heads: int = 12
    inter: int = 3072
    region_dim: int = 2048
    clip_dim: int = 512
    T: int = 15           # 14 labels + pad 0 (My_cross_attention.py:215, 641)
    L: int = 1            # layer_num1 (ctor default CMIM:888; script default 5, MCA:603)
    eps: float = 1e-12    # BertConfig.layer_norm_eps default (CMIM:60)


STD = Shape()
HIRES = Shape(S=256, R=196)


def _layer_normed(x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.layer_norm(x, (x.shape[-1],))


def lengths(B: int, S: int, gen: torch.Generator, median: float = 28.0, sigma: float = 0.45) -> torch.Tensor:
    """Tweet-shaped sentence lengths: round(LogNormal(ln median, sigma)) clipped to [3, S]."""
    z = torch.randn(B, generator=gen)
    return torch.clamp(torch.round(torch.exp(math.log(median) + sigma * z)), 3, S).long()


def prefix_mask(lens: torch.Tensor, S: int) -> torch.Tensor:
    return (torch.arange(S).unsqueeze(0) < lens.unsqueeze(1)).long()


def fusion_inputs(B: int, shape: Shape = STD, seed: int = BASE_SEED, median_len: float = 28.0):
    """CPU fp32 tensors for one batch of the fusion segment."""
    g = torch.Generator().manual_seed(seed)
    grid = int(round(math.sqrt(shape.R)))
    assert grid * grid == shape.R
    text = _layer_normed(torch.randn(B, shape.S, shape.H, generator=g))
    regions = torch.relu(torch.randn(B, shape.region_dim, grid, grid, generator=g)) * 0.5
    clip = torch.randn(B, 1, shape.clip_dim, generator=g)
    tok = _layer_normed(torch.randn(B, shape.S, shape.H, generator=g))
    lens = lengths(B, shape.S, g, median_len)
    text_mask = prefix_mask(lens, shape.S)
    img_mask = torch.ones(B, shape.R, dtype=torch.long)      # MCA:373 sets the first 49 entries to 1
    return dict(text_states=text, visual_embeds_att=regions, clip_features=clip, token_embedding=tok,
                img_mask=img_mask, text_mask=text_mask, lens=lens)


def crf_params(T: int, seed: int = BASE_SEED, kind: str = 'uniform'):
    """``uniform`` = pytorch-crf's init U(-0.1, 0.1); ``normal`` = N(0,1) stress set."""
    g = torch.Generator().manual_seed(seed + 7)
    if kind == 'uniform':
        f = lambda *s: torch.rand(*s, generator=g) * 0.2 - 0.1
    else:
        f = lambda *s: torch.randn(*s, generator=g)
    return dict(start_transitions=f(T), end_transitions=f(T), transitions=f(T, T))


def emissions(B: int, S: int, T: int, seed: int = BASE_SEED, kind: str = 'normal') -> torch.Tensor:
    """fp32 emission scores [B,S,T].

    normal    N(0, 2^2)
    ties      quantised to multiples of 0.25 -> many exact ties (first-index rule decides)
    near_ties pairs of tags one ulp apart -> ties appear only after adding e[t][j] (rounding)
    """
    g = torch.Generator().manual_seed(seed + 13)
    e = torch.randn(B, S, T, generator=g) * 2.0
    if kind == 'ties':
        e = torch.round(e * 4.0) / 4.0
    elif kind == 'near_ties':
        base = torch.round(e[..., ::2] * 8.0) / 8.0
        e[..., ::2] = base
        n_odd = e[..., 1::2].shape[-1]
        e[..., 1::2] = torch.nextafter(base[..., :n_odd], torch.full_like(base[..., :n_odd], 1e9))
    elif kind != 'normal':
        raise ValueError(kind)
    return e.float().contiguous()


def crf_batch(B: int, shape: Shape = STD, seed: int = BASE_SEED, kind: str = 'normal',
              median_len: float = 28.0):
    g = torch.Generator().manual_seed(seed + 29)
    lens = lengths(B, shape.S, g, median_len)
    mask = prefix_mask(lens, shape.S).bool()
    e = emissions(B, shape.S, shape.T, seed, kind)
    tags = (torch.randint(1, shape.T, (B, shape.S), generator=g) if shape.T > 1
            else torch.zeros(B, shape.S, dtype=torch.long)) * mask.long()
    return dict(emissions=e, mask=mask, tags=tags, lens=lens)


def forbid_cells(emissions: torch.Tensor, seed: int, frac: float = 0.3, value: float = float('-inf')) -> torch.Tensor:
    """Constrained decoding as callers do it: a random subset of (step, tag) cells is forbidden (-inf or -10000)."""
    g = torch.Generator().manual_seed(seed)
    e = emissions.clone()
    e[torch.rand(e.shape, generator=g) < frac] = value
    return e
