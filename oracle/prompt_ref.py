"""CPU restatement of the prompt mapping networks + prefix assembly (TEST INFRASTRUCTURE ONLY).

  modules   mapping_network_alignment / mapping_network_vision = Sequential(Dropout 0.3, Linear(in, 756*5), Tanh,
            Dropout 0.3, Linear(756*5, H*5)), lastproj = Linear(H, 1024)   Cross_Modal_Interaction_Module.py:913-930
  forward   Alignment_prompt, prefix_vision, prefix_emb, prompt_mask                           CMIM:995-1009

Inference form (dropout is the identity in eval mode).  Parameters use the reference's own state_dict keys
(`mapping_network_alignment.1.weight`, ...), drawn from a seeded generator with nn.Linear's default bounds so that
tests can rebuild them anywhere.  Pinned: tests/golden/prompt_prefix.npz holds the outputs of the reference's OWN
modules (constructed through oracle/reference_shim.py by oracle/make_golden_prompt.py) for these parameters.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

PROMPT_LEN = 5          # CMIM:913
INNER = 756             # CMIM:916, 924
OUT_DIM = 1024          # CMIM:930, 1004


def make_params(H: int = 768, vision_dim: int = 2048, seed: int = 0, prompt_len: int = PROMPT_LEN,
                inner: int = INNER) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f):
        k = 1.0 / math.sqrt(in_f)
        return (torch.empty(out_f, in_f).uniform_(-k, k, generator=g), torch.empty(out_f).uniform_(-k, k, generator=g))

    p = {}
    for name, in_f in (('mapping_network_alignment', H), ('mapping_network_vision', vision_dim)):
        p[f'{name}.1.weight'], p[f'{name}.1.bias'] = lin(inner * prompt_len, in_f)
        p[f'{name}.4.weight'], p[f'{name}.4.bias'] = lin(H * prompt_len, inner * prompt_len)
    p['lastproj.weight'], p['lastproj.bias'] = lin(OUT_DIM, H)
    return p


def prompt_prefix(clip_features: torch.Tensor, visual_embeds_mean: torch.Tensor, input_mask: torch.Tensor,
                  p: Dict[str, torch.Tensor], prompt_len: int = PROMPT_LEN, drop=None):
    """clip_features [B,1,H] (output of the image->text encoders), visual_embeds_mean [B,2048], input_mask [B,L]
    -> (prefix_emb [B, 2*prompt_len, 1024], prompt_mask [B, 2*prompt_len]).
    ``drop`` (training mode, CMIM:915, :918, :923, :926): dict(p=0.3, masks={'<network>.0': keep [B,in], '<network>.3':
    keep [B,inner]}) -- 0/1 keep masks applied as nn.Dropout would (x * keep / (1 - p)), so a test can replay the masks
    the CUDA kernels drew."""
    B = clip_features.shape[0]

    def dropped(x, key):
        if drop is None:
            return x
        return x * drop['masks'][key].to(x.dtype).view(x.shape) / (1.0 - drop['p'])

    def mlp(x, name):
        h = torch.tanh(dropped(x, f'{name}.0') @ p[f'{name}.1.weight'].t() + p[f'{name}.1.bias'])
        return dropped(h, f'{name}.3') @ p[f'{name}.4.weight'].t() + p[f'{name}.4.bias']

    alignment = mlp(clip_features, 'mapping_network_alignment').unsqueeze(1).view(B, prompt_len, -1)   # CMIM:995
    vision = mlp(visual_embeds_mean, 'mapping_network_vision').reshape(B, prompt_len, -1)               # CMIM:998-999
    prefix = torch.cat([vision, alignment], dim=1)                                                      # CMIM:1002
    if prefix.size(2) != 1024:                                                                          # CMIM:1003-1004
        prefix = prefix @ p['lastproj.weight'].t() + p['lastproj.bias']
    align_mask = input_mask[:, :1].repeat(1, prompt_len)                                                # CMIM:1007
    vision_mask = input_mask[:, :1].repeat(1, prompt_len)                                               # CMIM:1008
    return prefix, torch.cat([vision_mask, align_mask], dim=1)                                          # CMIM:1009


GOLDEN_CASES = {'std': dict(B=3, H=768, seed=51, L=20)}


def golden_inputs(B, H, L, seed):
    g = torch.Generator().manual_seed(seed + 1)
    clip = torch.randn(B, 1, H, generator=g)
    vmean = torch.relu(torch.randn(B, 2048, generator=g)) * 0.5
    mask = torch.ones(B, L, dtype=torch.long)
    mask[-1, 0] = 0                      # exercise the mask copy
    return clip, vmean, mask


def load_golden_case(name: str = 'std'):
    """(params, clip, vmean, mask, reference prefix, reference prompt_mask) of tests/golden/prompt_prefix.npz; inputs and
    parameters are rebuilt from their seeds and checked against the stored checksums."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden',
                             'prompt_prefix.npz'))
    c = GOLDEN_CASES[name]
    p = make_params(c['H'], seed=c['seed'])
    clip, vmean, mask = golden_inputs(c['B'], c['H'], c['L'], c['seed'])
    chk = g[f'{name}_checksum']
    assert abs(float(sum(v.double().abs().sum() for v in p.values())) - chk[0]) <= 1e-6 * chk[0], 'parameters drifted'
    assert abs(float(clip.double().abs().sum() + vmean.double().abs().sum()) - chk[1]) <= 1e-6 * chk[1], 'inputs drifted'
    return p, clip, vmean, mask, torch.from_numpy(g[f'{name}_prefix']), torch.from_numpy(g[f'{name}_mask'])
