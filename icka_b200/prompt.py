"""Prompt mapping networks + prefix assembly (SURVEY 8f "next" row 4): the remaining non-encoder compute between the
image->text encoders and the gate of ``MTCCMBertForMMTokenClassificationCRF``:

    self.mapping_network_alignment / self.mapping_network_vision / self.lastproj            CMIM:913-930
    Alignment_prompt, prefix_vision, prefix_emb (cat + lastproj), prompt_mask               CMIM:995-1009

``PromptMapping`` keeps the reference's attribute names and ``nn.Sequential`` indices, so the checkpoint keys
(``mapping_network_alignment.1.weight`` ... ``lastproj.bias``) load unchanged.  When autograd is recording (mode='train',
CMIM:1046-1048) the same five layers are ``autograd.DenseActFn`` nodes (tcgen05 forward, dgrad and wgrad GEMMs; tanh'
through ``icka_act_bwd``) and the Dropout(0.3) layers (CMIM:915, :918, :923, :926) are Philox ``DropoutFn`` nodes whose
masks backward regenerates; a module in training mode called with autograd disabled raises instead of silently skipping
the dropout.

Five tensor-core GEMMs (``icka_linear_fwd``), the Tanh fused into the epilogue of the first layer of each network
(ICKA_ACT_TANH).  756 * 5 = 3780 is not a multiple of 8 (16-byte bf16 rows for TMA): the bf16 operand copies are
zero-padded once to 3784 columns / rows -- tanh(0) = 0 meets zero weights, so the padding is exact.  Both networks
write straight into the two halves of one [B, 10 * H] buffer (no torch.cat), which lastproj reads as [B * 10, H].
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import nn

from . import modules, ops
from ._lib import ACT_NONE, ACT_TANH
from .autograd import DenseActFn, DropoutFn
from .modules import _OperandCache, _seed


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class PromptMapping(nn.Module):
    def __init__(self, config, prompt_len: int = 5, inner: int = 756, vision_dim: int = 2048, out_dim: int = 1024):
        super().__init__()
        H = config.hidden_size
        self.hidden_size, self.prompt_len, self.out_dim = H, prompt_len, out_dim
        self.mapping_network_alignment = nn.Sequential(                                      # CMIM:914-920
            nn.Dropout(p=0.3), nn.Linear(H, inner * prompt_len, bias=True), nn.Tanh(), nn.Dropout(p=0.3),
            nn.Linear(inner * prompt_len, H * prompt_len, bias=True))
        self.mapping_network_vision = nn.Sequential(                                         # CMIM:922-928
            nn.Dropout(p=0.3), nn.Linear(vision_dim, inner * prompt_len, bias=True), nn.Tanh(), nn.Dropout(p=0.3),
            nn.Linear(inner * prompt_len, H * prompt_len, bias=True))
        self.lastproj = nn.Linear(H, out_dim)                                                # CMIM:930
        self._cache = _OperandCache()

    def _operands(self, net: nn.Sequential, key: str):
        """(W1 [inner_p, in], b1 [inner_p], W2 [out, inner_p], b2) in the compute dtype, inner padded to 8."""
        l1, l2 = net[1], net[4]

        def build():
            inner = l1.out_features
            ip = _pad8(inner)
            w1 = torch.zeros(ip, l1.in_features, device=l1.weight.device)
            w1[:inner] = l1.weight.detach()
            b1 = torch.zeros(ip, device=l1.weight.device)
            b1[:inner] = l1.bias.detach()
            w2 = torch.zeros(l2.out_features, ip, device=l1.weight.device)
            w2[:, :inner] = l2.weight.detach()
            if modules.get_precision() == 'bf16':
                w1, w2 = ops.cast_bf16(w1), ops.cast_bf16(w2)
            return w1, b1, w2, l2.bias.detach().float().contiguous()
        return self._cache.get(key, (l1.weight, l1.bias, l2.weight, l2.bias), build)

    def forward(self, clip_features: torch.Tensor, visual_embeds_mean: torch.Tensor,
                input_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """clip_features [B,1,H] or [B,H], visual_embeds_mean [B,vision_dim], input_mask [B,L]
        -> (prefix_emb [B, 2*prompt_len, out_dim] fp32, prompt_mask [B, 2*prompt_len])."""
        rec = torch.is_grad_enabled() and (clip_features.requires_grad or visual_embeds_mean.requires_grad
                                           or any(p.requires_grad for p in self.parameters()))
        if rec:
            return self._forward_recorded(clip_features, visual_embeds_mean, input_mask)
        if self.training:
            raise NotImplementedError('PromptMapping in training mode (Dropout(0.3) of CMIM:915, 918) runs on the '
                                      'autograd-recording path only: call .eval() for forward-only use')
        B, H, P = clip_features.shape[0], self.hidden_size, self.prompt_len
        lp = modules.get_precision() == 'bf16'
        cdt = torch.bfloat16 if lp else torch.float32

        def operand(x):
            x = x.reshape(B, -1).float().contiguous()
            return ops.cast_bf16(x) if lp else x

        both = torch.empty(B, 2 * P * H, dtype=cdt, device=clip_features.device)    # [vision | alignment], CMIM:1002
        for x, net, key, half in ((visual_embeds_mean, self.mapping_network_vision, 'vision', 0),
                                  (clip_features, self.mapping_network_alignment, 'alignment', 1)):
            w1, b1, w2, b2 = self._operands(net, key)
            hidden = ops.linear(operand(x), w1, b1, act=ACT_TANH, out_dtype=cdt)
            ops.linear(hidden, w2, b2, act=ACT_NONE, out=both[:, half * P * H:(half + 1) * P * H])
        prefix = both.view(B * 2 * P, H)
        if H != 1024:                                                                  # CMIM:1003-1004
            w = self.lastproj.weight.detach()
            wl = self._cache.get('lastproj', (self.lastproj.weight,),
                                 lambda: ops.cast_bf16(w.contiguous()) if lp else w.float().contiguous())
            prefix = ops.linear(prefix, wl, self.lastproj.bias.detach().float().contiguous(), out_dtype=torch.float32)
            prefix = prefix.view(B, 2 * P, self.out_dim)
        else:
            prefix = prefix.float().view(B, 2 * P, H)
        return prefix, self._prompt_mask(input_mask)

    def _prompt_mask(self, input_mask: torch.Tensor) -> torch.Tensor:
        first = input_mask[:, :1]
        return torch.cat([first.repeat(1, self.prompt_len), first.repeat(1, self.prompt_len)], dim=1)   # CMIM:1007-1009

    def _forward_recorded(self, clip_features, visual_embeds_mean, input_mask):
        """Training pass (autograd recording).  bf16: the hidden width 3780 is zero-padded to 3840 -- the backward GEMMs read
        both operands MN-major in 64-element chunks; tanh(0) = 0 meets zero weight columns, so the padding is exact and its
        gradient rows are sliced away by autograd's view of the pad."""
        B, H, P = clip_features.shape[0], self.hidden_size, self.prompt_len
        bf = modules.get_precision() == 'bf16'
        pad = torch.nn.functional.pad

        def run(x, net):
            l1, l2 = net[1], net[4]
            p1, p2 = (net[0].p, net[3].p) if self.training else (0.0, 0.0)
            x32 = x.reshape(B, -1).float().contiguous()
            if p1 > 0:
                x32 = DropoutFn.apply(x32, p1, _seed())
            inner = l1.out_features
            extra = ((inner + 63) // 64 * 64 - inner) if bf else 0
            w1 = pad(l1.weight, (0, 0, 0, extra)) if extra else l1.weight
            b1 = pad(l1.bias, (0, extra)) if extra else l1.bias
            w2 = pad(l2.weight, (0, extra)) if extra else l2.weight
            hidden = DenseActFn.apply(x32, w1, b1, None, ACT_TANH, bf)
            if p2 > 0:
                hidden = DropoutFn.apply(hidden, p2, _seed())
            return DenseActFn.apply(hidden, w2, l2.bias, None, ACT_NONE, bf)

        prefix_vision = run(visual_embeds_mean, self.mapping_network_vision).view(B, P, H)          # CMIM:998-999
        alignment = run(clip_features, self.mapping_network_alignment).view(B, P, H)                 # CMIM:995
        prefix = torch.cat([prefix_vision, alignment], dim=1)                                        # CMIM:1002
        if H != 1024:                                                                                # CMIM:1003-1004
            prefix = DenseActFn.apply(prefix.reshape(B * 2 * P, H), self.lastproj.weight, self.lastproj.bias, None,
                                      ACT_NONE, bf).view(B, 2 * P, self.out_dim)
        return prefix, self._prompt_mask(input_mask)
