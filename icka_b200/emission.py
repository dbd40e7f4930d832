"""Emission head between the fusion stack and the CRF (SURVEY 8f "next" row 1): the reference's

    self.lstm = nn.LSTM(input_size=H, hidden_size=H, batch_first=True, bidirectional=True)      CMIM:905-908
    self.classifier = torch.nn.Linear(H * 2, num_labels)                                        CMIM:910
    x, _ = self.lstm(result); emissions = self.classifier(x)                                    CMIM:1042-1043

``LSTM`` keeps nn.LSTM's constructor arguments (the subset the reference uses), parameter names
(``weight_ih_l0`` ... ``bias_hh_l0_reverse`` -> the same state_dict keys) and return value
``(output, (h_n, c_n))``; ``EmissionHead`` owns ``lstm`` + ``classifier`` under the reference's attribute names.
Inference runs on the persistent kernel (parameters detached); when autograd is recording, ``autograd.BiLstmFn`` /
``ClassifierFn`` take over (per-step kernels with backpropagation through time).

bf16 mode, H = 768:  x -> [icka_linear_fwd: Gx = x . W_ih^T + b for both directions, slice-ordered columns, bf16]
                       -> [icka_lstm_rec_fwd: ONE persistent weight-stationary tcgen05 kernel, all S steps]
                       -> [icka_emission_head_fwd]
otherwise (fp32 parity mode, other H): the per-step path  icka_linear_fwd (h . W_hh^T) + icka_lstm_cell_fwd.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import nn

from . import modules, ops
from .autograd import BiLstmFn, ClassifierFn, LinearFn
from .modules import _OperandCache

REC_H = 768          # hidden size the persistent kernel is built for (csrc/lstm_sm100.cu)
REC_CHUNK = 2048     # sentences per persistent launch (4 sentence tiles per CTA); larger batches are walked chunk by chunk
                     # so that the Gx buffer of a chunk (3.2 GB) is reused


def slice_order(H: int = REC_H, variant: int = 2) -> torch.Tensor:
    """perm[8H]: row of the slice-ordered weights -> row of cat(weight_*_l0, weight_*_l0_reverse).
    variant 2 (CTA pairs, 48-unit slices; columns 0..95 of a slice live in CTA 0, 96..191 in CTA 1):
      col = ((dir*16 + slice)*4 + blk)*48 + jg*16 + gate*4 + jj  <->  dir*4H + gate*H + slice*48 + blk*12 + jg*4 + jj
    variant 1 (single CTAs, 24-unit slices): the same with 32 slices of 2 blocks."""
    units = 48 if variant == 2 else 24
    n_slices, n_blk = H // units, units // 12
    d = torch.arange(2).view(2, 1, 1, 1, 1, 1)
    s = torch.arange(n_slices).view(1, n_slices, 1, 1, 1, 1)
    blk = torch.arange(n_blk).view(1, 1, n_blk, 1, 1, 1)
    jg = torch.arange(3).view(1, 1, 1, 3, 1, 1)
    g = torch.arange(4).view(1, 1, 1, 1, 4, 1)
    jj = torch.arange(4).view(1, 1, 1, 1, 1, 4)
    return (d * 4 * H + g * H + s * units + blk * 12 + jg * 4 + jj).reshape(-1)


class LSTM(nn.Module):
    """Single-layer nn.LSTM mirror (CMIM:905-908).  ``forward(x) -> (output [B,S,2H] fp32, (h_n, c_n) [2,B,H])``."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1, bias: bool = True,
                 batch_first: bool = False, dropout: float = 0.0, bidirectional: bool = False):
        super().__init__()
        if num_layers != 1 or not bias or dropout != 0.0 or not bidirectional:
            raise NotImplementedError('icka_b200.LSTM covers the reference configuration: one bidirectional layer '
                                      'with biases and no dropout (CMIM:905-908)')
        self.input_size, self.hidden_size, self.batch_first = input_size, hidden_size, batch_first
        self.num_layers, self.bidirectional = 1, True
        k = 1.0 / math.sqrt(hidden_size)
        for suffix in ('', '_reverse'):
            for name, shape in (('weight_ih_l0', (4 * hidden_size, input_size)),
                                ('weight_hh_l0', (4 * hidden_size, hidden_size)),
                                ('bias_ih_l0', (4 * hidden_size,)), ('bias_hh_l0', (4 * hidden_size,))):
                self.register_parameter(name + suffix, nn.Parameter(torch.empty(shape).uniform_(-k, k)))
        self._cache = _OperandCache()

    # ---- cached operands -------------------------------------------------------------------------
    def _params(self):
        return [getattr(self, n + s) for s in ('', '_reverse')
                for n in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0')]

    def _prepared(self, variant: int):
        """(W_ih [8H, I], bias [8H] = b_ih + b_hh, W_hh [8H, H]) for both directions, in the compute dtype; rows in the
        slice order of persistent-kernel ``variant`` (1 or 2), in PyTorch order for 0 (the per-step path)."""
        def build():
            wi = torch.cat([self.weight_ih_l0.detach(), self.weight_ih_l0_reverse.detach()]).contiguous()
            wh = torch.cat([self.weight_hh_l0.detach(), self.weight_hh_l0_reverse.detach()]).contiguous()
            b = ops.add_f32(torch.cat([self.bias_ih_l0.detach(), self.bias_ih_l0_reverse.detach()]).contiguous(),
                            torch.cat([self.bias_hh_l0.detach(), self.bias_hh_l0_reverse.detach()]).contiguous())
            if variant:
                perm = slice_order(self.hidden_size, variant).to(wi.device)
                wi, wh, b = wi[perm].contiguous(), wh[perm].contiguous(), b[perm].contiguous()
            if modules.get_precision() == 'bf16':
                wi, wh = ops.cast_bf16(wi), ops.cast_bf16(wh)
            return wi, b, wh
        return self._cache.get(f'variant{variant}', self._params(), build)

    def uses_persistent_kernel(self) -> bool:
        return modules.get_precision() == 'bf16' and self.hidden_size == REC_H and self.input_size % 8 == 0

    # ---- forward -----------------------------------------------------------------------------------
    def _recording(self, x: torch.Tensor) -> bool:
        """True when this call must build an autograd graph (BiLstmFn: per-step kernels with BPTT)."""
        return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))

    def _check_inference(self):
        pass

    def _train_forward(self, x: torch.Tensor) -> torch.Tensor:
        return BiLstmFn.apply(x, self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0,
                              self.weight_ih_l0_reverse, self.weight_hh_l0_reverse, self.bias_ih_l0_reverse,
                              self.bias_hh_l0_reverse, modules.get_precision() == 'bf16')

    def states(self, x: torch.Tensor, want_state: bool = False):
        """x [B,S,I] (batch-first) -> (y [B,S,2H] in the compute dtype, h_n, c_n | None).  On the persistent-kernel
        path y is a batch-first VIEW of the time-major [S,B,2H] tensor the kernel writes."""
        self._check_inference()
        B, S, I = x.shape
        H = self.hidden_size
        lp = modules.get_precision() == 'bf16'
        if self.uses_persistent_kernel():
            variant = ops.lstm_rec_variant(B)
            wi, b, wh = self._prepared(variant)
            xc = x.contiguous()
            x_tm = ops.cast_bf16_time_major(xc if xc.dtype in (torch.float32, torch.bfloat16) else xc.float())
            gx = ops.linear(x_tm.view(S * B, I), wi, b, out_dtype=torch.bfloat16)
            out = ops.lstm_rec(gx, wh, B, S, H, variant=variant, want_state=want_state)
            if want_state:
                return out[0].transpose(0, 1), out[1], out[2]
            return out.transpose(0, 1), None, None
        x2 = x.reshape(B * S, I)
        if lp:
            x2 = x2.contiguous() if x2.dtype == torch.bfloat16 else ops.cast_bf16(x2.float().contiguous())
        else:
            x2 = x2.float().contiguous()
        # per-step path
        wi, b, wh = self._prepared(0)
        cdt = torch.bfloat16 if lp else torch.float32
        gx = ops.linear(x2, wi, b, out_dtype=cdt).view(B, S, 8 * H)
        y = torch.empty(B, S, 2 * H, dtype=cdt, device=x.device)
        h_n = torch.empty(2, B, H, dtype=torch.float32, device=x.device) if want_state else None
        c_n = torch.zeros(2, B, H, dtype=torch.float32, device=x.device)
        for d in range(2):
            h_prev = torch.empty(B, H, dtype=cdt, device=x.device)
            h_next = torch.empty(B, H, dtype=cdt, device=x.device)
            c = c_n[d]
            w = wh[d * 4 * H:(d + 1) * 4 * H]
            for t in range(S):
                pos = S - 1 - t if d else t
                gates = None if t == 0 else ops.linear(h_prev, w, None, out_dtype=torch.float32)
                ops.lstm_cell(gates, gx[:, pos, d * 4 * H:(d + 1) * 4 * H], c, h_next, y[:, pos, d * H:(d + 1) * H],
                              h_n[d] if (want_state and t == S - 1) else None)
                h_prev, h_next = h_next, h_prev
        return y, h_n, (c_n if want_state else None)

    def forward(self, input: torch.Tensor, hx=None) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        if hx is not None:
            raise NotImplementedError('initial states are not supported (the reference passes none, CMIM:1042)')
        if input.dim() != 3:
            raise ValueError(f'LSTM: Expected input to be 3D (batched), got {input.dim()}D instead')
        if input.shape[-1] != self.input_size:
            raise RuntimeError(f'input.size(-1) must be equal to input_size. Expected {self.input_size}, '
                               f'got {input.shape[-1]}')
        x = input if self.batch_first else input.transpose(0, 1)
        B = x.shape[0]
        if self._recording(x):                                  # training: autograd node, final states not tracked
            y = self._train_forward(x.float())
            H = self.hidden_size
            h_n = torch.stack([y[:, -1, :H], y[:, 0, H:]]).detach()
            return (y if self.batch_first else y.transpose(0, 1)), (h_n, None)
        if B > REC_CHUNK and self.uses_persistent_kernel():
            parts = [self.states(x[b0:b0 + REC_CHUNK], want_state=True) for b0 in range(0, B, REC_CHUNK)]
            y = torch.cat([p[0].float() for p in parts], dim=0)
            h_n = torch.cat([p[1] for p in parts], dim=1)
            c_n = torch.cat([p[2] for p in parts], dim=1)
            return (y if self.batch_first else y.transpose(0, 1)), (h_n, c_n)
        y, h_n, c_n = self.states(x, want_state=True)
        y = y.float()
        return (y if self.batch_first else y.transpose(0, 1)), (h_n, c_n)


class EmissionHead(nn.Module):
    """``lstm`` + ``classifier`` of MTCCMBertForMMTokenClassificationCRF (CMIM:905-910) and their use at
    CMIM:1042-1043: ``forward(result [B,S,H]) -> emissions [B,S,num_labels]`` fp32."""

    def __init__(self, config, num_labels: int = 2):
        super().__init__()
        self.lstm = LSTM(input_size=config.hidden_size, hidden_size=config.hidden_size, batch_first=True,
                         bidirectional=True)
        self.classifier = nn.Linear(config.hidden_size * 2, num_labels)

    def forward(self, result: torch.Tensor) -> torch.Tensor:
        B, S, _ = result.shape
        if self.lstm._recording(result) or (torch.is_grad_enabled() and self.classifier.weight.requires_grad):
            y = self.lstm._train_forward(result.float())                     # [B,S,2H] fp32, autograd node (BPTT)
            if self.classifier.out_features > 16 or y.shape[-1] % 8:
                return LinearFn.apply(y.reshape(B * S, -1), self.classifier.weight, self.classifier.bias).view(B, S, -1)
            y_tm = y.transpose(0, 1)
            if y_tm.is_contiguous():          # the fused step kernels write the sequence time-major: classify it in place
                return ClassifierFn.apply(y_tm.reshape(S * B, -1), self.classifier.weight, self.classifier.bias,
                                          S).view(B, S, -1)
            return ClassifierFn.apply(y.reshape(B * S, -1), self.classifier.weight, self.classifier.bias).view(B, S, -1)
        if B > REC_CHUNK and self.lstm.uses_persistent_kernel():
            out = torch.empty(B, S, self.classifier.out_features, dtype=torch.float32, device=result.device)
            for b0 in range(0, B, REC_CHUNK):
                out[b0:b0 + REC_CHUNK] = self.forward(result[b0:b0 + REC_CHUNK])
            return out
        y, _, _ = self.lstm.states(result)
        w = self.classifier.weight.detach().float().contiguous()
        b = self.classifier.bias.detach().float().contiguous()
        if y.is_contiguous():
            e = ops.emission_head(y.view(B * S, -1), w, b)
        else:                                                   # time-major states of the persistent kernel
            e = ops.emission_head(y.transpose(0, 1).view(S * B, -1), w, b, time_major_S=S)
        return e.view(B, S, -1)
