"""CPU restatement of the emission head between fusion and CRF (TEST INFRASTRUCTURE ONLY; SURVEY 8f "next" row 1).

  bilstm          `self.lstm = nn.LSTM(input_size=H, hidden_size=H, batch_first=True, bidirectional=True)`
                  (Cross_Modal_Interaction_Module.py:905-908), called as `x, _ = self.lstm(result)` (CMIM:1042) on the
                  FULL 128 positions (no packing: the backward direction starts on the padding)
  emission_head   + `emissions = self.classifier(x)` with `nn.Linear(2H, num_labels)` (CMIM:910, 1043)

The arithmetic of the reference here is torch's own nn.LSTM / nn.Linear; this file spells the cell out step by step
(PyTorch gate order i, f, g, o) so the CUDA kernels have an op-for-op statement to be compared with.
Pinned: tests/test_oracle_lstm.py checks it against torch.nn.LSTM / nn.Linear themselves -- the very calls the
reference makes -- on seeded inputs (output sequence, h_n, c_n, emissions).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def lstm_direction(x: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor,
                   reverse: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x [B,S,I] -> (states [B,S,H], h_n [B,H], c_n [B,H]); zero initial state."""
    B, S, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_zeros(B, S, H)
    steps = range(S - 1, -1, -1) if reverse else range(S)
    for t in steps:
        gates = x[:, t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, g, o = gates.chunk(4, dim=1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c = f * c + i * g
        h = o * torch.tanh(c)
        out[:, t] = h
    return out, h, c


def bilstm(x: torch.Tensor, p: Dict[str, torch.Tensor]):
    """p: nn.LSTM's own parameter names (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, *_reverse).
    -> (output [B,S,2H], (h_n [2,B,H], c_n [2,B,H])) as nn.LSTM(batch_first=True, bidirectional=True) returns."""
    f, hf, cf = lstm_direction(x, p['weight_ih_l0'], p['weight_hh_l0'], p['bias_ih_l0'], p['bias_hh_l0'], False)
    b, hb, cb = lstm_direction(x, p['weight_ih_l0_reverse'], p['weight_hh_l0_reverse'], p['bias_ih_l0_reverse'],
                               p['bias_hh_l0_reverse'], True)
    return torch.cat([f, b], dim=2), (torch.stack([hf, hb]), torch.stack([cf, cb]))


def emission_head(x: torch.Tensor, lstm_params: Dict[str, torch.Tensor], w_cls: torch.Tensor, b_cls: torch.Tensor):
    """CMIM:1042-1043: emissions [B,S,T]."""
    out, _ = bilstm(x, lstm_params)
    return out @ w_cls.t() + b_cls
