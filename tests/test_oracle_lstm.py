"""The emission-head oracle (oracle/lstm_ref.py) against torch's own nn.LSTM / nn.Linear on the CPU -- the calls the
reference itself makes (CMIM:905-910, 1042-1043)."""
import pytest
import torch
from torch import nn

from oracle import lstm_ref


@pytest.mark.parametrize('B,S,I,H,T', [(3, 7, 16, 16, 5), (2, 128, 48, 48, 15), (1, 1, 8, 24, 3)])
def test_oracle_matches_torch_lstm(B, S, I, H, T):
    torch.manual_seed(B * 100 + S)
    lstm = nn.LSTM(input_size=I, hidden_size=H, batch_first=True, bidirectional=True).double()
    cls = nn.Linear(2 * H, T).double()
    x = torch.randn(B, S, I, dtype=torch.float64)
    with torch.no_grad():
        want, (hn, cn) = lstm(x)
        want_e = cls(want)
        p = {k: v.detach() for k, v in lstm.named_parameters()}
        got, (ghn, gcn) = lstm_ref.bilstm(x, p)
        got_e = lstm_ref.emission_head(x, p, cls.weight.detach(), cls.bias.detach())
    assert torch.allclose(got, want, atol=1e-12, rtol=0)
    assert torch.allclose(ghn, hn, atol=1e-12, rtol=0) and torch.allclose(gcn, cn, atol=1e-12, rtol=0)
    assert torch.allclose(got_e, want_e, atol=1e-12, rtol=0)


def test_oracle_fp32_close_to_torch_fp32():
    torch.manual_seed(5)
    lstm = nn.LSTM(input_size=32, hidden_size=32, batch_first=True, bidirectional=True)
    x = torch.randn(4, 64, 32)
    with torch.no_grad():
        want, _ = lstm(x)
        got, _ = lstm_ref.bilstm(x, {k: v.detach() for k, v in lstm.named_parameters()})
    assert (got - want).abs().max() < 2e-6
