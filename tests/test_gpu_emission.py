"""GPU parity of the emission head (BiLSTM + classifier, CMIM:905-910, 1042-1043) against the
oracle restatement (oracle/lstm_ref.py, pinned to torch.nn.LSTM in tests/test_oracle_lstm.py).

Tolerances: fp32 mode max|a-b| / max(|b|,1) <= 1e-5 (the north star's fp32 bar); bf16 mode max|a-b| <= 2e-2 on the
LSTM states (|h| < 1) and on the emissions (the north star's bf16 bar for activations)."""
import pytest
import torch
from torch import nn

import icka_b200
from icka_b200 import ops
from icka_b200.config import FusionConfig
from oracle import lstm_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = icka_b200.get_precision()
    yield
    icka_b200.set_precision(prev)


def make(I, H, T, seed):
    torch.manual_seed(seed)
    ref_lstm = nn.LSTM(input_size=I, hidden_size=H, batch_first=True, bidirectional=True)
    ref_cls = nn.Linear(2 * H, T)
    ours = icka_b200.LSTM(input_size=I, hidden_size=H, batch_first=True, bidirectional=True)
    ours.load_state_dict(ref_lstm.state_dict())                 # same keys as nn.LSTM
    return ref_lstm, ref_cls, ours.cuda().eval()


def oracle(ref_lstm, ref_cls, x):
    p = {k: v.detach().double() for k, v in ref_lstm.named_parameters()}
    with torch.no_grad():
        y, (hn, cn) = lstm_ref.bilstm(x.double(), p)
        e = y @ ref_cls.weight.detach().double().t() + ref_cls.bias.detach().double()
    return y, hn, cn, e


def rel(a, b):
    return ((a.double().cpu() - b).abs() / b.abs().clamp(min=1.0)).max().item()


@pytest.mark.parametrize('B,S,I,H', [(3, 9, 32, 32), (2, 128, 24, 40), (2, 12, 768, 768)])
def test_fp32_per_step_path(B, S, I, H):
    icka_b200.set_precision('fp32')
    ref_lstm, ref_cls, ours = make(I, H, 15, seed=B + S)
    x = torch.randn(B, S, I)
    y, hn, cn, _ = oracle(ref_lstm, ref_cls, x)
    with torch.no_grad():
        out, (h_n, c_n) = ours(x.cuda())
    assert out.dtype == torch.float32 and out.shape == (B, S, 2 * H)
    assert rel(out, y) <= 1e-5 and rel(h_n, hn) <= 1e-5 and rel(c_n, cn) <= 1e-5


@pytest.mark.parametrize('variant', [1, 2])
@pytest.mark.parametrize('B,S', [(5, 128), (300, 24), (128, 3), (1, 1), (640, 5), (1100, 3), (2048, 4), (1500, 2)])
def test_bf16_persistent_kernel(B, S, variant, monkeypatch):
    """Both kernels behind icka_lstm_rec_fwd (1: single CTAs / 24-unit slices, 2: CTA pairs / 48-unit slices) at every
    batch shape (the library picks 1 for B <= 256, 2 above; ICKA_LSTM_VARIANT forces one)."""
    monkeypatch.setenv('ICKA_LSTM_VARIANT', str(variant))
    assert ops.lstm_rec_variant(B) == variant
    icka_b200.set_precision('bf16')
    H = 768
    ref_lstm, ref_cls, ours = make(H, H, 15, seed=B * 7 + S)
    assert ours.uses_persistent_kernel()
    x = torch.randn(B, S, H)
    y, hn, cn, _ = oracle(ref_lstm, ref_cls, x)
    with torch.no_grad():
        out, (h_n, c_n) = ours(x.cuda())
    err_y = (out.double().cpu() - y).abs().max().item()
    err_h = (h_n.double().cpu() - hn).abs().max().item()
    err_c = (c_n.double().cpu() - cn).abs().max().item()
    print(f'bf16 persistent variant {variant} B={B} S={S}: max|dy|={err_y:.2e} max|dh_n|={err_h:.2e} max|dc_n|={err_c:.2e}')
    assert err_y <= 2e-2 and err_h <= 2e-2 and err_c <= 4e-2
    with torch.no_grad():
        out2, _ = ours(x.cuda())                                # no atomics on the data path: reruns are identical
    assert torch.equal(out, out2)


def test_variant_choice():
    assert ops.lstm_rec_variant(1) == 1 and ops.lstm_rec_variant(256) == 1 and ops.lstm_rec_variant(257) == 2


def test_bf16_per_step_path_other_hidden_size():
    icka_b200.set_precision('bf16')
    ref_lstm, ref_cls, ours = make(64, 64, 15, seed=3)
    assert not ours.uses_persistent_kernel()
    x = torch.randn(4, 20, 64)
    y, hn, cn, _ = oracle(ref_lstm, ref_cls, x)
    with torch.no_grad():
        out, (h_n, c_n) = ours(x.cuda())
    assert (out.double().cpu() - y).abs().max().item() <= 2e-2
    assert (h_n.double().cpu() - hn).abs().max().item() <= 2e-2


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,K,T', [(1, 8, 1), (37, 1536, 15), (1000, 1536, 16), (9, 264, 7)])
def test_emission_head_kernel(dtype, M, K, T):
    torch.manual_seed(M + K + T)
    x = torch.randn(M, K).to(dtype)
    w = torch.randn(T, K) / K ** 0.5
    b = torch.randn(T)
    want = x.double() @ w.double().t() + b.double()
    got = ops.emission_head(x.cuda(), w.cuda(), b.cuda())
    # bf16 states take the tensor-core kernel with hi + lo split weights (16 mantissa bits): 1e-4; fp32 stays 1e-5
    assert got.shape == (M, T) and rel(got, want) <= (1e-5 if dtype == torch.float32 else 1e-4)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,K,T', [(1, 8, 1), (37, 1536, 15), (5000, 1536, 16), (9, 264, 7), (130, 2048, 3)])
def test_emission_head_bwd_kernel(dtype, M, K, T):
    """Classifier backward (autograd of nn.Linear(2H, T), CMIM:910): dx = dout . W and dW = dout^T . x against fp64."""
    torch.manual_seed(M + K + T)
    x = torch.randn(M, K).to(dtype)
    w = torch.randn(T, K) / K ** 0.5
    dout = torch.randn(M, T)
    dx, dw = ops.emission_head_bwd(dout.cuda(), x.cuda(), w.cuda())
    assert dx.shape == (M, K) and dw.shape == (T, K)
    assert rel(dx, dout.double() @ w.double()) <= 1e-5
    want_dw = dout.double().t() @ x.double()
    scale = want_dw.abs().max().item()                   # sums of M products: fp32 rounding scales with the largest entries
    assert (dw.double().cpu() - want_dw).abs().max().item() <= 1e-5 * max(scale, 1.0)
    if M % 5 == 0:                                       # time-major states (row t*B + b) against batch-major dout (row b*S + t)
        S_, B_ = 5, M // 5
        x_tm = x.view(B_, S_, K).transpose(0, 1).reshape(M, K).contiguous()
        dx_tm, dw_tm = ops.emission_head_bwd(dout.cuda(), x_tm.cuda(), w.cuda(), time_major_S=S_)
        assert rel(dx_tm.view(S_, B_, K).transpose(0, 1).reshape(M, K), dout.double() @ w.double()) <= 1e-5
        assert (dw_tm.double().cpu() - want_dw).abs().max().item() <= 1e-5 * max(scale, 1.0)
    only_dw = ops.emission_head_bwd(dout.cuda(), x.cuda(), w.cuda(), want_dx=False)
    assert only_dw[0] is None and (only_dw[1].double().cpu() - want_dw).abs().max().item() <= 1e-5 * max(scale, 1.0)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 2e-2)])
def test_emission_head_module(precision, tol):
    icka_b200.set_precision(precision)
    H, T, B, S = 768, 15, 6, 32
    torch.manual_seed(42)
    ref_lstm = nn.LSTM(input_size=H, hidden_size=H, batch_first=True, bidirectional=True)
    ref_cls = nn.Linear(2 * H, T)
    head = icka_b200.EmissionHead(FusionConfig(hidden_size=H), num_labels=T)
    head.lstm.load_state_dict(ref_lstm.state_dict())
    head.classifier.load_state_dict(ref_cls.state_dict())
    head = head.cuda().eval()
    x = torch.randn(B, S, H)
    _, _, _, e = oracle(ref_lstm, ref_cls, x)
    with torch.no_grad():
        got = head(x.cuda())
    assert got.shape == (B, S, T) and got.dtype == torch.float32
    err = rel(got, e) if precision == 'fp32' else (got.double().cpu() - e).abs().max().item()
    print(f'emission head {precision}: err {err:.2e}')
    assert err <= tol


def test_batches_above_one_launch_are_chunked():
    """B > 2048: EmissionHead / LSTM walk 2048-sentence chunks; same result as the chunks on their own."""
    icka_b200.set_precision('bf16')
    H, T, B, S = 768, 15, 2100, 3
    torch.manual_seed(2)
    head = icka_b200.EmissionHead(FusionConfig(hidden_size=H), num_labels=T).cuda().eval()
    x = torch.randn(B, S, H, device='cuda')
    with torch.no_grad():
        whole = head(x)
        parts = torch.cat([head(x[:2048]), head(x[2048:])])
        out, (h_n, c_n) = head.lstm(x)
        out_a, (h_a, _) = head.lstm(x[:2048])
    assert torch.equal(whole, parts)
    assert out.shape == (B, S, 2 * H) and h_n.shape == (2, B, H) and c_n.shape == (2, B, H)
    assert torch.equal(out[:2048], out_a) and torch.equal(h_n[:, :2048], h_a)


def test_lstm_rejects_what_it_does_not_cover():
    with pytest.raises(NotImplementedError):
        icka_b200.LSTM(8, 8, batch_first=True, bidirectional=False)
    m = icka_b200.LSTM(8, 8, batch_first=True, bidirectional=True).cuda().eval()
    with pytest.raises(ValueError):
        m(torch.zeros(3, 8, device='cuda'))
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 9, device='cuda'))
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 3, 8))                                 # CPU tensor: no fallback


def _grad_case(precision, B, S, H, T, seed):
    icka_b200.set_precision(precision)
    torch.manual_seed(seed)
    ref_lstm = nn.LSTM(input_size=H, hidden_size=H, batch_first=True, bidirectional=True)
    ref_cls = nn.Linear(2 * H, T)
    head = icka_b200.EmissionHead(FusionConfig(hidden_size=H), num_labels=T)
    head.lstm.load_state_dict(ref_lstm.state_dict())
    head.classifier.load_state_dict(ref_cls.state_dict())
    head = head.cuda().train()
    x = torch.randn(B, S, H)
    w_out = torch.randn(B, S, T)                                 # fixed upstream gradient
    # oracle: autograd through the restatement (fp64)
    xo = x.double().requires_grad_(True)
    po = {k: v.detach().double().requires_grad_(True) for k, v in ref_lstm.named_parameters()}
    wc, bc = ref_cls.weight.detach().double().requires_grad_(True), ref_cls.bias.detach().double().requires_grad_(True)
    e = lstm_ref.emission_head(xo, po, wc, bc)
    (e * w_out.double()).sum().backward()
    xg = x.cuda().requires_grad_(True)
    got = head(xg)
    (got * w_out.cuda()).sum().backward()
    return head, xg, got, e.detach(), xo, po, wc, bc


@pytest.mark.parametrize('precision,B,S,H,tol', [('fp32', 3, 7, 32, 2e-4), ('fp32', 2, 5, 768, 2e-4), ('bf16', 4, 12, 64, 4e-2),
                                                 ('bf16', 3, 6, 768, 4e-2), ('bf16', 40, 9, 768, 4e-2), ('bf16', 32, 1, 768, 4e-2)])
def test_training_gradients_match_oracle_autograd(precision, B, S, H, tol):
    """BiLstmFn / LinearFn (BPTT on per-step kernels) against autograd through the oracle: every gradient within `tol`
    of its own scale (max |g|), as in tests/test_gpu_training.py."""
    head, xg, got, want_e, xo, po, wc, bc = _grad_case(precision, B, S, H, 15, seed=B * 10 + S)
    assert (got.detach().double().cpu() - want_e).abs().max().item() <= (1e-5 if precision == 'fp32' else 2e-2)

    def close(name, g, w):
        scale = w.abs().max().item()
        err = (g.double().cpu() - w).abs().max().item()
        assert err <= tol * max(scale, 1e-6), f'{name}: err {err:.3e} vs scale {scale:.3e}'

    close('dx', xg.grad, xo.grad)
    for k, v in head.lstm.named_parameters():
        close(k, v.grad, po[k].grad)
    close('classifier.weight', head.classifier.weight.grad, wc.grad)
    close('classifier.bias', head.classifier.bias.grad, bc.grad)


def test_fused_step_kernels_agree_with_the_per_step_path(monkeypatch):
    """bf16, H = 768: one launch per step for both directions (csrc/lstm_train.cu) against the GEMM + cell kernels per
    direction and step it replaces -- same operands and rounding points, so outputs and gradients agree to bf16 rounding."""
    from icka_b200 import autograd
    outs = []
    for fused in (True, False):
        monkeypatch.setattr(autograd, '_FUSED_STEPS', fused)
        monkeypatch.setattr(autograd, '_FUSED_STEP_MAX_B', 1 << 30)       # 70 sentences: three 32-row blocks, the last ragged
        head, xg, got, *_ = _grad_case('bf16', 70, 11, 768, 15, seed=5)
        outs.append((got.detach(), xg.grad.clone(), {k: v.grad.clone() for k, v in head.named_parameters()}))
    (ea, xa, pa), (eb, xb, pb) = outs
    assert (ea - eb).abs().max().item() <= 2e-2
    assert (xa - xb).abs().max().item() <= 3e-2 * xb.abs().max().item()
    for k in pa:
        assert (pa[k] - pb[k]).abs().max().item() <= 3e-2 * pb[k].abs().max().item() + 1e-6, k
