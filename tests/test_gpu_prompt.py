"""GPU parity of the prompt mapping networks + prefix assembly (CMIM:913-930, 995-1009) against the
oracle and the golden outputs of the reference's own modules.  fp32: max|a-b|/max(|b|,1) <= 1e-5; bf16: max|a-b| <= 2e-2."""
import pytest
import torch

import icka_b200
from icka_b200.prompt import PromptMapping
from oracle import prompt_ref

load_case = prompt_ref.load_golden_case

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = icka_b200.get_precision()
    yield
    icka_b200.set_precision(prev)


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 2e-2)])
def test_prefix_matches_reference_golden(precision, tol):
    icka_b200.set_precision(precision)
    p, clip, vmean, mask, want, want_mask = load_case('std')
    m = PromptMapping(icka_b200.FusionConfig(hidden_size=768))
    m.load_state_dict(p, strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        got, got_mask = m(clip.cuda(), vmean.cuda(), mask.cuda())
    assert got.shape == want.shape and got.dtype == torch.float32
    diff = (got.cpu() - want).abs()
    err = float((diff / want.abs().clamp(min=1.0)).max()) if precision == 'fp32' else float(diff.max())
    print(f'prompt prefix {precision}: err {err:.2e}')
    assert err <= tol
    assert torch.equal(got_mask.cpu(), want_mask)


def test_batch_1024_bf16_against_oracle():
    icka_b200.set_precision('bf16')
    torch.manual_seed(3)
    p = prompt_ref.make_params(768, seed=9)
    B = 1024
    clip, vmean = torch.randn(B, 1, 768), torch.relu(torch.randn(B, 2048)) * 0.5
    mask = torch.ones(B, 40, dtype=torch.long)
    want, _ = prompt_ref.prompt_prefix(clip, vmean, mask, p)
    m = PromptMapping(icka_b200.FusionConfig(hidden_size=768))
    m.load_state_dict(p)
    m = m.cuda().eval()
    with torch.no_grad():
        got, _ = m(clip.cuda(), vmean.cuda(), mask.cuda())
    assert (got.cpu() - want).abs().max().item() <= 2e-2


def test_training_mode_refuses():
    m = PromptMapping(icka_b200.FusionConfig(hidden_size=768)).cuda().train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 1, 768, device='cuda'), torch.zeros(1, 2048, device='cuda'), torch.ones(1, 4, device='cuda'))
