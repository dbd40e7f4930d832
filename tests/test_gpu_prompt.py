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


def test_training_mode_without_autograd_refuses():
    m = PromptMapping(icka_b200.FusionConfig(hidden_size=768)).cuda().train()
    with torch.no_grad(), pytest.raises(NotImplementedError):
        m(torch.zeros(1, 1, 768, device='cuda'), torch.zeros(1, 2048, device='cuda'), torch.ones(1, 4, device='cuda'))


@pytest.mark.parametrize('precision,tol,ftol', [('fp32', 3e-4, 1e-5), ('bf16', 5e-2, 3e-2)])
@pytest.mark.parametrize('training', [False, True], ids=['eval', 'dropout'])
def test_prompt_networks_train(precision, tol, ftol, training, monkeypatch):
    """mode='train' (CMIM:1046-1048) reaches the prompt mapping networks: forward and every gradient against autograd
    through the oracle, with the Dropout(0.3) keep masks the kernels drew (Philox, icka_dropout_mask) replayed there."""
    from icka_b200 import ops
    icka_b200.set_precision(precision)
    B = 6
    p = prompt_ref.make_params(768, seed=21)
    m = PromptMapping(icka_b200.FusionConfig(hidden_size=768))
    m.load_state_dict(p, strict=True)
    m = m.cuda().train(training)
    g = torch.Generator().manual_seed(22)
    clip, vmean = torch.randn(B, 1, 768, generator=g), torch.relu(torch.randn(B, 2048, generator=g)) * 0.5
    mask = torch.ones(B, 12, dtype=torch.long)
    wgt = torch.randn(B, 10, 1024, generator=g) / (10 * 1024) ** 0.5
    seeds = []
    real_randint = torch.randint

    def spy(*a, **k):
        out = real_randint(*a, **k)
        seeds.append(int(out.item()))
        return out
    monkeypatch.setattr(torch, 'randint', spy)
    clip_d = clip.cuda().requires_grad_(True)
    got, got_mask = m(clip_d, vmean.cuda(), mask.cuda())
    monkeypatch.undo()
    assert got.grad_fn is not None
    (got * wgt.cuda()).sum().backward()
    torch.cuda.synchronize()
    drop = None
    if training:
        assert len(seeds) == 4                           # vision.0, vision.3, alignment.0, alignment.3 -- in this order
        inner = 3840 if precision == 'bf16' else 3780    # bf16: the hidden width is zero-padded to a multiple of 64
        keep = lambda shape, seed: ops.dropout_mask(shape, 0.3, seed, 'cuda').cpu()
        drop = dict(p=0.3, masks={'mapping_network_vision.0': keep((B, 2048), seeds[0]),
                                  'mapping_network_vision.3': keep((B, inner), seeds[1])[:, :3780],
                                  'mapping_network_alignment.0': keep((B, 768), seeds[2]),
                                  'mapping_network_alignment.3': keep((B, inner), seeds[3])[:, :3780]})
        for k, v in drop['masks'].items():
            assert 0.6 <= float(v.float().mean()) <= 0.8, k
    else:
        assert not seeds
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    clip_r = clip.clone().requires_grad_(True)
    want, want_mask = prompt_ref.prompt_prefix(clip_r, vmean, mask, pr, drop=drop)
    (want * wgt).sum().backward()
    assert torch.equal(got_mask.cpu(), want_mask)
    assert float((got.detach().cpu() - want.detach()).abs().max()) <= ftol * max(1.0, float(want.abs().max()))
    assert float((clip_d.grad.cpu() - clip_r.grad).abs().max()) <= tol * float(clip_r.grad.abs().max())
    for k, v in m.named_parameters():
        assert v.grad is not None, k
        assert float((v.grad.cpu() - pr[k].grad).abs().max()) <= tol * float(pr[k].grad.abs().max()) + 1e-7, k
