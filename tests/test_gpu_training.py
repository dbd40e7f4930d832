"""GPU parity of the training path: gradients of the fusion segment and of the CRF loss, computed by the
kernel-backed autograd nodes (icka_b200/autograd.py), against torch autograd through the CPU oracle
(oracle/fusion_ref.py, oracle/crf_ref.py) on identical weights and inputs.

Tolerances (relative to the largest entry of each gradient tensor):
  fp32 mode  2e-4   (same graph on FFMA kernels; reassociation only)
  bf16 mode  4e-2   (bf16 GEMM operands, fp32 accumulation / residual stream / statistics)
"""
import pytest
import torch

import icka_b200
from icka_b200 import synth
from oracle import crf_ref, fusion_ref

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
KEYS = ('text_states', 'visual_embeds_att', 'clip_features', 'token_embedding', 'img_mask', 'text_mask')


def build(shape, seed):
    params = fusion_ref.make_params(shape.H, shape.heads, shape.inter, shape.L, seed=seed)
    cfg = icka_b200.FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads,
                                 intermediate_size=shape.inter, layer_norm_eps=shape.eps)
    model = icka_b200.CrossModalFusion(cfg, layer_num1=shape.L).to(DEV).eval()
    model.load_state_dict(params, strict=True)
    return params, model


def oracle_grads(inp, params, shape, w_res, w_clip):
    """Oracle gradients, sentence by sentence (sentences are independent on this path).  Returns the summed
    gradient and, per tensor, the largest entry of the sum of |per-sentence gradients|: the magnitude of what
    is being added up, i.e. the scale rounding errors are proportional to when the batch sum cancels."""
    B = inp['text_states'].shape[0]
    total, scale, dtext, dtok, loss_sum = {}, {}, [], [], 0.0
    for b in range(B):
        sl = slice(b, b + 1)
        p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        text = inp['text_states'][sl].clone().requires_grad_(True)
        tok = inp['token_embedding'][sl].clone().requires_grad_(True)
        out = fusion_ref.fusion_segment(text, inp['visual_embeds_att'][sl], inp['clip_features'][sl], tok,
                                        inp['img_mask'][sl], inp['text_mask'][sl], p, num_layers=shape.L,
                                        num_heads=shape.heads, layer_norm_eps=shape.eps)
        loss = (out['result'] * w_res[sl]).sum() + (out['clip'] * w_clip[sl]).sum()
        loss.backward()
        loss_sum += float(loss.detach())
        for k, v in p.items():
            if v.grad is not None:
                total[k] = total.get(k, 0) + v.grad
                scale[k] = scale.get(k, 0) + v.grad.abs()
        dtext.append(text.grad)
        dtok.append(tok.grad)
    scale = {k: float(v.max()) for k, v in scale.items()}
    for k in list(scale):
        # d(key.bias) is identically zero in exact arithmetic (softmax ignores a shift common to all keys): what
        # is left on both sides is rounding noise of sums whose terms are as large as those of d(query.bias)
        if k.endswith('key.bias'):
            scale[k] = scale[k.replace('key.bias', 'query.bias')]
    return loss_sum, total, scale, torch.cat(dtext), torch.cat(dtok)


@pytest.mark.parametrize('mode,tol', [('fp32', 2e-4), ('bf16', 4e-2)])
@pytest.mark.parametrize('L', [1, 2])
def test_fusion_gradients_match_oracle_autograd(mode, tol, L):
    shape = synth.Shape(L=L)
    B = 3
    params, model = build(shape, seed=51 + L)
    inp = synth.fusion_inputs(B, shape, seed=52)
    g = torch.Generator().manual_seed(53)
    w_res = torch.randn(B, shape.S, shape.H, generator=g) / (shape.S * shape.H) ** 0.5
    w_clip = torch.randn(B, 1, shape.H, generator=g) / shape.H ** 0.5
    want_loss, want, scale, want_dtext, want_dtok = oracle_grads(inp, params, shape, w_res, w_clip)

    icka_b200.set_precision(mode)
    try:
        args = [inp[k].to(DEV) for k in KEYS]
        args[0].requires_grad_(True)
        args[3].requires_grad_(True)
        result, clip = model(*args)
        loss = (result * w_res.to(DEV)).sum() + (clip * w_clip.to(DEV)).sum()
        loss.backward()
        torch.cuda.synchronize()
    finally:
        icka_b200.set_precision('bf16')
    assert abs(float(loss) - want_loss) <= (1e-4 if mode == 'fp32' else 5e-2) * max(1.0, abs(want_loss))

    def rel(got, ref, denom):
        return float((got.detach().cpu() - ref).abs().max()) / denom

    worst = {}
    for name, p in model.named_parameters():
        assert name in want, f'oracle has no gradient for {name}'
        assert p.grad is not None, f'no gradient reached {name}'
        worst[name] = rel(p.grad, want[name], scale[name])
    worst['d text_states'] = rel(args[0].grad, want_dtext, float(want_dtext.abs().max()))
    worst['d token_embedding'] = rel(args[3].grad, want_dtok, float(want_dtok.abs().max()))
    bad = {k: v for k, v in worst.items() if not v <= tol}
    assert not bad, f'gradient mismatch ({mode}, L={L}): {bad}'


def test_crf_loss_gradients_and_sgd_step():
    sh = synth.STD
    batch = synth.crf_batch(48, sh, seed=61)
    cp = synth.crf_params(sh.T, 62, 'uniform')
    crf = icka_b200.CRF(sh.T, batch_first=True).to(DEV)
    crf.load_state_dict(cp)
    e = batch['emissions'].to(DEV).requires_grad_(True)
    tags, mask = batch['tags'].to(DEV), batch['mask'].to(DEV)
    loss = -crf(e, tags, mask, reduction='token_mean')          # CMIM:1047-1048
    loss.backward()

    ed = batch['emissions'].double().requires_grad_(True)
    pd = {k: v.double().requires_grad_(True) for k, v in cp.items()}
    want = -crf_ref.log_likelihood(ed, batch['tags'], batch['mask'], pd['start_transitions'], pd['end_transitions'],
                                   pd['transitions'], 'token_mean')
    want.backward()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    assert float((e.grad.cpu().double() - ed.grad).abs().max()) <= 1e-6
    for name, p in crf.named_parameters():
        assert float((p.grad.cpu().double() - pd[name].grad).abs().max()) <= 2e-5, name

    # a few plain SGD steps on the CRF parameters lower the loss
    opt = torch.optim.SGD(crf.parameters(), lr=0.5)
    first = float(loss)
    for _ in range(5):
        opt.zero_grad()
        loss = -crf(e.detach(), tags, mask, reduction='token_mean')
        loss.backward()
        opt.step()
    assert float(loss) < first


def test_encoder_module_api_records_a_graph():
    """BertCrossEncoder called directly (the reference's own module boundary) is differentiable too."""
    shape = synth.Shape(L=1, H=128, heads=2, inter=256)
    cfg = icka_b200.FusionConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256, layer_norm_eps=1e-12)
    torch.manual_seed(3)
    enc = icka_b200.BertCrossEncoder(cfg, 2).to(DEV).eval()
    s1 = torch.randn(2, 16, 128, device=DEV, requires_grad=True)
    s2 = torch.randn(2, 9, 128, device=DEV, requires_grad=True)
    mask = torch.zeros(2, 1, 1, 9, device=DEV)
    icka_b200.set_precision('fp32')
    try:
        wgt = torch.randn(2, 16, 128, device=DEV)       # (a plain sum of squares of a LayerNorm output is constant)
        out = enc(s1, s2, mask)[-1]
        (out * wgt).sum().backward()
    finally:
        icka_b200.set_precision('bf16')
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    a, b = s1.detach().cpu().requires_grad_(True), s2.detach().cpu().requires_grad_(True)
    ref = fusion_ref.cross_encoder(a, b, mask.cpu(), {'e.' + k: v for k, v in p.items()}, 'e', 2, 2, 1e-12)[-1]
    (ref * wgt.cpu()).sum().backward()
    assert float((s1.grad.cpu() - a.grad).abs().max()) <= 2e-4 * float(a.grad.abs().max())
    assert float((s2.grad.cpu() - b.grad).abs().max()) <= 2e-4 * float(b.grad.abs().max())
    for k, v in enc.named_parameters():
        ref_k = p[k.replace('key.bias', 'query.bias')].grad      # d(key.bias) == 0 in exact arithmetic
        assert float((v.grad.cpu() - p[k].grad).abs().max()) <= 2e-4 * float(ref_k.abs().max()) + 1e-7, k


@pytest.mark.parametrize('Sq,Skv', [(128, 49), (1, 128)], ids=['text2image', 'image2text'])
@pytest.mark.parametrize('mode,tol', [('fp32', 3e-4), ('bf16', 5e-2)])
def test_cross_layer_with_dropout_replays_the_same_philox_masks(mode, tol, Sq, Skv, monkeypatch):
    """Training-mode layer with p = 0.1 on all three dropout sites: the keep masks the kernels draw are exported
    (icka_dropout_mask) and replayed in the oracle, so forward and every gradient must match as without dropout."""
    from icka_b200 import ops
    B, H, nh, I = 3 if Sq > 1 else 40, 768, 12, 3072
    cfg = icka_b200.FusionConfig(hidden_size=H, num_attention_heads=nh, intermediate_size=I, layer_norm_eps=1e-12,
                                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    torch.manual_seed(5)
    layer = icka_b200.BertCrossAttentionLayer(cfg).to(DEV).train()
    g = torch.Generator().manual_seed(6)
    s1 = torch.nn.functional.layer_norm(torch.randn(B, Sq, H, generator=g), (H,))
    s2 = torch.randn(B, Skv, H, generator=g)
    m01 = (torch.rand(B, Skv, generator=g) > 0.2).long()
    ext = fusion_ref.additive_mask(m01, torch.float32)
    wgt = torch.randn(B, Sq, H, generator=g) / (Sq * H) ** 0.5
    seeds = []
    real_randint = torch.randint

    def spy(*a, **k):
        out = real_randint(*a, **k)
        seeds.append(int(out.item()))
        return out
    monkeypatch.setattr(torch, 'randint', spy)
    icka_b200.set_precision(mode)
    try:
        a, b = s1.to(DEV).requires_grad_(True), s2.to(DEV).requires_grad_(True)
        out = layer(a, b, ext.to(DEV))
        (out * wgt.to(DEV)).sum().backward()
        torch.cuda.synchronize()
    finally:
        icka_b200.set_precision('bf16')
    monkeypatch.undo()
    assert len(seeds) == 1
    seed = seeds[0]
    drop = dict(p_attn=0.1, p_hid=0.1,
                attn=ops.dropout_mask((B, nh, Sq, Skv), 0.1, 3 * seed, DEV, attention=True).cpu(),
                h1=ops.dropout_mask((B, Sq, H), 0.1, 3 * seed + 1, DEV).cpu(),
                h2=ops.dropout_mask((B, Sq, H), 0.1, 3 * seed + 2, DEV).cpu())
    for k in ('attn', 'h1', 'h2'):
        assert 0.87 <= float(drop[k].float().mean()) <= 0.93, k          # keep rate 1 - p
    p = {'l.' + k: v.detach().cpu().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    ar, br = s1.clone().requires_grad_(True), s2.clone().requires_grad_(True)
    ref = fusion_ref.cross_layer(ar, br, ext, p, 'l', nh, 1e-12, drop=drop)
    (ref * wgt).sum().backward()
    ftol = 1e-5 if mode == 'fp32' else 3e-2
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= ftol * max(1.0, float(ref.abs().max()))
    assert float((a.grad.cpu() - ar.grad).abs().max()) <= tol * float(ar.grad.abs().max())
    assert float((b.grad.cpu() - br.grad).abs().max()) <= tol * float(br.grad.abs().max())
    for k, v in layer.named_parameters():
        ref_k = p['l.' + k.replace('key.bias', 'query.bias')].grad      # d(key.bias) == 0 in exact arithmetic
        assert float((v.grad.cpu() - p['l.' + k].grad).abs().max()) <= tol * float(ref_k.abs().max()) + 1e-7, k
