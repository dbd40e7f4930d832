"""The small drop-in blocks called on their own build an autograd graph too (they used to return detached
tensors): BertLayerNorm, BertSelfOutput, BertIntermediate (gelu / relu / swish), BertOutput, BertCoAttention,
BertCrossAttention and cls_layer_both against torch autograd through the oracle's restatement of the same blocks.
fp32 mode: 2e-4 of each gradient's largest entry; bf16 mode: 4e-2."""
import pytest
import torch

import icka_b200
from oracle import fusion_ref

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
H, NH, I = 128, 2, 256
MODES = [('fp32', 2e-4, 1e-5), ('bf16', 4e-2, 2e-2)]


def cfg(**kw):
    return icka_b200.FusionConfig(hidden_size=H, num_attention_heads=NH, intermediate_size=I, layer_norm_eps=1e-12,
                                  hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)


def rnd(*shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def gerr(got, want, scale, l2):
    """Gradient error relative to `scale` (max |want| by default).  ``l2``: relative L2 error instead -- for relu, whose
    derivative jumps at 0: a pre-activation within rounding distance of 0 may land on the other side of the kink than in
    the oracle, which changes single gradient entries by their full size without being an error of the kernels."""
    d = got.detach().cpu().double() - want.double()
    if l2:
        return float(d.norm() / want.double().norm().clamp(min=1e-30))
    return float(d.abs().max()) / scale


def compare(module, ours_fn, ref_fn, inputs, tol, ftol, prefix='m', l2=False):
    """ours_fn(module, *cuda inputs) and ref_fn(params dict with `prefix.` keys, *cpu inputs) -> output tensor."""
    dev_in = [t.to(DEV).requires_grad_(True) for t in inputs]
    out = ours_fn(module, *dev_in)
    assert out.grad_fn is not None, 'no autograd graph was recorded'
    wgt = rnd(*out.shape, seed=99)
    (out * wgt.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    p = {f'{prefix}.{k}': v.detach().cpu().clone().requires_grad_(True) for k, v in module.state_dict().items()}
    cpu_in = [t.clone().requires_grad_(True) for t in inputs]
    ref = ref_fn(p, *cpu_in)
    (ref * wgt).sum().backward()
    ferr = float((out.detach().cpu() - ref.detach()).abs().max())
    assert ferr <= ftol * max(1.0, float(ref.abs().max())), ('forward', ferr)
    for a, b in zip(dev_in, cpu_in):
        assert a.grad is not None
        assert gerr(a.grad, b.grad, float(b.grad.abs().max()) + 1e-30, l2) <= tol + 1e-7, 'input gradient'
    seen = set()
    for k, v in module.named_parameters():
        if id(v) in seen:
            continue
        seen.add(id(v))
        want = p[f'{prefix}.{k}'].grad
        if want is None:
            continue
        scale = float(want.abs().max())
        if k.endswith('key.bias'):            # identically zero in exact arithmetic: gate it at the query bias' scale
            scale = float(p[f'{prefix}.{k.replace("key.bias", "query.bias")}'].grad.abs().max())
        assert v.grad is not None, f'no gradient reached {k}'
        assert gerr(v.grad, want, scale + 1e-30, l2) <= tol + 1e-7, k


@pytest.mark.parametrize('mode,tol,ftol', MODES)
def test_layernorm_and_dense_residual_blocks(mode, tol, ftol):
    torch.manual_seed(1)
    with icka_b200.precision(mode):
        ln = icka_b200.BertLayerNorm(H, eps=1e-12).to(DEV)
        with torch.no_grad():
            ln.weight.add_(0.1 * torch.randn(H, device=DEV))
            ln.bias.add_(0.1 * torch.randn(H, device=DEV))
        compare(ln, lambda m, x: m(x),
                lambda p, x: fusion_ref.bert_layer_norm(x, p['m.weight'], p['m.bias'], 1e-12),
                [rnd(3, 8, H, seed=2)], 2e-4, 1e-5)          # LayerNorm is fp32 in both modes
        for cls, width in ((icka_b200.BertSelfOutput, H), (icka_b200.BertOutput, I)):
            blk = cls(cfg()).to(DEV).eval()
            compare(blk, lambda m, h, x: m(h, x),
                    lambda p, h, x: fusion_ref.bert_layer_norm(fusion_ref.linear(h, p, 'm.dense') + x,
                                                               p['m.LayerNorm.weight'], p['m.LayerNorm.bias'], 1e-12),
                    [rnd(3, 8, width, seed=3), rnd(3, 8, H, seed=4)], tol, ftol)


@pytest.mark.parametrize('act', ['gelu', 'relu', 'swish'])
@pytest.mark.parametrize('mode,tol,ftol', MODES)
def test_intermediate_block_and_its_activations(mode, tol, ftol, act):
    torch.manual_seed(2)
    with icka_b200.precision(mode):
        blk = icka_b200.BertIntermediate(cfg(hidden_act=act)).to(DEV).eval()
        x = rnd(3, 8, H, seed=5)
        compare(blk, lambda m, t: m(t), lambda p, t: fusion_ref.ACT2FN[act](fusion_ref.linear(t, p, 'm.dense')), [x], tol, ftol,
                l2=act == 'relu')
        with torch.no_grad():                      # forward-only path: the activation fused into the GEMM epilogue
            got = blk(x.to(DEV)).cpu()
        want = fusion_ref.ACT2FN[act](torch.nn.functional.linear(x, blk.dense.weight.detach().cpu(), blk.dense.bias.detach().cpu()))
        assert float((got - want).abs().max()) <= ftol * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize('mode,tol,ftol', MODES)
def test_coattention_and_cross_attention_blocks(mode, tol, ftol):
    torch.manual_seed(3)
    B, Sq, Skv = 2, 16, 9
    m01 = (torch.rand(B, Skv, generator=torch.Generator().manual_seed(6)) > 0.3).long()
    m01[:, 0] = 1
    ext = fusion_ref.additive_mask(m01, torch.float32)
    with icka_b200.precision(mode):
        co = icka_b200.BertCoAttention(cfg()).to(DEV).eval()
        compare(co, lambda m, a, b: m(a, b, ext.to(DEV)),
                lambda p, a, b: fusion_ref.co_attention(a, b, ext, p, 'm', NH),
                [rnd(B, Sq, H, seed=7), rnd(B, Skv, H, seed=8)], tol, ftol)
        ca = icka_b200.BertCrossAttention(cfg()).to(DEV).eval()

        def ref(p, a, b):
            ctx = fusion_ref.co_attention(a, b, ext, p, 'm.self', NH)
            return fusion_ref.bert_layer_norm(fusion_ref.linear(ctx, p, 'm.output.dense') + a,
                                              p['m.output.LayerNorm.weight'], p['m.output.LayerNorm.bias'], 1e-12)
        compare(ca, lambda m, a, b: m(a, b, ext.to(DEV)), ref, [rnd(B, Sq, H, seed=9), rnd(B, Skv, H, seed=10)], tol, ftol)


def test_cls_layer_both_records_and_matches():
    torch.manual_seed(4)
    blk = icka_b200.cls_layer_both(H, H).to(DEV)
    with torch.no_grad():
        blk.proj_norm.weight.add_(0.1 * torch.randn(H, device=DEV))

    def ref(p, a, b):
        n = torch.nn.functional.layer_norm(a + b, (H,), p['m.proj_norm.weight'], p['m.proj_norm.bias'], 1e-5)
        return torch.nn.functional.linear(n, p['m.proj.weight'], p['m.proj.bias'])
    compare(blk, lambda m, a, b: m(a, b), ref, [rnd(5, H, seed=11), rnd(5, H, seed=12)], 2e-4, 1e-5)
    with torch.no_grad():
        got = blk(rnd(5, H, seed=11).to(DEV), rnd(5, H, seed=12).to(DEV))
    assert got.grad_fn is None


@pytest.mark.parametrize('act', ['relu', 'swish'])
@pytest.mark.parametrize('mode,tol', [('fp32', 2e-4), ('bf16', 4e-2)])
def test_cross_layer_trains_with_the_other_act2fn_entries(mode, tol, act):
    """CrossLayerFn with config.hidden_act = relu / swish: the fused node against autograd through the oracle layer."""
    torch.manual_seed(5)
    c = icka_b200.FusionConfig(hidden_size=768, num_attention_heads=12, intermediate_size=3072, layer_norm_eps=1e-12,
                               hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, hidden_act=act)
    B, Sq, Skv = 2, 128, 49
    ext = fusion_ref.additive_mask(torch.ones(B, Skv, dtype=torch.long), torch.float32)
    s1 = torch.nn.functional.layer_norm(rnd(B, Sq, 768, seed=13), (768,))
    s2 = rnd(B, Skv, 768, seed=14)
    wgt = rnd(B, Sq, 768, seed=15) / (Sq * 768) ** 0.5
    with icka_b200.precision(mode):
        layer = icka_b200.BertCrossAttentionLayer(c).to(DEV).eval()
        a, b = s1.to(DEV).requires_grad_(True), s2.to(DEV).requires_grad_(True)
        out = layer(a, b, ext.to(DEV))
        (out * wgt.to(DEV)).sum().backward()
        torch.cuda.synchronize()
    p = {'l.' + k: v.detach().cpu().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    ar, br = s1.clone().requires_grad_(True), s2.clone().requires_grad_(True)
    ref = fusion_ref.cross_layer(ar, br, ext, p, 'l', 12, 1e-12, hidden_act=act)
    (ref * wgt).sum().backward()
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= (1e-5 if mode == 'fp32' else 2e-2) * max(1.0, float(ref.abs().max()))
    l2 = act == 'relu'
    if l2:
        tol *= 6.0          # a handful of the 786k pre-activations sit within rounding distance of the kink (see gerr)
    assert gerr(a.grad, ar.grad, float(ar.grad.abs().max()), l2) <= tol
    assert gerr(b.grad, br.grad, float(br.grad.abs().max()), l2) <= tol
    for k, v in layer.named_parameters():
        if k.endswith('key.bias'):
            continue                         # identically zero in exact arithmetic
        assert gerr(v.grad, p['l.' + k].grad, float(p['l.' + k].grad.abs().max()), l2) <= tol + 1e-7, k
