"""GPU parity of the full fusion segment (region proj -> t2i -> i2t -> gate) through the drop-in modules.

Gates (BASELINE.json north_star / SURVEY 8d):
  fp32 path: max |a-b| / max(|b|, 1) <= 1e-5 on every post-LayerNorm tensor
  bf16 path: max |a-b| <= 2e-2 on post-LayerNorm activations (vs the fp32 oracle)
Checked against oracle/fusion_ref.py on the same seeded inputs AND against the committed golden vectors
(outputs of the reference's own classes, oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import icka_b200
from oracle import fusion_ref
from oracle.make_golden import CASES, build_case, case_extras

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
PENDING = set()           # golden cases not yet confirmed on a B200 (tools/parity_case.py prints their errors)
GPU_CASES = [c for c in CASES if c not in PENDING]


def rel(a, b):
    return float(((a - b).abs() / b.abs().clamp(min=1.0)).max())


def run_ours(name, precision):
    B, shape, params, inp, stride = build_case(name)
    extras = case_extras(name)
    cfg = icka_b200.FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads,
                                 intermediate_size=shape.inter, layer_norm_eps=shape.eps,
                                 hidden_act=extras.get('hidden_act', 'gelu'))
    model = icka_b200.CrossModalFusion(cfg, layer_num1=shape.L, region_dim=shape.region_dim, clip_dim=shape.clip_dim,
                                       num_i2t_encoders=extras.get('num_i2t_encoders', 2), precision=precision).to(DEV).eval()
    model.load_state_dict(params, strict=True)
    with torch.no_grad():
        out = model(inp['text_states'].to(DEV), inp['visual_embeds_att'].to(DEV), inp['clip_features'].to(DEV),
                    inp['token_embedding'].to(DEV), inp['img_mask'].to(DEV), inp['text_mask'].to(DEV),
                    return_dict=True)
    torch.cuda.synchronize()
    out = {k: v.float().cpu() for k, v in out.items()}
    return B, shape, params, inp, stride, out


def oracle(shape, params, inp, name=None):
    return fusion_ref.fusion_segment(inp['text_states'], inp['visual_embeds_att'], inp['clip_features'],
                                     inp['token_embedding'], inp['img_mask'], inp['text_mask'], params,
                                     num_layers=shape.L, num_heads=shape.heads, layer_norm_eps=shape.eps,
                                     **(case_extras(name) if name else {}))


@pytest.mark.parametrize('name', GPU_CASES)
def test_fusion_fp32_parity(name):
    B, shape, params, inp, stride, out = run_ours(name, 'fp32')
    want = oracle(shape, params, inp, name)
    for k in ('regions', 'fused', 'clip', 'result', 'gate'):
        assert rel(out[k].reshape(want[k].shape), want[k]) <= 1e-5, k
    g = np.load(os.path.join(GOLDEN, f'fusion_{name}.npz'))
    assert rel(out['fused'][:, ::stride], torch.from_numpy(g['fused'])) <= 1e-5
    assert rel(out['result'][:, ::stride], torch.from_numpy(g['result'])) <= 1e-5
    assert rel(out['clip'], torch.from_numpy(g['clip'])) <= 1e-5
    assert rel(out['gate'], torch.from_numpy(g['gate'])) <= 1e-5


@pytest.mark.parametrize('name', GPU_CASES)
def test_fusion_bf16_parity(name):
    B, shape, params, inp, stride, out = run_ours(name, 'bf16')
    g = np.load(os.path.join(GOLDEN, f'fusion_{name}.npz'))
    for k, ref in (('fused', g['fused']), ('result', g['result'])):
        err = float((out[k][:, ::stride] - torch.from_numpy(ref)).abs().max())
        assert err <= 2e-2, (k, err)
    assert float((out['clip'] - torch.from_numpy(g['clip'])).abs().max()) <= 2e-2
    assert float((out['gate'] - torch.from_numpy(g['gate'])).abs().max()) <= 2e-2


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-5), ('bf16', 2e-2)])
def test_cross_encoder_dropin_signature(precision, tol):
    """BertCrossEncoder(config, layer_num).forward(s1, s2, mask[B,1,1,R]) -> list of per-layer outputs (CMIM:659-667)."""
    B, shape, params, inp, stride = build_case('std_L2_eps5')
    cfg = icka_b200.FusionConfig(layer_norm_eps=shape.eps)
    enc = icka_b200.BertCrossEncoder(cfg, shape.L).to(DEV).eval()
    sub = {k[len('txt2img_attention.'):]: v for k, v in params.items() if k.startswith('txt2img_attention.')}
    enc.load_state_dict(sub, strict=True)
    g = torch.Generator().manual_seed(5)
    s2 = torch.randn(B, shape.R, shape.H, generator=g)
    m01 = (torch.rand(B, shape.R, generator=g) > 0.3).long()
    ext = fusion_ref.additive_mask(m01, torch.float32)
    want = fusion_ref.cross_encoder(inp['text_states'], s2, ext, params, 'txt2img_attention', shape.L, shape.heads, shape.eps)
    icka_b200.set_precision(precision)
    try:
        with torch.no_grad():
            got = enc(inp['text_states'].to(DEV), s2.to(DEV), ext.to(DEV))
            last_only = enc(inp['text_states'].to(DEV), s2.to(DEV), ext.to(DEV), output_all_encoded_layers=False)
    finally:
        icka_b200.set_precision('bf16')
    assert isinstance(got, list) and len(got) == shape.L and len(last_only) == 1
    for a, b in zip(got, want):
        err = rel(a.cpu(), b) if precision == 'fp32' else float((a.cpu() - b).abs().max())
        assert err <= tol, err
    assert torch.equal(last_only[0], got[-1])


def test_weight_update_refreshes_operand_cache():
    cfg = icka_b200.FusionConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256)
    enc = icka_b200.BertCrossEncoder(cfg, 1).to(DEV).eval()
    x, y = torch.randn(2, 8, 128, device=DEV), torch.randn(2, 5, 128, device=DEV)
    m = torch.zeros(2, 1, 1, 5, device=DEV)
    with torch.no_grad():
        a = enc(x, y, m)[-1].clone()
        enc.layer[0].output.dense.weight.mul_(0.5)
        b = enc(x, y, m)[-1]
    assert (a - b).abs().max() > 1e-3


def test_cast_f32_exact():
    """icka_cast_bf16_to_f32 is an exact widening, tails and odd sizes included."""
    from icka_b200 import ops
    for n in (0, 1, 7, 8, 4096 + 5, 128 * 768 * 3):
        x = torch.randn(n, device=DEV).to(torch.bfloat16)
        y = ops.cast_f32(x)
        assert y.dtype == torch.float32 and torch.equal(y, x.float())


@pytest.mark.parametrize('name', ['std_L1', 'hires_L1'])
@pytest.mark.parametrize('rows', [False, True])
def test_fusion_bf16_resident_inputs(name, rows):
    """bf16 path with inputs that ARRIVE in bf16 (the caller's encoders ran in bf16; `rows`: regions as the producer
    tail's K-major bf16 rows [B,R,2048] instead of the fp32 grid).  The bf16 states are the GEMM operand as is and the
    residual stream is their exact widening, so the result must (a) equal, bit for bit, the run on the same values
    handed over as fp32, and (b) stay within the bf16 gate (2e-2) of the fp32 oracle on those values."""
    B, shape, params, inp, stride = build_case(name)
    cfg = icka_b200.FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads,
                                 intermediate_size=shape.inter, layer_norm_eps=shape.eps)
    model = icka_b200.CrossModalFusion(cfg, layer_num1=shape.L, region_dim=shape.region_dim,
                                       clip_dim=shape.clip_dim).to(DEV).eval()
    model.load_state_dict(params, strict=True)
    icka_b200.set_precision('bf16')
    lp = dict(inp)
    for k in ('text_states', 'token_embedding', 'visual_embeds_att'):
        lp[k] = inp[k].to(torch.bfloat16)
    grid = lp['visual_embeds_att']
    R = grid.shape[2] * grid.shape[3]
    regions = grid.view(B, grid.shape[1], R).permute(0, 2, 1).contiguous() if rows else grid     # CMIM:956
    keys = ('clip_features', 'token_embedding', 'img_mask', 'text_mask')
    with torch.no_grad():
        got = model(lp['text_states'].to(DEV), regions.to(DEV), *[lp[k].to(DEV) for k in keys], return_dict=True)
        same = model(lp['text_states'].float().to(DEV), grid.float().to(DEV),
                     *[lp[k].float().to(DEV) if lp[k].is_floating_point() else lp[k].to(DEV) for k in keys],
                     return_dict=True)
    torch.cuda.synchronize()
    for k in ('result', 'fused', 'clip', 'gate'):
        assert torch.equal(got[k], same[k]), k
    widened = {k: (v.float() if v.is_floating_point() else v) for k, v in lp.items()}
    want = oracle(shape, params, widened)
    for k in ('result', 'fused', 'clip', 'gate'):
        err = float((got[k].float().cpu().reshape(want[k].shape) - want[k]).abs().max())
        assert err <= 2e-2, (k, err)
