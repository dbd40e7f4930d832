"""Pin oracle/fusion_ref.py: against the reference's own classes (container only) and the golden vectors."""
import os

import numpy as np
import pytest
import torch

from icka_b200 import synth
from oracle import fusion_ref, reference_shim
from oracle.make_golden import CASES, build_case, case_extras, checksum

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def rel_err(a, b):
    """SURVEY 8d parity metric: max|a-b| / max(|b|, 1)."""
    return float(((a - b).abs() / b.abs().clamp(min=1.0)).max())


def run_oracle(name):
    B, shape, params, inp, stride = build_case(name)
    out = fusion_ref.fusion_segment(
        inp['text_states'], inp['visual_embeds_att'], inp['clip_features'], inp['token_embedding'],
        inp['img_mask'], inp['text_mask'], params, num_layers=shape.L, num_heads=shape.heads,
        layer_norm_eps=shape.eps, **case_extras(name))
    return B, shape, params, inp, stride, out


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_golden(name):
    B, shape, params, inp, stride, out = run_oracle(name)
    g = np.load(os.path.join(GOLDEN, f'fusion_{name}.npz'))
    assert int(g['row_stride']) == stride
    assert abs(float(g['params_checksum']) - sum(checksum(v) for k, v in sorted(params.items()))) < 1e-6 * float(g['params_checksum'])
    assert abs(float(g['inputs_checksum']) - sum(checksum(inp[k]) for k in (
        'text_states', 'visual_embeds_att', 'clip_features', 'token_embedding'))) < 1e-6 * float(g['inputs_checksum'])
    assert rel_err(out['regions'], torch.from_numpy(g['regions'])) <= 1e-5
    assert rel_err(out['fused'][:, ::stride], torch.from_numpy(g['fused'])) <= 1e-5
    assert rel_err(out['clip'], torch.from_numpy(g['clip'])) <= 1e-5
    assert rel_err(out['result'][:, ::stride], torch.from_numpy(g['result'])) <= 1e-5
    assert rel_err(out['gate'], torch.from_numpy(g['gate'])) <= 1e-5


@pytest.mark.skipif(not reference_shim.available(), reason='/root/reference not present (GPU box)')
@pytest.mark.parametrize('name', ['tiny', 'std_L1', 'tiny_relu', 'tiny_swish'])
def test_oracle_matches_reference_classes(name):
    B, shape, params, inp, stride, out = run_oracle(name)
    mods = reference_shim.build_reference_modules(
        params, hidden=shape.H, heads=shape.heads, inter=shape.inter, num_layers=shape.L,
        layer_norm_eps=shape.eps, **case_extras(name))
    ref = reference_shim.reference_fusion_segment(
        mods, inp['text_states'], inp['visual_embeds_att'], inp['clip_features'],
        inp['token_embedding'], inp['img_mask'], inp['text_mask'])
    for k in ('regions', 'fused', 'clip', 'result', 'gate'):
        assert rel_err(out[k], ref[k]) <= 2e-6, k


@pytest.mark.skipif(not reference_shim.available(), reason='/root/reference not present (GPU box)')
def test_reference_state_dict_keys():
    """The drop-in keeps the reference's parameter names."""
    cmim = reference_shim.load()
    cfg = cmim.BertConfig(30522, hidden_size=128, num_attention_heads=2, intermediate_size=256)
    ref_keys = set(cmim.BertCrossEncoder(cfg, 2).state_dict().keys())
    p = fusion_ref.make_params(128, 2, 256, 2)
    ours = {k[len('txt2img_attention.'):] for k in p if k.startswith('txt2img_attention.')}
    assert ours == ref_keys
    both = set(cmim.cls_layer_both(128, 128).state_dict().keys())
    assert both == {k[len('cls_layer.'):] for k in p if k.startswith('cls_layer.')}


def test_masked_regions_change_output():
    """A -10000 additive mask on some regions must move the result (mask is really applied)."""
    B, shape, params, inp, stride = build_case('tiny')
    args = (inp['text_states'], inp['visual_embeds_att'], inp['clip_features'], inp['token_embedding'])
    kw = dict(num_layers=shape.L, num_heads=shape.heads, layer_norm_eps=shape.eps)
    a = fusion_ref.fusion_segment(*args, inp['img_mask'], inp['text_mask'], params, **kw)
    m2 = inp['img_mask'].clone(); m2[:, ::2] = 0
    b = fusion_ref.fusion_segment(*args, m2, inp['text_mask'], params, **kw)
    assert (a['fused'] - b['fused']).abs().max() > 1e-3
    # fully masked row == unmasked softmax (uniform shift), SURVEY 7.3 #8
    m3 = torch.zeros_like(inp['img_mask'])
    c = fusion_ref.fusion_segment(*args, m3, inp['text_mask'], params, **kw)
    assert rel_err(c['fused'], a['fused']) < 1e-3


def test_fp64_noise_floor():
    """fp32 oracle vs the same oracle in fp64: the 1e-5 gate sits above reassociation noise."""
    B, shape, params, inp, stride = build_case('std_L1')
    kw = dict(num_layers=shape.L, num_heads=shape.heads, layer_norm_eps=shape.eps)
    f32 = fusion_ref.fusion_segment(inp['text_states'], inp['visual_embeds_att'], inp['clip_features'],
                                    inp['token_embedding'], inp['img_mask'], inp['text_mask'], params, **kw)
    p64 = {k: v.double() for k, v in params.items()}
    f64 = fusion_ref.fusion_segment(inp['text_states'].double(), inp['visual_embeds_att'].double(),
                                    inp['clip_features'].double(), inp['token_embedding'].double(),
                                    inp['img_mask'], inp['text_mask'], p64, **kw)
    assert rel_err(f32['result'].double(), f64['result']) < 5e-6
