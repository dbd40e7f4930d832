"""Host-side mirror of the reference interface: names, state_dict keys, error behaviour (no GPU needed)."""
import pytest
import torch

import icka_b200
from oracle import fusion_ref

# state_dict keys of the reference's BertCrossEncoder(config, 1) (verified against the reference class in
# tests/test_oracle_fusion.py::test_reference_state_dict_keys)
LAYER_KEYS = [
    'attention.self.query.weight', 'attention.self.query.bias', 'attention.self.key.weight',
    'attention.self.key.bias', 'attention.self.value.weight', 'attention.self.value.bias',
    'attention.output.dense.weight', 'attention.output.dense.bias', 'attention.output.LayerNorm.weight',
    'attention.output.LayerNorm.bias', 'intermediate.dense.weight', 'intermediate.dense.bias',
    'output.dense.weight', 'output.dense.bias', 'output.LayerNorm.weight', 'output.LayerNorm.bias']


def cfg(**kw):
    return icka_b200.FusionConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256, **kw)


def test_cross_encoder_state_dict_keys():
    enc = icka_b200.BertCrossEncoder(cfg(), 2)
    assert sorted(enc.state_dict()) == sorted(f'layer.{i}.{k}' for i in range(2) for k in LAYER_KEYS)
    assert sum(p.numel() for p in icka_b200.BertCrossEncoder(icka_b200.FusionConfig(), 1).parameters()) == 7087872


def test_layers_start_identical_like_the_reference():
    enc = icka_b200.BertCrossEncoder(cfg(), 3)
    sd = enc.state_dict()
    assert torch.equal(sd['layer.0.output.dense.weight'], sd['layer.2.output.dense.weight'])


def test_fusion_loads_oracle_params_and_aliases_layernorm():
    m = icka_b200.CrossModalFusion(cfg(), layer_num1=2, region_dim=64, clip_dim=32)
    p = fusion_ref.make_params(128, 2, 256, 2, region_dim=64, clip_dim=32)
    res = m.load_state_dict(p, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert m.cls_layer.proj_norm is m.cls_layer.LayerNorm                     # CMIM:876
    assert {'cls_layer.proj_norm.weight', 'cls_layer.LayerNorm.weight'} <= set(m.state_dict())
    n = sum(p.numel() for p in icka_b200.CrossModalFusion(icka_b200.FusionConfig(), 1).parameters())
    assert n == 1573632 + 393984 + 3 * 7087872 + 592128 + 769                 # SURVEY 8e, minus the CRF's 255


def test_hidden_size_not_multiple_of_heads():
    with pytest.raises(ValueError, match='not a multiple of the number of attention'):
        icka_b200.BertCoAttention(icka_b200.FusionConfig(hidden_size=100, num_attention_heads=12))


def test_dropout_needs_the_recording_path():
    """Training mode with p > 0: dropout is applied only when autograd records (the masks are regenerated in backward);
    forward-only use refuses rather than skip it."""
    enc = icka_b200.BertCrossEncoder(cfg(), 1).train()
    with torch.no_grad(), pytest.raises(NotImplementedError):
        enc(torch.zeros(1, 4, 128), torch.zeros(1, 3, 128), torch.zeros(1, 1, 1, 3))
    so = icka_b200.BertSelfOutput(cfg()).train()
    with torch.no_grad(), pytest.raises(NotImplementedError):
        so(torch.zeros(1, 4, 128), torch.zeros(1, 4, 128))
    pm = icka_b200.PromptMapping(cfg()).train()
    with torch.no_grad(), pytest.raises(NotImplementedError):
        pm(torch.zeros(1, 1, 128), torch.zeros(1, 2048), torch.ones(1, 8, dtype=torch.long))


def test_hidden_act_follows_act2fn():
    """config.hidden_act names an ACT2FN entry (CMIM:43); an unknown name is the reference's KeyError (CMIM:544)."""
    for name in ('gelu', 'relu', 'swish'):
        c = cfg()
        c.hidden_act = name
        icka_b200.BertIntermediate(c)
    c = cfg()
    c.hidden_act = 'mish'
    with pytest.raises(KeyError):
        icka_b200.BertIntermediate(c)


def test_precision_override_is_thread_local():
    import threading
    icka_b200.set_precision('bf16')
    seen = []
    with icka_b200.precision('fp32'):
        assert icka_b200.get_precision() == 'fp32'
        t = threading.Thread(target=lambda: seen.append(icka_b200.get_precision()))
        t.start()
        t.join()
        with icka_b200.precision(None):
            assert icka_b200.get_precision() == 'fp32'
    assert seen == ['bf16'] and icka_b200.get_precision() == 'bf16'
    with pytest.raises(ValueError):
        icka_b200.precision('fp8')


def test_crf_tag_range_is_an_index_error():
    crf = icka_b200.CRF(4, batch_first=True)
    e = torch.zeros(2, 3, 4)
    for bad in (-100, 4):
        tags = torch.zeros(2, 3, dtype=torch.long)
        tags[1, 2] = bad
        with pytest.raises(IndexError):
            crf(e, tags)


def test_precision_switch():
    icka_b200.set_precision('fp32')
    assert icka_b200.get_precision() == 'fp32'
    icka_b200.set_precision('bf16')
    with pytest.raises(ValueError):
        icka_b200.set_precision('fp8')


def test_crf_constructor_and_validation():
    with pytest.raises(ValueError, match='invalid number of tags'):
        icka_b200.CRF(0)
    crf = icka_b200.CRF(4, batch_first=True)
    assert sorted(crf.state_dict()) == ['end_transitions', 'start_transitions', 'transitions']
    assert crf.transitions.abs().max() <= 0.1
    e = torch.zeros(2, 3, 4)
    with pytest.raises(ValueError, match='dimension of 3'):
        crf.decode(torch.zeros(2, 3))
    with pytest.raises(ValueError, match='expected last dimension'):
        crf.decode(torch.zeros(2, 3, 5))
    with pytest.raises(ValueError, match='emissions and mask must match'):
        crf.decode(e, mask=torch.ones(3, 3, dtype=torch.bool))
    m = torch.ones(2, 3, dtype=torch.bool); m[1, 0] = False
    with pytest.raises(ValueError, match='first timestep'):
        crf.decode(e, mask=m)
    with pytest.raises(ValueError, match='emissions and tags must match'):
        crf(e, torch.zeros(2, 4, dtype=torch.long))
    with pytest.raises(ValueError, match='invalid reduction'):
        crf(e, torch.zeros(2, 3, dtype=torch.long), reduction='bogus')


def test_full_model_state_dict_keys_are_the_reference_names():
    """MTCCMBertForMMTokenClassificationCRF keeps the reference's flat parameter names (CMIM:886-935)."""
    import icka_b200
    from oracle import reference_shim
    cfg = icka_b200.FusionConfig(hidden_size=768)
    ours = icka_b200.MTCCMBertForMMTokenClassificationCRF(cfg, None, None, 1, 1, 1, num_labels=15)
    keys = set(ours.state_dict())
    for k in ('vismap2text.weight', 'txt2img_attention.layer.0.attention.self.query.weight', 'cls_layer_Y.1.layer.0.output.dense.bias',
              'cls_layer.proj.weight', 'aux_head.bias', 'lstm.weight_hh_l0_reverse', 'classifier.weight', 'crf.transitions',
              'mapping_network_alignment.1.weight', 'mapping_network_vision.4.bias', 'lastproj.weight'):
        assert k in keys
    assert not any(k.startswith(('_prompt', '_head', 'fusion.')) for k in keys)
    if reference_shim.available():
        import torch
        cmim = reference_shim.load()

        class _CRF(torch.nn.Module):
            def __init__(self, num_tags, batch_first=False):
                super().__init__()
                self.start_transitions = torch.nn.Parameter(torch.zeros(num_tags))
                self.end_transitions = torch.nn.Parameter(torch.zeros(num_tags))
                self.transitions = torch.nn.Parameter(torch.zeros(num_tags, num_tags))
        cmim.CRF = _CRF
        rcfg = cmim.BertConfig(30522, hidden_size=768, num_hidden_layers=1, num_attention_heads=12, intermediate_size=3072)
        ref = cmim.MTCCMBertForMMTokenClassificationCRF(rcfg, None, None, 1, 1, 1, num_labels=15)
        ref_sd = ref.state_dict()
        assert keys <= set(ref_sd), sorted(keys - set(ref_sd))
        for k in keys:
            assert tuple(ref_sd[k].shape) == tuple(ours.state_dict()[k].shape), k
        # what the reference has and the drop-in leaves out: only the members its forward never touches
        unused = {k.split('.')[0] for k in set(ref_sd) - keys}
        assert unused <= {'self_attention', 'self_attention_v2', 'embedding_layer', 'LayerNorm'}, unused


def test_host_batch_bf16_states_layout():
    """make_host_batch(bf16_states=True): the same values, bf16, regions as CMIM:956 rows [B, R, 2048]."""
    import torch
    from icka_b200 import synth
    from icka_b200.pipeline import FusionViterbiPipeline as P
    a = P.make_host_batch(3, synth.STD, 9, pin=False)
    b = P.make_host_batch(3, synth.STD, 9, pin=False, bf16_states=True)
    assert b['text_states'].dtype == b['token_embedding'].dtype == b['visual_embeds_att'].dtype == torch.bfloat16
    assert torch.equal(b['text_states'], a['text_states'].to(torch.bfloat16))
    rows = a['visual_embeds_att'].view(-1, 2048, 49).permute(0, 2, 1)
    assert b['visual_embeds_att'].shape == (3, 49, 2048) and b['visual_embeds_att'].is_contiguous()
    assert torch.equal(b['visual_embeds_att'], rows.to(torch.bfloat16))
    for k in ('clip_features', 'img_mask', 'text_mask', 'emissions', 'crf_mask'):
        assert torch.equal(a[k], b[k])
    assert P.h2d_bytes(b) < 0.51 * P.h2d_bytes(a)
