"""GPU parity of the region-producer tail (icka_region_tail_fwd, icka_b200.myResnet; resnet/resnet_utils.py:36-43 +
CMIM:956) against the oracle; and CrossModalFusion fed with the rows equals CrossModalFusion fed with the grid."""
import pytest
import torch

import icka_b200
from icka_b200 import ops, synth
from oracle import region_tail_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_precision():
    prev = icka_b200.get_precision()
    yield
    icka_b200.set_precision(prev)


@pytest.mark.parametrize('B,C,g,a', [(3, 2048, 7, 7), (2, 2048, 14, 7), (2, 2048, 14, 14), (1, 100, 7, 14), (5, 64, 3, 2)])
def test_tail_kernel(B, C, g, a):
    torch.manual_seed(B + C + g + a)
    x = torch.relu(torch.randn(B, C, g, g)) * 0.5
    fc, att, rows = region_tail_ref.region_tail(x.double(), a)
    gfc, gatt, grows = ops.region_tail(x.cuda(), a, rows_dtype=torch.float32)
    assert (gfc.cpu().double() - fc).abs().max().item() <= 1e-6
    assert (gatt.cpu().double() - att).abs().max().item() <= 1e-6
    assert (grows.cpu().double() - rows).abs().max().item() <= 1e-6
    _, _, brows = ops.region_tail(x.cuda(), a, want_fc=False, want_att=False, rows_dtype=torch.bfloat16)
    assert torch.equal(brows.cpu(), rows.float().to(torch.bfloat16)) or \
        (brows.float().cpu() - rows.float()).abs().max().item() <= 4e-3


def test_myresnet_mirror_and_fusion_rows_path():
    icka_b200.set_precision('bf16')
    shape = synth.Shape(L=1)
    B = 4
    inp = synth.fusion_inputs(B, shape, seed=8)
    dev = torch.device('cuda')
    net = icka_b200.myResnet(region_tail_ref.IdentityBackbone(), False, dev)
    grid = inp['visual_embeds_att'].to(dev)
    pooled, fc, att = net(grid, att_size=7)
    assert torch.equal(att, grid) and pooled.shape == (B, 2048)
    assert (fc.cpu() - inp['visual_embeds_att'].mean(3).mean(2)).abs().max().item() <= 1e-6
    fc2, rows = net.forward_rows(grid, att_size=7)
    assert rows.shape == (B, 49, 2048) and rows.dtype == torch.bfloat16
    cfg = icka_b200.FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads, intermediate_size=shape.inter,
                                 layer_norm_eps=shape.eps)
    torch.manual_seed(1)
    model = icka_b200.CrossModalFusion(cfg, layer_num1=1).to(dev).eval()
    args = [inp[k].to(dev) for k in ('text_states', 'visual_embeds_att', 'clip_features', 'token_embedding', 'img_mask',
                                     'text_mask')]
    with torch.no_grad():
        want, _ = model(*args)
        args[1] = rows
        got, _ = model(*args)
    assert torch.equal(got, want)                                # same bf16 rows -> bit-identical downstream
    with pytest.raises(NotImplementedError):
        icka_b200.myResnet(region_tail_ref.IdentityBackbone(), True, dev)
