"""Region-tail oracle against the reference's own myResnet.forward (identity backbone) where /root/reference exists."""
import os
import sys

import pytest
import torch

from oracle import reference_shim, region_tail_ref


@pytest.mark.skipif(not reference_shim.available(), reason='/root/reference not present')
@pytest.mark.parametrize('g,a', [(7, 7), (14, 7), (14, 14), (7, 14)])
def test_oracle_matches_reference_myresnet(g, a):
    sys.dont_write_bytecode = True
    if reference_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, reference_shim.REFERENCE_ROOT)
    from resnet.resnet_utils import myResnet as RefResnet
    torch.manual_seed(g * 10 + a)
    x = torch.relu(torch.randn(2, 32, g, g))
    ref = RefResnet(region_tail_ref.IdentityBackbone(), False, 'cpu')
    pooled, fc, att = ref(x, att_size=a)
    mfc, matt, rows = region_tail_ref.region_tail(x, a)
    assert torch.equal(fc, mfc) and torch.equal(att, matt)
    assert torch.equal(rows, att.view(-1, 32, a * a).permute(0, 2, 1))          # CMIM:956
    if g == 7:
        assert torch.allclose(pooled, fc, atol=1e-6)


def test_adaptive_bins_by_hand():
    x = torch.arange(16.0).view(1, 1, 4, 4)
    _, att, rows = region_tail_ref.region_tail(x, 2)
    assert att.view(-1).tolist() == [2.5, 4.5, 10.5, 12.5]
    assert rows.shape == (1, 4, 1)
