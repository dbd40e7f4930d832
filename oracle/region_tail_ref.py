"""CPU restatement of the region-producer tail (TEST INFRASTRUCTURE ONLY).

  resnet/resnet_utils.py:36-43   fc = x.mean(3).mean(2); att = F.adaptive_avg_pool2d(x, [att_size, att_size]);
                                 x = resnet.avgpool(x).view(B, -1)
  Cross_Modal_Interaction_Module.py:956   rows = att.view(-1, C, R).permute(0, 2, 1)

The arithmetic is torch's own (the reference is pure PyTorch).  Pinned: tests/test_oracle_region_tail.py runs the
reference's OWN ``myResnet.forward`` (imported from /root/reference when present) on a backbone whose layers are
identities, so its tail statements execute on a synthetic layer4 map, and compares.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def region_tail(x: torch.Tensor, att_size: int = 7):
    """x [B,C,g,g] -> (fc [B,C], att [B,C,a,a], rows [B,a*a,C])."""
    fc = x.mean(3).mean(2)
    att = F.adaptive_avg_pool2d(x, [att_size, att_size])
    B, C = x.shape[:2]
    rows = att.reshape(B, C, att_size * att_size).permute(0, 2, 1)
    return fc, att, rows


class IdentityBackbone(torch.nn.Module):
    """Stands in for torchvision's ResNet so that ``myResnet.forward`` exercises only its tail."""

    def __init__(self):
        super().__init__()
        for name in ('conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4'):
            setattr(self, name, torch.nn.Identity())
        self.avgpool = torch.nn.AvgPool2d(7, stride=1)          # resnet/resnet.py:110
