"""Region-producer tail (SURVEY 8f "next" row 3): drop-in for ``resnet/resnet_utils.py:myResnet``.

Same constructor ``myResnet(resnet, if_fine_tune, device)`` and ``forward(x, att_size=7) -> (x, fc, att)``.  The
convolutional backbone stays the caller's torchvision ResNet-152 (cuDNN; outside the hot path); what
changes is everything after ``layer4``: the two means, the adaptive pooling and -- new -- the K-major bf16 region rows
``[B, R, 2048]`` the region projection wants are produced by ONE pass over the layer4 output (``icka_region_tail_fwd``).
``forward_rows`` returns those rows; ``CrossModalFusion`` takes them in place of ``visual_embeds_att`` and skips its
relayout kernel.  ``att_size`` other than the map size (a 448-px input gives a 14 x 14 map) works for fc / att / rows;
the reference's first return value ``resnet.avgpool(x).view(B, -1)`` is only defined for the 7 x 7 map.
Inference only: ``if_fine_tune=True`` is refused (no backward through the tail).
"""
from __future__ import annotations

import torch
from torch import nn

from . import modules, ops


class myResnet(nn.Module):
    def __init__(self, resnet, if_fine_tune, device):
        super().__init__()
        if if_fine_tune:
            raise NotImplementedError('icka_b200.myResnet is inference-only: if_fine_tune must be False')
        self.resnet = resnet
        self.if_fine_tune = if_fine_tune
        self.device = device

    def _layer4(self, x):
        r = self.resnet                                             # resnet_utils.py:18-34 (torch / cuDNN backbone)
        x = r.maxpool(r.relu(r.bn1(r.conv1(x))))
        return r.layer4(r.layer3(r.layer2(r.layer1(x)))).float().contiguous()

    @torch.no_grad()
    def forward(self, x, att_size=7):
        x = self._layer4(x)
        fc, att, _ = ops.region_tail(x, att_size)                    # resnet_utils.py:37, 39
        if x.shape[2] == 7:
            pooled = fc                                              # AvgPool2d(7) of a 7 x 7 map = the same mean (:42-43)
        else:
            pooled = self.resnet.avgpool(x).view(x.size(0), -1)
        return pooled, fc, att

    @torch.no_grad()
    def forward_rows(self, x, att_size=7):
        """-> (fc [B,2048] fp32, rows [B, att_size**2, 2048] in the compute dtype): feed ``rows`` to CrossModalFusion."""
        dt = torch.bfloat16 if modules.get_precision() == 'bf16' else torch.float32
        fc, _, rows = ops.region_tail(self._layer4(x), att_size, want_att=False, rows_dtype=dt)
        return fc, rows
