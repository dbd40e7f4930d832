"""Generate tests/golden/fusion_*.npz by running the reference's OWN classes (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Each fixture stores the case description (shape, seeds) and the reference outputs; inputs and
parameters are regenerated from the seeds at test time (``icka_b200.synth.fusion_inputs``,
``oracle.fusion_ref.make_params``) and guarded by checksums stored next to the outputs.  Large outputs
keep every ``row_stride``-th text row so the fixtures stay small.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from icka_b200 import synth                                     # noqa: E402
from oracle import fusion_ref, reference_shim                   # noqa: E402

CASES = {
    # name: (B, Shape kwargs, param seed, input seed, row_stride, median_len)
    'tiny':  (3, dict(S=16, R=9, H=128, heads=2, inter=256, region_dim=64, clip_dim=32, L=2, eps=1e-12), 11, 12, 1, 9.0),
    'std_L1': (1, dict(L=1), 21, 22, 4, 28.0),
    'std_L2_eps5': (2, dict(L=2, eps=1e-5), 31, 32, 8, 28.0),
    'hires_L1': (1, dict(S=256, R=196, L=1), 41, 42, 8, 60.0),
    'std_L5': (1, dict(L=5), 51, 52, 8, 28.0),      # the training script's default depth (MCA:603)
    # the width the shipped script loads (roberta-large config: hidden 1024, 16 heads, intermediate 4096; MCA:660-672)
    'h1024_L1': (1, dict(H=1024, heads=16, inter=4096, L=1), 61, 62, 8, 28.0),
    # the `_bert` clone's FIVE image->text encoders (CMIM:1075)
    'std_L1_y5': (1, dict(L=1), 71, 72, 8, 28.0),
    # config.hidden_act = the other two ACT2FN entries (CMIM:43)
    'tiny_relu': (3, dict(S=16, R=9, H=128, heads=2, inter=256, region_dim=64, clip_dim=32, L=2, eps=1e-12), 81, 82, 1, 9.0),
    'tiny_swish': (3, dict(S=16, R=9, H=128, heads=2, inter=256, region_dim=64, clip_dim=32, L=2, eps=1e-12), 91, 92, 1, 9.0),
}
# per-case arguments beyond the Shape: number of image->text encoders, FFN activation
EXTRAS = {'std_L1_y5': dict(num_i2t_encoders=5), 'tiny_relu': dict(hidden_act='relu'), 'tiny_swish': dict(hidden_act='swish')}


def case_extras(name):
    return dict(EXTRAS.get(name, {}))


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def build_case(name):
    B, kw, pseed, iseed, stride, med = CASES[name]
    shape = synth.Shape(**kw)
    params = fusion_ref.make_params(shape.H, shape.heads, shape.inter, shape.L, seed=pseed,
                                    region_dim=shape.region_dim, clip_dim=shape.clip_dim,
                                    num_i2t_encoders=case_extras(name).get('num_i2t_encoders', 2))
    inp = synth.fusion_inputs(B, shape, seed=iseed, median_len=med)
    return B, shape, params, inp, stride


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for name in (sys.argv[1:] or CASES):      # optional: only the named cases
        B, shape, params, inp, stride = build_case(name)
        mods = reference_shim.build_reference_modules(
            params, hidden=shape.H, heads=shape.heads, inter=shape.inter, num_layers=shape.L,
            layer_norm_eps=shape.eps, **case_extras(name))
        ref = reference_shim.reference_fusion_segment(
            mods, inp['text_states'], inp['visual_embeds_att'], inp['clip_features'],
            inp['token_embedding'], inp['img_mask'], inp['text_mask'])
        path = os.path.join(out_dir, f'fusion_{name}.npz')
        np.savez_compressed(
            path,
            row_stride=np.int64(stride),
            params_checksum=np.float64(sum(checksum(v) for k, v in sorted(params.items()))),
            inputs_checksum=np.float64(sum(checksum(inp[k]) for k in
                                           ('text_states', 'visual_embeds_att', 'clip_features', 'token_embedding'))),
            regions=ref['regions'].numpy(),
            fused=ref['fused'][:, ::stride].numpy(),
            clip=ref['clip'].numpy(),
            result=ref['result'][:, ::stride].numpy(),
            gate=ref['gate'].numpy(),
        )
        print(f'{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)')


if __name__ == '__main__':
    main()
