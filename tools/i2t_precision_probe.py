"""Where does the bf16 error of the image->text (single CLIP query) chain come from?  Per-layer max |z - oracle| at
B = 256, L = 5 (std shape), for: the folded single-query form (default), the unfolded generic kernels, each fed with our
own `fused` and with the oracle's `fused` (isolates the chain from the text->image encoder's error)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icka_b200                                                   # noqa: E402
from icka_b200 import modules, ops, synth                          # noqa: E402
from oracle import fusion_ref                                      # noqa: E402

DEV = 'cuda:0'
B, L = int(os.environ.get('PROBE_B', 256)), int(os.environ.get('PROBE_L', 5))
shape = synth.Shape(L=L)
params = fusion_ref.make_params(shape.H, shape.heads, shape.inter, L, seed=100 + L)
inp = synth.fusion_inputs(B, shape, seed=200 + B)
KEYS = ('text_states', 'visual_embeds_att', 'clip_features', 'token_embedding', 'img_mask', 'text_mask')

with torch.no_grad():
    dt = torch.float32
    regions = fusion_ref.region_projection(inp['visual_embeds_att'], params)
    fused_layers = fusion_ref.cross_encoder(inp['text_states'], regions, fusion_ref.additive_mask(inp['img_mask'], dt), params,
                                            'txt2img_attention', L, shape.heads, shape.eps)
    fused = fused_layers[-1]
    z = fusion_ref.linear(inp['clip_features'].squeeze(1), params, 'vismapping').unsqueeze(1)
    tmask = fusion_ref.additive_mask(inp['text_mask'], dt)
    z0_want = z.clone()
    want_z = []
    for e in range(2):
        outs = fusion_ref.cross_encoder(z, fused, tmask, params, f'cls_layer_Y.{e}', L, shape.heads, shape.eps)
        want_z += outs
        z = outs[-1]

cfg = icka_b200.FusionConfig(layer_norm_eps=shape.eps)
model = icka_b200.CrossModalFusion(cfg, layer_num1=L, precision='bf16').to(DEV).eval()
model.load_state_dict(params, strict=True)


def chain(fused_lp, fold, exact_z0=False):
    orig = modules.BertCrossAttention._can_fold
    if not fold:
        modules.BertCrossAttention._can_fold = lambda self: False
    try:
        with torch.no_grad(), icka_b200.precision('bf16'):
            clip_in = ops.cast_bf16(inp['clip_features'].to(DEV).float().reshape(B, -1).contiguous())
            w = ops.cast_bf16(model.vismapping.weight.detach().contiguous())
            z32 = ops.linear(clip_in, w, model.vismapping.bias.detach(), out_dtype=torch.float32)
            if exact_z0:
                z32 = ops.linear(inp['clip_features'].to(DEV).float().reshape(B, -1).contiguous(),
                                 model.vismapping.weight.detach().contiguous(), model.vismapping.bias.detach())
            print('   z0 err', float((z32.cpu().view(B, 1, -1) - z0_want).abs().max()), 'z0 std', float(z0_want.std()))
            z_lp = ops.cast_bf16(z32)
            txt_mask = ops.mask_additive(inp['text_mask'].to(DEV), shape.S)
            errs = []
            i = 0
            for enc in model.cls_layer_Y:
                for layer in enc.layer:
                    z32, z_lp = layer._run(z32, z_lp, fused_lp, txt_mask, B, 1, shape.S)
                    errs.append(float((z32.cpu().view(B, 1, -1) - want_z[i]).abs().max()))
                    i += 1
            return errs
    finally:
        modules.BertCrossAttention._can_fold = orig


with torch.no_grad():
    out = model(*[inp[k].to(DEV) for k in KEYS], return_dict=True)
    ours_fused = out['fused']
    print('fused err', float((ours_fused.cpu() - fused).abs().max()), 'clip err (model)', float((out['clip'].cpu() - want_z[-1]).abs().max()))
    ours_lp = ops.cast_bf16(ours_fused.reshape(B * shape.S, -1).contiguous())
    oracle_lp = ops.cast_bf16(fused.to(DEV).reshape(B * shape.S, -1).contiguous())
    for name, y, fold in (('folded, our fused', ours_lp, True), ('folded, oracle fused', oracle_lp, True),
                          ('unfolded, our fused', ours_lp, False), ('unfolded, oracle fused', oracle_lp, False)):
        print(f'{name:26s}', ' '.join(f'{e:.2e}' for e in chain(y, fold)))
    print('exact z0, folded, our fused', ' '.join(f'{e:.2e}' for e in chain(ours_lp, True, exact_z0=True)))
    # magnitude of the tensors involved
    print('|z| max per layer (oracle)', ' '.join(f'{float(t.abs().max()):.1f}' for t in want_z))
