"""Which rounding of the single-query (image->text) layer costs the most?  Re-runs the 10-layer chain at B=256, L=5 with
ONE stage at a time moved to fp32 (exact z0 throughout), printing the per-layer max |z - oracle|."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icka_b200                                                   # noqa: E402
from icka_b200 import ops, synth                                   # noqa: E402
from icka_b200._lib import ACT_GELU_ERF                            # noqa: E402
from oracle import fusion_ref                                      # noqa: E402

DEV = 'cuda:0'
B, L = int(os.environ.get('PROBE_B', 256)), int(os.environ.get('PROBE_L', 5))
shape = synth.Shape(L=L)
params = fusion_ref.make_params(shape.H, shape.heads, shape.inter, L, seed=100 + L)
inp = synth.fusion_inputs(B, shape, seed=200 + B)
H, nh, S = shape.H, shape.heads, shape.S
with torch.no_grad():
    regions = fusion_ref.region_projection(inp['visual_embeds_att'], params)
    fused = fusion_ref.cross_encoder(inp['text_states'], regions, fusion_ref.additive_mask(inp['img_mask'], torch.float32),
                                     params, 'txt2img_attention', L, nh, shape.eps)[-1]
    z = fusion_ref.linear(inp['clip_features'].squeeze(1), params, 'vismapping').unsqueeze(1)
    tmask = fusion_ref.additive_mask(inp['text_mask'], torch.float32)
    want_z = []
    for e in range(2):
        outs = fusion_ref.cross_encoder(z, fused, tmask, params, f'cls_layer_Y.{e}', L, nh, shape.eps)
        want_z += outs
        z = outs[-1]
P = {k: v.to(DEV) for k, v in params.items()}
bf = ops.cast_bf16
fused32 = fused.to(DEV).reshape(B * S, H).contiguous()
fused16 = bf(fused32)
txt_mask = ops.mask_additive(inp['text_mask'].to(DEV), S)


def lin(x32, w, b, fp32, **kw):
    if fp32:
        return ops.linear(x32, w.contiguous(), b, out_dtype=torch.float32, **kw)
    return ops.linear(bf(x32), bf(w.contiguous()), b, out_dtype=torch.float32, **kw)


def layer(z32, pre, flags):
    """One single-query cross layer, unfolded, every GEMM either bf16-operand or fp32 according to `flags`."""
    g = lambda n: P[f'{pre}.{n}']
    q = lin(z32, g('attention.self.query.weight'), g('attention.self.query.bias'), 'q' in flags)
    wkv = torch.cat([g('attention.self.key.weight'), g('attention.self.value.weight')])
    bkv = torch.cat([g('attention.self.key.bias'), g('attention.self.value.bias')])
    kv = lin(fused32, wkv, bkv, 'kv' in flags)
    if 'core' in flags:
        ctx = ops.cross_attn_core(q, kv[:, :H], kv[:, H:], txt_mask, B, 1, S, nh, H // nh)
    else:
        q16, kv16 = bf(q), bf(kv)
        ctx = ops.cast_f32(ops.cross_attn_core(q16, kv16[:, :H], kv16[:, H:], txt_mask, B, 1, S, nh, H // nh))
    pre1 = lin(ctx, g('attention.output.dense.weight'), g('attention.output.dense.bias'), 'out' in flags, residual=z32)
    a32, _ = ops.layernorm(pre1, g('attention.output.LayerNorm.weight'), g('attention.output.LayerNorm.bias'), shape.eps)
    if 'up' in flags:
        f = ops.linear(a32, g('intermediate.dense.weight'), g('intermediate.dense.bias'), act=ACT_GELU_ERF, out_dtype=torch.float32)
    else:
        f = ops.cast_f32(ops.linear(bf(a32), bf(g('intermediate.dense.weight')), g('intermediate.dense.bias'), act=ACT_GELU_ERF,
                                    out_dtype=torch.bfloat16))
    pre2 = lin(f, g('output.dense.weight'), g('output.dense.bias'), 'down' in flags, residual=a32)
    o32, _ = ops.layernorm(pre2, g('output.LayerNorm.weight'), g('output.LayerNorm.bias'), shape.eps)
    return o32


with torch.no_grad():
    z0 = ops.linear(inp['clip_features'].to(DEV).float().reshape(B, -1).contiguous(), P['vismapping.weight'], P['vismapping.bias'])
    for flags in ((), ('q',), ('kv',), ('core',), ('out',), ('up',), ('down',), ('up', 'down'), ('q', 'kv', 'core', 'out'),
                  ('q', 'kv', 'core', 'out', 'up', 'down')):
        z32, errs, i = z0, [], 0
        for e in range(2):
            for l in range(L):
                z32 = layer(z32, f'cls_layer_Y.{e}.layer.{l}', flags)
                errs.append(float((z32.cpu().view(B, 1, -1) - want_z[i]).abs().max()))
                i += 1
        print(f'fp32 stages {"+".join(flags) or "none":22s}', ' '.join(f'{x:.2e}' for x in errs))
