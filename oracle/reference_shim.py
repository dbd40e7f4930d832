"""Import the reference's OWN classes from /root/reference (TEST INFRASTRUCTURE ONLY).

Only usable where /root/reference exists (the build container).  Nothing run on the GPU box
(``-m gpu`` tests, ``smoke()``, ``bench.py``) may call this; those use ``oracle/fusion_ref.py`` and the
golden vectors under ``tests/golden/`` instead.

``import Cross_Modal_Interaction_Module`` needs three third-party packages that are not in this image
and play no part in the fusion arithmetic: ``torchcrf`` (CMIM:3), ``sparsemax``
(CMIM:17, referenced only from a comment) and ``boto3``/``botocore`` (my_bert/file_utils.py:20-22).
They are stubbed in ``sys.modules``; the reference tree is never written to.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('ICKA_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'Cross_Modal_Interaction_Module.py'))


def load():
    """Return the imported reference module ``Cross_Modal_Interaction_Module``."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    sys.dont_write_bytecode = True
    for name in ('torchcrf', 'sparsemax', 'boto3', 'botocore', 'botocore.exceptions'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['torchcrf'].CRF = getattr(sys.modules['torchcrf'], 'CRF', object)
    sys.modules['sparsemax'].Sparsemax = getattr(sys.modules['sparsemax'], 'Sparsemax', object)
    sys.modules['botocore.exceptions'].ClientError = getattr(
        sys.modules['botocore.exceptions'], 'ClientError', Exception)
    sys.modules['botocore'].exceptions = sys.modules['botocore.exceptions']
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import Cross_Modal_Interaction_Module as cmim  # noqa: E402
    return cmim


def build_reference_modules(params, *, hidden, heads, inter, num_layers, layer_norm_eps,
                            num_i2t_encoders=2, hidden_act='gelu'):
    """Instantiate the reference's classes and load ``params`` (reference key names) into them."""
    import torch
    cmim = load()
    cfg = cmim.BertConfig(30522, hidden_size=hidden, num_hidden_layers=12, num_attention_heads=heads,
                          intermediate_size=inter, hidden_act=hidden_act, hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0, layer_norm_eps=layer_norm_eps)
    mods = torch.nn.ModuleDict({
        'vismap2text': torch.nn.Linear(params['vismap2text.weight'].shape[1], hidden),   # CMIM:897
        'vismapping': torch.nn.Linear(params['vismapping.weight'].shape[1], hidden),     # CMIM:899
        'txt2img_attention': cmim.BertCrossEncoder(cfg, num_layers),                     # CMIM:900
        'cls_layer_Y': torch.nn.ModuleList([cmim.BertCrossEncoder(cfg, num_layers)       # CMIM:901
                                            for _ in range(num_i2t_encoders)]),
        'cls_layer': cmim.cls_layer_both(hidden, hidden),                                # CMIM:933
        'aux_head': torch.nn.Linear(hidden, 1),                                          # CMIM:934
    })
    missing, unexpected = mods.load_state_dict(params, strict=True)
    assert not missing and not unexpected
    return mods.eval()


def reference_fusion_segment(mods, text_states, visual_embeds_att, clip_features, token_embedding,
                             img_mask01, text_mask01):
    """Run the reference's statements CMIM:954-989, 1029-1036 on its own modules."""
    import torch
    with torch.no_grad():
        dt = text_states.dtype
        clip = mods['vismapping'](clip_features.to(dt).squeeze(1))                              # :954
        R = visual_embeds_att.shape[2] * visual_embeds_att.shape[3]
        vis = visual_embeds_att.view(-1, visual_embeds_att.shape[1], R).permute(0, 2, 1)        # :956
        regions = mods['vismap2text'](vis)                                                      # :958
        ext_img = img_mask01.unsqueeze(1).unsqueeze(2).to(dt)                                   # :962-964
        ext_img = (1.0 - ext_img) * -10000.0                                                    # :965
        fused = mods['txt2img_attention'](text_states, regions, ext_img)[-1]                    # :968-969
        ext_txt = (1.0 - text_mask01.to(dt)) * -10000.0                                         # :976-977
        clip = clip.unsqueeze(1)                                                                # :981
        ext_txt = ext_txt.unsqueeze(1).unsqueeze(2)                                             # :982
        for enc in mods['cls_layer_Y']:                                                         # :984-989
            clip = enc(clip, fused, ext_txt)[-1]
        feat = mods['cls_layer'](fused[:, 0, :], token_embedding[:, 0, :])                      # :1029-1033
        gate = torch.sigmoid(mods['aux_head'](feat)).view(token_embedding.size(0), 1, 1)        # :1034-1035
        result = gate * token_embedding + (1 - gate) * fused                                    # :1036
    return dict(regions=regions, fused=fused, clip=clip, result=result, gate=gate.view(-1))
