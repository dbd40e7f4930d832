"""Generate tests/golden/prompt_prefix.npz from the reference's OWN modules (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_prompt           (build container only: needs /root/reference)

Constructs MTCCMBertForMMTokenClassificationCRF through the import shim (its `embedding` / `last_encoder` arguments are
None and torchcrf.CRF is a permissive stub: neither takes part here), loads the seeded parameters of
oracle/prompt_ref.make_params into its mapping_network_alignment / mapping_network_vision / lastproj, and runs the
statements CMIM:995-1009 on them in eval mode.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import prompt_ref, reference_shim                   # noqa: E402

CASES = prompt_ref.GOLDEN_CASES
inputs = prompt_ref.golden_inputs


def main():
    cmim = reference_shim.load()

    class _CRF(torch.nn.Module):         # torchcrf is absent; the CRF plays no part in these statements
        def __init__(self, *a, **k):
            super().__init__()
    cmim.CRF = _CRF
    out = {}
    for name, c in CASES.items():
        cfg = cmim.BertConfig(30522, hidden_size=c['H'], num_hidden_layers=1, num_attention_heads=12,
                              intermediate_size=3072, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        model = cmim.MTCCMBertForMMTokenClassificationCRF(cfg, None, None, 1, 1, 1, num_labels=15).eval()
        p = prompt_ref.make_params(c['H'], seed=c['seed'])
        missing, unexpected = model.load_state_dict(p, strict=False)
        assert not unexpected, unexpected
        clip, vmean, mask = inputs(c['B'], c['H'], c['L'], c['seed'])
        with torch.no_grad():
            B = clip.shape[0]
            alignment = model.mapping_network_alignment(clip).unsqueeze(1).view(B, model.prompt_len, -1)     # :995
            vision = model.mapping_network_vision(vmean).reshape(B, model.prompt_len, -1)                    # :998-999
            prefix = torch.cat([vision, alignment], dim=1)                                                   # :1002
            if prefix.size(2) != 1024:
                prefix = model.lastproj(prefix)                                                              # :1004
            pm = torch.cat([mask[:, :1].repeat(1, model.prompt_len)] * 2, dim=1)                             # :1007-1009
            mine, mine_mask = prompt_ref.prompt_prefix(clip, vmean, mask, p)
        err = float((mine - prefix).abs().max())
        assert err < 1e-5 and torch.equal(mine_mask, pm), err
        out[f'{name}_prefix'] = prefix.numpy()
        out[f'{name}_mask'] = pm.numpy()
        out[f'{name}_checksum'] = np.array([float(sum(v.double().abs().sum() for v in p.values())),
                                            float(clip.double().abs().sum() + vmean.double().abs().sum())])
        print(name, 'restatement vs reference modules: max |diff|', err)
    path = os.path.join(ROOT, 'tests', 'golden', 'prompt_prefix.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
