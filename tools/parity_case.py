"""Print the GPU-vs-golden errors of one fusion case (tests/golden/fusion_<name>.npz) in both precision modes.

    python tools/parity_case.py std_L5
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_fusion as t      # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'std_L5'
g = np.load(os.path.join(ROOT, 'tests', 'golden', f'fusion_{name}.npz'))
ok = True
for precision, gate in (('fp32', 1e-5), ('bf16', 2e-2)):
    B, shape, params, inp, stride, out = t.run_ours(name, precision)
    for k in ('fused', 'result', 'clip', 'gate'):
        ref = torch.from_numpy(g[k])
        got = out[k][:, ::stride] if k in ('fused', 'result') else out[k]
        got = got.reshape(ref.shape)
        err = t.rel(got, ref) if precision == 'fp32' else float((got - ref).abs().max())
        ok &= err <= gate
        print(f'{name} {precision} {k}: err {err:.3e} (gate {gate})')
print('PASS' if ok else 'FAIL')
