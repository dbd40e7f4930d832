"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every declared symbol,
and the product path refuses to run without a CUDA device (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import icka_b200
from icka_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'icka_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(icka_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/icka_b200.h but not exported'
    assert set(syms) == set(_lib.SIGNATURES), 'ctypes prototypes out of sync with the header'
    assert lib.icka_version() >= 100


def test_library_is_plain_c_abi_without_torch():
    out = os.popen(f'ldd {_lib.LIB_PATH}').read()
    assert 'torch' not in out and 'c10' not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only check')
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        _lib.handle(0)
    with pytest.raises(RuntimeError):
        ops.cast_bf16(torch.zeros(8))
    crf = icka_b200.CRF(5, batch_first=True)
    with pytest.raises(RuntimeError):
        crf.decode(torch.zeros(2, 3, 5))
    cfg = icka_b200.FusionConfig(hidden_size=128, num_attention_heads=2, intermediate_size=256)
    enc = icka_b200.BertCrossEncoder(cfg, 1).eval()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 4, 128), torch.zeros(1, 3, 128), torch.zeros(1, 1, 1, 3))


def test_error_channel():
    lib = _lib.load()
    rc = lib.icka_create(0, None)
    assert rc < 0 and 'null' in _lib.last_error()
