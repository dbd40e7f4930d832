"""ctypes loader for oracle/viterbi_ref.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libicka_oracle.so')
        if not os.path.isfile(path):
            subprocess.check_call(['make', '-s', '-C', _HERE])
        _LIB = ctypes.CDLL(path)
        _LIB.icka_oracle_viterbi.restype = ctypes.c_int
    return _LIB


def viterbi(emissions: np.ndarray, mask, start, end, trans):
    """numpy in, (tags [B,S] int32 with -1 padding, lens [B] int32) out."""
    e = np.ascontiguousarray(emissions, dtype=np.float32)
    B, S, T = e.shape
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    st = np.ascontiguousarray(start, dtype=np.float32)
    en = np.ascontiguousarray(end, dtype=np.float32)
    tr = np.ascontiguousarray(trans, dtype=np.float32)
    tags = np.empty((B, S), dtype=np.int32)
    lens = np.empty((B,), dtype=np.int32)
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().icka_oracle_viterbi(p(e), p(m), p(st), p(en), p(tr), p(tags), p(lens),
                                   ctypes.c_int(B), ctypes.c_int(S), ctypes.c_int(T))
    if rc != 0:
        raise RuntimeError(f'icka_oracle_viterbi failed: {rc}')
    return tags, lens


def to_lists(tags: np.ndarray, lens: np.ndarray):
    return [tags[b, :lens[b]].tolist() for b in range(tags.shape[0])]
