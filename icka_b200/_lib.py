"""ctypes binding of libicka_b200.so (the C ABI declared in include/icka_b200.h).

The library has no CPU fallback: importing works anywhere (so CPU-only tests can check that every
symbol is exported), but ``handle()`` -- needed by every compute call -- raises unless a sm_100 GPU is
present, and every non-zero status becomes a ``RuntimeError`` carrying ``icka_last_error()``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libicka_b200.so')

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU_ERF, ACT_GELU_ERF_BWD, ACT_TANH, ACT_RELU, ACT_SWISH = 0, 1, 2, 3, 4, 5

# name -> (restype, argtypes); must list every symbol of include/icka_b200.h
SIGNATURES = {
    'icka_version': (c_int, []),
    'icka_last_error': (c_char_p, []),
    'icka_create': (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    'icka_destroy': (c_int, [c_void_p]),
    'icka_set_seed_base': (c_int, [c_void_p, c_void_p]),
    'icka_launch_count': (c_int64, [c_void_p]),
    'icka_cast_f32_to_bf16': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'icka_cast_bf16_to_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'icka_split_bf16x3': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    'icka_region_rows': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_mask_additive': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    'icka_linear_fwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_linear_ln_fwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_linear_fwd_ex': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_linear_dgrad': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                  c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_linear_wgrad': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_int,
                                  c_int, c_int, c_void_p]),
    'icka_act_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    'icka_colsum': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    'icka_layernorm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'icka_cross_attn_core_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                         c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                         c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_gate_blend_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                    c_int, c_int, c_void_p]),
    'icka_gate_fold_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, c_void_p]),
    'icka_crf_llh_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'icka_set_gemm_mode': (c_int, [c_int]),
    'icka_set_ln_mode': (c_int, [c_int]),
    'icka_set_attn_mode': (c_int, [c_int]),
    'icka_layernorm_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                                   c_int, c_void_p]),
    'icka_cross_attn_core_fwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                         c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_cross_attn_core_fwd_drop': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                              c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                              c_uint64, c_void_p]),
    'icka_cross_attn_core_bwd_drop': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                              c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                              c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                              c_uint64, c_void_p]),
    'icka_dropout_fwd': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_float, c_uint64,
                                 c_void_p]),
    'icka_dropout_mask': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_uint64, c_void_p]),
    'icka_i2t_pool_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_gate_fold': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_void_p]),
    'icka_gate_blend_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'icka_ln_gate_blend_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                       c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_int, c_void_p]),
    'icka_viterbi_decode': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_void_p]),
    'icka_crf_llh_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_void_p]),
    'icka_lstm_rec_workspace_bytes': (c_int64, [c_int, c_int]),
    'icka_lstm_rec_variant': (c_int, [c_int]),
    'icka_lstm_rec_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                  c_int, c_int, c_int, c_void_p]),
    'icka_lstm_cell_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_void_p, c_int, c_int, c_int, c_void_p]),
    'icka_lstm_cell_fwd_save': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    'icka_lstm_cell_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_int, c_int, c_int, c_void_p]),
    'icka_lstm_dir_fwd_save': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_lstm_dir_bwd': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_lstm_bidir_fwd_save': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_void_p]),
    'icka_lstm_bidir_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_void_p]),
    'icka_emission_head_fwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int,
                                       c_int, c_int, c_void_p]),
    'icka_emission_head_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                       c_int, c_int, c_int, c_int, c_void_p]),
    'icka_cast_bf16_time_major': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'icka_add_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'icka_region_tail_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    'icka_ner_chunk_counts': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_int, c_int, c_void_p]),
}

_lib = None
_handles = {}
_seed_base = {}          # device index -> (pointer, keep-alive object): applied to every handle (slot) of the device
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """dlopen the library and attach prototypes.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m icka_b200.build` '
                '(icka_b200 has no CPU or PyTorch fallback)')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return (load().icka_last_error() or b'').decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f'{what} failed ({rc}): {last_error()}')


_slot = threading.local()


class use_slot:
    """Context manager: calls made inside use handle `slot` of each device (its own split-K workspace).  The C library's
    rule is one stream at a time per handle; work that runs CONCURRENTLY on several streams (two captured steps in flight)
    therefore takes one slot per stream.  Slot 0 is the default."""

    def __init__(self, slot: int):
        self.slot = int(slot)

    def __enter__(self):
        self.prev = getattr(_slot, 'value', 0)
        _slot.value = self.slot
        return self

    def __exit__(self, *exc):
        _slot.value = self.prev
        return False


def handle(device_index: int) -> c_void_p:
    """Per-device library handle of the current slot (created on first use, kept for the life of the process)."""
    key = (device_index, getattr(_slot, 'value', 0))
    with _lock:
        h = _handles.get(key)
        if h is None:
            lib = load()
            out = c_void_p()
            check(lib.icka_create(int(device_index), ctypes.byref(out)), 'icka_create')
            h = _handles[key] = out
            base = _seed_base.get(device_index)
            if base is not None:
                check(lib.icka_set_seed_base(h, base[0]), 'icka_set_seed_base')
        return h


def set_seed_base(device_index: int, ptr, keep_alive=None) -> None:
    """Device-resident dropout seed base for every handle (slot) of ``device_index``; ``ptr`` None clears it."""
    lib = load()
    with _lock:
        if ptr is None:
            _seed_base.pop(device_index, None)
        else:
            _seed_base[device_index] = (int(ptr), keep_alive)
        for (d, _), h in _handles.items():
            if d == device_index:
                check(lib.icka_set_seed_base(h, None if ptr is None else int(ptr)), 'icka_set_seed_base')


def launch_count(device_index: int = 0) -> int:
    """Kernels launched so far on `device_index` through every slot's handle."""
    lib = load()
    return sum(int(lib.icka_launch_count(h)) for (d, _), h in list(_handles.items()) if d == device_index)
