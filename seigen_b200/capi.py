"""ctypes binding of ``include/seigen_b200.h`` (``libseigen_b200.so``, built in-tree by ``csrc/Makefile``).

Host code stays Python (as in the reference, where ``seigen/elastic.py`` is pure Python over
PyOP2-generated C); everything numerical happens behind this C ABI.  There is deliberately no
fallback: if the shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["lib", "load", "SgError", "SgAsymmetric", "MeshDesc", "LIB_PATH", "check", "pinned_zeros", "PeerDesc",
           "FIELD_U", "FIELD_S", "FIELD_UH", "FIELD_SH", "PART_ALL", "PART_BOUNDARY", "PART_INTERIOR"]

LIB_PATH = os.environ.get("SG_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libseigen_b200.so")

FIELD_U, FIELD_S, FIELD_UH, FIELD_SH = 0, 1, 2, 3
PART_ALL, PART_BOUNDARY, PART_INTERIOR = 0, 1, 2


class SgError(RuntimeError):
    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


class SgAsymmetric(SgError):
    """SG_EASYM: asymmetric stress / source handed to a solver created with ``symmetric_stress = 1``."""


EASYM = -4


class MeshDesc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("degree", C.c_int32),
        ("n_owned", C.c_int64),
        ("n_total", C.c_int64),
        ("nbr", C.c_void_p),
        ("code", C.c_void_p),
        ("jinv", C.c_void_p),
        ("device", C.c_int32),
        ("n_boundary", C.c_int32),
        ("geom_classes", C.c_int32),
        ("symmetric_stress", C.c_int32),
    ]


class PeerDesc(C.Structure):
    _fields_ = [
        ("rank", C.c_int32),
        ("flag_slot", C.c_int32),
        ("send_offset", C.c_int64),
        ("send_count", C.c_int64),
        ("remote_first_cell", C.c_int64),
        ("handles", (C.c_ubyte * 64) * 5),
    ]


_P = C.c_void_p
_SIGNATURES = {
    "sg_last_error": (C.c_char_p, []),
    "sg_version": (C.c_int, []),
    "sg_create": (C.c_int, [C.POINTER(_P), C.POINTER(MeshDesc)]),
    "sg_destroy": (None, [_P]),
    "sg_set_material": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _P, _P]),
    "sg_set_absorption": (C.c_int, [_P, C.c_int64, _P, _P]),
    "sg_set_source": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P]),
    "sg_set_state": (C.c_int, [_P, _P, _P]),
    "sg_set_state_async": (C.c_int, [_P, _P, _P]),
    "sg_set_state_finish": (C.c_int, [_P]),
    "sg_get_state": (C.c_int, [_P, _P, _P]),
    "sg_get_field": (C.c_int, [_P, C.c_int, _P]),
    "sg_step": (C.c_int, [_P, C.c_int64, C.c_double, C.c_int64]),
    "sg_synchronize": (C.c_int, [_P]),
    "sg_last_step_ms": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "sg_mark": (C.c_int, [_P, C.c_int]),
    "sg_set_receivers": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64]),
    "sg_get_receivers": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "sg_stage": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_int64]),
    "sg_record_receivers": (C.c_int, [_P, C.c_int64]),
    "sg_time_stage": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "sg_set_halo_plan": (C.c_int, [_P, C.c_int64, _P]),
    "sg_pack": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "sg_unpack": (C.c_int, [_P, C.c_int, _P, C.c_int64, C.c_int64, C.c_int]),
    "sg_comm_wait_compute": (C.c_int, [_P]),
    "sg_compute_wait_comm": (C.c_int, [_P]),
    "sg_ipc_export": (C.c_int, [_P, _P]),
    "sg_peer_connect": (C.c_int, [_P, C.c_int32, C.POINTER(PeerDesc)]),
    "sg_exchange": (C.c_int, [_P, C.c_int]),
    "sg_peer_error": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "sg_stream": (_P, [_P, C.c_int]),
    "sg_field_ptr": (_P, [_P, C.c_int]),
    "sg_nodes_per_cell": (C.c_int, [C.c_int, C.c_int]),
    "sg_tile_cells": (C.c_int, [C.c_int, C.c_int]),
    "sg_host_alloc": (_P, [C.c_int64]),
    "sg_host_free": (None, [_P]),
}

_lib = None


def load():
    """Load the C-ABI library; raises ``SgError`` if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SgError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "or `make -C seigen_b200/csrc` (there is no CPU fallback)")
    lib_ = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib_, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib_
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


class _Lazy:
    def __getattr__(self, name):
        return getattr(load(), name)


lib = _Lazy()


def check(rc: int):
    if rc != 0:
        msg = load().sg_last_error()
        cls = SgAsymmetric if rc == EASYM else SgError
        raise cls(f"seigen_b200 error {rc}: {msg.decode() if msg else '?'}", rc)


def pinned_zeros(shape):
    """Zero-filled float64 array in page-locked host memory (``sg_host_alloc``), freed with the array.  The state
    Functions of ``ElasticLF4`` live in such arrays so that ``sg_set_state`` / ``sg_get_state`` run at PCIe speed."""
    import weakref
    n = int(np.prod(shape, dtype=np.int64))
    if n == 0:
        return np.zeros(shape)
    l = load()
    p = l.sg_host_alloc(n * 8)
    if not p:
        raise SgError("sg_host_alloc failed: " + (l.sg_last_error() or b"?").decode())
    buf = (C.c_double * n).from_address(p)
    a = np.ctypeslib.as_array(buf).reshape(shape)
    weakref.finalize(buf, l.sg_host_free, C.c_void_p(p))
    a[...] = 0.0
    return a


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
