"""k-way partition of the mesh's dual graph (cells = vertices, shared facets = edges) with METIS.

The role PETSc DMPlex distribution plays for the reference's MPI runs (seigen/elastic.py:404-414).
``libsg_metis.so`` (built in-tree by ``csrc/Makefile`` from ``csrc/sg_metis.c`` + the METIS static library of the
CUDA toolkit) exports ``sg_partition_graph`` (include/seigen_b200.h).  Deterministic: every rank calls it on the
same graph and obtains the same partition, so no broadcast is needed.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .mesh import BOUNDARY

__all__ = ["dual_graph", "metis_partition", "edge_cut", "LIB_PATH", "available"]

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsg_metis.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} not found: build it with `make -C seigen_b200/csrc metis` "
                               "(needs libmetis_static.a of the CUDA toolkit); use partition method 'rcb' otherwise")
        lib = C.CDLL(LIB_PATH)
        lib.sg_partition_graph.restype = C.c_int
        lib.sg_partition_graph.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                           C.POINTER(C.c_int64)]
        _lib = lib
    return _lib


def dual_graph(topology):
    """CSR (xadj, adjncy) of the dual graph: one edge per interior facet, both directions."""
    nbr, code = topology.nbr, topology.code
    E, nf = nbr.shape
    interior = (code & BOUNDARY) == 0
    deg = interior.sum(axis=1)
    xadj = np.zeros(E + 1, dtype=np.int64)
    np.cumsum(deg, out=xadj[1:])
    adjncy = np.ascontiguousarray(nbr[interior].astype(np.int64))       # row-major: grouped by cell
    return xadj, adjncy


def edge_cut(topology, part):
    """Number of interior facets whose two cells belong to different parts."""
    nbr, code = topology.nbr, topology.code
    interior = (code & BOUNDARY) == 0
    part = np.asarray(part)
    return int(((part[:, None] != part[nbr]) & interior).sum() // 2)


def metis_partition(topology, nparts, recursive=False):
    """Owner rank (int32) of every cell."""
    E = topology.nbr.shape[0]
    if nparts <= 1:
        return np.zeros(E, dtype=np.int32)
    xadj, adjncy = dual_graph(topology)
    part = np.zeros(E, dtype=np.int64)
    cut = C.c_int64()
    rc = _load().sg_partition_graph(E, xadj.ctypes.data, adjncy.ctypes.data, int(nparts), int(bool(recursive)),
                                    part.ctypes.data, C.byref(cut))
    if rc != 0:
        raise RuntimeError(f"sg_partition_graph failed (rc = {rc})")
    if len(np.unique(part)) != nparts:
        raise RuntimeError("METIS returned an empty part")
    return part.astype(np.int32)
