"""ctypes driver of ``oracle/elastic_fused_c.c``  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The fused six-pass CPU variant of BASELINE.md section 3 (the algorithm the GPU runs, in C/OpenMP).  Built from plain
arrays: the reference-element tables and facet adjacency in the form ``oracle/nodal.py`` takes them, plus an
``ElasticOracle`` for the sponge quadrature.  Used by ``bench.py``'s CPU legs (second baseline) and checked against the
literal oracle in ``tests/test_oracle_fused.py``.  Parity unpinned at the Firedrake boundary (see elastic_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_fused_c.so")
_P = C.c_void_p


class _Ctx(C.Structure):
    _fields_ = [("d", C.c_int), ("nd", C.c_int), ("nfp", C.c_int), ("nperm", C.c_int), ("E", C.c_int64),
                ("Dr", _P), ("Lift", _P), ("fnodes", _P), ("ftab", _P), ("nbr", _P), ("code", _P), ("jinv", _P),
                ("lam", _P), ("mu", _P), ("absidx", _P), ("absmat", _P), ("density", C.c_double)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "elastic_fused_c.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", HERE, "-s", "-B", "liboracle_fused_c.so"])
        _lib = C.CDLL(LIB)
        _lib.fused_num_threads.restype = C.c_int
        _lib.fused_set_threads.argtypes = [C.c_int]
        _lib.fused_step.argtypes = [C.POINTER(_Ctx), _P, _P, _P, C.c_double, _P, _P]
        _lib.fused_pass_f.argtypes = [C.POINTER(_Ctx)] + [_P] * 4 + [C.c_double] * 3 + [_P]
        _lib.fused_pass_g.argtypes = [C.POINTER(_Ctx)] + [_P] * 4 + [C.c_double] * 3 + [_P]
    return _lib


def _p(a):
    return a.ctypes.data_as(_P) if a is not None else None


class CFused:
    """``Dr, Lift, fnodes, ftab, nbr, code, jinv`` as for ``oracle.nodal.NodalOperator``; ``lam, mu`` scalars or per
    cell; ``sigma_mats``: optional (cells, A) with A[k] the nd x nd sponge matrix of cell cells[k]."""

    def __init__(self, Dr, Lift, fnodes, ftab, nbr, code, jinv, lam, mu, density=1.0, sigma_mats=None):
        self.lib = _load()
        Dr = np.ascontiguousarray(Dr, dtype=np.float64)
        d, nd = Dr.shape[0], Dr.shape[1]
        nbr = np.ascontiguousarray(nbr, dtype=np.int32)
        E = nbr.shape[0]
        fnodes = np.ascontiguousarray(fnodes, dtype=np.int32)
        nfp = fnodes.shape[1]
        ftab = np.ascontiguousarray(np.asarray(ftab).reshape(-1, nfp), dtype=np.int32)
        self.a = dict(Dr=Dr, Lift=np.ascontiguousarray(Lift, dtype=np.float64), fnodes=fnodes, ftab=ftab, nbr=nbr,
                      code=np.ascontiguousarray(code, dtype=np.uint8), jinv=np.ascontiguousarray(jinv, dtype=np.float64),
                      lam=np.ascontiguousarray(np.broadcast_to(np.asarray(lam, dtype=float), (E,))),
                      mu=np.ascontiguousarray(np.broadcast_to(np.asarray(mu, dtype=float), (E,))))
        self.d, self.nd, self.E = d, nd, E
        self.ctx = _Ctx(d=d, nd=nd, nfp=nfp, nperm=ftab.shape[0] // (d + 1), E=E, density=float(density))
        for k, v in self.a.items():
            setattr(self.ctx, k, v.ctypes.data)
        if sigma_mats is not None and len(sigma_mats[0]):
            cells, A = sigma_mats
            idx = np.full(E, -1, dtype=np.int32)
            idx[np.asarray(cells)] = np.arange(len(cells), dtype=np.int32)
            self.a["absidx"] = idx
            self.a["absmat"] = np.ascontiguousarray(A, dtype=np.float64)
            self.ctx.absidx = idx.ctypes.data
            self.ctx.absmat = self.a["absmat"].ctypes.data
        self.uh = np.zeros((E, nd, d))
        self.sh = np.zeros((E, nd, d, d))

    @staticmethod
    def sponge_matrices(orc):
        """(cells, A) from an ElasticOracle's absorption field: A = Mref^-1 int phi_a sigma phi_c (|detJ| cancels)."""
        if orc.sigma is None:
            return None
        cells = np.flatnonzero(np.any(orc.sigma != 0.0, axis=1))
        sigq = np.einsum("qb,eb->eq", orc.psi, orc.sigma[cells])
        Mref = np.einsum("q,qa,qb->ab", orc.wq, orc.phi, orc.phi)
        T = np.einsum("q,qa,eq,qc->eac", orc.wq, orc.phi, sigq, orc.phi)
        return cells, np.einsum("ab,ebc->eac", np.linalg.inv(Mref), T)

    @property
    def threads(self):
        return int(self.lib.fused_num_threads())

    def set_threads(self, n):
        self.lib.fused_set_threads(int(n))
        return self.threads

    def step_inplace(self, u, s, src, dt):
        assert u.flags["C_CONTIGUOUS"] and s.flags["C_CONTIGUOUS"] and u.dtype == np.float64 and s.dtype == np.float64
        src = np.ascontiguousarray(src, dtype=float) if src is not None else None
        self.lib.fused_step(C.byref(self.ctx), _p(u), _p(s), _p(src), float(dt), _p(self.uh), _p(self.sh))
