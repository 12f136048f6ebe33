"""ctypes driver of ``oracle/elastic_c.c``  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Builds the context of the C restatement from an ``ElasticOracle`` (the literal NumPy restatement): same
quadrature, same tabulations, same facet normals; only the loops run in C with OpenMP.  Used as the timed
CPU baseline (``bench.py``: ``cpu_baseline`` and ``--impl reference``) and as a faster checker for long
parity runs.  Parity unpinned at the Firedrake boundary (see ``elastic_oracle.py``).
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
import subprocess

import numpy as np

from .elastic_oracle import ElasticOracle, simplex_quadrature

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_c.so")

_P = C.c_void_p


class _Ctx(C.Structure):
    _fields_ = [("d", C.c_int), ("nd", C.c_int), ("nq", C.c_int), ("nfq", C.c_int), ("nperm", C.c_int),
                ("E", C.c_int64), ("nif", C.c_int64), ("nef", C.c_int64),
                ("wq", _P), ("phi", _P), ("dphi", _P), ("fw", _P), ("phif", _P), ("jinv", _P), ("detj", _P),
                ("minv", _P), ("mass", _P), ("ifac", _P), ("inrm", _P), ("imeas", _P), ("efac", _P),
                ("enrm", _P), ("emeas", _P), ("lam", _P), ("mu", _P), ("sigq", _P), ("density", C.c_double)]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "elastic_c.c")):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "liboracle_c.so"])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_threads.argtypes = [C.c_int]
        _lib.oracle_set_threads.restype = None
        _lib.oracle_solve_f.argtypes = [C.POINTER(_Ctx), _P, _P, _P, _P]
        _lib.oracle_solve_g.argtypes = [C.POINTER(_Ctx), _P, _P, _P, _P]
        _lib.oracle_step.argtypes = [C.POINTER(_Ctx), _P, _P, _P, C.c_double] + [_P] * 6
    return _lib


def _p(a):
    return a.ctypes.data_as(_P) if a is not None else None


class COracle:
    """Same scheme as ``ElasticOracle`` with the loops in C/OpenMP; parameters are copied from ``orc``."""

    def __init__(self, orc: ElasticOracle):
        self.orc = orc
        d, nd, E = orc.dim, orc.nd, orc.E
        self.d, self.nd, self.E = d, nd, E
        self._keep = []
        k = self._keep.append
        perms = list(itertools.permutations(range(d)))
        pidx = {p: i for i, p in enumerate(perms)}
        fq, fw = simplex_quadrature(d - 1, 2 * orc.p)
        # reference vertices of the cell, facet f = all but vertex f (increasing local index)
        rv = np.vstack([np.zeros(d), np.eye(d)])
        vid = [[i for i in range(d + 1) if i != f] for f in range(d + 1)]
        phif = np.zeros((d + 1, len(perms), len(fw), nd))
        for f in range(d + 1):
            for pi, perm in enumerate(perms):
                vs = rv[[vid[f][j] for j in perm]]                     # (d, d) facet vertices in order perm
                if d == 2:
                    x = vs[0][None] + fq[:, 0:1] * (vs[1] - vs[0])[None]
                else:
                    x = vs[0][None] + fq[:, 0:1] * (vs[1] - vs[0])[None] + fq[:, 1:2] * (vs[2] - vs[0])[None]
                phif[f, pi] = orc.el.tab(x)
        cells = orc.cells

        # perm-: the order in which the '-' cell's facet vertices must be taken to follow the '+' cell's facet vertices
        # (matched through their global ids), as an index into `perms`
        ident = pidx[tuple(range(d))]
        vid_a = np.array(vid)                                                   # (d+1, d) local vertices of facet f
        ifac = np.zeros((len(orc.int_facets), 6), dtype=np.int32)
        if len(ifac):
            ep, fp, em, fm = orc.int_facets.T
            gp = cells[ep[:, None], vid_a[fp]]                                  # (n, d) global ids, '+' side order
            gm = cells[em[:, None], vid_a[fm]]                                  # (n, d) global ids, '-' side order
            loc = np.argmax(gm[:, None, :] == gp[:, :, None], axis=2)           # loc[n, j]: where gp[n, j] sits in gm[n]
            assert np.all(np.take_along_axis(gm, loc, axis=1) == gp)
            code = (loc * (d ** np.arange(d))[None, :]).sum(axis=1)
            lut = np.full(d ** d, -1, dtype=np.int32)
            for perm, i in pidx.items():
                lut[sum(perm[j] * d ** j for j in range(d))] = i
            pm = lut[code]
            assert (pm >= 0).all()
            ifac[:, 0], ifac[:, 1], ifac[:, 2] = ep, fp, ident
            ifac[:, 3], ifac[:, 4], ifac[:, 5] = em, fm, pm
        efac = np.zeros((len(orc.ext_facets), 3), dtype=np.int32)
        if len(efac):
            efac[:, 0], efac[:, 1], efac[:, 2] = orc.ext_facets[:, 0], orc.ext_facets[:, 1], ident
        if orc.iF is not None:
            inrm = np.ascontiguousarray(orc.iF["n_plus"])
            imeas = np.ascontiguousarray(orc.iF["w"][:, 0] / orc.fw_ref[0])
        else:
            inrm, imeas = np.zeros((0, d)), np.zeros(0)
        enrm = np.ascontiguousarray(orc.eF["n"])
        emeas = np.ascontiguousarray(orc.eF["w"][:, 0] / orc.fw_ref[0])
        arrs = dict(wq=np.ascontiguousarray(orc.wq), phi=np.ascontiguousarray(orc.phi),
                    dphi=np.ascontiguousarray(orc.dphi), fw=np.ascontiguousarray(fw),
                    phif=np.ascontiguousarray(phif), jinv=np.ascontiguousarray(orc.Jinv),
                    detj=np.ascontiguousarray(orc.detJ), minv=np.ascontiguousarray(orc.Minv_cell),
                    mass=np.ascontiguousarray(orc.Mcell), ifac=ifac, inrm=inrm, imeas=imeas, efac=efac,
                    enrm=enrm, emeas=emeas)
        self.arrs = arrs
        self.ctx = _Ctx(d=d, nd=nd, nq=len(orc.wq), nfq=len(fw), nperm=len(perms), E=E, nif=len(ifac),
                        nef=len(efac), density=1.0)
        for name, a in arrs.items():
            setattr(self.ctx, name, a.ctypes.data)
        self.sync_parameters()
        nU, nS = E * nd * d, E * nd * d * d
        self.bufU = [np.zeros(nU) for _ in range(3)]
        self.bufS = [np.zeros(nS) for _ in range(3)]
        self.lib = _load()

    @property
    def threads(self):
        return int(self.lib.oracle_num_threads())

    def set_threads(self, n):
        """Use ``n`` OpenMP threads from now on, whatever OMP_NUM_THREADS said at start-up."""
        self.lib.oracle_set_threads(int(n))
        return self.threads

    def sync_parameters(self):
        orc = self.orc
        self.lam = np.ascontiguousarray(orc._cellwise(orc.l)) if orc.l is not None else np.zeros(self.E)
        self.mu = np.ascontiguousarray(orc._cellwise(orc.mu)) if orc.mu is not None else np.zeros(self.E)
        self.ctx.lam = self.lam.ctypes.data
        self.ctx.mu = self.mu.ctypes.data
        self.ctx.density = float(orc.density)
        if orc.sigma is not None:
            self.sigq = np.ascontiguousarray(np.einsum("qb,eb->eq", orc.psi, orc.sigma))
            self.ctx.sigq = self.sigq.ctypes.data
        else:
            self.sigq = None
            self.ctx.sigq = None

    def solve_f(self, s, u0):
        s = np.ascontiguousarray(s, dtype=float)
        u0 = np.ascontiguousarray(u0, dtype=float)
        out = np.zeros((self.E, self.nd, self.d))
        self.lib.oracle_solve_f(C.byref(self.ctx), _p(s), _p(u0), _p(self.bufU[2]), _p(out))
        return out

    def solve_g(self, u, src):
        u = np.ascontiguousarray(u, dtype=float)
        src = np.ascontiguousarray(src, dtype=float) if src is not None else None
        out = np.zeros((self.E, self.nd, self.d, self.d))
        self.lib.oracle_solve_g(C.byref(self.ctx), _p(u), _p(src), _p(self.bufS[2]), _p(out))
        return out

    def step_inplace(self, u, s, src, dt):
        """u, s: C-contiguous float64 arrays updated in place (one time step)."""
        assert u.flags["C_CONTIGUOUS"] and s.flags["C_CONTIGUOUS"]
        src = np.ascontiguousarray(src, dtype=float) if src is not None else None
        self.lib.oracle_step(C.byref(self.ctx), _p(u), _p(s), _p(src), float(dt),
                             _p(self.bufU[0]), _p(self.bufU[1]), _p(self.bufU[2]),
                             _p(self.bufS[0]), _p(self.bufS[1]), _p(self.bufS[2]))

    def run(self, u0, s0, nsteps, dt, source=None, times=None):
        u = np.array(u0, dtype=float).reshape(self.E, self.nd, self.d).copy()
        s = np.array(s0, dtype=float).reshape(self.E, self.nd, self.d, self.d).copy()
        for n in range(nsteps):
            src = source(times[n] if times is not None else (n + 1) * dt) if source is not None else None
            self.step_inplace(u, s, src, dt)
        return u, s
