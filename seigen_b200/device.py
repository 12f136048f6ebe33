"""Thin object wrapper over the C ABI: one ``DeviceSolver`` = one ``sg_solver`` = one GPU's share of the mesh.

It owns nothing numerical: it hands the rank plan (adjacency, geometry) and the user's parameters
to ``libseigen_b200.so`` and moves fields across the boundary in Firedrake's ``dat.data`` layout
(``(n_cells*nd, d)`` / ``(n_cells*nd, d, d)``), in the *global* cell numbering of the mesh.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib, ptr
from .layout import RankPlan, build_rank_plan
from .mesh import Mesh
from .refelem import get_refelem

__all__ = ["DeviceSolver"]


class DeviceSolver:
    def __init__(self, mesh: Mesh, degree: int, device: int = 0, plan: RankPlan | None = None,
                 geom_classes: bool = True, symmetric: bool = False):
        self.mesh = mesh
        self.dim = mesh.dim
        self.degree = int(degree)
        self.elem = get_refelem(self.dim, self.degree)
        self.nd = self.elem.nd
        if plan is None:
            from .compat import _tile_cells
            plan = build_rank_plan(mesh, np.zeros(mesh.num_cells(), dtype=np.int32), 0, 1,
                                   tile=_tile_cells(mesh.dim, degree))
        self.plan = plan
        self._h = C.c_void_p()
        nbr = np.ascontiguousarray(plan.nbr, dtype=np.int32)
        code = np.ascontiguousarray(plan.code, dtype=np.uint8)
        jinv = np.ascontiguousarray(plan.jinv, dtype=np.float64)
        desc = capi.MeshDesc(dim=self.dim, degree=self.degree, n_owned=plan.n_owned, n_total=plan.n_total,
                             nbr=nbr.ctypes.data, code=code.ctypes.data, jinv=jinv.ctypes.data,
                             device=int(device), n_boundary=int(plan.n_boundary),
                             geom_classes=int(bool(geom_classes)), symmetric_stress=int(bool(symmetric)))
        self.symmetric = bool(symmetric)
        check(lib.sg_create(C.byref(self._h), C.byref(desc)))
        self.n_owned = plan.n_owned
        self.n_total = plan.n_total
        if plan.send_cells is not None and len(plan.send_cells):
            sc = np.ascontiguousarray(plan.send_cells, dtype=np.int64)
            check(lib.sg_set_halo_plan(self._h, len(sc), ptr(sc)))
        self._g2l = None

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.sg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- numbering ----------------------------------------------------------------------------
    @property
    def local_to_global(self):
        return self.plan.local_to_global

    def _to_local(self, a, ncomp_shape):
        """Global-order field (E*nd, ...) -> this rank's local order (n_total*nd, ...)."""
        a = np.asarray(a, dtype=np.float64).reshape((self.mesh.num_cells(), self.nd) + ncomp_shape)
        return np.ascontiguousarray(a[self.plan.local_to_global])

    def _from_local(self, loc, out, ncomp_shape, owned_only=True):
        loc = loc.reshape((self.n_total, self.nd) + ncomp_shape)
        out = out.reshape((self.mesh.num_cells(), self.nd) + ncomp_shape)
        n = self.n_owned if owned_only else self.n_total
        out[self.plan.local_to_global[:n]] = loc[:n]

    # -- parameters ---------------------------------------------------------------------------
    def set_material(self, density, lam, mu):
        lam_a = np.asarray(lam, dtype=np.float64)
        mu_a = np.asarray(mu, dtype=np.float64)
        if lam_a.ndim == 0 and mu_a.ndim == 0:
            check(lib.sg_set_material(self._h, float(density), float(lam_a), float(mu_a), None, None))
            return
        E = self.mesh.num_cells()
        lam_c = np.broadcast_to(lam_a, (E,))[self.plan.local_to_global[:self.n_owned]]
        mu_c = np.broadcast_to(mu_a, (E,))[self.plan.local_to_global[:self.n_owned]]
        lam_c = np.ascontiguousarray(lam_c, dtype=np.float64)
        mu_c = np.ascontiguousarray(mu_c, dtype=np.float64)
        check(lib.sg_set_material(self._h, float(density), 0.0, 0.0, ptr(lam_c), ptr(mu_c)))

    def set_absorption(self, sigma, sigma_degree):
        """``sigma``: nodal values (E, nd_sigma) of the absorption field in global cell order, or None."""
        if sigma is None:
            check(lib.sg_set_absorption(self._h, 0, None, None))
            return
        sigma = np.asarray(sigma, dtype=np.float64).reshape(self.mesh.num_cells(), -1)
        loc = sigma[self.plan.local_to_global[:self.n_owned]]
        cells = np.flatnonzero(np.any(loc != 0.0, axis=1)).astype(np.int64)
        if len(cells) == 0:
            check(lib.sg_set_absorption(self._h, 0, None, None))
            return
        W = self.elem.absorption_tensor(int(sigma_degree))              # (nd, nd_sigma, nd)
        mats = np.ascontiguousarray(np.einsum("abc,eb->eac", W, loc[cells]))
        check(lib.sg_set_absorption(self._h, len(cells), ptr(cells), ptr(mats)))

    def set_source(self, sdof_global, amp):
        """``sdof_global``: flat indices into the global stress array (E*nd*d*d); ``amp``: (nsteps, nsrc)."""
        if sdof_global is None or len(sdof_global) == 0:
            check(lib.sg_set_source(self._h, 0, None, 0, None))
            return
        sdof_global = np.asarray(sdof_global, dtype=np.int64)
        amp = np.asarray(amp, dtype=np.float64)
        per_cell = self.nd * self.dim * self.dim
        gcell = sdof_global // per_cell
        if self._g2l is None:
            g2l = np.full(self.mesh.num_cells(), -1, dtype=np.int64)
            g2l[self.plan.local_to_global] = np.arange(self.n_total)
            self._g2l = g2l
        lcell = self._g2l[gcell]
        keep = (lcell >= 0) & (lcell < self.n_owned)
        if not keep.any():
            check(lib.sg_set_source(self._h, 0, None, 0, None))
            return
        ldof = np.ascontiguousarray(lcell[keep] * per_cell + sdof_global[keep] % per_cell)
        a = np.ascontiguousarray(amp[:, keep])
        check(lib.sg_set_source(self._h, len(ldof), ptr(ldof), a.shape[0], ptr(a)))

    # -- state ----------------------------------------------------------------------------------
    def set_state(self, u=None, s=None):
        d = self.dim
        ul = self._to_local(u, (d,)) if u is not None else None
        sl = self._to_local(s, (d, d)) if s is not None else None
        check(lib.sg_set_state(self._h, ptr(ul), ptr(sl)))

    def get_state(self, u_out=None, s_out=None):
        """Owned cells of (u, s) written into global-order arrays (allocated if not given)."""
        d, E = self.dim, self.mesh.num_cells()
        ul = np.empty((self.n_total * self.nd, d))
        sl = np.empty((self.n_total * self.nd, d, d))
        check(lib.sg_get_state(self._h, ptr(ul), ptr(sl)))
        if u_out is None:
            u_out = np.zeros((E * self.nd, d))
        if s_out is None:
            s_out = np.zeros((E * self.nd, d, d))
        self._from_local(ul, u_out, (d,))
        self._from_local(sl, s_out, (d, d))
        return u_out, s_out

    def get_field(self, which, owned_only=True):
        d, E = self.dim, self.mesh.num_cells()
        shape = (d,) if which in (capi.FIELD_U, capi.FIELD_UH) else (d, d)
        loc = np.empty((self.n_total * self.nd,) + shape)
        check(lib.sg_get_field(self._h, which, ptr(loc)))
        out = np.zeros((E * self.nd,) + shape)
        self._from_local(loc, out, shape, owned_only=owned_only)
        return out

    # -- stepping ---------------------------------------------------------------------------------
    def step(self, nsteps, dt, first_step=0):
        check(lib.sg_step(self._h, int(nsteps), float(dt), int(first_step)))

    def stage(self, stage, dt, step=0, part=capi.PART_ALL):
        check(lib.sg_stage(self._h, int(stage), int(part), float(dt), int(step)))

    def time_stage(self, stage, dt, reps=10, part=capi.PART_ALL):
        ms = C.c_double()
        check(lib.sg_time_stage(self._h, int(stage), int(part), float(dt), int(reps), C.byref(ms)))
        return ms.value

    def synchronize(self):
        check(lib.sg_synchronize(self._h))

    def last_step_ms(self):
        ms = C.c_double()
        check(lib.sg_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    @property
    def handle(self):
        return self._h
