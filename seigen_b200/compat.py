"""The sliver of the Firedrake API the ElasticLF4 scripts touch, backed by NumPy host arrays.

``tests/eigenmode/eigenmode_2d.py``, ``tests/explosive_source/explosive_source_lf4.py`` and
``tests/pulse/pulse_1d_lf4.py`` build function spaces, interpolate ``Expression``s into ``Function``s, assign
them to ``elastic.u0`` / ``elastic.s0`` and read ``u1.dat.data`` back.  That is all this module provides: there
is no UFL, no assembly and no solver here -- the time stepping lives behind the C ABI.

As in Firedrake, ``Function.dat.data`` holds the cells *this rank owns*, in the rank's own cell order (Firedrake
renumbers cells through DMPlex, so scripts never rely on the order); ``FunctionSpace.cell_order`` gives the
global cell id of every local cell for code that needs it (the parity tests).
"""
from __future__ import annotations

import time
from contextlib import contextmanager

import numpy as np

from .expression import Expression
from .layout import build_rank_plan, partition_cells
from .mesh import Mesh
from .refelem import get_refelem

__all__ = ["FunctionSpace", "VectorFunctionSpace", "TensorFunctionSpace", "Function", "File", "timed_region",
           "get_timers", "reset_timers", "errornorm_l2", "norm", "mesh_plan"]

_timers: dict = {}


@contextmanager
def timed_region(name):
    """``pyop2.profiling.timed_region`` stand-in (seigen/elastic.py:4)."""
    t0 = time.perf_counter()
    try:
        yield
    finally:
        _timers[name] = _timers.get(name, 0.0) + time.perf_counter() - t0


def get_timers(reset=False):
    out = dict(_timers)
    if reset:
        _timers.clear()
    return out


def reset_timers():
    _timers.clear()


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # pragma: no cover
        pass
    return 0, 1


def _tile_cells(dim, degree):
    """Cells per tile of the stage kernels for this element (``sg_tile_cells``), or None: the cell order inside tiles is
    an optimisation (layout._order_within_tiles), so a missing library or SG_TILE_ORDER=0 only switches it off."""
    import os
    if os.environ.get("SG_TILE_ORDER", "1") == "0":
        return None
    try:
        from .capi import lib
        tile = int(lib.sg_tile_cells(int(dim), int(degree)))
    except Exception:
        return None
    if tile <= 0:
        return None
    # measured (profiles/r02_tune_order_within_tiles.log): +0.5 % (2D P2) ... +3 % (3D P1) where one thread owns a whole
    # cell, -1.5 % for the 3-D elements that run one tensor row per thread in small tiles (3D P2 / P3: TILE 64 / 32) --
    # there the cells with out-of-tile neighbours end up in the same warps, which then trail the others
    if dim == 3 and tile < 128:
        return None
    return tile


def mesh_plan(mesh: Mesh, degree=None):
    """The rank plan of ``mesh`` for the current process group (created once per mesh, by its first function space:
    every space on the mesh shares the cell order, which is tuned to the tile size of that first space's degree --
    ElasticLF4 creates its own spaces before anything else)."""
    plan = getattr(mesh, "_plan", None)
    if plan is None:
        rank, size = _dist()
        part = getattr(mesh, "_partition", None)
        if part is None:
            import os
            method = os.environ.get("SG_PARTITION") or getattr(mesh, "partition_method", "rcb")
            part = partition_cells(mesh, size, method)
            mesh._partition = part
        plan = build_rank_plan(mesh, part, rank, size, tile=_tile_cells(mesh.dim, degree) if degree else None)
        mesh._plan = plan
    return plan


class FunctionSpace:
    """Scalar DG space; ``value_shape`` () / (d,) / (d, d) for the vector and tensor variants."""

    def __init__(self, mesh: Mesh, family, degree, name=None, shape=()):
        family = {"Discontinuous Lagrange": "DG", "DP": "DG"}.get(family, family)
        if family != "DG":
            # ExplicitElasticLF4 inverts the mass matrix cell by cell (elastic.py:376-381): only valid for DG
            raise NotImplementedError("only discontinuous Lagrange ('DG') spaces are supported")
        self.mesh_ = mesh
        self.family = family
        self.degree = int(degree)
        self.name = name
        self.shape = tuple(shape)
        self.elem = get_refelem(mesh.dim, self.degree)
        self.plan = mesh_plan(mesh, self.degree)

    def mesh(self):
        return self.mesh_

    def ufl_element(self):
        return self

    @property
    def cell_order(self):
        """Global cell id of each local (owned) cell."""
        return self.plan.local_to_global[:self.plan.n_owned]

    @property
    def node_count(self):
        return self.plan.n_owned * self.elem.nd

    @property
    def dof_count(self):
        return self.node_count * int(np.prod(self.shape, dtype=np.int64))

    @property
    def value_size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    def node_coords(self):
        """(n_owned*nd, d) physical coordinates of the owned nodes, local order."""
        m = self.mesh_
        v = m.coords[m.cells[self.cell_order]]                                  # (n_owned, d+1, d)
        lam = self.elem.lattice.astype(np.float64) / self.elem.degree          # (nd, d+1) barycentric
        return np.einsum("av,evk->eak", lam, v).reshape(-1, m.dim)


def VectorFunctionSpace(mesh, family, degree, name=None, dim=None):
    return FunctionSpace(mesh, family, degree, name=name, shape=(dim or mesh.dim,))


def TensorFunctionSpace(mesh, family, degree, name=None, shape=None):
    return FunctionSpace(mesh, family, degree, name=name, shape=shape or (mesh.dim, mesh.dim))


class _Dat:
    """``Function.dat``: ``data`` is the NumPy array of the owned nodes.  ``defer_copy_from(other)`` makes this Dat
    a lazy copy of another one (the ``u0.assign(u1)`` at the end of a time step, elastic.py:296): the bytes are
    copied when ``data`` is next looked at, not before."""

    def __init__(self, data):
        self._data = data
        self._lazy_from = None

    @property
    def data(self):
        src = self._lazy_from
        if src is not None:
            self._lazy_from = None
            np.copyto(self._data, src.data)
        return self._data

    @data.setter
    def data(self, value):
        self._lazy_from = None
        self._data = value

    @property
    def data_ro(self):
        return self.data

    def defer_copy_from(self, other):
        self._lazy_from = other if other is not self else None

    def current_source(self):
        """The Dat whose array currently holds this Dat's values (itself unless a deferred copy is pending)."""
        return self._lazy_from if self._lazy_from is not None else self


class Function:
    def __init__(self, function_space, val=None, name=None, alloc=np.zeros):
        if isinstance(function_space, Function):
            val = function_space.dat.data
            function_space = function_space.function_space()
        self._fs = function_space
        self._name = name
        data = alloc((function_space.node_count,) + function_space.shape)
        if val is not None:
            data[...] = np.asarray(val).reshape(data.shape)
        self.dat = _Dat(data)

    def function_space(self):
        return self._fs

    def name(self):
        return self._name

    def assign(self, other):
        if isinstance(other, Function):
            if other.dat.data.shape != self.dat.data.shape:
                raise ValueError("assign: function spaces differ")
            self.dat.data[...] = other.dat.data
        else:
            self.dat.data[...] = float(other)
        return self

    def interpolate(self, expression):
        if isinstance(expression, Function):
            return self.assign(expression)
        if not isinstance(expression, Expression):
            expression = Expression(expression)
        eshape, fshape = tuple(expression.value_shape()), tuple(self._fs.shape)
        if eshape != fshape and not (eshape == () and all(n == 1 for n in fshape)):
            raise ValueError(f"interpolate: expression shape {expression.value_shape()} != space shape {self._fs.shape}")
        # (a scalar expression fills the one-component vector / tensor spaces of a 1-D problem:
        #  tests/pulse/pulse_1d_lf4.py:27-30)
        self.dat.data[...] = np.asarray(expression.evaluate(self._fs.node_coords())).reshape(self.dat.data.shape)
        return self

    def copy(self, deepcopy=True):
        return Function(self._fs, val=self.dat.data.copy(), name=self._name)

    # pointwise expressions for the post-processing forms of the reference's scripts (forms.py): u1 - uexact, abs(..)
    def __sub__(self, other):
        from .forms import as_field
        return as_field(self) - other

    def __add__(self, other):
        from .forms import as_field
        return as_field(self) + other

    def __neg__(self):
        from .forms import as_field
        return -as_field(self)

    def __abs__(self):
        from .forms import as_field
        return abs(as_field(self))


class File:
    """``File("velocity.pvd")`` (elastic.py:123-124).  Every ``write`` adds one snapshot ``<base>_<n>.vtu`` holding
    all nd nodes of every owned cell (``vtkout.write_vtu``) and rewrites the ``.pvd`` collection, so the series can
    be opened while the run is still going, as with Firedrake's File.  With several ranks each writes its own piece
    ``<base>_<n>_<rank>.vtu`` and rank 0 adds the ``<base>_<n>.pvtu`` that ties them together (the reference's
    parallel output) -- no two ranks ever write the same path."""

    def __init__(self, name):
        self.name = name
        self.count = 0
        self.entries = []

    def write(self, f, time=None):
        from .vtkout import write_pvd, write_pvtu, write_vtu
        base = self.name.rsplit(".", 1)[0]
        rank, size = _dist()
        t = float(self.count if time is None else time)
        if size == 1:
            snap = f"{base}_{self.count}.vtu"
            write_vtu(snap, f)
        else:
            pieces = [f"{base}_{self.count}_{r}.vtu" for r in range(size)]
            name, nc = write_vtu(pieces[rank], f)
            snap = f"{base}_{self.count}.pvtu"
            if rank == 0:
                write_pvtu(snap, pieces, name, nc)
        self.entries.append((t, snap))
        if rank == 0:
            write_pvd(base + ".pvd", self.entries)
        self.count += 1

    def __lshift__(self, f):
        self.write(f)
        return self


def _quadrature(dim, degree):
    """Collapsed Gauss rule on the unit simplex (host-side error norms only)."""
    n = (degree + dim) // 2 + 1
    x, w = np.polynomial.legendre.leggauss(n)
    g, w = 0.5 * (x + 1.0), 0.5 * w
    if dim == 1:
        return g[:, None], w
    if dim == 2:
        U, V = np.meshgrid(g, g, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        return np.stack([U.ravel(), (V * (1 - U)).ravel()], 1), (WU * WV * (1 - U)).ravel()
    U, V, W = np.meshgrid(g, g, g, indexing="ij")
    WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
    pts = np.stack([U.ravel(), (V * (1 - U)).ravel(), (W * (1 - U) * (1 - V)).ravel()], 1)
    return pts, (WU * WV * WW * (1 - U) ** 2 * (1 - V)).ravel()


def errornorm_l2(f: Function, exact: Expression, degree_rise=3):
    """||f - exact||_L2 over the owned cells (sum over ranks is the caller's job), exact evaluated at quadrature
    points.  Plays the role of the DG6 projection norm of tests/eigenmode/eigenmode_2d.py:49-63."""
    fs = f.function_space()
    mesh, el = fs.mesh(), fs.elem
    d = mesh.dim
    xq, wq = _quadrature(d, 2 * (el.degree + degree_rise))
    phi = el.tabulate(xq)                                                   # (nq, nd)
    cells = mesh.cells[fs.cell_order]
    v = mesh.coords[cells]                                                  # (E, d+1, d)
    lam = np.concatenate([1.0 - xq.sum(1, keepdims=True), xq], axis=1)      # (nq, d+1)
    xphys = np.einsum("qv,evk->eqk", lam, v)
    detj = np.abs(mesh.topology.detj[fs.cell_order])
    vals = f.dat.data.reshape((len(cells), el.nd, -1))
    fq = np.einsum("qb,ebc->eqc", phi, vals)
    ex = exact.evaluate(xphys.reshape(-1, d)).reshape(fq.shape)
    diff = fq - ex
    return float(np.sqrt(np.einsum("e,q,eqc,eqc->", detj, wq, diff, diff)))


def norm(f: Function):
    """L2 norm of a Function over the owned cells."""
    fs = f.function_space()
    el = fs.elem
    detj = np.abs(fs.mesh().topology.detj[fs.cell_order])
    vals = f.dat.data.reshape((fs.plan.n_owned, el.nd, -1))
    return float(np.sqrt(np.einsum("e,ab,eac,ebc->", detj, el.M, vals, vals)))
