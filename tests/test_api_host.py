"""Host-side mirror of the reference interface (no GPU): Expression parser, spaces, time-loop bookkeeping,
constructor error behaviour, and that the C-ABI library exports every symbol include/seigen_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from seigen_b200 import (ElasticLF4, Expression, Function, FunctionSpace, RectangleMesh, TensorFunctionSpace,
                         UnitSquareMesh, VectorFunctionSpace, Vp, Vs, cfl_dt, step_times)
from seigen_b200 import capi
from seigen_b200.expression import ExpressionSyntaxError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_expression_eigenmode_ic():
    # tests/eigenmode/eigenmode_2d.py:30-31
    a = 1.234
    e = Expression(('a*cos(pi*x[0])*sin(pi*x[1])*cos(a*t)', '-a*sin(pi*x[0])*cos(pi*x[1])*cos(a*t)'), a=a, t=0)
    x = np.random.default_rng(0).uniform(size=(7, 2))
    v = e.evaluate(x)
    assert v.shape == (7, 2)
    assert np.allclose(v[:, 0], a * np.cos(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]))
    e.t = 0.5
    assert np.allclose(e.evaluate(x)[:, 1], -a * np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 1]) * np.cos(a * 0.5))


def test_expression_ternary_logic_tensor():
    # tests/explosive_source/explosive_source_lf4.py:36-45
    src = "x[0] >= 44.5 && x[0] <= 45.5 && x[1] >= 148.5 && x[1] <= 149.5 ? (-1.0 + 2*a*pow(t - 0.3, 2))*exp(-a*pow(t - 0.3, 2)) : 0.0"
    e = Expression(((src, "0.0"), ("0.0", src)), a=159.42, t=0.25)
    x = np.array([[45.0, 149.0], [10.0, 149.0], [45.0, 10.0]])
    v = e.evaluate(x)
    assert v.shape == (3, 2, 2)
    r = (-1.0 + 2 * 159.42 * (0.25 - 0.3) ** 2) * np.exp(-159.42 * (0.25 - 0.3) ** 2)
    assert np.allclose(v[0], [[r, 0], [0, r]]) and np.all(v[1:] == 0)
    sponge = Expression("x[0] <= 20 || x[0] >= 280 || x[1] <= 20.0 ? 1000 : 0")
    assert list(sponge.evaluate(np.array([[10., 100.], [150., 100.], [150., 5.], [290., 100.]]))) == [1000, 0, 1000, 1000]


def test_expression_precedence_and_errors():
    e = Expression("-x[0]*2 + 3 < 1 ? 1 : (x[0] > 5 ? 2 : 3)")
    assert list(e.evaluate(np.array([[2.0], [0.5], [0.0]]))) == [1.0, 3.0, 3.0]
    assert Expression("1.5e1 - 2/4").evaluate(np.zeros((1, 1)))[0] == 14.5
    assert Expression("!(x[0] > 0)").evaluate(np.array([[1.0], [-1.0]])).tolist() == [0.0, 1.0]
    with pytest.raises(ExpressionSyntaxError):
        Expression("foo(x[0])")
    with pytest.raises(ExpressionSyntaxError):
        Expression("x[0] +* 2")
    with pytest.raises(ExpressionSyntaxError):
        Expression("undefined_name * 2")


def test_spaces_and_interpolation():
    mesh = UnitSquareMesh(4, 4)
    U = VectorFunctionSpace(mesh, "DG", 2)
    S = TensorFunctionSpace(mesh, "DG", 2)
    F = FunctionSpace(mesh, "DG", 4)
    assert U.dof_count == 32 * 6 * 2 and S.dof_count == 32 * 6 * 4 and F.dof_count == 32 * 15
    f = Function(U).interpolate(Expression(("x[0]", "2*x[1]")))
    x = U.node_coords()
    assert np.allclose(f.dat.data[:, 0], x[:, 0]) and np.allclose(f.dat.data[:, 1], 2 * x[:, 1])
    g = Function(U)
    g.assign(f)
    assert np.array_equal(g.dat.data, f.dat.data)
    with pytest.raises(ValueError):
        Function(S).interpolate(Expression(("1", "2")))
    with pytest.raises(NotImplementedError):
        FunctionSpace(mesh, "CG", 1)
    # cell_order is a permutation of the cells
    assert sorted(U.cell_order.tolist()) == list(range(32))


def test_step_times_match_reference_loop():
    # T=5, dt=0.125*4/N gives exactly 40*N/4 steps
    for N in (4, 8, 16, 32):
        assert len(step_times(5.0, 0.5 * (1.0 / N))) == 40 * N // 4
    assert step_times(0.35, 0.1) == pytest.approx([0.1, 0.2, 0.30000000000000004])


def test_helpers():
    assert Vp(0.25, 0.5, 1.0) == 1.0 and Vs(0.25, 1.0) == 0.5
    assert cfl_dt(2.5, 100.0, 0.05) == pytest.approx(0.00125)


def test_create_error_behaviour():
    mesh = RectangleMesh(2, 2, 1.0, 1.0)
    with pytest.raises(ValueError, match="Unknown solver mode"):
        ElasticLF4.create(mesh, "DG", 1, dimension=2, solver="bogus", output=False)
    with pytest.raises(NotImplementedError):
        ElasticLF4.create(mesh, "DG", 1, dimension=2, solver="implicit", output=False)
    for mode in ("explicit", "parloop", "fusion", "tiling"):
        el = ElasticLF4.create(mesh, "DG", 1, dimension=2, solver=mode, output=False)
        assert el.u0.dat.data.shape == (8 * 3, 2) and el.s0.dat.data.shape == (8 * 3, 2, 2)
    with pytest.raises(ValueError):
        el.run(1.0)          # density / dt / mu / l not set


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "seigen_b200.h")).read()
    declared = set(re.findall(r"\b(sg_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(capi.exported_symbols())
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert capi.load().sg_version() >= 1
    assert capi.load().sg_nodes_per_cell(2, 2) == 6 and capi.load().sg_nodes_per_cell(3, 3) == 20
    assert capi.load().sg_nodes_per_cell(2, 7) < 0


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 3), (3, 2), (3, 3)])
def test_vtu_output_carries_every_node(tmp_path, dim, p):
    """File.write (seigen/elastic.py:120-124, 221-232): the snapshot holds all nd nodes of every cell, cut into p^d
    linear sub-cells that tile the cell exactly; a .pvd collection indexes the series."""
    from seigen_b200 import BoxMesh, File, Function, RectangleMesh, VectorFunctionSpace
    from seigen_b200.vtkout import read_vtu_arrays, subcells
    mesh = RectangleMesh(3, 2, 1.5, 1.0) if dim == 2 else BoxMesh(2, 1, 2, 1.0, 0.5, 1.0)
    V = VectorFunctionSpace(mesh, "DG", p)
    f = Function(V, name="VelocityNew")
    rng = np.random.default_rng(0)
    f.dat.data[...] = rng.standard_normal(f.dat.data.shape)
    out = File(str(tmp_path / "velocity.pvd"))
    out.write(f, time=0.25)
    out.write(f, time=0.5)
    a = read_vtu_arrays(tmp_path / "velocity_1.vtu")
    E, nd = mesh.num_cells(), V.elem.nd
    assert a["points"].shape == (E * nd, 3) and np.allclose(a["points"][:, :dim], V.node_coords())
    assert np.array_equal(a["VelocityNew"][:, :dim], f.dat.data) and not a["VelocityNew"][:, dim:].any()
    conn = a["connectivity"].reshape(-1, dim + 1)
    assert len(conn) == E * p ** dim and set(a["types"]) == {5 if dim == 2 else 10}
    # the sub-cells tile every cell: their volumes add up to the cell's
    sub = subcells(V.elem)
    x = V.elem.nodes[sub]                                         # (nsub, dim+1, dim) reference coordinates
    vol = np.abs(np.linalg.det(x[:, 1:] - x[:, :1])).sum()
    assert vol == pytest.approx(1.0)                               # |det| of the reference cell = 1 (x d!)
    assert sorted(np.unique(sub)) == list(range(nd))               # every node is used
    pvd = open(tmp_path / "velocity.pvd").read()
    assert 'timestep="0.25"' in pvd and 'file="velocity_1.vtu"' in pvd
