"""Oracle self-check: literal quadrature assembly of the UFL forms (seigen/elastic.py:204-219) followed by the
block inverse mass (elastic.py:358-367) == the quadrature-free nodal operator the CUDA kernels implement."""
import numpy as np
import pytest

from oracle.elastic_oracle import ElasticOracle
from oracle.nodal import NodalOperator
from tests.util import nodal_from_mesh, random_state, rel_err, small_mesh

CASES = [(2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3)]


@pytest.mark.parametrize("dim,p", CASES)
def test_literal_forms_equal_nodal_operator(dim, p):
    mesh = small_mesh(dim)
    orc = ElasticOracle(mesh.coords, mesh.cells, p)
    orc.l, orc.mu = 0.7, 0.3
    op, el = nodal_from_mesh(mesh, p)
    u, s = random_state(mesh, p)
    assert rel_err(op.Dv(s), orc.solve_f(s, u)) < 2e-12
    assert rel_err(op.Ds(u, 0.7, 0.3), orc.solve_g(u, None)) < 2e-12


@pytest.mark.parametrize("dim,p,q", [(2, 2, 4), (2, 1, 1), (3, 1, 1), (2, 2, 2)])
def test_absorption_projection(dim, p, q):
    mesh = small_mesh(dim)
    orc = ElasticOracle(mesh.coords, mesh.cells, p, sigma_degree=q)
    rng = np.random.default_rng(3)
    orc.sigma = rng.uniform(0, 5, size=(mesh.num_cells(), orc.sel.nd))
    op, el = nodal_from_mesh(mesh, p)
    u, s = random_state(mesh, p)
    W = el.absorption_tensor(q)
    expect = orc.solve_f(s, u)
    got = op.Dv(s) - NodalOperator.absorb(W, orc.sigma, u)
    assert rel_err(got, expect) < 2e-12


def test_per_cell_material():
    mesh = small_mesh(2)
    orc = ElasticOracle(mesh.coords, mesh.cells, 2)
    rng = np.random.default_rng(5)
    lam = rng.uniform(1, 2, mesh.num_cells())
    mu = rng.uniform(1, 2, mesh.num_cells())
    orc.l, orc.mu = lam, mu
    op, _ = nodal_from_mesh(mesh, 2)
    u, s = random_state(mesh, 2)
    assert rel_err(op.Ds(u, lam, mu), orc.solve_g(u, None)) < 2e-12
