"""The C/OpenMP restatement (timed CPU baseline) reproduces the literal NumPy oracle."""
import numpy as np
import pytest

from oracle.c_oracle import COracle
from oracle.elastic_oracle import ElasticOracle
from tests.util import random_state, rel_err, small_mesh


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2)])
def test_c_oracle_matches_numpy_oracle(dim, p):
    mesh = small_mesh(dim)
    q = 1 if dim == 3 else p
    orc = ElasticOracle(mesh.coords, mesh.cells, p, sigma_degree=q)
    rng = np.random.default_rng(0)
    E = mesh.num_cells()
    orc.l, orc.mu = rng.uniform(0.4, 0.6, E), rng.uniform(0.2, 0.3, E)
    orc.sigma = rng.uniform(0, 2, (E, orc.sel.nd))
    orc.dt = 1e-2
    co = COracle(orc)
    u, s = random_state(mesh, p)
    src = rng.standard_normal(s.shape)
    assert rel_err(co.solve_f(s, u), orc.solve_f(s, u)) < 1e-13
    assert rel_err(co.solve_g(u, src), orc.solve_g(u, src)) < 1e-13
    orc.source = lambda t: src
    u1, s1, _ = orc.step(u, s, 0.0)
    uc, sc = u.copy(), s.copy()
    co.step_inplace(uc, sc, src, orc.dt)
    assert rel_err(uc, u1) < 1e-13
    assert rel_err(sc, s1) < 1e-13
    assert co.threads >= 1
