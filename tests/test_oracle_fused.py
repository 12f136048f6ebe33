"""The fused six-pass CPU variant (oracle/elastic_fused_c.c, the second CPU baseline of BASELINE.md section 3) against the
literal oracle (quadrature assembly of seigen/elastic.py:204-219 + block inverse mass): <= 1e-12 per step with sponge,
source and per-cell materials, 2D and 3D."""
import numpy as np
import pytest

from oracle.c_fused import CFused
from oracle.elastic_oracle import ElasticOracle
from tests.util import nodal_from_mesh, random_state, rel_err, small_mesh


@pytest.mark.parametrize("dim,p,q", [(2, 1, 1), (2, 2, 4), (2, 3, 3), (3, 1, 1), (3, 2, 1)])
def test_fused_c_step_matches_literal_oracle(dim, p, q):
    mesh = small_mesh(dim)
    E = mesh.num_cells()
    rng = np.random.default_rng(2)
    orc = ElasticOracle(mesh.coords, mesh.cells, p, sigma_degree=q)
    lam, mu = rng.uniform(0.4, 0.6, E), rng.uniform(0.2, 0.3, E)
    orc.l, orc.mu, orc.density, orc.dt = lam, mu, 1.0, 1e-3
    sig = rng.uniform(0, 3, size=(E, orc.sel.nd))
    sig[rng.uniform(size=E) < 0.5] = 0.0
    orc.sigma = sig
    u, s = random_state(mesh, p)
    src = np.zeros_like(s)
    src[rng.choice(E, 3, replace=False), 0] = rng.standard_normal((3, dim, dim))
    orc.source = lambda t: src
    op, el = nodal_from_mesh(mesh, p)
    cf = CFused(el.Dr, el.Lift, el.fnodes, el.ftab, op.nbr, op.code, op.jinv, lam, mu, 1.0,
                sigma_mats=CFused.sponge_matrices(orc))
    uo, so = u, s
    uf, sf = u.copy(), s.copy()
    for _ in range(3):
        uo, so, _ = orc.step(uo, so, 0.0)
        cf.step_inplace(uf, sf, src, orc.dt)
    assert rel_err(uf, uo) < 1e-12 and rel_err(sf, so) < 1e-12
