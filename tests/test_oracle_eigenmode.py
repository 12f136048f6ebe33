"""Pins the CPU oracle to the reference's analytic standing-wave solution (tests/eigenmode/eigenmode_2d.py:30-46,
eigenmode_3d.py:30-50): true L2 errors at T = 5 with the reference's dt rule, and the convergence rates the central
flux gives (u ~ p+1, s ~ p; the reference records no expected values).  The expected error
values below were obtained with an independent NumPy restatement during the survey."""
import numpy as np
import pytest

from oracle.c_oracle import COracle
from oracle.elastic_oracle import ElasticOracle, step_count
from seigen_b200.mesh import UnitCubeMesh, UnitSquareMesh
from tests.scenarios import LAM, MU, eigenmode_dt, eigenmode_expressions, rates


def run_eigenmode(dim, N, p, T=5.0, fields=False):
    mesh = UnitSquareMesh(N, N) if dim == 2 else UnitCubeMesh(N, N, N)
    orc = ElasticOracle(mesh.coords, mesh.cells, p)
    orc.l, orc.mu, orc.density = LAM, MU, 1.0
    orc.dt = dt = eigenmode_dt(N, p)
    x = orc.node_coords().reshape(-1, dim)
    uic, sic = eigenmode_expressions(dim, dt, 0.0, dt / 2.0)
    u0 = uic.evaluate(x).reshape(orc.E, orc.nd, dim)
    s0 = sic.evaluate(x).reshape(orc.E, orc.nd, dim, dim)
    n = step_count(T, dt)
    u, s = COracle(orc).run(u0, s0, n, dt)
    uex, sex = eigenmode_expressions(dim, dt, 5.0, 5.0 + dt / 2.0)       # t = 5 hard-coded, eigenmode_2d.py:42, 46
    eu = orc.l2_error(u, lambda xq: uex.evaluate(xq.reshape(-1, dim)).reshape(xq.shape[:2] + (dim,)))
    es = orc.l2_error(s, lambda xq: sex.evaluate(xq.reshape(-1, dim)).reshape(xq.shape[:2] + (dim, dim)))
    return (n, eu, es, u, s) if fields else (n, eu, es)


EXPECTED_2D = {   # (p, N): (steps, u_err, s_err)   SURVEY.md Appendix C
    (1, 4): (40, 3.19e-1, 3.84e-1), (1, 8): (80, 1.37e-1, 2.08e-1), (1, 16): (160, 4.43e-2, 1.01e-1),
    (2, 4): (80, 1.66e-2, 4.32e-2), (2, 8): (160, 1.69e-3, 8.41e-3),
    (3, 4): (160, 8.83e-4, 5.06e-3),
}


@pytest.mark.parametrize("p,N", sorted(EXPECTED_2D))
def test_eigenmode_2d_errors(p, N):
    steps, eu0, es0 = EXPECTED_2D[(p, N)]
    n, eu, es = run_eigenmode(2, N, p)
    assert n == steps
    # the survey's restatement used its own triangulation: coarse P1 meshes differ by a few per cent, the resolved
    # cases agree to < 1 %
    tol = 0.12 if p == 1 and N <= 8 else 0.03
    assert eu == pytest.approx(eu0, rel=tol) and es == pytest.approx(es0, rel=tol)


def test_eigenmode_2d_rates():
    for p, (ru, rs) in {1: (1.5, 0.9), 2: (2.8, 2.2)}.items():
        Ns = [4, 8, 16] if p == 1 else [4, 8]
        res = [run_eigenmode(2, N, p) for N in Ns]
        hs = [1.0 / N for N in Ns]
        r_u, r_s = rates([r[1] for r in res], hs), rates([r[2] for r in res], hs)
        assert r_u[-1] > ru and r_s[-1] > rs, (p, r_u, r_s)


def test_eigenmode_3d_errors():
    # P1 N=2,4: u 4.75e-1/2.55e-1, s 4.29e-1/1.71e-1; P2 N=2: u 1.01e-1, s 1.19e-1
    n, eu, es = run_eigenmode(3, 2, 1)
    assert eu == pytest.approx(4.75e-1, rel=0.05) and es == pytest.approx(4.29e-1, rel=0.05)
    n, eu, es = run_eigenmode(3, 4, 1)
    assert eu == pytest.approx(2.55e-1, rel=0.05) and es == pytest.approx(1.71e-1, rel=0.10)
    n, eu, es = run_eigenmode(3, 2, 2)
    assert eu == pytest.approx(1.01e-1, rel=0.05) and es == pytest.approx(1.19e-1, rel=0.05)


def test_stress_stays_symmetric():
    mesh = UnitSquareMesh(4, 4)
    orc = ElasticOracle(mesh.coords, mesh.cells, 2)
    orc.l, orc.mu, orc.density, orc.dt = LAM, MU, 1.0, eigenmode_dt(4, 2)
    x = orc.node_coords().reshape(-1, 2)
    uic, sic = eigenmode_expressions(2, orc.dt, 0.0, orc.dt / 2)
    u, s = orc.run(uic.evaluate(x).reshape(orc.E, orc.nd, 2), sic.evaluate(x).reshape(orc.E, orc.nd, 2, 2), 5 * orc.dt)
    assert np.abs(s - np.swapaxes(s, 2, 3)).max() < 1e-14
