"""The reference's scenarios through the public API on the GPU (mirrors of tests/eigenmode/eigenmode_2d.py,
eigenmode_3d.py and tests/explosive_source/explosive_source_lf4.py written against ``from seigen_b200 import *``),
checked against the CPU oracle on the same mesh, degree, dt and source: relative L2 error <= 1e-10 per field
(BASELINE.json's tolerance; also tests/tiling/explosive_source.py:659-660), the eigenmode errors / rates, and the
REF-C1 sensor trace."""
import numpy as np
import pytest

from tests.scenarios import (EXPL_LAM, EXPL_MU, LAM, MU, eigenmode_dt, eigenmode_expressions, explosive_dt,
                             explosive_expressions, explosive_oracle, locate, rates)
from tests.util import rel_err

pytestmark = pytest.mark.gpu


class EigenmodeLF4:
    """tests/eigenmode/eigenmode_2d.py:7-65 / eigenmode_3d.py:7-69 with the firedrake import swapped."""

    def __init__(self, dim, N, degree, dt, solver='explicit', output=False):
        from seigen_b200 import ElasticLF4, UnitCubeMesh, UnitSquareMesh
        self.dim = dim
        self.mesh = UnitSquareMesh(N, N) if dim == 2 else UnitCubeMesh(N, N, N)
        self.elastic = ElasticLF4.create(self.mesh, "DG", degree, dimension=dim, solver=solver, output=output)
        self.elastic.density = 1.0
        self.elastic.dt = dt
        self.elastic.mu = MU
        self.elastic.l = LAM

    def run(self, T=5.0):
        from seigen_b200 import Function
        el = self.elastic
        uic, sic = eigenmode_expressions(self.dim, el.dt, 0, el.dt / 2.0)
        el.u0.assign(Function(el.U).interpolate(uic))
        el.s0.assign(Function(el.S).interpolate(sic))
        return el.run(T)

    def error(self, u1, s1):
        from seigen_b200 import errornorm_l2
        uex, sex = eigenmode_expressions(self.dim, self.elastic.dt, 5, 5 + self.elastic.dt / 2.0)
        return errornorm_l2(u1, uex), errornorm_l2(s1, sex)

    def error_as_reference(self, u1, s1):
        """The reference's own functional (eigenmode_2d.py:40-65, eigenmode_3d.py:42-69): exact solution interpolated
        into U / S, |difference| L2-projected into DG6 (2D) / DG3 (3D), norm of the projection."""
        from seigen_b200 import (Function, TensorFunctionSpace, TestFunction, TrialFunction, VectorFunctionSpace, dx,
                                 inner, lhs, norm, rhs, solve)
        el = self.elastic
        uex, sex = eigenmode_expressions(self.dim, el.dt, 5, 5 + el.dt / 2.0)
        uexact = Function(el.U).interpolate(uex)
        sexact = Function(el.S).interpolate(sex)
        q = 6 if self.dim == 2 else 3
        out = []
        for space, f, exact in ((VectorFunctionSpace(self.mesh, "DG", q), u1, uexact),
                                (TensorFunctionSpace(self.mesh, "DG", q), s1, sexact)):
            temp = Function(space)
            temp_test, temp_trial = TestFunction(space), TrialFunction(space)
            G = inner(temp_test, temp_trial)*dx - inner(temp_test, abs(f - exact))*dx
            solve(lhs(G) == rhs(G), temp)
            out.append(norm(temp))
        return tuple(out)


MIN_RATES = {(2, 1): (1.5, 0.9), (2, 2): (2.8, 2.2), (2, 3): (3.7, 2.7), (3, 1): (0.8, 1.0), (3, 2): (3.0, 2.2)}


def oracle_eigenmode(dim, N, p):
    from tests.test_oracle_eigenmode import run_eigenmode
    return run_eigenmode(dim, N, p, fields=True)


@pytest.mark.parametrize("dim,p,Ns", [(2, 1, (4, 8, 16)), (2, 2, (4, 8)), (2, 3, (4, 8)), (2, 4, (4,)),
                                      (3, 1, (2, 4)), (3, 2, (2, 4)), (3, 3, (2,))])
def test_eigenmode_errors_and_rates(dim, p, Ns):
    errs = []
    for N in Ns:
        em = EigenmodeLF4(dim, N, p, eigenmode_dt(N, p))
        u1, s1 = em.run()
        eu, es = em.error(u1, s1)
        if (dim, p) in ((2, 1), (2, 2), (2, 3), (3, 1)) or N == 2:
            n, eu_o, es_o, u_o, s_o = oracle_eigenmode(dim, N, p)
            assert em.elastic.steps_done == n
            order = em.elastic.S.cell_order
            assert rel_err(u1.dat.data.reshape(u_o.shape), u_o[order]) < 1e-10
            assert rel_err(s1.dat.data.reshape(s_o.shape), s_o[order]) < 1e-10
            # (the two error norms use different quadrature rules for the non-polynomial exact solution)
            assert eu == pytest.approx(eu_o, rel=2e-3) and es == pytest.approx(es_o, rel=2e-3)
        if N <= 8:
            # the reference's DG6 / DG3 projection norm measures || u1 - I_p u_exact || (interpolated exact solution):
            # same order of magnitude as the true L2 error, never far above it
            ru_, rs_ = em.error_as_reference(u1, s1)
            assert 0.3 * eu < ru_ < 1.5 * eu and 0.3 * es < rs_ < 1.5 * es, (eu, ru_, es, rs_)
        errs.append((eu, es))
    if len(Ns) > 1:
        hs = [1.0 / N for N in Ns]
        ru, rs = rates([e[0] for e in errs], hs)[-1], rates([e[1] for e in errs], hs)[-1]
        # central flux: u ~ p+1, s ~ p asymptotically; N = 2 -> 4 in 3D is
        # pre-asymptotic (Appendix C: P1 u 4.75e-1 -> 2.55e-1)
        ru_min, rs_min = MIN_RATES[(dim, p)]
        assert ru > ru_min and rs > rs_min, (errs, ru, rs)


def _explosive_gpu(Lx, Ly, h, T, receivers):
    from seigen_b200 import ElasticLF4, Function, FunctionSpace, RectangleMesh
    mesh = RectangleMesh(int(Lx / h), int(Ly / h), Lx, Ly)
    el = ElasticLF4.create(mesh, "DG", 2, dimension=2, solver="explicit", output=False)
    el.density, el.mu, el.l = 1.0, EXPL_MU, EXPL_LAM
    el.dt = explosive_dt(h)
    source, sponge = explosive_expressions(Lx, Ly)
    el.source_expression = source
    el.source_function = Function(el.S)
    el.source = el.source_expression
    el.absorption_function = Function(FunctionSpace(mesh, "DG", 4))
    el.absorption = sponge
    el.receivers = receivers
    u1, s1 = el.run(T)
    return el, u1, s1


def test_explosive_source_1000_steps_parity_and_ref_c1():
    from oracle.c_oracle import COracle
    from oracle.elastic_oracle import step_times
    from tests.test_oracle_refc import compare_with_ref
    Lx, Ly, h = 100.0, 50.0, 2.5
    dt = explosive_dt(h)
    T = 1000.5 * dt
    el, u1, s1 = _explosive_gpu(Lx, Ly, h, T, [(45.3, Ly - 1.0), (45.0, Ly - 1.0)])
    assert el.steps_done == 1000

    mesh, orc, src = explosive_oracle(Lx, Ly, h)
    co = COracle(orc)
    u = np.zeros((orc.E, orc.nd, 2))
    s = np.zeros((orc.E, orc.nd, 2, 2))
    e, xi = locate(mesh.coords, mesh.cells, (45.3, Ly - 1.0))      # strictly inside a cell: DG values are unique there
    phi = orc.el.tab(xi[None])[0]
    times = step_times(T, dt)
    trace = []
    for t in times:
        co.step_inplace(u, s, src(t), dt)
        trace.append(phi @ u[e])
    trace = np.array(trace)
    order = el.S.cell_order
    assert rel_err(u1.dat.data.reshape(orc.E, orc.nd, 2), u[order]) < 1e-10
    assert rel_err(s1.dat.data.reshape(orc.E, orc.nd, 2, 2), s[order]) < 1e-10
    # device-side receivers == oracle trace, and both track the reference's external solution
    rec = el.receiver_data
    assert rec.shape == (1000, 2, 2) and np.isfinite(rec).all()
    assert rel_err(rec[:, 0], trace) < 1e-9
    w = np.array(times) <= 0.45
    rel, peak_ratio, dt_peak = compare_with_ref(np.array(times)[w], -rec[w, 1, 1])     # sensor C1 = (45, 149)
    assert rel < 0.25 and 0.8 < peak_ratio < 1.2 and abs(dt_peak) < 0.01


def test_run_twice_continues_and_output_mode(tmp_path, monkeypatch):
    """run(T) twice == run(2T) without source (u0 <- u1, s0 <- s1 after each run, elastic.py:296, 304); output=True
    writes one snapshot per step plus the initial one (elastic.py:273, 310)."""
    monkeypatch.chdir(tmp_path)
    a = EigenmodeLF4(2, 4, 2, eigenmode_dt(4, 2))
    ua, sa = a.run(T=1.0)
    ua = ua.dat.data.copy()
    b = EigenmodeLF4(2, 4, 2, eigenmode_dt(4, 2), output=True)
    b.run(T=0.5)
    n1 = b.elastic.steps_done
    ub, sb = b.elastic.run(0.5)
    assert n1 + b.elastic.steps_done == a.elastic.steps_done
    assert rel_err(ub.dat.data, ua) < 1e-12
    assert np.array_equal(b.elastic.u0.dat.data, ub.dat.data)
    assert len(list(tmp_path.glob("velocity_*.vtu"))) == 2 * (n1 + 1)
    assert (tmp_path / "velocity.pvd").exists() and (tmp_path / "stress.pvd").exists()
    # the last snapshot holds every node of the final field
    from seigen_b200.vtkout import read_vtu_arrays
    last = read_vtu_arrays(tmp_path / ("velocity_%d.vtu" % (2 * (n1 + 1) - 1)))
    assert np.array_equal(last["VelocityNew"][:, :2], ub.dat.data)


def test_output_every_k_steps(tmp_path, monkeypatch):
    """``output_every = k``: snapshots after every k-th step and after the last one; the result does not change."""
    monkeypatch.chdir(tmp_path)
    a = EigenmodeLF4(2, 4, 1, eigenmode_dt(4, 1))
    ua, _ = a.run(T=1.0)
    b = EigenmodeLF4(2, 4, 1, eigenmode_dt(4, 1), output=True)
    b.elastic.output_every = 3
    ub, _ = b.run(T=1.0)
    n = b.elastic.steps_done
    assert n == a.elastic.steps_done and np.array_equal(ub.dat.data, ua.dat.data)
    assert len(list(tmp_path.glob("velocity_*.vtu"))) == 1 + -(-n // 3)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_pulse_3d_parity(p):
    """BASELINE.json configs[2]: 3D Gaussian pulse on a tetrahedral box, DG P1-P3, DG1 sponge, through the public API,
    against the C restatement of the oracle on the same mesh (box [0,4]x[0,1]x[0,1] at a quarter of the headline
    resolution so that the oracle finishes in seconds; the full-size case runs in test_gpu_properties.py)."""
    from oracle.c_oracle import COracle
    from oracle.elastic_oracle import ElasticOracle
    from seigen_b200 import BoxMesh, ElasticLF4, Function, FunctionSpace
    from tests.scenarios import PULSE_LAM, PULSE_MU, pulse_dt, pulse_expressions
    nx = 16
    mesh = BoxMesh(nx, nx // 4, nx // 4, 4.0, 1.0, 1.0)
    u_e, s_e, sponge = pulse_expressions()
    el = ElasticLF4.create(mesh, "DG", p, dimension=3, solver="explicit", output=False)
    el.density, el.mu, el.l = 1.0, PULSE_MU, PULSE_LAM
    el.dt = pulse_dt(4.0 / nx, p)
    el.absorption_function = Function(FunctionSpace(mesh, "DG", 1))
    el.absorption = sponge
    el.u0.interpolate(u_e)
    el.s0.interpolate(s_e)
    nsteps = {1: 200, 2: 100, 3: 60}[p]
    u0 = el.u0.dat.data.copy()
    s0 = el.s0.dat.data.copy()
    u1, s1 = el.run((nsteps + 0.5) * el.dt)
    assert el.steps_done == nsteps and el._dev.symmetric           # a symmetric stress: packed storage is kept

    order = el.S.cell_order
    orc = ElasticOracle(mesh.coords, mesh.cells[order], p, sigma_degree=1)
    orc.l, orc.mu, orc.density, orc.dt = PULSE_LAM, PULSE_MU, 1.0, el.dt
    orc.sigma = el.absorption_function.dat.data.reshape(orc.E, -1)
    co = COracle(orc)
    u = u0.reshape(orc.E, orc.nd, 3).copy()
    s = s0.reshape(orc.E, orc.nd, 3, 3).copy()
    for _ in range(nsteps):
        co.step_inplace(u, s, None, el.dt)
    assert np.abs(u).max() < 10.0                                  # stable
    assert rel_err(u1.dat.data.reshape(u.shape), u) < 1e-10
    assert rel_err(s1.dat.data.reshape(s.shape), s) < 1e-10


def test_pulse_1d_through_the_public_api():
    """tests/pulse/pulse_1d_lf4.py:7-33 (IntervalMesh(400, 4.0), DG P1, dimension=1, DG1 sponge sigma = 100, scalar
    Gaussian ICs, dt = 0.0025, T = 2 => 800 steps) through ElasticLF4.run on the device, against the literal oracle."""
    from oracle.elastic_oracle import ElasticOracle
    from seigen_b200 import ElasticLF4, Expression, Function, FunctionSpace, IntervalMesh
    mesh = IntervalMesh(int(4.0 / 1e-2), 4.0)
    el = ElasticLF4.create(mesh, "DG", 1, dimension=1, output=False)
    el.density, el.dt, el.mu, el.l = 1.0, 0.0025, 0.25, 0.5
    el.absorption_function = Function(FunctionSpace(el.mesh, "DG", 1))
    el.absorption = Expression("x[0] >= 3.5 || x[0] <= 0.5 ? 100.0 : 0")
    el.u0.assign(Function(el.U).interpolate(Expression('exp(-50*pow((x[0]-1), 2))')))
    el.s0.assign(Function(el.S).interpolate(Expression('-exp(-50*pow((x[0]-1), 2))')))
    u0, s0 = el.u0.dat.data.copy(), el.s0.dat.data.copy()
    u1, s1 = el.run(2.0)
    assert el.steps_done == 800

    order = el.S.cell_order
    orc = ElasticOracle(mesh.coords, mesh.cells[order], 1, sigma_degree=1)
    orc.l, orc.mu, orc.density, orc.dt = 0.5, 0.25, 1.0, 0.0025
    orc.sigma = el.absorption_function.dat.data.reshape(orc.E, -1)
    u, s = u0.reshape(orc.E, orc.nd, 1), s0.reshape(orc.E, orc.nd, 1, 1)
    for _ in range(800):
        u, s, _ = orc.step(u, s, 0.0)
    assert rel_err(u1.dat.data.reshape(u.shape), u) < 1e-10
    assert rel_err(s1.dat.data.reshape(s.shape), s) < 1e-10
    x = el.U.node_coords()[:, 0]
    assert abs(x[np.argmax(u1.dat.data[:, 0])] - 3.0) <= 0.02     # the right-going pulse has moved from x = 1 to x = 3
