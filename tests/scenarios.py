"""The reference's own test scenarios, stated once for the oracle (CPU) and the B200 path (test infrastructure).

* eigenmode 2D / 3D: tests/eigenmode/eigenmode_2d.py:9-65, eigenmode_3d.py:9-69 (analytic standing wave);
* explosive source: tests/explosive_source/explosive_source_lf4.py:9-56 on a sub-domain, with the stable time step
  (Courant 0.05) and the sensor of tests/explosive_source/uy.py:36.
"""
import math

import numpy as np

from seigen_b200 import Expression

MU, LAM, RHO = 0.25, 0.5, 1.0
VS = math.sqrt(MU / RHO)


def eigenmode_expressions(dim, dt, t_u, t_s):
    """(u, s) Expressions of the standing wave at times t_u / t_s (eigenmode_2d.py:30-35, eigenmode_3d.py:30-38)."""
    if dim == 2:
        a, b = math.sqrt(2) * math.pi * VS, 2 * math.pi * MU
        u = Expression(('a*cos(pi*x[0])*sin(pi*x[1])*cos(a*t)', '-a*sin(pi*x[0])*cos(pi*x[1])*cos(a*t)'), a=a, t=t_u)
        s = Expression((('-b*sin(pi*x[0])*sin(pi*x[1])*sin(a*t)', '0'),
                        ('0', 'b*sin(pi*x[0])*sin(pi*x[1])*sin(a*t)')), a=a, b=b, t=t_s)
        return u, s
    A = math.sqrt(2 * RHO * MU)                               # eigenmode_3d.py:25-26
    O = math.pi * math.sqrt(2 * MU / RHO)
    u = Expression(('cos(pi*x[0])*(sin(pi*x[1]) - sin(pi*x[2]))*cos(O*t)',
                    'cos(pi*x[1])*(sin(pi*x[2]) - sin(pi*x[0]))*cos(O*t)',
                    'cos(pi*x[2])*(sin(pi*x[0]) - sin(pi*x[1]))*cos(O*t)'), O=O, t=t_u)
    s = Expression((('-A*sin(pi*x[0])*(sin(pi*x[1]) - sin(pi*x[2]))*sin(O*t)', '0', '0'),
                    ('0', '-A*sin(pi*x[1])*(sin(pi*x[2]) - sin(pi*x[0]))*sin(O*t)', '0'),
                    ('0', '0', '-A*sin(pi*x[2])*(sin(pi*x[0]) - sin(pi*x[1]))*sin(O*t)')), A=A, O=O, t=t_s)
    return u, s


def eigenmode_dt(N, p):
    return 0.5 * (1.0 / N) / (2.0 ** (p - 1))                 # eigenmode_2d.py:75


def rates(err, hs):
    return [math.log(err[i] / err[i + 1]) / math.log(hs[i] / hs[i + 1]) for i in range(len(err) - 1)]


# ---- explosive source (sub-domain Lx x Ly whose top edge is the free surface at y = Ly) ------------------------
EXPL_MU, EXPL_LAM = 3600.0, 3599.3664
EXPL_A = 159.42


def explosive_expressions(Lx, Ly):
    """Source box centred 1 m below the surface at x = 45 (explosive_source_lf4.py:36-38 shifted from Ly = 150) and the
    sponge of :43-45 with the right strip moved to the sub-domain's edge."""
    box = f"x[0] >= 44.5 && x[0] <= 45.5 && x[1] >= {Ly - 1.5} && x[1] <= {Ly - 0.5}"
    ric = "(-1.0 + 2*a*pow(t - 0.3, 2))*exp(-a*pow(t - 0.3, 2))"
    src = f"{box} ? {ric} : 0.0"
    source = Expression(((src, "0.0"), ("0.0", src)), a=EXPL_A, t=0)
    sponge = Expression(f"x[0] <= 20 || x[0] >= {Lx - 20} || x[1] <= 20.0 ? 1000 : 0")
    return source, sponge


def explosive_dt(h, courant=0.05):
    vp = math.sqrt((EXPL_LAM + 2 * EXPL_MU) / 1.0)
    return courant * h / vp


def locate(coords, cells, point):
    """(cell, reference coordinates) of the simplex containing `point` (first hit; brute force, small meshes)."""
    v = coords[cells]                                        # (E, d+1, d)
    J = np.swapaxes(v[:, 1:] - v[:, :1], 1, 2)
    xi = np.linalg.solve(J, (np.asarray(point, dtype=float)[None] - v[:, 0])[..., None])[..., 0]
    ok = (xi >= -1e-12).all(axis=1) & (xi.sum(axis=1) <= 1 + 1e-12)
    e = int(np.flatnonzero(ok)[0])
    return e, xi[e]


def explosive_oracle(Lx, Ly, h, courant=0.05, degree=2, sigma_degree=4, nx=None, ny=None):
    """(oracle, source(t) callable, dt) for the explosive-source scenario on a Lx x Ly sub-domain (test infrastructure)."""
    from oracle.elastic_oracle import ElasticOracle
    from seigen_b200.mesh import RectangleMesh
    mesh = RectangleMesh(nx or int(Lx / h), ny or int(Ly / h), Lx, Ly)
    orc = ElasticOracle(mesh.coords, mesh.cells, degree, sigma_degree=sigma_degree, lite=True)
    orc.l, orc.mu, orc.density = EXPL_LAM, EXPL_MU, 1.0
    orc.dt = explosive_dt(h, courant)
    source, sponge = explosive_expressions(Lx, Ly)
    orc.sigma = sponge.evaluate(orc.sigma_node_coords().reshape(-1, 2)).reshape(orc.E, -1)
    xs = orc.node_coords().reshape(-1, 2)
    active = np.flatnonzero(np.any(source.evaluate(xs, t=0.3).reshape(len(xs), -1) != 0, axis=1))

    def src(t):
        out = np.zeros((len(xs), 2, 2))
        out[active] = source.evaluate(xs[active], t=t)
        return out.reshape(orc.E, orc.nd, 2, 2)
    return mesh, orc, src


# ---- 3D Gaussian pulse on a tetrahedral box (BASELINE.json configs[2]) --------------------
# 3-D lift of tests/pulse/pulse_1d_lf4.py: rho = 1, mu = 0.25, lambda = 0.5 (:14-17), u = (G, 0, 0), s = -G I with
# G = exp(-50 (x-1)^2) (:27-30), DG1 sponge sigma = 100 near both ends in x (:22-24).
PULSE_MU, PULSE_LAM = 0.25, 0.5


def pulse_expressions():
    g = "exp(-50*pow(x[0] - 1.0, 2))"
    u0 = Expression((g, "0.0", "0.0"))
    s0 = Expression((("-" + g, "0.0", "0.0"), ("0.0", "-" + g, "0.0"), ("0.0", "0.0", "-" + g)))
    sponge = Expression("x[0] >= 3.5 || x[0] <= 0.5 ? 100.0 : 0")
    return u0, s0, sponge


def pulse_dt(h, p):
    """min(0.16 h / (2^(p-1) Vp), 1/sigma_max) with Vp = 1, sigma_max = 100.  SURVEY.md 8d config 3 proposed Courant 0.5
    "to be confirmed in the oracle before freezing": at the full resolution h = 1/16 the explicit sponge (sigma dt = 1)
    on top of the wave operator is unstable at P2 for dt > 0.00625 and at P3 for dt > 0.005 (growth 1e40-1e94 in 300
    steps in the CPU oracle), so the Courant number is 0.16: dt = 0.01 / 0.005 / 0.0025 for P1 / P2 / P3 at h = 1/16."""
    return min(0.16 * h / (2 ** (p - 1) * 1.0), 1.0 / 100.0)
