"""BASELINE.json configs[1..3] at their stated sizes, defined once for the golden generator
(scripts/make_golden_fullsize.py: CPU oracle, run in the build container) and the GPU tests
(tests/test_gpu_fullsize.py: CUDA path through ElasticLF4.run).  Test infrastructure.

Each case gives the mesh builder arguments, degree, dt, number of steps and the way the inputs are produced; inputs
are expressions / seeded random fields, evaluated by both sides with the same code on the same node coordinates.
"""
import math

import numpy as np

from tests.scenarios import (EXPL_LAM, EXPL_MU, PULSE_LAM, PULSE_MU, explosive_dt, explosive_expressions, pulse_dt,
                             pulse_expressions)

NSAMPLE = 20000


def sample_indices(n, seed):
    return np.sort(np.random.default_rng(seed).choice(n, size=min(NSAMPLE, n), replace=False))


CASES = {
    # configs[1]: explosive source, 2D P2, 1 016 064 DoF, 1000 steps
    "explosive_168x84_p2": dict(kind="explosive", nx=168, ny=84, Lx=300.0, Ly=150.0, degree=2, steps=1000),
    # configs[3] down-scaled as BASELINE.md section 3 states: Marmousi grid at its native h = 24 m, 3.34 M DoF, 200 steps
    "marmousi_383x121_p2": dict(kind="marmousi", nx=383, ny=121, Lx=9192.0, Ly=2904.0, degree=2, steps=200),
    # configs[2]: 3D Gaussian pulse on the (64,16,16) x 6 tetrahedra box, DG P1-P3
    "pulse3d_64x16x16_p1": dict(kind="pulse", n=(64, 16, 16), L=(4.0, 1.0, 1.0), degree=1, steps=200),
    "pulse3d_64x16x16_p2": dict(kind="pulse", n=(64, 16, 16), L=(4.0, 1.0, 1.0), degree=2, steps=200),
    "pulse3d_64x16x16_p3": dict(kind="pulse", n=(64, 16, 16), L=(4.0, 1.0, 1.0), degree=3, steps=60),
}


def build_mesh(case):
    from seigen_b200 import BoxMesh, RectangleMesh
    if case["kind"] == "pulse":
        return BoxMesh(*case["n"], *case["L"])
    return RectangleMesh(case["nx"], case["ny"], case["Lx"], case["Ly"])


def case_dt(case):
    if case["kind"] == "explosive":
        return explosive_dt(case["Lx"] / case["nx"])
    if case["kind"] == "marmousi":
        return 0.5 * (case["Lx"] / case["nx"]) / (2 ** (case["degree"] - 1) * 5500.0)
    return pulse_dt(case["L"][0] / case["n"][0], case["degree"])


MARMOUSI_RICKER = ("x[0] >= 4584.0 && x[0] <= 4608.0 && x[1] >= 2868.0 && x[1] <= 2892.0 ? "
                   "(-1.0 + 2*a*pow(t - t0, 2))*exp(-a*pow(t - t0, 2)) : 0.0")


def case_expressions(case):
    """dict of Expressions: u0, s0 (or None = zero), sponge (+ its DG degree), source."""
    from seigen_b200 import Expression
    if case["kind"] == "explosive":
        source, sponge = explosive_expressions(case["Lx"], case["Ly"])
        return dict(u0=None, s0=None, sponge=sponge, sponge_degree=4, source=source, lam=EXPL_LAM, mu=EXPL_MU)
    if case["kind"] == "pulse":
        u0, s0, sponge = pulse_expressions()
        return dict(u0=u0, s0=s0, sponge=sponge, sponge_degree=1, source=None, lam=PULSE_LAM, mu=PULSE_MU)
    a = (math.pi * 10.0) ** 2
    src = MARMOUSI_RICKER
    source = Expression(((src, "0.0"), ("0.0", src)), a=a, t0=0.012, t=0.0)
    return dict(u0=None, s0=None, sponge=None, sponge_degree=None, source=source, lam="marmousi", mu="marmousi")
