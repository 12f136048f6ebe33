"""Parity at the sizes BASELINE.json states (VERDICT r01 "row p"): the CUDA path, through ``ElasticLF4.run``, against
golden digests of the CPU oracle at full size (tests/golden/fullsize_*.npz, made by scripts/make_golden_fullsize.py
from the same case definitions, tests/fullsize_cases.py) -- relative L2 error <= 1e-10 per field after the stated
number of steps (BASELINE.json's tolerance; tests/tiling/explosive_source.py:659-660) -- and, for the explosive source
on its shipped mesh run to T = 2.5 s, against the oracle's sensor traces and the reference's own REF-C1, REF-C2 and
REF-C3 (tests/explosive_source/uy.py:36-43)."""
import os

import numpy as np
import pytest

from tests.fullsize_cases import CASES, build_mesh, case_dt, case_expressions
from tests.scenarios import EXPL_LAM, EXPL_MU, explosive_dt, explosive_expressions

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_case_gpu(name):
    from seigen_b200 import ElasticLF4, Function, FunctionSpace
    case = CASES[name]
    mesh = build_mesh(case)
    ex = case_expressions(case)
    d = mesh.dim
    el = ElasticLF4.create(mesh, "DG", case["degree"], dimension=d, solver="explicit", output=False)
    el.density, el.dt = 1.0, case_dt(case)
    if ex["lam"] == "marmousi":
        from seigen_b200.marmousi import marmousi_lame
        lam, mu = marmousi_lame(mesh)
        order = el.S.cell_order
        el.l, el.mu = lam[order], mu[order]
    else:
        el.l, el.mu = ex["lam"], ex["mu"]
    if ex["sponge"] is not None:
        el.absorption_function = Function(FunctionSpace(mesh, "DG", ex["sponge_degree"]))
        el.absorption = ex["sponge"]
    if ex["source"] is not None:
        el.source_expression = ex["source"]
        el.source_function = Function(el.S)
        el.source = el.source_expression
    if ex["u0"] is not None:
        el.u0.interpolate(ex["u0"])
    if ex["s0"] is not None:
        el.s0.interpolate(ex["s0"])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)        # the source-support probe is sub-sampled at this size
        u1, s1 = el.run((case["steps"] + 0.5) * el.dt)
    assert el.steps_done == case["steps"]
    return el, u1, s1


def to_global(el, f):
    """Owned-cell data of a Function scattered into the global [cell][node][comp] order (single rank: all cells)."""
    order = el.S.cell_order
    loc = f.dat.data.reshape(len(order), -1)
    out = np.empty_like(loc)
    out[order] = loc
    return out.reshape(-1)


@pytest.mark.parametrize("name", sorted(CASES))
def test_full_size_parity_with_oracle_digest(name):
    path = os.path.join(GOLDEN, f"fullsize_{name}.npz")
    assert os.path.exists(path), f"{path} missing: run scripts/make_golden_fullsize.py {name}"
    z = np.load(path)
    el, u1, s1 = run_case_gpu(name)
    assert el.mesh.num_cells() == int(z["cells"]) and el.dt == pytest.approx(float(z["dt"]), rel=1e-15)
    for f, idx, val, nrm, cn in ((u1, z["iu"], z["u"], float(z["u_norm"]), z["u_comp_norm"]),
                                 (s1, z["is_"], z["s"], float(z["s_norm"]), z["s_comp_norm"])):
        g = to_global(el, f)
        assert np.isfinite(g).all()
        # 20 000 sampled entries: relative L2 over the sample <= 1e-10 (the sample's norm is that of the field up to
        # the sampling ratio), and no entry off by more than 1e-10 of the largest sampled value
        assert np.linalg.norm(g[idx] - val) <= 1e-10 * np.linalg.norm(val)
        assert np.abs(g[idx] - val).max() <= 1e-10 * np.abs(val).max()
        # norms of the whole field and of every component
        assert abs(np.linalg.norm(g) - nrm) <= 1e-10 * nrm
        ncomp = len(np.ravel(cn))
        comp = np.sqrt((g.reshape(-1, ncomp) ** 2).sum(axis=0))
        assert np.allclose(comp, np.ravel(cn), rtol=1e-9, atol=1e-10 * nrm)


def correlate(a, b):
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def test_explosive_source_shipped_mesh_to_T25_traces_and_ref_c123():
    """tests/explosive_source/explosive_source_lf4.py as shipped (120 x 60 mesh, h = 2.5, 518 400 DoF) with the stable
    time step, to T = 2.5 (2078 steps).  Device-side receivers vs (i) the oracle's traces at the same interior points
    (tests/golden/oracle_refc_traces.npz) and (ii) REF-C1..3 at the sensors of uy.py:36-43.

    What REF-C pins, measured with the CPU oracle (scripts/make_golden_traces.py; same numbers on the GPU): C1 (above
    the source) agrees to 15 % relative L2 with amplitude ratio 0.97 (mean of the two cells that share the sensor's
    mesh line; 14 % / 1.04 and 20 % / 0.90 taken singly); at C2 / C3 the Rayleigh-wave train has the right
    arrival time and waveform (correlation 0.97-0.99) but 2.2 x the amplitude of the external solution -- the surface
    wave excited by a source 1 m below the surface is not resolved by nodal interpolation of a 1 m box on an h = 2.5 m
    mesh.  The reference itself only compares by eye (uy.py:46-80); its plot windows for C2 / C3 are +-8e-6."""
    from seigen_b200 import ElasticLF4, Function, FunctionSpace, RectangleMesh
    Lx, Ly, h = 300.0, 150.0, 2.5
    zt = np.load(os.path.join(GOLDEN, "oracle_refc_traces.npz"))
    interior = [tuple(p) for p in zt["sensors"]]
    # a DG field is double-valued on the mesh lines x = 45, 90, 140 the reference's sensors sit on (its VTK probe
    # picks one side): sample 1 cm to either side and average
    sensors = [(x + dx, 149.0) for x in (45.0, 90.0, 140.0) for dx in (-0.01, 0.01)]
    mesh = RectangleMesh(120, 60, Lx, Ly)
    el = ElasticLF4.create(mesh, "DG", 2, dimension=2, solver="explicit", output=False)
    el.density, el.mu, el.l = 1.0, EXPL_MU, EXPL_LAM
    el.dt = explosive_dt(h)
    source, sponge = explosive_expressions(Lx, Ly)
    el.source_expression = source
    el.source_function = Function(el.S)
    el.source = el.source_expression
    el.absorption_function = Function(FunctionSpace(mesh, "DG", 4))
    el.absorption = sponge
    el.receivers = interior + sensors
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        el.run(2.5)
    rec = el.receiver_data
    assert el.steps_done == len(zt["t"]) == 2078 and np.isfinite(rec).all()
    # (i) the oracle's traces, both velocity components, all three interior points
    for k in range(3):
        assert np.linalg.norm(rec[:, k] - zt["u"][:, k]) <= 1e-9 * np.linalg.norm(zt["u"][:, k])
    # (ii) the reference's external solution
    zr = np.load(os.path.join(GOLDEN, "ref_c123.npz"))
    tt = zt["t"]
    windows = [(0.1, 1.0), (0.5, 1.5), (1.0, 2.5)]                      # the time axes of uy.py:52, 65, 78
    stats = []
    for k, (lo, hi) in enumerate(windows):
        ref = np.interp(tt, zr["t"], zr["uy"][k])
        w = (tt >= lo) & (tt <= hi)
        sim = -0.5 * (rec[w, 3 + 2 * k, 1] + rec[w, 4 + 2 * k, 1])
        rel = np.linalg.norm(sim - ref[w]) / np.linalg.norm(ref[w])
        stats.append((rel, correlate(sim, ref[w]), (sim @ ref[w]) / (ref[w] @ ref[w])))
    # measured with the CPU oracle at the same points: C1 (0.154, 0.988, 0.971), C2 (1.30, 0.985, 2.24), C3 (1.29, 0.971, 2.17)
    rel1, corr1, amp1 = stats[0]
    assert rel1 < 0.18 and corr1 > 0.98 and 0.92 < amp1 < 1.02, stats
    for rel, corr, amp in stats[1:]:
        assert corr > 0.95 and 1.9 < amp < 2.5, stats
