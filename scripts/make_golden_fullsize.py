"""Golden digests of the CPU oracle at BASELINE.json's stated sizes (configs[1], [2] and the down-scaled [3]).

    OMP_NUM_THREADS=6 python scripts/make_golden_fullsize.py [case ...]      # ~45 min on 8 cores for all five

For every case of tests/fullsize_cases.py the C/OpenMP oracle (oracle/elastic_c.c) advances the stated number of steps
from the stated inputs; what is stored (tests/golden/fullsize_<case>.npz) is a digest the GPU tests can check without
re-running the oracle: 20 000 randomly chosen entries of the final velocity and of the final stress (flat indices into
the global ``[cell][node][comp]`` arrays), the L2 norms of both fields and of each component.  20 000 entries agreeing
to 1e-10 of the field's scale plus equal norms leave no room for a field that differs.  Test infrastructure.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.c_oracle import COracle  # noqa: E402
from oracle.elastic_oracle import ElasticOracle  # noqa: E402
from tests.fullsize_cases import CASES, build_mesh, case_dt, case_expressions, sample_indices  # noqa: E402


def run_case(name):
    case = CASES[name]
    t0 = time.time()
    mesh = build_mesh(case)
    ex = case_expressions(case)
    d, p = mesh.dim, case["degree"]
    orc = ElasticOracle(mesh.coords, mesh.cells, p, sigma_degree=ex["sponge_degree"], lite=True)
    if ex["lam"] == "marmousi":
        from seigen_b200.marmousi import marmousi_lame
        orc.l, orc.mu = marmousi_lame(mesh)
    else:
        orc.l, orc.mu = ex["lam"], ex["mu"]
    orc.density, orc.dt = 1.0, case_dt(case)
    xs = orc.node_coords().reshape(-1, d)
    if ex["sponge"] is not None:
        orc.sigma = ex["sponge"].evaluate(orc.sigma_node_coords().reshape(-1, d)).reshape(orc.E, -1)
    co = COracle(orc)
    E, nd = orc.E, orc.nd
    u = (ex["u0"].evaluate(xs) if ex["u0"] is not None else np.zeros((len(xs), d))).reshape(E, nd, d).copy()
    s = (ex["s0"].evaluate(xs) if ex["s0"] is not None else np.zeros((len(xs), d, d))).reshape(E, nd, d, d).copy()
    src = None
    if ex["source"] is not None:
        source = ex["source"]
        active = np.zeros(len(xs), dtype=bool)
        for n in range(case["steps"]):
            v = source.evaluate(xs, t=(n + 1) * orc.dt) if n % 10 == 0 else None
            if v is not None:
                active |= np.any(v.reshape(len(xs), -1) != 0, axis=1)
        active = np.flatnonzero(active)
        assert len(active), "source selects no node"

        def src(t):
            out = np.zeros((len(xs), d, d))
            out[active] = source.evaluate(xs[active], t=t)
            return out.reshape(E, nd, d, d)
    print(f"{name}: {E} cells, {E * nd * (d + d * d)} DoF, dt {orc.dt:.6g}, setup {time.time() - t0:.1f} s", flush=True)
    t0 = time.time()
    t = 0.0
    for n in range(case["steps"]):
        t += orc.dt                                   # the accumulated t of elastic.py:279-280, 313
        co.step_inplace(u, s, src(t) if src is not None else None, orc.dt)
    wall = time.time() - t0
    assert np.isfinite(u).all() and np.isfinite(s).all()
    iu, is_ = sample_indices(u.size, 11), sample_indices(s.size, 12)
    out = os.path.join(ROOT, "tests", "golden", f"fullsize_{name}.npz")
    np.savez_compressed(out, iu=iu, u=u.reshape(-1)[iu], is_=is_, s=s.reshape(-1)[is_],
                        u_norm=np.linalg.norm(u), s_norm=np.linalg.norm(s),
                        u_comp_norm=np.sqrt((u ** 2).sum(axis=(0, 1))), s_comp_norm=np.sqrt((s ** 2).sum(axis=(0, 1))),
                        u_absmax=np.abs(u).max(), s_absmax=np.abs(s).max(), steps=case["steps"], dt=orc.dt,
                        cells=E, threads=co.threads, wall_s=wall)
    print(f"{name}: {case['steps']} steps in {wall:.1f} s ({E * nd * (d + d * d) * case['steps'] / wall / 1e6:.2f} M upd/s), "
          f"|u| {np.linalg.norm(u):.6e} |s| {np.linalg.norm(s):.6e} max|u| {np.abs(u).max():.3e} -> {out}", flush=True)


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name)
