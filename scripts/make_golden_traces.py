"""Golden sensor traces of the CPU oracle for the FULL explosive-source domain (BASELINE.json configs[1]).

    python scripts/make_golden_traces.py          # ~3 min on 8 cores; writes tests/golden/oracle_refc_traces.npz

The scenario of tests/explosive_source/explosive_source_lf4.py on its shipped mesh (generate_mesh defaults, :9-10:
RectangleMesh(120, 60, 300, 150), h = 2.5, DG P2, DG4 sponge) with the stable time step (Courant 0.05, SURVEY.md
Appendix B-7), run by the C/OpenMP oracle to T = 2.5; -u_y is recorded after every step at the three sensors of
tests/explosive_source/uy.py:36-43 shifted by +0.3 m in x so that they lie strictly inside a cell (DG fields are
double-valued on the mesh lines x = 45, 90, 140).  tests/test_gpu_fullsize.py compares the device-side receivers of
the CUDA path with these traces without running the oracle, and the oracle itself with REF-C1..3
(tests/test_oracle_refc.py uses the stored traces too).  Test infrastructure, not product code.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.c_oracle import COracle  # noqa: E402
from oracle.elastic_oracle import step_times  # noqa: E402
from tests.scenarios import explosive_oracle, locate  # noqa: E402

NX, NY, LX, LY, T = 120, 60, 300.0, 150.0, 2.5
SENSORS = [(45.3, 149.0), (90.3, 149.0), (140.3, 149.0)]


def main():
    h = LX / NX
    mesh, orc, src = explosive_oracle(LX, LY, h)
    co = COracle(orc)
    loc = [locate(mesh.coords, mesh.cells, p) for p in SENSORS]
    phis = [orc.el.tab(xi[None])[0] for e, xi in loc]
    u = np.zeros((orc.E, orc.nd, 2))
    s = np.zeros((orc.E, orc.nd, 2, 2))
    times = step_times(T, orc.dt)
    tr = np.zeros((len(times), len(SENSORS), 2))
    t0 = time.time()
    for n, t in enumerate(times):
        co.step_inplace(u, s, src(t), orc.dt)
        for k, (e, xi) in enumerate(loc):
            tr[n, k] = phis[k] @ u[e]
    print("steps", len(times), "wall %.1f s" % (time.time() - t0))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_refc_traces.npz"), t=np.array(times), u=tr,
                        sensors=np.array(SENSORS), dt=orc.dt, nx=NX, ny=NY,
                        u_final_norm=np.linalg.norm(u), s_final_norm=np.linalg.norm(s),
                        source="oracle/elastic_c.c via scripts/make_golden_traces.py")


if __name__ == "__main__":
    main()
