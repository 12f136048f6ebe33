"""Seeded small cases behind tests/golden/oracle_states.npz (test infrastructure; may import ``oracle``).

``build`` makes the inputs of a case, ``run_oracle`` what the literal CPU oracle (oracle/elastic_oracle.py, the
quadrature restatement of seigen/elastic.py:204-219, 156-202, 341-352, 267-315) produces from them.  The fixture is
written by scripts/make_golden_states.py, pinned on the CPU by tests/test_oracle_golden.py and compared with the CUDA
path by tests/test_gpu_vectors.py.
"""
import os

import numpy as np

from oracle.elastic_oracle import ElasticOracle
from seigen_b200.mesh import BoxMesh, RectangleMesh, perturb_vertices

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_states.npz")
INPUTS = ("coords", "cells", "degree", "sigma_degree", "lam", "mu", "sigma", "sdof", "amp", "u0", "s0", "dt")
OUTPUTS = ("uh1", "stemp", "u1", "sh1", "utemp", "s1", "u_end", "s_end")
CASES = [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2)]
NSTEPS = 4
DT = 2e-3


def build(dim, p):
    seed = 100 * dim + p
    mesh = RectangleMesh(4, 5, 1.3, 1.0) if dim == 2 else BoxMesh(2, 3, 2, 1.0, 1.2, 0.9)
    mesh = perturb_vertices(mesh, 0.15, seed)
    rng = np.random.default_rng(seed)
    E, d = mesh.num_cells(), dim
    q = {1: 1, 2: 4, 3: 3}[p] if dim == 2 else 1
    orc = ElasticOracle(mesh.coords, mesh.cells, p, sigma_degree=q)
    nd = orc.nd
    lam, mu = rng.uniform(0.4, 0.6, E), rng.uniform(0.2, 0.3, E)
    sig = rng.uniform(0.0, 3.0, size=(E, orc.sel.nd))
    sig[rng.uniform(size=E) < 0.5] = 0.0
    cells = rng.choice(E, size=3, replace=False)
    sdof = [((c * nd + node) * d + i) * d + i for c in cells for node in range(2) for i in range(d)]
    base = (cells[0] * nd) * d * d
    sdof = np.array(sdof + [base + 1, base + d], dtype=np.int64)           # + one symmetric off-diagonal pair
    amp = rng.standard_normal((NSTEPS, len(sdof)))
    amp[:, -1] = amp[:, -2]
    u0 = rng.standard_normal((E, nd, d))
    s0 = rng.standard_normal((E, nd, d, d))
    s0 = 0.5 * (s0 + np.swapaxes(s0, 2, 3))
    return dict(coords=mesh.coords, cells=mesh.cells, degree=p, sigma_degree=q, lam=lam, mu=mu, sigma=sig, sdof=sdof,
                amp=amp, u0=u0, s0=s0, dt=DT)


def run_oracle(c):
    """Outputs of the oracle for the inputs of one case (shared with tests/test_oracle_golden.py)."""
    orc = ElasticOracle(c["coords"], c["cells"], int(c["degree"]), sigma_degree=int(c["sigma_degree"]))
    orc.l, orc.mu, orc.density, orc.dt = c["lam"], c["mu"], 1.0, float(c["dt"])
    orc.sigma = c["sigma"]
    E, nd, d = orc.E, orc.nd, c["u0"].shape[-1]

    def src_at(step):
        out = np.zeros(E * nd * d * d)
        out[c["sdof"]] = c["amp"][step]
        return out.reshape(E, nd, d, d)
    out = {}
    u, s = c["u0"], c["s0"]
    for n in range(len(c["amp"])):
        orc.source = (lambda t, n=n: src_at(n))
        u, s, st = orc.step(u, s, 0.0)
        if n == 0:
            out.update(uh1=st["uh1"], stemp=st["stemp"], u1=u, sh1=st["sh1"], utemp=st["utemp"], s1=s)
    out.update(u_end=u, s_end=s)
    return out




def load_case(dim, p):
    """Inputs and stored oracle outputs of one case from the committed fixture."""
    z = np.load(GOLDEN)
    return {k: z[f"d{dim}p{p}_{k}"] for k in INPUTS + OUTPUTS}
