"""Regenerates tests/golden/oracle_states.npz: seeded inputs and the literal CPU oracle's outputs for small cases.

    python scripts/make_golden_states.py

Each case holds everything needed to rebuild it -- mesh (coords, cells), degree, per-cell lambda/mu, sponge nodal values
(and their degree), source DoFs and amplitudes, dt, symmetric initial data -- plus what oracle/elastic_oracle.py produces
from them: the six stage fields of the first step and the state after NSTEPS steps (tests/golden_cases.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.golden_cases import CASES, GOLDEN, build, run_oracle          # noqa: E402


def main():
    data = {}
    for dim, p in CASES:
        c = build(dim, p)
        c.update(run_oracle(c))
        for k, v in c.items():
            data[f"d{dim}p{p}_{k}"] = np.asarray(v)
        print(f"d{dim}p{p}: {len(c['cells'])} cells, |u_end| = {np.linalg.norm(c['u_end']):.6e}")
    np.savez_compressed(GOLDEN, **data)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
