"""Regenerates the fixtures under tests/golden/ from the reference checkout (run in the build container only; the
GPU box has no /root/reference).

    python scripts/make_golden.py [/root/reference]

* ref_c1.npz  -- the external sensor trace tests/explosive_source/REF-C1 (t, ux, uy at (45, 149)) that
  tests/explosive_source/uy.py:7-43 overlays on the simulated -u_y; first 600 rows (t <= 0.6 s) as float64.
* marmousi: see scripts/make_marmousi_fixture.py.
"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(out, exist_ok=True)
rows = []
with open(os.path.join(ref, "tests", "explosive_source", "REF-C1")) as f:
    for line in f:
        rows.append([float(x) for x in line.split()])
a = np.array(rows)[:600]
np.savez_compressed(os.path.join(out, "ref_c1.npz"), t=a[:, 0], ux=a[:, 1], uy=a[:, 2],
                    source="devitocodes/seigen tests/explosive_source/REF-C1 rows 1-600; sensor (45, 149)")
print("wrote ref_c1.npz", a.shape, "peak |uy| %.3e at t=%.3f" % (np.abs(a[:, 2]).max(), a[np.abs(a[:, 2]).argmax(), 0]))
