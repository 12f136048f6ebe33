"""Regenerates the fixtures under tests/golden/ from the reference checkout (run in the build container only; the
GPU box has no /root/reference).

    python scripts/make_golden.py [/root/reference]

* ref_c1.npz  -- the external sensor trace tests/explosive_source/REF-C1 (t, ux, uy at (45, 149)) that
  tests/explosive_source/uy.py:7-43 overlays on the simulated -u_y; first 600 rows (t <= 0.6 s) as float64.
* ref_c123.npz -- all three traces REF-C1, REF-C2, REF-C3 (sensors (45, 149), (90, 149), (140, 149), uy.py:36-43), all 2500
  rows (t = 0.001 .. 2.5 s), for the full-domain run of BASELINE.json configs[1] (tests/test_gpu_fullsize.py).
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

tr = []
for i in (1, 2, 3):
    with open(os.path.join(ref, "tests", "explosive_source", "REF-C%d" % i)) as f:
        tr.append(np.array([[float(x) for x in line.split()] for line in f]))
assert all(np.array_equal(tr[0][:, 0], x[:, 0]) for x in tr)
np.savez_compressed(os.path.join(out, "ref_c123.npz"), t=tr[0][:, 0], ux=np.stack([x[:, 1] for x in tr]),
                    uy=np.stack([x[:, 2] for x in tr]), sensors=np.array([[45.0, 149.0], [90.0, 149.0], [140.0, 149.0]]),
                    source="devitocodes/seigen tests/explosive_source/REF-C1..3, all rows")
for i, x in enumerate(tr):
    k = np.abs(x[:, 2]).argmax()
    print("REF-C%d peak |uy| %.3e at t=%.3f" % (i + 1, abs(x[k, 2]), x[k, 0]))
