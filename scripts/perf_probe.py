"""Quick single-GPU throughput probe of sg_step on synthetic structured meshes (development aid)."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from seigen_b200.device import DeviceSolver
from seigen_b200.mesh import BoxMesh, RectangleMesh
from seigen_b200.refelem import get_refelem


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--nx", type=int, default=1532)
    ap.add_argument("--ny", type=int, default=484)
    ap.add_argument("--nz", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sym", type=int, default=1, help="symmetric stress storage on the device")
    a = ap.parse_args()
    t0 = time.time()
    if a.dim == 2:
        mesh = RectangleMesh(a.nx, a.ny, 9192.0, 2904.0)
    else:
        mesh = BoxMesh(a.nx, a.ny, a.nz, 4.0, 1.0, 1.0)
    el = get_refelem(a.dim, a.degree)
    E = mesh.num_cells()
    d = a.dim
    ndof = E * el.nd * (d + d * d)
    dev = DeviceSolver(mesh, a.degree, symmetric=bool(a.sym))
    t1 = time.time()
    dev.set_material(1.0, 0.5, 0.25)
    rng = np.random.default_rng(0)
    u = rng.standard_normal((E * el.nd, d)) * 1e-3
    s = rng.standard_normal((E * el.nd, d, d)) * 1e-3
    if a.sym:
        s = 0.5 * (s + np.swapaxes(s, 1, 2))
    dev.set_state(u, s)
    dt = 1e-6
    dev.step(3, dt)
    dev.synchronize()
    for r in range(a.reps):
        dev.step(a.steps, dt)
        ms = dev.last_step_ms() / a.steps
        import os
        tag = " ".join(f"{k}={os.environ[k]}" for k in ("SG_TILE", "SG_SPLIT", "SG_MINB") if k in os.environ)
        print(f"{tag} sym={a.sym} dim={d} p={a.degree} cells={E} dof={ndof} setup={t1 - t0:.1f}s  {ms:.4f} ms/step  "
              f"{ndof / ms / 1e6:.2f} Gupd/s  {64 * ndof / ms / 1e6:.0f} GB/s algorithmic")
    dev.close()


if __name__ == "__main__":
    main()
