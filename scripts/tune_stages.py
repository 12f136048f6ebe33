"""Per-pass timing of every compiled kernel variant of one element (development aid; C ABI driven directly).

    python scripts/tune_stages.py --dim 3 --degree 3 --nx 64 --ny 32 --nz 16

For each variant in the library's table that matches (dim, degree) -- selected through SG_TILE / SG_SPLIT / SG_MINB /
SG_MINBA / SG_NS, which sg_create reads -- prints ms of the six passes (sg_time_stage) and of a whole step (sg_step).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from seigen_b200.device import DeviceSolver  # noqa: E402
from seigen_b200.layout import build_rank_plan  # noqa: E402
from seigen_b200.mesh import BoxMesh, RectangleMesh  # noqa: E402
from seigen_b200.refelem import get_refelem  # noqa: E402

VARIANTS = {   # (tile, split, minb, minba, ns, xreg, axs) as compiled in csrc/sg_inst_*.cu; split >= 11: composed variant
    (2, 1): [(128, 1, 4, 2, 22, 1, 1), (64, 1, 8, 4, 22, 1, 1)],
    (2, 2): [(128, 1, 4, 4, 21, 1, 1), (128, 1, 4, 2, 22, 1, 1), (128, 1, 4, 3, 22, 1, 0), (64, 1, 8, 3, 22, 1, 1),
             (256, 1, 2, 1, 22, 1, 1)],
    (2, 3): [(64, 1, 4, 3, 21, 1, 1), (64, 1, 4, 2, 22, 1, 0), (64, 1, 4, 2, 22, 1, 1), (32, 1, 8, 4, 22, 1, 1)],
    (2, 4): [(32, 12, 4, 3, 22, 1, 0), (32, 2, 3, 3, 22, 0, 0), (32, 1, 4, 3, 22, 1, 0), (32, 2, 3, 3, 22, 0, 1)],
    (3, 1): [(128, 11, 2, 2, 22, 1, 1), (128, 1, 2, 2, 22, 1, 1), (128, 1, 2, 3, 21, 1, 1), (64, 1, 4, 4, 22, 1, 1),
             (32, 1, 8, 4, 22, 1, 1)],
    (3, 2): [(64, 3, 2, 2, 22, 0, 0), (64, 13, 4, 2, 22, 0, 0), (64, 1, 4, 2, 22, 0, 0), (32, 3, 3, 3, 22, 0, 1)],
    (3, 3): [(32, 3, 2, 2, 22, 0, 0), (32, 3, 2, 2, 21, 0, 0)],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--nx", type=int, default=1532)
    ap.add_argument("--ny", type=int, default=484)
    ap.add_argument("--nz", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--tag", default="")
    ap.add_argument("--cube", type=int, default=0, help="3D: UnitCubeMesh(N) (the box3d workload's mesh) instead of the box")
    ap.add_argument("--fake-intile", action="store_true",
                    help="measurement only (results are wrong): every out-of-tile facet neighbour is replaced by the cell "
                         "itself, so all facet gathers hit the shared-memory tile -- the time without any L2 gather")
    a = ap.parse_args()
    mesh = RectangleMesh(a.nx, a.ny, 9192.0, 2904.0) if a.dim == 2 else BoxMesh(a.nx, a.ny, a.nz, 4.0, 1.0, 1.0)
    if a.dim == 3 and a.cube:
        mesh = BoxMesh(a.cube, a.cube, a.cube, 1.0, 1.0, 1.0)
    el = get_refelem(a.dim, a.degree)
    E, d = mesh.num_cells(), a.dim
    ndof = E * el.nd * (d + d * d)
    rng = np.random.default_rng(0)
    u = rng.standard_normal((E * el.nd, d)) * 1e-3
    s = rng.standard_normal((E * el.nd, d, d)) * 1e-3
    s = 0.5 * (s + np.swapaxes(s, 1, 2))
    variants = VARIANTS[(a.dim, a.degree)]
    if os.environ.get("SG_ONLY_DEFAULT"):
        variants = variants[:1]
    for tile, split, minb, minba, ns, xreg, axs in variants:
        os.environ.update(SG_TILE=str(tile), SG_SPLIT=str(split), SG_MINB=str(minb), SG_MINBA=str(minba), SG_NS=str(ns),
                          SG_XREG=str(xreg), SG_AXS=str(axs))
        plan = build_rank_plan(mesh, np.zeros(E, dtype=np.int32), 0, 1,
                               tile=None if os.environ.get("SG_TILE_ORDER") == "0" else tile)
        if a.fake_intile:
            me = np.arange(plan.n_owned, dtype=plan.nbr.dtype)[:, None]
            plan.nbr = np.ascontiguousarray(np.where(plan.nbr // tile == me // tile, plan.nbr, me))
        dev = DeviceSolver(mesh, a.degree, symmetric=True, plan=plan)
        dev.set_material(1.0, 0.5, 0.25)
        dev.set_state(u, s)
        dt = 1e-6
        dev.step(3, dt)
        dev.synchronize()
        st = [dev.time_stage(k, dt, 10) for k in range(1, 7)]
        best = 1e9
        for _ in range(2):
            dev.step(a.steps, dt)
            best = min(best, dev.last_step_ms() / a.steps)
        print(f"{a.tag} d{d}p{a.degree} tile={tile} split={split} minb={minb}/{minba} ns={ns} xreg={xreg} axs={axs}: "
              + " ".join(f"K{k + 1}={1e3 * t:.1f}" for k, t in enumerate(st))
              + f" | step {best:.4f} ms {ndof / best / 1e6:.2f} Gupd/s", flush=True)
        dev.close()


if __name__ == "__main__":
    main()
