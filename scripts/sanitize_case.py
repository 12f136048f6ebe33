"""Small cases for compute-sanitizer (memcheck / racecheck / synccheck), SURVEY.md section 5:

    compute-sanitizer --tool memcheck python scripts/sanitize_case.py single
    compute-sanitizer --tool memcheck --target-processes all python scripts/sanitize_case.py peers

single: 2D P2 (sponge, source, per-cell material, receivers) and 3D P1 / 3D P2, a few steps through ElasticLF4.run.
peers:  two ranks on cuda:0 exchanging halos inside the stage kernels (CUDA IPC), 2D P2 and 3D P1.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def problem(dim, p, sponge=True, source=True):
    from seigen_b200 import (BoxMesh, ElasticLF4, Expression, Function, FunctionSpace, RectangleMesh)
    mesh = RectangleMesh(12, 9, 30.0, 15.0) if dim == 2 else BoxMesh(4, 3, 3, 4.0, 3.0, 3.0)
    el = ElasticLF4.create(mesh, "DG", p, dimension=dim, solver="explicit", output=False)
    rng = np.random.default_rng(3)
    n = el.S.plan.n_owned
    el.density, el.dt = 1.0, 1e-3
    el.l, el.mu = rng.uniform(0.4, 0.6, n), rng.uniform(0.2, 0.3, n)
    if sponge:
        el.absorption_function = Function(FunctionSpace(mesh, "DG", 1))
        el.absorption = Expression("x[0] <= 5 ? 10.0 : 0")
    if source:
        z = ", ".join(['"0.0"'] * dim)
        box = "x[0] >= 1.0 && x[0] <= 2.9 && x[1] >= 1.0 && x[1] <= 2.9 ? sin(40*t) : 0.0"
        rows = [[box if i == j else "0.0" for j in range(dim)] for i in range(dim)]
        el.source_expression = Expression(tuple(tuple(r) for r in rows), t=0.0)
        el.source_function = Function(el.S)
        el.source = el.source_expression
        del z
    el.u0.dat.data[...] = 1e-3 * rng.standard_normal(el.u0.dat.data.shape)
    s0 = 1e-3 * rng.standard_normal(el.s0.dat.data.shape)
    el.s0.dat.data[...] = 0.5 * (s0 + np.swapaxes(s0, 1, 2))
    el.receivers = [(2.3, 2.1) if dim == 2 else (2.3, 2.1, 1.2)]
    return el


def run(dim, p):
    el = problem(dim, p)
    u1, s1 = el.run(4.5 * el.dt)
    assert el.steps_done == 4 and np.isfinite(u1.dat.data).all() and np.isfinite(s1.dat.data).all()
    return el


def peer_worker(rank, world, port):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for dim, p in ((2, 2), (3, 1)):
            el = run(dim, p)
            assert el.halo_mode == "peer"
            el.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "single"
    if mode == "single":
        for dim, p in ((2, 2), (3, 1), (3, 2)):
            run(dim, p).close()
        print("sanitize_case single: ok")
    else:
        import torch.multiprocessing as mp
        mp.spawn(peer_worker, args=(2, 29533), nprocs=2, join=True)
        print("sanitize_case peers: ok")
