"""Peer-memory halo exchange on the GPU (C-ABI sg_ipc_export / sg_peer_connect / sg_exchange + the per-step graph
with six push/signal/wait exchanges).  Two or three ranks share cuda:0 -- CUDA IPC works between processes on one
device, so the multi-rank path is covered on a single-GPU box; torch.distributed (gloo) is only the bootstrap.
The gathered result must match the literal CPU oracle on the whole mesh to the 1e-10 tolerance of BASELINE.json."""
import os
import socket

import numpy as np
import pytest

from tests.util import rel_err, small_mesh

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dim, p, nsteps, out_dir, mode, symmetric):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SG_HALO"] = mode
    # SG_TEST_SPREAD=1 (multi-GPU box): one rank per GPU, so the rows really cross NVLink; default: all on cuda:0
    torch.cuda.set_device(rank % torch.cuda.device_count() if os.environ.get("SG_TEST_SPREAD") else 0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from seigen_b200 import ElasticLF4
        mesh = small_mesh(dim, n=12 if dim == 2 else 4)
        el = ElasticLF4.create(mesh, "DG", p, dimension=dim, solver="explicit", output=False)
        el.density, el.l, el.mu, el.dt = 1.0, 0.5, 0.25, 2e-3
        g = el.S.cell_order
        rng = np.random.default_rng(5)
        E, nd = mesh.num_cells(), el.S.elem.nd
        u0 = rng.standard_normal((E, nd, dim))
        s0 = rng.standard_normal((E, nd, dim, dim))
        if symmetric:
            s0 = 0.5 * (s0 + np.swapaxes(s0, 2, 3))
        el.u0.dat.data[...] = u0[g].reshape(el.u0.dat.data.shape)
        el.s0.dat.data[...] = s0[g].reshape(el.s0.dat.data.shape)
        u1, s1 = el.run((nsteps + 0.5) * el.dt)
        # a second run() continues from the state of the first (exercises the epoch counters across calls)
        u1, s1 = el.run((nsteps + 0.5) * el.dt)
        from seigen_b200.capi import lib, check
        import ctypes
        err = ctypes.c_int64()
        check(lib.sg_peer_error(el._dev.handle, ctypes.byref(err)))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), g=g, u=u1.dat.data, s=s1.dat.data, err=err.value,
                 steps=el.steps_done, mode=el.halo_mode, packed=el._dev.symmetric)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,p,world,symmetric", [(2, 2, 2, False), (2, 1, 3, False), (3, 1, 2, False),
                                                   (2, 2, 2, True), (3, 1, 2, True)])
def test_peer_exchange_matches_oracle(tmp_path, dim, p, world, symmetric):
    """symmetric = False: random s0, so every rank starts with packed stress storage, finds its share asymmetric
    and all ranks fall back to full storage together; True: the packed layout travels through the halo exchange."""
    import torch.multiprocessing as mp
    from oracle.elastic_oracle import ElasticOracle
    nsteps = 3
    mp.spawn(_worker, args=(world, _free_port(), dim, p, nsteps, str(tmp_path), "peer", symmetric), nprocs=world,
             join=True)
    mesh = small_mesh(dim, n=12 if dim == 2 else 4)
    orc = ElasticOracle(mesh.coords, mesh.cells, p)
    orc.l, orc.mu, orc.density, orc.dt = 0.5, 0.25, 1.0, 2e-3
    rng = np.random.default_rng(5)
    E, nd = mesh.num_cells(), orc.nd
    u = rng.standard_normal((E, nd, dim))
    s = rng.standard_normal((E, nd, dim, dim))
    if symmetric:
        s = 0.5 * (s + np.swapaxes(s, 2, 3))
    for _ in range(2 * nsteps):
        u, s, _ = orc.step(u, s, 0.0)
    seen = np.zeros(E, dtype=int)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert int(z["err"]) == 0 and int(z["steps"]) == nsteps and str(z["mode"]) == "peer"
        assert bool(z["packed"]) == symmetric
        g = z["g"]
        seen[g] += 1
        assert rel_err(z["u"].reshape(len(g), nd, dim), u[g]) < 1e-10
        assert rel_err(z["s"].reshape(len(g), nd, dim, dim), s[g]) < 1e-10
    assert (seen == 1).all()
