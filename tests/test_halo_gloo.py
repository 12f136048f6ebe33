"""N > 1 host logic on CPU: rank plans + HaloExchanger over torch.distributed/gloo (world_size 2 and 3).

Each rank advances the LF4 scheme with the nodal CPU operator (oracle/nodal.py) on its own cells, exchanging the
one-layer DG halo after every pass exactly as ``ExplicitElasticLF4._step_multi`` does on GPUs (pack -> send/recv
-> unpack); the gathered result must equal the single-rank run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.nodal import NodalOperator
from seigen_b200.halo import HaloExchanger
from seigen_b200.layout import build_rank_plan, partition_cells
from seigen_b200.refelem import get_refelem
from tests.util import random_state, small_mesh


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _lf4_steps(op, u, s, lam, mu, dt, nsteps, exchange):
    """u, s: (n_total, nd, ...) local arrays whose first op.E cells are owned; exchange(a) refreshes halo cells."""
    E = op.E
    c3 = dt ** 3 / 24.0
    uh, sh = np.zeros_like(u), np.zeros_like(s)
    for _ in range(nsteps):
        uh[:E] = op.Dv(s); exchange(uh)                                   # K1
        sh[:E] = op.Ds(uh, lam, mu); exchange(sh)                         # K2
        u[:E] = u[:E] + dt * uh[:E] + c3 * op.Dv(sh); exchange(u)         # K3
        sh[:E] = op.Ds(u, lam, mu); exchange(sh)                          # K4
        uh[:E] = op.Dv(sh); exchange(uh)                                  # K5
        s[:E] = s[:E] + dt * sh[:E] + c3 * op.Ds(uh, lam, mu); exchange(s)  # K6
    return u, s


def _worker(rank, world, port, dim, p, method, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = small_mesh(dim, n=5 if dim == 2 else 3)
        el = get_refelem(dim, p)
        part = partition_cells(mesh, world, method)
        plan = build_rank_plan(mesh, part, rank, world)
        op = NodalOperator(el.Dr, el.Lift, el.fnodes, el.ftab, plan.nbr, plan.code, plan.jinv, n_owned=plan.n_owned)
        u0, s0 = random_state(mesh, p)
        g = plan.local_to_global
        u, s = u0[g].copy(), s0[g].copy()
        u[plan.n_owned:] = np.nan                       # halo values must come from the exchange
        s[plan.n_owned:] = np.nan
        halo = HaloExchanger(plan, el.nd * dim * dim, "cpu")

        def exchange(a):
            K = int(np.prod(a.shape[1:]))
            flat = a.reshape(a.shape[0], K)
            halo.sendbuf[:halo.nsend * K] = torch.from_numpy(flat[plan.send_cells].reshape(-1))
            halo.exchange(K)
            flat[plan.n_owned:] = halo.recvbuf[:plan.n_halo * K].numpy().reshape(plan.n_halo, K)

        exchange(u)
        exchange(s)
        lam = np.linspace(0.4, 0.6, mesh.num_cells())[g[:plan.n_owned]]
        mu = np.linspace(0.2, 0.3, mesh.num_cells())[g[:plan.n_owned]]
        u, s = _lf4_steps(op, u, s, lam, mu, 1e-2, 3, exchange)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), g=g[:plan.n_owned], u=u[:plan.n_owned], s=s[:plan.n_owned],
                 nb=plan.n_boundary, nh=plan.n_halo)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,p,world,method", [(2, 2, 2, "rcb"), (3, 1, 2, "rcb"), (2, 1, 3, "rcb")])
def test_multi_rank_equals_single_rank(tmp_path, dim, p, world, method):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dim, p, method, str(tmp_path)), nprocs=world, join=True)
    mesh = small_mesh(dim, n=5 if dim == 2 else 3)
    el = get_refelem(dim, p)
    t = mesh.topology
    op = NodalOperator(el.Dr, el.Lift, el.fnodes, el.ftab, t.nbr, t.code, t.jinv)
    u0, s0 = random_state(mesh, p)
    E = mesh.num_cells()
    lam, mu = np.linspace(0.4, 0.6, E), np.linspace(0.2, 0.3, E)
    u_ref, s_ref = _lf4_steps(op, u0.copy(), s0.copy(), lam, mu, 1e-2, 3, lambda a: None)
    seen = np.zeros(E, dtype=int)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        seen[z["g"]] += 1
        assert z["nh"] > 0 and z["nb"] > 0
        assert np.allclose(z["u"], u_ref[z["g"]], rtol=0, atol=1e-13 * np.abs(u_ref).max())
        assert np.allclose(z["s"], s_ref[z["g"]], rtol=0, atol=1e-13 * np.abs(s_ref).max())
    assert (seen == 1).all()                            # every cell owned by exactly one rank


def test_rank_plan_invariants():
    mesh = small_mesh(3, n=3)
    topo = mesh.topology
    for world in (2, 4):
        part = partition_cells(mesh, world)
        plans = [build_rank_plan(mesh, part, r, world) for r in range(world)]
        for r, pl in enumerate(plans):
            g = pl.local_to_global
            assert np.array_equal(g[pl.nbr], topo.nbr[g[:pl.n_owned]])
            assert np.array_equal(pl.code, topo.code[g[:pl.n_owned]])
            assert np.allclose(pl.jinv, topo.jinv[g[:pl.n_owned]], rtol=1e-13)
            reads_halo = (pl.nbr >= pl.n_owned).any(axis=1)
            assert reads_halo[:pl.n_boundary].all() and not reads_halo[pl.n_boundary:].any()
            for q, (o, n) in pl.recv.items():
                so, sn = plans[q].send_offsets[r]
                assert n == sn
                assert np.array_equal(g[pl.n_owned + o:pl.n_owned + o + n],
                                      plans[q].local_to_global[plans[q].send_cells[so:so + sn]])
