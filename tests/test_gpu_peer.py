"""Peer-memory halo exchange on the GPU (C-ABI sg_ipc_export / sg_peer_connect / sg_exchange + the per-step graph whose
six stage kernels carry the six exchanges).  Two or three ranks share cuda:0 -- CUDA IPC works between processes on
one device, so the multi-rank path is covered on a single-GPU box; torch.distributed (gloo) is only the bootstrap.
The gathered result must match the literal CPU oracle on the whole mesh to the 1e-10 tolerance of BASELINE.json, and
the single-rank result of the same library BIT FOR BIT when every cell keeps its own geometry record (the arithmetic
of a cell does not depend on who computes it; tests/tiling/explosive_source.py:659-660 asks for rtol 1e-10)."""
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


def _problem(dim, p, symmetric, unstructured=False):
    if unstructured:
        from tests.test_partition import delaunay_mesh
        mesh = delaunay_mesh(700, seed=9)
    else:
        mesh = small_mesh(dim, n={1: 300, 2: 12, 3: 4}[dim])
    from seigen_b200.refelem import get_refelem
    nd = get_refelem(dim, p).nd
    rng = np.random.default_rng(5)
    E = mesh.num_cells()
    u0 = rng.standard_normal((E, nd, dim))
    s0 = rng.standard_normal((E, nd, dim, dim))
    if symmetric:
        s0 = 0.5 * (s0 + np.swapaxes(s0, 2, 3))
    return mesh, u0, s0


def _run(mesh, dim, p, u0, s0, nsteps):
    from seigen_b200 import ElasticLF4
    el = ElasticLF4.create(mesh, "DG", p, dimension=dim, solver="explicit", output=False)
    el.density, el.l, el.mu, el.dt = 1.0, 0.5, 0.25, 2e-3
    g = el.S.cell_order
    el.u0.dat.data[...] = u0[g].reshape(el.u0.dat.data.shape)
    el.s0.dat.data[...] = s0[g].reshape(el.s0.dat.data.shape)
    u1, s1 = el.run((nsteps + 0.5) * el.dt)
    # a second run() continues from the state of the first (exercises the epoch counters across calls)
    u1, s1 = el.run((nsteps + 0.5) * el.dt)
    return el, g, u1, s1


def _worker(rank, world, port, dim, p, nsteps, out_dir, symmetric, env, unstructured):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.update(env)
    # SG_TEST_SPREAD=1 (multi-GPU box): one rank per GPU, so the rows really cross NVLink; default: all on cuda:0
    torch.cuda.set_device(rank % torch.cuda.device_count() if os.environ.get("SG_TEST_SPREAD") else 0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh, u0, s0 = _problem(dim, p, symmetric, unstructured)
        el, g, u1, s1 = _run(mesh, dim, p, u0, s0, nsteps)
        from seigen_b200.capi import lib, check
        import ctypes
        err = ctypes.c_int64()
        check(lib.sg_peer_error(el._dev.handle, ctypes.byref(err)))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), g=g, u=u1.dat.data, s=s1.dat.data, err=err.value,
                 steps=el.steps_done, mode=el.halo_mode, packed=el._dev.symmetric,
                 nb=el._dev.plan.n_boundary, method=getattr(mesh, "partition_method", "rcb"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


CASES = [
    # dim, p, world, symmetric s0, extra environment, unstructured mesh
    (2, 2, 2, False, {}, False),
    (2, 1, 3, False, {}, False),
    (3, 1, 2, False, {}, False),
    (2, 2, 2, True, {}, False),
    (3, 1, 2, True, {}, False),
    (3, 2, 2, True, {}, False),                                  # SPLIT = 3 kernels through the exchange
    (3, 3, 2, True, {}, False),
    (3, 2, 3, False, {}, False),
    (2, 3, 2, True, {}, False),
    (2, 2, 3, True, {"SG_PARTITION": "metis"}, True),            # graph partitioner on a Delaunay mesh
    (2, 2, 2, True, {"SG_PEER_SCHED_SPLIT": "1"}, False),        # the two-stream schedule (comparison path)
    (1, 2, 2, True, {}, False),                                  # 1-D (tests/pulse/pulse_1d_lf4.py): a facet is a point
]


@pytest.mark.parametrize("dim,p,world,symmetric,env,unstructured", CASES)
def test_peer_exchange_matches_oracle_and_single_rank(tmp_path, dim, p, world, symmetric, env, unstructured):
    """symmetric = False: random s0, so every rank starts with packed stress storage, finds its share asymmetric
    and all ranks fall back to full storage together; True: the packed layout travels through the halo exchange."""
    import torch.multiprocessing as mp
    from oracle.elastic_oracle import ElasticOracle
    nsteps = 3
    env = dict(env, SG_HALO="peer", SG_GEOM_CLASSES="0")
    mp.spawn(_worker, args=(world, _free_port(), dim, p, nsteps, str(tmp_path), symmetric, env, unstructured),
             nprocs=world, join=True)
    mesh, u0, s0 = _problem(dim, p, symmetric, unstructured)
    E = mesh.num_cells()
    # single rank, same library, per-cell geometry: the N-rank result must equal it bit for bit
    os.environ["SG_GEOM_CLASSES"] = "0"
    try:
        el1, g1, u1, s1 = _run(mesh, dim, p, u0, s0, nsteps)
    finally:
        del os.environ["SG_GEOM_CLASSES"]
    nd = el1.S.elem.nd
    ref_u = np.empty((E, nd * dim))
    ref_s = np.empty((E, nd * dim * dim))
    ref_u[g1] = u1.dat.data.reshape(E, -1)
    ref_s[g1] = s1.dat.data.reshape(E, -1)
    # the literal oracle (skipped for the larger elements: the single-rank result above is checked against it in
    # tests/test_gpu_parity.py)
    u = s = None
    if nd <= 10:
        orc = ElasticOracle(mesh.coords, mesh.cells, p)
        orc.l, orc.mu, orc.density, orc.dt = 0.5, 0.25, 1.0, 2e-3
        u, s = u0, s0
        for _ in range(2 * nsteps):
            u, s, _ = orc.step(u, s, 0.0)
    seen = np.zeros(E, dtype=int)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert int(z["err"]) == 0 and int(z["steps"]) == nsteps and str(z["mode"]) == "peer"
        assert bool(z["packed"]) == symmetric
        assert int(z["nb"]) > 0
        if "SG_PARTITION" in env:
            assert str(z["method"]) in ("metis", "rcb")
        g = z["g"]
        seen[g] += 1
        assert np.array_equal(z["u"].reshape(len(g), -1), ref_u[g])
        assert np.array_equal(z["s"].reshape(len(g), -1), ref_s[g])
        if u is not None:
            assert rel_err(z["u"].reshape(len(g), nd, dim), u[g]) < 1e-10
            assert rel_err(z["s"].reshape(len(g), nd, dim, dim), s[g]) < 1e-10
    assert (seen == 1).all()


def test_geometry_classes_change_results_only_at_roundoff(tmp_path):
    """Default mode (cells that are translates of one another share one Jinv record): N ranks vs one rank agree to
    1e-12 -- the class representative depends on who owns which cells, nothing else does."""
    import torch.multiprocessing as mp
    dim, p, world, nsteps = 2, 2, 2, 3
    mp.spawn(_worker, args=(world, _free_port(), dim, p, nsteps, str(tmp_path), True, {"SG_HALO": "peer"}, False),
             nprocs=world, join=True)
    mesh, u0, s0 = _problem(dim, p, True)
    el1, g1, u1, s1 = _run(mesh, dim, p, u0, s0, nsteps)
    E = mesh.num_cells()
    ref_u = np.empty((E, el1.S.elem.nd * dim))
    ref_u[g1] = u1.dat.data.reshape(E, -1)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert int(z["err"]) == 0
        assert rel_err(z["u"].reshape(len(z["g"]), -1), ref_u[z["g"]]) < 1e-12


def _nccl_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SG_HALO"] = "nccl"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        mesh, u0, s0 = _problem(2, 2, True)
        from seigen_b200 import ElasticLF4
        el = ElasticLF4.create(mesh, "DG", 2, dimension=2, solver="explicit", output=False)
        el.density, el.l, el.mu, el.dt = 1.0, 0.5, 0.25, 2e-3
        g = el.S.cell_order
        el.u0.dat.data[...] = u0[g].reshape(el.u0.dat.data.shape)
        el.s0.dat.data[...] = s0[g].reshape(el.s0.dat.data.shape)
        c = mesh.cell_centroids()
        el.receivers = [tuple(c[5]), tuple(c[mesh.num_cells() - 7])]
        u1, s1 = el.run(4.5 * el.dt)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), g=g, u=u1.dat.data, s=s1.dat.data, mode=el.halo_mode,
                 rec=el.receiver_data)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_library_transport_nccl_with_receivers(tmp_path):
    """SG_HALO=nccl (pack -> NCCL send/recv -> unpack, driven pass by pass from the host: what north_star names, kept
    as the baseline transport): same result as one rank, and the receivers are recorded on this path too
    (sg_record_receivers).  NCCL needs one GPU per rank: skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL cannot place two ranks on one device)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mesh, u0, s0 = _problem(2, 2, True)
    from seigen_b200 import ElasticLF4
    el = ElasticLF4.create(mesh, "DG", 2, dimension=2, solver="explicit", output=False)
    el.density, el.l, el.mu, el.dt = 1.0, 0.5, 0.25, 2e-3
    g1 = el.S.cell_order
    el.u0.dat.data[...] = u0[g1].reshape(el.u0.dat.data.shape)
    el.s0.dat.data[...] = s0[g1].reshape(el.s0.dat.data.shape)
    c = mesh.cell_centroids()
    el.receivers = [tuple(c[5]), tuple(c[mesh.num_cells() - 7])]
    u1, s1 = el.run(4.5 * el.dt)
    E = mesh.num_cells()
    ref_u = np.empty((E, u1.dat.data.size // E))
    ref_u[g1] = u1.dat.data.reshape(E, -1)
    rec = np.full_like(el.receiver_data, np.nan)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        assert str(z["mode"]) == "nccl"
        assert rel_err(z["u"].reshape(len(z["g"]), -1), ref_u[z["g"]]) < 1e-12
        mine = np.isfinite(z["rec"])
        rec[mine] = z["rec"][mine]
    # every receiver was recorded by the rank that owns its cell, with the single-rank values (not zeros)
    assert np.isfinite(rec).all() and np.abs(rec).max() > 0
    assert rel_err(rec, el.receiver_data) < 1e-12
