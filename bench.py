#!/usr/bin/env python
"""Headline benchmark: DoF-time-step updates/s of the ElasticLF4 explicit update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Workload at every N (weak scaling): BASELINE.json configs[3], the Marmousi 2D heterogeneous model at DG P2 --
``RectangleMesh(1532*N, 484, 9192*N, 2904)`` (h = 6 m; 1 482 976 cells = 53 387 136 DoF *per GPU*; the Vp grid
repeats with period 9192 m in x), per-cell Lame parameters mu = lambda = Vp^2/3, rho = 1, dt = 0.5*h/(2*Vp_max),
a Ricker point source one cell row below the surface of every tile, random 1e-3 initial data.  configs[1]
(explosive source, ~1 M DoF) is 8 MB of state -- it lives in L2 and measures launch latency, not the HBM
roofline the metric asks for -- so it is a parity-test case (tests/), not the bench line.  The state (427 MB per
GPU) is larger than L2 (126 MB), which is what keeps the timed iterations cold (``config.l2``).

A "step" is one LF4 time step = six fused passes.  ``value`` times K steps on the device with CUDA
events on the solver's stream (state resident in HBM); ``e2e`` times ``ElasticLF4.run(T)`` -- the call a user of
the reference makes -- K steps per call, host (page-locked) u0/s0 copied in and u1/s1 copied out inside the timed
region.  ``roofline`` is for the dominant kernel (pass K6, ``stage_g_kernel<AXPY>``); ``roofline_step`` for the
whole step (64 B per DoF per step).  ``cpu_baseline`` / ``--impl reference`` time the C/OpenMP
restatement of the reference's PyOP2 loop structure (``oracle/``; Firedrake itself cannot be installed here,
DESIGN.md) on the host cores, all of them (OMP_NUM_THREADS=1 exported by torchrun is overridden): ``--impl reference``
on the SAME 1532 x 484 P2 mesh, materials, source and dt as one GPU's share of this arm (53.4 M DoF, a few time
steps); the ``cpu_baseline`` leg inside the GPU arm on the h = 24 m sample of it, to keep the default run short.

``extra`` carries what the headline number does not show (``--extras none`` skips it): at N = 1 ``box3d``
(BASELINE.json configs[4], UnitCubeMesh(26) P3) and ``elements`` (device-resident throughput of every supported
element on a structured mesh); at N > 1 ``strong`` (the ONE 53.4 M-DoF model cut N ways: north_star's ">= 85 % at 8
GPUs on a >= 50 M-DoF mesh") and ``box3d`` weak-scaled.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dof_timestep_updates_per_sec"
UNIT = "DoF-updates/s"
NX, NY, LX, LY = 1532, 484, 9192.0, 2904.0
DEGREE = 2
VP_MAX = 5500.0


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as exc:  # pragma: no cover
            self.nv = None
            self.error = repr(exc)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:  # pragma: no cover
                    pass
            time.sleep(0.01)

    def __enter__(self):
        self._active.set()
        return self

    def __exit__(self, *a):
        self._active.clear()

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "error", "?")]}
        self._stop.set()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------------
RICKER = ("x[0] >= {x0} && x[0] <= {x1} && x[1] >= {y0} && x[1] <= {y1} ? "
          "(-1.0 + 2*a*pow(t - t0, 2))*exp(-a*pow(t - t0, 2)) : 0.0")


def marmousi_problem(nranks, scale=1.0, strong=False):
    """ElasticLF4 set up through the public API exactly as a reference script would (tests/explosive_source/
    explosive_source_lf4.py:9-52), on this rank's share of the weak-scaled Marmousi mesh."""
    from seigen_b200 import ElasticLF4, Expression, Function, RectangleMesh
    from seigen_b200.marmousi import marmousi_lame_at

    nx, ny = max(2, int(round(NX * scale))), max(2, int(round(NY * scale)))
    ntile = 1 if strong else nranks          # strong scaling: the one 53 M-DoF model is cut into nranks parts
    mesh = RectangleMesh(nx * ntile, ny, LX * ntile, LY)
    el = ElasticLF4.create(mesh, "DG", DEGREE, dimension=2, solver="explicit", output=False)
    order = el.S.cell_order
    cent = mesh.coords[mesh.cells[order]].mean(axis=1)
    lam, mu = marmousi_lame_at(cent)
    el.density, el.l, el.mu = 1.0, lam, mu
    h = LX / nx
    el.dt = 0.5 * h / (2 ** (DEGREE - 1) * VP_MAX)
    # Ricker source (explosive_source_lf4.py:36-40) in a one-cell box below the surface of every tile
    fpeak = 10.0
    a = (np.pi * fpeak) ** 2
    boxes = []
    for r in range(ntile):
        xc = LX * (r + 0.5)
        boxes.append(RICKER.format(x0=xc - 0.5 * h, x1=xc + 0.5 * h, y0=LY - 1.5 * h, y1=LY - 0.5 * h))
    src = " + ".join("(" + b + ")" for b in boxes)
    el.source_expression = Expression(((src, "0.0"), ("0.0", src)), a=a, t0=0.1, t=0.0)
    el.source_function = Function(el.S)
    el.source = el.source_expression
    rng = np.random.default_rng(1234 + el.S.plan.rank)
    el.u0.dat.data[...] = 1e-3 * rng.standard_normal(el.u0.dat.data.shape)
    s0 = 1e-3 * rng.standard_normal(el.s0.dat.data.shape)
    el.s0.dat.data[...] = 0.5 * (s0 + np.swapaxes(s0, 1, 2))      # a stress tensor: symmetric
    name = f"marmousi_2d_p{DEGREE}_{nx}x{ny}_" + ("total" if strong else "per_gpu")
    return el, name


BOX3D_N = {1: 26, 2: 33, 4: 41, 8: 52}     # 25.3 / 51.7 / 99.2 / 202.5 M DoF at P3


def box3d_problem(nranks):
    """BASELINE.json configs[4] (--workload box3d): UnitCubeMesh(N) of Kuhn tetrahedra, DG P3 (240 DoF per cell),
    weak-scaled with N = 26 / 33 / 41 / 52 for 1 / 2 / 4 / 8 GPUs (~25 M DoF per GPU), rho = 1, mu = 0.25,
    lambda = 0.5, dt = 0.5 h / (2^(p-1) Vp) as tests/eigenmode/eigenmode_3d.py:72-79, random 1e-3 initial data."""
    from seigen_b200 import ElasticLF4, UnitCubeMesh
    p = 3
    N = BOX3D_N.get(nranks) or int(round(26 * nranks ** (1.0 / 3.0)))
    mesh = UnitCubeMesh(N, N, N)
    el = ElasticLF4.create(mesh, "DG", p, dimension=3, solver="explicit", output=False)
    el.density, el.l, el.mu = 1.0, 0.5, 0.25
    el.dt = 0.5 * (1.0 / N) / (2 ** (p - 1) * 1.0)
    rng = np.random.default_rng(4321 + el.S.plan.rank)
    el.u0.dat.data[...] = 1e-3 * rng.standard_normal(el.u0.dat.data.shape)
    s0 = 1e-3 * rng.standard_normal(el.s0.dat.data.shape)
    el.s0.dat.data[...] = 0.5 * (s0 + np.swapaxes(s0, 1, 2))
    return el, f"box3d_p{p}_unitcube_{N}"


def sample_problem():
    """Bounded CPU sample of the same workload: the Marmousi grid at its native h = 24 m (seigen/marmousi.py:18-21),
    P2 -- 92 686 cells, 3 336 696 DoF, same materials rule, same dt rule."""
    from oracle.c_oracle import COracle
    from oracle.elastic_oracle import ElasticOracle
    from seigen_b200.marmousi import marmousi_lame
    from seigen_b200.mesh import RectangleMesh

    mesh = RectangleMesh(383, 121, LX, LY)
    orc = ElasticOracle(mesh.coords, mesh.cells, DEGREE)
    lam, mu = marmousi_lame(mesh)
    orc.l, orc.mu, orc.density = lam, mu, 1.0
    orc.dt = 0.5 * 24.0 / (2 ** (DEGREE - 1) * VP_MAX)
    co = COracle(orc)
    E, nd = orc.E, orc.nd
    rng = np.random.default_rng(7)
    u = 1e-3 * rng.standard_normal((E, nd, 2))
    s = 1e-3 * rng.standard_normal((E, nd, 2, 2))
    ndof = E * nd * 6
    return co, u, s, orc.dt, ndof, f"Marmousi 2D P{DEGREE} at h=24 m: {E} cells, {ndof} DoF", fused_from(mesh, orc)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # pragma: no cover
        return os.cpu_count() or 1


def full_problem_cpu(scale=1.0):
    """The GPU arm's per-GPU workload for the CPU arm: same mesh (1532 x 484, h = 6 m), degree, per-cell Lame
    parameters, Ricker source box, dt and kind of initial data as ``marmousi_problem(1)`` -- 53 387 136 DoF."""
    from oracle.c_oracle import COracle
    from oracle.elastic_oracle import ElasticOracle
    from seigen_b200 import Expression
    from seigen_b200.marmousi import marmousi_lame
    from seigen_b200.mesh import RectangleMesh

    nx, ny = max(2, int(round(NX * scale))), max(2, int(round(NY * scale)))
    mesh = RectangleMesh(nx, ny, LX, LY)
    orc = ElasticOracle(mesh.coords, mesh.cells, DEGREE, lite=True)
    lam, mu = marmousi_lame(mesh)
    orc.l, orc.mu, orc.density = lam, mu, 1.0
    h = LX / nx
    orc.dt = 0.5 * h / (2 ** (DEGREE - 1) * VP_MAX)
    co = COracle(orc)
    E, nd = orc.E, orc.nd
    a = (np.pi * 10.0) ** 2
    xc = LX * 0.5
    box = RICKER.format(x0=xc - 0.5 * h, x1=xc + 0.5 * h, y0=LY - 1.5 * h, y1=LY - 0.5 * h)
    expr = Expression(((box, "0.0"), ("0.0", box)), a=a, t0=0.1, t=0.0)
    xs = orc.node_coords().reshape(-1, 2)
    near = np.flatnonzero((np.abs(xs[:, 0] - xc) <= h) & (xs[:, 1] >= LY - 2 * h))
    active = near[np.any(expr.evaluate(xs[near], t=0.1).reshape(len(near), -1) != 0, axis=1)]
    src = np.zeros((E * nd, 2, 2))
    src[active] = expr.evaluate(xs[active], t=orc.dt)
    src = src.reshape(E, nd, 2, 2)
    rng = np.random.default_rng(1234)
    u = 1e-3 * rng.standard_normal((E, nd, 2))
    s0 = 1e-3 * rng.standard_normal((E, nd, 2, 2))
    s = np.ascontiguousarray(0.5 * (s0 + np.swapaxes(s0, 2, 3)))
    ndof = E * nd * 6
    return co, u, s, src, orc.dt, ndof, f"marmousi_2d_p{DEGREE}_{nx}x{ny}", fused_from(mesh, orc)


def fused_from(mesh, orc):
    """The second CPU baseline of BASELINE.md section 3 for the same problem: the six-pass algorithm the GPU runs, in
    C/OpenMP (oracle/elastic_fused_c.c)."""
    from oracle.c_fused import CFused
    from seigen_b200.refelem import get_refelem
    el = get_refelem(mesh.dim, orc.p)
    t = mesh.topology
    return CFused(el.Dr, el.Lift, el.fnodes, el.ftab, t.nbr, t.code, t.jinv, orc.l, orc.mu, orc.density,
                  sigma_mats=CFused.sponge_matrices(orc))


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    --set full capture of this workload (profiles/dominant_kernel_traffic.json), or None if not captured yet."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))["bytes_per_launch"])
    except Exception:
        return None


def time_cpu(co, u, s, dt, steps, warmup, src=None):
    for _ in range(warmup):
        co.step_inplace(u, s, src, dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        co.step_inplace(u, s, src, dt)
    return time.perf_counter() - t0


# ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    t0 = time.perf_counter()
    co, u, s, src, dt, ndof, wname, cf = full_problem_cpu(args.scale)
    cores = co.set_threads(host_threads())          # torchrun exports OMP_NUM_THREADS=1: use every core we may
    cf.set_threads(cores)
    setup = time.perf_counter() - t0
    # each bench step = one LF4 time step of the full per-GPU mesh (~1 s on 16+ cores); capped so that the whole run
    # stays within a few minutes
    t_probe = time_cpu(co, u, s, dt, 1, 0, src)
    steps = max(1, min(args.steps, int(60.0 / max(t_probe, 1e-6))))
    warm = max(0, min(args.warmup, int(15.0 / max(t_probe, 1e-6))))
    wall = time_cpu(co, u, s, dt, steps, warm, src)
    val = ndof * steps / wall
    # second CPU baseline (reported beside the first, not instead of it): the fused six-pass algorithm on the host
    tf = time_cpu(cf, u, s, dt, max(2, steps), 1, src)
    fused = {"value": ndof * max(2, steps) / tf, "unit": UNIT, "cores": cf.threads, "kind": "port-fused",
             "note": "the six-pass algorithm the GPU runs (SURVEY.md 8a K1-K6) in C/OpenMP, oracle/elastic_fused_c.c"}
    world = env_int("WORLD_SIZE", 1)
    sample = (f"{wname}: the GPU arm's mesh, degree, per-cell materials, Ricker source and dt for ONE GPU "
              f"({ndof} DoF), {steps} time steps" + ("" if world == 1 else f" -- 1/{world} of the {world}-GPU workload"))
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": wname + "_per_gpu", "degree": DEGREE, "dim": 2, "dof": int(ndof), "dt": dt,
                      "sample": sample, "setup_s": setup,
                      "material": "per-cell lambda=mu=Vp^2/3 from the Marmousi grid, rho=1",
                      "note": "Firedrake/PyOP2 cannot be installed here; this is the C/OpenMP restatement of the "
                              "reference's PyOP2 loop structure (oracle/elastic_c.c: 25 sweeps per step), all host "
                              "threads"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "cpu_baseline_fused": fused,
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


ELEMENT_MESHES = [   # (dim, degree, builder arguments): structured meshes of ~25-55 M DoF, as scripts/perf_probe.py
    (2, 1, (1532, 484)), (2, 2, (1532, 484)), (2, 3, (1000, 400)), (2, 4, (800, 300)),
    (3, 1, (128, 32, 32)), (3, 2, (64, 32, 32)), (3, 3, (64, 32, 16)),
]


def element_table(steps, peak):
    """Device-resident DoF-updates/s of every supported element (constant material, no source, symmetric stress
    storage), C ABI driven directly: the kernels behind BASELINE.json configs[0]-[4] beside the headline one."""
    from seigen_b200.device import DeviceSolver
    from seigen_b200.mesh import BoxMesh, RectangleMesh
    from seigen_b200.refelem import get_refelem
    rows = []
    for d, p, n in ELEMENT_MESHES:
        mesh = RectangleMesh(n[0], n[1], LX, LY) if d == 2 else BoxMesh(n[0], n[1], n[2], 4.0, 1.0, 1.0)
        nd = get_refelem(d, p).nd
        E = mesh.num_cells()
        ndof = E * nd * (d + d * d)
        dev = DeviceSolver(mesh, p, symmetric=True)
        dev.set_material(1.0, 0.5, 0.25)
        rng = np.random.default_rng(0)
        u = 1e-3 * rng.standard_normal((E * nd, d))
        s = 1e-3 * rng.standard_normal((E * nd, d, d))
        dev.set_state(u, 0.5 * (s + np.swapaxes(s, 1, 2)))
        dev.step(3, 1e-6)
        dev.synchronize()
        best = None
        for _ in range(2):
            dev.step(steps, 1e-6)
            ms = dev.last_step_ms() / steps
            best = ms if best is None else min(best, ms)
        dev.close()
        rows.append({"dim": d, "degree": p, "cells": int(E), "dof": int(ndof), "ms_per_step": best,
                     "value": ndof / (best * 1e-3), "roofline_step_frac": 64.0 * ndof / (best * 1e-3) / 1e9 / peak})
    return rows


def device_value(el, K, W, barrier, max_over_ranks, sum_over_ranks):
    """K steps after W warm-up steps, state resident in HBM, CUDA events on the solver's stream, max over ranks."""
    dt = float(el.dt)
    el.run((W + 0.5) * dt)                      # plan, upload of geometry / material / source table, graph (untimed)
    dev = el._dev
    nd, d = el.S.elem.nd, el.dimension
    ndof_local = dev.n_owned * nd * (d + d * d)
    ndof = sum_over_ranks(ndof_local)
    el.setup([dt * (i + 1) for i in range(K)])
    el._upload_state()
    el._advance(W, 0)
    barrier()
    el._advance(K, 0)
    barrier()
    ms = max_over_ranks(dev.last_step_ms())
    el._check_peers("bench")
    return {"value": ndof * K / (ms * 1e-3), "ms_per_step": ms / K, "dof_total": int(ndof),
            "dof_per_gpu": int(ndof_local), "cells_per_gpu": int(dev.n_owned), "steps": K}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from seigen_b200 import helpers
    helpers.LOG_STREAM = sys.stderr            # stdout carries the one JSON line and nothing else

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (seigen_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    from seigen_b200 import capi

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_setup = time.perf_counter()
    strong = args.scaling == "strong"
    if args.workload == "box3d":
        el, wname = box3d_problem(world)
    else:
        el, wname = marmousi_problem(world, args.scale, strong)
    K, W = args.steps, max(3, args.warmup)
    dt = float(el.dt)
    T = (K + 0.5) * dt
    # first run(): builds the rank plan, uploads geometry/material/source table, captures the graph (untimed)
    el.run((W + 0.5) * dt)
    dev = el._dev
    nd, d = el.S.elem.nd, el.dimension
    ndof_local = dev.n_owned * nd * (d + d * d)
    ndof = sum_over_ranks(ndof_local)
    t_setup = time.perf_counter() - t_setup
    sampler = ClockSampler(local)

    # ---- value: K steps, state resident in HBM, CUDA events on the solver's stream, max over ranks -----------
    el.setup([dt * (i + 1) for i in range(K)])
    el._upload_state()
    el._advance(W, 0)
    barrier()
    with sampler:
        el._advance(K, 0)
        barrier()
    ms_local = dev.last_step_ms()
    ms = max_over_ranks(ms_local)
    value = ndof * K / (ms * 1e-3)
    step_counter = 1 if el.source_function is not None else 0     # the source is added inside the G-type passes
    if world == 1 or not os.environ.get("SG_PEER_SCHED_SPLIT"):
        launches = K * (6 + step_counter)       # six fused passes (+ device-side step counter); with peers the halo
    else:                                       # exchange rides inside the same six kernels
        launches = K * (6 * 5 + step_counter)   # two-stream schedule: boundary, interior, push, signal, wait per pass

    # ---- per-pass timing of the six kernels (roofline of the dominant one) ----------------------------------
    reps = max(10, min(K, 50))
    stage_ms = [dev.time_stage(k, dt * 1e-3, reps) for k in range(1, 7)]
    E = dev.n_owned
    # field doubles per node per pass: algorithmic = the reference's full d*d stress storage (SURVEY.md 8d,
    # DESIGN.md section 4); moved = what this solver really reads + writes (upper triangle when the stress is symmetric)
    ncs = d * (d + 1) // 2 if dev.symmetric else d * d

    def pass_doubles(S, U):
        return {1: S + U, 2: S + U, 3: S + 3 * U, 4: S + U, 5: S + U, 6: U + 3 * S}
    cell_doubles, moved_doubles = pass_doubles(d * d, d), pass_doubles(ncs, d)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    stages = []
    for k in range(1, 7):
        b = cell_doubles[k] * nd * 8.0 * E
        mb = moved_doubles[k] * nd * 8.0 * E
        stages.append({"pass": f"K{k}", "ms": stage_ms[k - 1], "alg_bytes": b, "gbs": b / stage_ms[k - 1] / 1e6,
                       "moved_bytes": mb, "moved_gbs": mb / stage_ms[k - 1] / 1e6})
    dom = stages[5]
    roofline = {"kernel": "stage_g_kernel<%d,%d,AXPY> (pass K6: s1 = s0 + dt*sh1 + dt^3/24*(Ds(utemp)+src))" % (d, el.S.degree),
                "bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak,
                "traffic": profiled_traffic() if args.workload == "marmousi" and dev.symmetric else None,
                "peak_source": peak_src,
                "alg_bytes_per_launch": dom["alg_bytes"], "ms_per_launch": dom["ms"],
                "moved_bytes_per_launch": dom["moved_bytes"], "moved_gbs": dom["moved_gbs"],
                "moved_frac": dom["moved_gbs"] / peak,
                "note": "achieved/frac use the ALGORITHMIC bytes of SURVEY.md 8d (full d*d stress: %d*nd*8 B per cell "
                        "for K6); with symmetric stress storage the kernel moves only moved_bytes_per_launch "
                        "(%d*nd*8 B per cell), so frac can exceed moved_frac -- moved_frac is the kernel's real "
                        "HBM efficiency, frac the speed-up-relevant one" % (cell_doubles[6], moved_doubles[6])}
    step_gbs = 64.0 * ndof_local * K / (ms_local * 1e-3) / 1e9
    moved_per_dof = 8.0 * sum(moved_doubles.values()) / (d + d * d)
    roofline_step = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                     "alg_bytes_per_dof_step": 64, "moved_bytes_per_dof_step": moved_per_dof,
                     "moved_frac": step_gbs * moved_per_dof / 64.0 / peak,
                     "note": "per GPU, whole step = 6 passes"}

    # ---- e2e: ElasticLF4.run(T), host page-locked state in and out inside the timed region -------------------
    el.run(T)                                   # warm: source table for these K steps, graph
    if os.environ.get("SG_BENCH_DEBUG"):
        from seigen_b200 import get_timers
        get_timers(reset=True)
    walls = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        with sampler:
            el.run(T)
            barrier()
        walls.append(max_over_ranks(time.perf_counter() - t0))
    e2e_wall = float(np.mean(walls))
    if os.environ.get("SG_BENCH_DEBUG") and rank == 0:
        from seigen_b200 import get_timers
        print("timers:", {k: round(v, 4) for k, v in get_timers().items()}, "walls:", walls, file=sys.stderr)
    state_bytes = sum_over_ranks(el.u0.dat.data.nbytes + el.s0.dat.data.nbytes)
    e2e = {"value": ndof * K / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": state_bytes / K,
           "d2h_bytes_per_step": state_bytes / K, "call": f"ElasticLF4.run(T) with T = {K} steps per call",
           "h2d_bytes_per_call": state_bytes, "d2h_bytes_per_call": state_bytes, "wall_s_per_call": e2e_wall}
    clocks = sampler.summary()

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu and args.workload == "marmousi":
        co, u, s, cdt, cndof, sample, cf = sample_problem()
        co.set_threads(host_threads())
        cf.set_threads(co.threads)
        t1 = time_cpu(co, u, s, cdt, 1, 1)
        csteps = max(2, min(60, int(15.0 / max(t1, 1e-6))))
        cw = time_cpu(co, u, s, cdt, csteps, 0)
        fw = time_cpu(cf, u, s, cdt, csteps, 1)
        cpu = {"value": cndof * csteps / cw, "unit": UNIT, "cores": co.threads, "kind": "port",
               "sample": f"{sample}, {csteps} steps",
               "fused_variant": {"value": cndof * csteps / fw, "unit": UNIT, "cores": cf.threads, "kind": "port-fused",
                                 "note": "second CPU baseline of BASELINE.md section 3: the six-pass algorithm in "
                                         "C/OpenMP (oracle/elastic_fused_c.c)"}}

    degree, halo_mode, symmetric = int(el.S.degree), el.halo_mode, bool(dev.symmetric)
    # ---- extras: what the headline does not show (3D, strong scaling, the other elements) -------------------------
    extra = {}
    if args.extras != "none":
        el.close()
        del el
        import gc
        gc.collect()
        Kx = max(10, min(K, 50))
        if args.workload == "marmousi":
            b3, b3name = box3d_problem(world)
            r = device_value(b3, Kx, W, barrier, max_over_ranks, sum_over_ranks)
            r.update(workload=b3name, scaling="weak", n_gpus=world,
                     roofline_step_frac=64.0 * r["dof_per_gpu"] / (r["ms_per_step"] * 1e-3) / 1e9 / peak,
                     note="BASELINE.json configs[4]: UnitCubeMesh(N) P3, N = 26/33/41/52 for 1/2/4/8 GPUs")
            extra["box3d"] = r
            b3.close()
            del b3
            gc.collect()
        if world > 1 and args.workload == "marmousi" and not strong:
            st, stname = marmousi_problem(world, args.scale, strong=True)
            r = device_value(st, Kx, W, barrier, max_over_ranks, sum_over_ranks)
            r.update(workload=stname, scaling="strong", n_gpus=world,
                     note="the ONE 53.4 M-DoF Marmousi model cut into n_gpus parts (north_star: >= 85 percent at 8 GPUs); "
                          "efficiency = value / (n_gpus * value of the N = 1 run); the per-GPU state is "
                          f"{8e-6 * r['dof_per_gpu']:.0f} MB (the L2 holds 126 MB)")
            extra["strong"] = r
            st.close()
            del st
            gc.collect()
        if world == 1:
            extra["elements"] = element_table(20, peak)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": wname, "degree": degree, "dim": int(d), "cells_per_gpu": int(E),
                          "dof_per_gpu": int(ndof_local), "dof_total": int(ndof), "dt": dt,
                          "material": ("per-cell lambda=mu=Vp^2/3 from the Marmousi grid, rho=1"
                                       if args.workload == "marmousi" else "constant lambda=0.5, mu=0.25, rho=1"),
                          "source": "Ricker, one cell box per tile" if args.workload == "marmousi" else "none",
                          "sponge": "none",
                          "initial_data": "random 1e-3 velocity, random 1e-3 symmetric stress",
                          "stress_storage": "symmetric (upper triangle)" if symmetric else "full",
                          "l2": ("state 8*dof_per_gpu bytes = %.0f MB > 126 MB L2 (no flush needed)"
                                 if 8 * ndof_local > 126e6 else
                                 "state 8*dof_per_gpu bytes = %.0f MB fits the 126 MB L2: not an HBM measurement "
                                 "(strong-scaling / reduced-scale run)") % (8e-6 * ndof_local),
                          "parallelism": f"mesh partition rcb x{world}, one-layer DG halo per pass, exchange={halo_mode}",
                          "setup_s": t_setup},
               "roofline": roofline, "roofline_step": roofline_step, "stages": stages,
               "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "extra": extra}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=env_int("WORLD_SIZE", 1))
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="mesh resolution factor (development aid; 1 = headline)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="marmousi", choices=["marmousi", "box3d"],
                    help="marmousi (default, the headline: BASELINE.json configs[3]); box3d: configs[4], 3D P3 weak scaling")
    ap.add_argument("--extras", default="auto", choices=["auto", "none"],
                    help="auto (default): add the `extra` records (box3d, strong scaling at N > 1, element table at N = 1)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's scaling run): 53.4 M DoF per GPU; strong: 53.4 M DoF in total")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
