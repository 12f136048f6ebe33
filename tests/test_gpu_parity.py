"""GPU parity: the fused sm_100a kernels, called through the C ABI, against the literal CPU oracle
(oracle/elastic_oracle.py = quadrature assembly of seigen/elastic.py:204-219 + block inverse mass).

Tolerances: per stage 1e-12 relative L2 (FP64 round-off only; both sides integrate exactly),
multi-step 1e-10 (the tolerance BASELINE.json states and tests/tiling/explosive_source.py:659 uses)."""
import numpy as np
import pytest

from tests.util import random_state, rel_err, small_mesh

pytestmark = pytest.mark.gpu

CASES = [(1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3)]


def _setup(dim, p, n=None, sponge=False, source=False, per_cell=False, seed=0, symmetric=False, packed=None):
    """symmetric: symmetric s0 and source; packed: symmetric stress storage on the device (default: = symmetric)."""
    from oracle.elastic_oracle import ElasticOracle
    from seigen_b200 import capi
    from seigen_b200.device import DeviceSolver

    mesh = small_mesh(dim, n=n, seed=seed)
    E = mesh.num_cells()
    rng = np.random.default_rng(seed + 10)
    q = {1: 1, 2: 4, 3: 3, 4: 4}[p] if dim == 2 else 1
    orc = ElasticOracle(mesh.coords, mesh.cells, p, sigma_degree=q if sponge else None)
    dev = DeviceSolver(mesh, p, symmetric=symmetric if packed is None else packed)
    if per_cell:
        lam, mu = rng.uniform(0.4, 0.6, E), rng.uniform(0.2, 0.3, E)
    else:
        lam, mu = 0.5, 0.25
    orc.l, orc.mu, orc.density = lam, mu, 1.0
    dev.set_material(1.0, lam, mu)
    if sponge:
        sig = rng.uniform(0.0, 3.0, size=(E, orc.sel.nd))
        sig[rng.uniform(size=E) < 0.5] = 0.0
        orc.sigma = sig
        dev.set_absorption(sig, q)
    nsteps = 8
    if source:
        nd, d = orc.nd, dim
        cells = rng.choice(E, size=3, replace=False)
        sdof = []
        for c in cells:
            for node in range(min(2, nd)):
                for i in range(d):
                    sdof.append(((c * nd + node) * d + i) * d + i)
        sdof = np.array(sdof, dtype=np.int64)
        amp = rng.standard_normal((nsteps, len(sdof)))
        if symmetric and d > 1:            # plus one off-diagonal pair carrying the same values
            base = (cells[0] * nd) * d * d
            sdof = np.concatenate([sdof, [base + 1, base + d]])
            pair = rng.standard_normal((nsteps, 1))
            amp = np.concatenate([amp, pair, pair], axis=1)

        def src_at(step):
            out = np.zeros(E * nd * d * d)
            out[sdof] = amp[step]
            return out.reshape(E, nd, d, d)
        dev.set_source(sdof, amp)
    else:
        src_at = None
    u0, s0 = random_state(mesh, p, seed=seed + 1)
    if symmetric:
        s0 = 0.5 * (s0 + np.swapaxes(s0, 2, 3))
    dev.set_state(u0.reshape(-1, dim), s0.reshape(-1, dim, dim))
    return mesh, orc, dev, u0, s0, src_at, capi


@pytest.mark.parametrize("dim,p", CASES)
@pytest.mark.parametrize("variant", ["plain", "full", "plain-sym", "full-sym"])
def test_stage_by_stage(dim, p, variant):
    """``-sym``: symmetric stress storage on the device (sg_mesh_desc.symmetric_stress), symmetric s0 and source."""
    full = variant.startswith("full")
    mesh, orc, dev, u0, s0, src_at, capi = _setup(dim, p, sponge=full, source=full, per_cell=full,
                                                  symmetric=variant.endswith("-sym"))
    dt = 0.01
    orc.dt = dt
    step = 3
    if src_at is not None:
        orc.source = lambda t: src_at(step)
    u1, s1, st = orc.step(u0, s0, 0.0)
    E, nd, d = mesh.num_cells(), orc.nd, dim
    tol = 1e-12
    dev.stage(1, dt, step)
    assert rel_err(dev.get_field(capi.FIELD_UH).reshape(E, nd, d), st["uh1"]) < tol
    dev.stage(2, dt, step)
    assert rel_err(dev.get_field(capi.FIELD_SH).reshape(E, nd, d, d), st["stemp"]) < tol
    dev.stage(3, dt, step)
    assert rel_err(dev.get_field(capi.FIELD_U).reshape(E, nd, d), u1) < tol
    dev.stage(4, dt, step)
    assert rel_err(dev.get_field(capi.FIELD_SH).reshape(E, nd, d, d), st["sh1"]) < tol
    dev.stage(5, dt, step)
    assert rel_err(dev.get_field(capi.FIELD_UH).reshape(E, nd, d), st["utemp"]) < tol
    dev.stage(6, dt, step)
    assert rel_err(dev.get_field(capi.FIELD_S).reshape(E, nd, d, d), s1) < tol
    dev.close()


@pytest.mark.parametrize("dim,p", CASES)
@pytest.mark.parametrize("symmetric", [False, True])
def test_multi_step_graph(dim, p, symmetric):
    mesh, orc, dev, u0, s0, src_at, capi = _setup(dim, p, sponge=True, source=True, symmetric=symmetric)
    # stable step for the random data: small dt, few steps; the point is the graph-replayed loop + source indexing
    dt = 2e-3
    orc.dt = dt
    nsteps = 8
    u, s = u0, s0
    for n in range(nsteps):
        orc.source = (lambda t, n=n: src_at(n))
        u, s, _ = orc.step(u, s, 0.0)
    dev.step(5, dt, 0)
    dev.step(3, dt, 5)           # second call continues the source table at step 5
    ug, sg = dev.get_state()
    assert rel_err(ug.reshape(u.shape), u) < 1e-10
    assert rel_err(sg.reshape(s.shape), s) < 1e-10
    assert dev.last_step_ms() > 0.0
    dev.close()


def test_set_get_state_roundtrip():
    mesh, orc, dev, u0, s0, _, _ = _setup(2, 2)
    ug, sg = dev.get_state()
    assert np.array_equal(ug.reshape(u0.shape), u0)
    assert np.array_equal(sg.reshape(s0.shape), s0)
    dev.close()


def test_larger_mesh_many_tiles():
    """More cells than one tile, unperturbed structured mesh: exercises cross-tile (global) neighbour reads."""
    mesh, orc, dev, u0, s0, _, capi = _setup(2, 2, n=12)
    orc.dt = dt = 1e-3
    u1, s1, st = orc.step(u0, s0, 0.0)
    dev.step(1, dt, 0)
    ug, sg = dev.get_state()
    assert rel_err(ug.reshape(u1.shape), u1) < 1e-12
    assert rel_err(sg.reshape(s1.shape), s1) < 1e-12
    dev.close()


@pytest.mark.parametrize("dim,p", [(2, 2), (3, 1), (3, 2)])
def test_symmetric_storage_is_bit_identical(dim, p):
    """Packed (upper-triangle) and full stress storage must give the same bits on symmetric data: the packed mode
    only drops loads/stores of values that are equal by construction."""
    out = []
    for packed in (False, True):
        dev = _setup(dim, p, sponge=True, source=True, per_cell=True, symmetric=True, packed=packed)[2]
        dev.step(6, 2e-3, 0)
        out.append(dev.get_state())
        dev.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[1][1], np.swapaxes(out[1][1], 1, 2))


def test_asymmetric_input_is_refused_by_symmetric_storage():
    from seigen_b200 import capi
    mesh, orc, dev, u0, s0, _, _ = _setup(2, 2, symmetric=True)
    bad = s0.copy()
    bad[3, 1, 0, 1] += 1e-9
    with pytest.raises(capi.SgAsymmetric):
        dev.set_state(u0.reshape(-1, 2), bad.reshape(-1, 2, 2))
    dev.set_state(u0.reshape(-1, 2), s0.reshape(-1, 2, 2))        # the solver stays usable
    nd = orc.nd
    sdof = np.array([(5 * nd) * 4 + 1], dtype=np.int64)             # a lone (0,1) entry
    with pytest.raises(capi.SgAsymmetric):
        dev.set_source(sdof, np.ones((4, 1)))
    dev.close()
