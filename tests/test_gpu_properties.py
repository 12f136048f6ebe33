"""Size-independent properties at BASELINE.json's full sizes (the oracle cannot run these in seconds):

* exact linearity under power-of-two scaling -- every operation of the update is linear in (u, s, source), and scaling
  by 2^k commutes with FP64 rounding, so run(2^k x0) == 2^k run(x0) bit for bit;
* symmetric (packed) stress storage == full storage, bit for bit, and the stress stays exactly symmetric;
* a zero state with no source stays zero;
* superposition of two sources to round-off.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solver(mesh, p, symmetric, lam=0.5, mu=0.25):
    from seigen_b200.device import DeviceSolver
    dev = DeviceSolver(mesh, p, symmetric=symmetric)
    dev.set_material(1.0, lam, mu)
    return dev


def _state(mesh, p, seed=0):
    from seigen_b200.refelem import get_refelem
    nd, d = get_refelem(mesh.dim, p).nd, mesh.dim
    rng = np.random.default_rng(seed)
    u = rng.standard_normal((mesh.num_cells() * nd, d))
    s = rng.standard_normal((mesh.num_cells() * nd, d, d))
    return u, 0.5 * (s + np.swapaxes(s, 1, 2))


def _full_size_mesh(name):
    from seigen_b200.mesh import BoxMesh, RectangleMesh
    if name == "marmousi":                       # configs[3]: RectangleMesh(1532, 484), P2 -> 53.4 M DoF
        return RectangleMesh(1532, 484, 9192.0, 2904.0), 2, 0.5 * 6.0 / (2 * 5500.0) * 100.0
    if name == "explosive":                      # configs[1]: RectangleMesh(168, 84, 300, 150), P2 -> 1.0 M DoF
        return RectangleMesh(168, 84, 300.0, 150.0), 2, 1e-3
    p = int(name[-1])                            # configs[2]: (64,16,16) cubes x 6 tets, P1-P3
    return BoxMesh(64, 16, 16, 4.0, 1.0, 1.0), p, 0.5 * (1.0 / 16) / 2 ** (p - 1)


@pytest.mark.parametrize("name", ["explosive", "marmousi", "pulse3d_p1", "pulse3d_p2", "pulse3d_p3"])
def test_full_size_linearity_and_storage_equivalence(name):
    mesh, p, dt = _full_size_mesh(name)
    u0, s0 = _state(mesh, p)
    nsteps = 4
    out = {}
    for key, symmetric, scale in (("packed", True, 1.0), ("full", False, 1.0), ("scaled", True, 2.0 ** 7)):
        dev = _solver(mesh, p, symmetric)
        dev.set_state(scale * u0, scale * s0)
        dev.step(nsteps, dt)
        out[key] = dev.get_state()
        dev.close()
    u, s = out["packed"]
    assert np.isfinite(u).all() and np.isfinite(s).all() and np.abs(u).max() > 0
    assert np.array_equal(s, np.swapaxes(s, 1, 2))                       # stays exactly symmetric
    assert np.array_equal(u, out["full"][0]) and np.array_equal(s, out["full"][1])
    assert np.array_equal(out["scaled"][0], 2.0 ** 7 * u) and np.array_equal(out["scaled"][1], 2.0 ** 7 * s)


def test_zero_stays_zero_and_source_superposition():
    from seigen_b200.mesh import RectangleMesh
    from seigen_b200.refelem import get_refelem
    mesh = RectangleMesh(168, 84, 300.0, 150.0)
    p, d, dt, nsteps = 2, 2, 1e-3, 6
    nd = get_refelem(2, p).nd
    E = mesh.num_cells()
    rng = np.random.default_rng(3)

    def run(sdof, amp):
        dev = _solver(mesh, p, True, lam=3599.3664, mu=3600.0)
        dev.set_state(np.zeros((E * nd, d)), np.zeros((E * nd, d, d)))
        if sdof is not None:
            dev.set_source(sdof, amp)
        dev.step(nsteps, dt)
        res = dev.get_state()
        dev.close()
        return res

    u, s = run(None, None)
    assert not u.any() and not s.any()
    cells = rng.choice(E, size=4, replace=False)
    diag = lambda c, node, i: ((c * nd + node) * d + i) * d + i          # noqa: E731
    sa = np.array([diag(c, 0, i) for c in cells[:2] for i in range(d)], dtype=np.int64)
    sb = np.array([diag(c, 1, i) for c in cells[2:] for i in range(d)], dtype=np.int64)
    aa, ab = rng.standard_normal((nsteps, len(sa))), rng.standard_normal((nsteps, len(sb)))
    ua, sa_ = run(sa, aa)
    ub, sb_ = run(sb, ab)
    uc, sc = run(np.concatenate([sa, sb]), np.concatenate([aa, ab], axis=1))
    assert np.abs(uc).max() > 0
    assert np.abs(uc - (ua + ub)).max() <= 1e-13 * np.abs(uc).max()
    assert np.abs(sc - (sa_ + sb_)).max() <= 1e-13 * np.abs(sc).max()
