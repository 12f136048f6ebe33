"""Host-side cell ordering (seigen_b200/layout.py): the Hilbert grid aligned with the mesh lattice, the resulting tile
quality, and the agreement of every rank on the order of the cells they share."""
import numpy as np
import pytest

from seigen_b200.layout import build_rank_plan, hilbert_key, mesh_lattice, partition_cells
from seigen_b200.mesh import BoxMesh, IntervalMesh, RectangleMesh, perturb_vertices


def in_tile_fraction(plan, tile):
    nb, E = plan.nbr[:plan.n_owned], plan.n_owned
    me = np.arange(E)[:, None]
    interior = nb != me
    same = (nb // tile == me // tile) & interior
    return same.sum() / interior.sum()


def test_mesh_lattice_is_the_spacing_of_a_structured_mesh():
    assert mesh_lattice(RectangleMesh(40, 10, 80.0, 5.0)) == pytest.approx([2.0, 0.5])
    assert mesh_lattice(BoxMesh(8, 4, 2, 4.0, 1.0, 1.0)) == pytest.approx([0.5, 0.25, 0.5])
    assert mesh_lattice(IntervalMesh(10, 4.0)) == pytest.approx([0.4])


def test_tiles_of_a_structured_mesh_are_aligned_blocks():
    """128 consecutive cells of RectangleMesh(64, 32) = one 8 x 8 block of quads (two triangles each): 24 of its 384
    facets leave the tile, i.e. 93.75 % stay inside; the unaligned curve of round 1 kept 87 %."""
    mesh = RectangleMesh(64, 32, 640.0, 320.0)
    plan = build_rank_plan(mesh, partition_cells(mesh, 1), 0, 1)
    cent = mesh.cell_centroids()[plan.local_to_global]
    for t in range(mesh.num_cells() // 128):
        q = np.floor(cent[t * 128:(t + 1) * 128] / 10.0).astype(int)          # quad coordinates of the tile's cells
        assert np.ptp(q[:, 0]) == 7 and np.ptp(q[:, 1]) == 7 and q[:, 0].min() % 8 == 0 and q[:, 1].min() % 8 == 0
    assert in_tile_fraction(plan, 128) > 0.93

    # a grid that is not a power of two (the Marmousi grid is 1532 x 484): the curve over the bounding box cuts blocks
    # that are shifted against the quads, the lattice-aligned one does not
    mesh = RectangleMesh(100, 44, 1000.0, 440.0)
    plan = build_rank_plan(mesh, partition_cells(mesh, 1), 0, 1)
    old = np.argsort(hilbert_key(mesh.cell_centroids()), kind="stable")       # bounding-box grid, no lattice
    pos = np.empty(len(old), dtype=np.int64)
    pos[old] = np.arange(len(old))
    nb = mesh.topology.nbr
    me = np.arange(len(old))[:, None]
    interior = nb != me
    unaligned = ((pos[nb] // 128 == pos[me] // 128) & interior).sum() / interior.sum()
    assert in_tile_fraction(plan, 128) > unaligned + 0.03 and in_tile_fraction(plan, 128) > 0.90


@pytest.mark.parametrize("mesh,tile,least", [(BoxMesh(12, 12, 12, 1.0, 1.0, 1.0), 32, 0.64),
                                             (BoxMesh(16, 8, 8, 4.0, 1.0, 1.0), 64, 0.70),
                                             (perturb_vertices(RectangleMesh(50, 30, 5.0, 3.0), 0.15, 3), 64, 0.80)])
def test_in_tile_fraction_of_other_meshes(mesh, tile, least):
    plan = build_rank_plan(mesh, partition_cells(mesh, 1), 0, 1)
    assert in_tile_fraction(plan, tile) >= least


def test_lattice_key_is_a_function_of_the_point_only():
    """Ranks hold different subsets of the mesh but pass the same bounding box and lattice: equal keys for equal
    centroids, so a rank's cut-adjacent cells and the halo copies its neighbours hold of them sort alike."""
    mesh = RectangleMesh(30, 20, 3.0, 2.0)
    bbox = (mesh.coords.min(axis=0), mesh.coords.max(axis=0))
    h = mesh_lattice(mesh)
    cent = mesh.cell_centroids()
    full = hilbert_key(cent, bbox=bbox, lattice=h)
    sub = np.flatnonzero(cent[:, 0] > 1.3)
    assert np.array_equal(hilbert_key(cent[sub], bbox=bbox, lattice=h), full[sub])
    # degenerate spacings fall back to the bounding-box grid instead of failing
    assert len(hilbert_key(cent, bbox=bbox, lattice=np.array([0.0, 1.0]))) == len(cent)


def test_send_and_halo_orders_match_between_ranks():
    mesh = BoxMesh(6, 5, 4, 1.0, 1.0, 1.0)
    part = partition_cells(mesh, 3)
    plans = [build_rank_plan(mesh, part, r, 3) for r in range(3)]
    for p in plans:
        for q, (first, count) in p.recv.items():
            mine = p.local_to_global[p.n_owned + first:p.n_owned + first + count]
            theirs = plans[q].local_to_global[plans[q].send[p.rank]]
            assert np.array_equal(mine, theirs)


def test_one_dimensional_facade_objects():
    """dimension = 1 (tests/pulse/pulse_1d_lf4.py): one-component vector / tensor spaces take scalar expressions."""
    from seigen_b200 import ElasticLF4, Expression, Function, IntervalMesh as IM
    mesh = IM(40, 4.0)
    el = ElasticLF4.create(mesh, "DG", 2, dimension=1, output=False)
    assert el.u0.dat.data.shape == (40 * 3, 1) and el.s0.dat.data.shape == (40 * 3, 1, 1)
    f = Function(el.U).interpolate(Expression('exp(-50*pow((x[0]-1), 2))'))
    x = el.U.node_coords()[:, 0]
    assert np.allclose(f.dat.data[:, 0], np.exp(-50 * (x - 1) ** 2))
    with pytest.raises(ValueError):
        Function(el.U).interpolate(Expression(('1.0', '2.0')))


def _lanes_per_sector(plan, tile):
    """Out-of-tile facet gathers per distinct (warp, facet slot, 32-byte sector of the neighbour's row, gluing code):
    lanes of a warp that read the same sector through the same facet slot cost one L2 request together."""
    nbr, code = plan.nbr[:plan.n_owned].astype(np.int64), plan.code[:plan.n_owned]
    E, nf = nbr.shape
    me = np.arange(E)
    acc = sec = 0
    for f in range(nf):
        n = nbr[:, f]
        out = (n // tile != me // tile) & (n != me)
        k = np.stack([me[out] // 32, n[out] // 4, code[out, f].astype(np.int64)], axis=1)
        acc += int(out.sum())
        sec += len(np.unique(k, axis=0))
    return acc / sec


@pytest.mark.parametrize("mesh,tile,gain", [(RectangleMesh(200, 96, 1200.0, 576.0), 128, 1.8),
                                            (BoxMesh(16, 16, 16, 1.0, 1.0, 1.0), 32, 1.15),
                                            (BoxMesh(16, 16, 8, 2.0, 2.0, 1.0), 64, 1.25)])
def test_order_within_tiles_shares_sectors_and_keeps_the_tiles(mesh, tile, gain):
    part = partition_cells(mesh, 1)
    a = build_rank_plan(mesh, part, 0, 1)
    b = build_rank_plan(mesh, part, 0, 1, tile=tile)
    E = mesh.num_cells()
    ta = np.empty(E, dtype=np.int64)
    tb = np.empty(E, dtype=np.int64)
    ta[a.local_to_global] = np.arange(E) // tile
    tb[b.local_to_global] = np.arange(E) // tile
    assert np.array_equal(ta, tb)                                    # every cell stays in its tile
    assert in_tile_fraction(a, tile) == in_tile_fraction(b, tile)
    assert _lanes_per_sector(b, tile) > gain * _lanes_per_sector(a, tile)
    # the plan is still a consistent renumbering of the same mesh
    g = b.local_to_global
    assert np.array_equal(g[b.nbr], mesh.topology.nbr[g])
    assert np.array_equal(b.code, mesh.topology.code[g])


def test_order_within_tiles_leaves_the_exchange_plan_alone():
    mesh = RectangleMesh(60, 40, 6.0, 4.0)
    part = partition_cells(mesh, 3)
    for r in range(3):
        a = build_rank_plan(mesh, part, r, 3)
        b = build_rank_plan(mesh, part, r, 3, tile=64)
        assert (a.n_owned, a.n_boundary, a.n_total) == (b.n_owned, b.n_boundary, b.n_total)
        assert np.array_equal(a.local_to_global[:a.n_boundary], b.local_to_global[:b.n_boundary])
        assert np.array_equal(a.local_to_global[a.n_owned:], b.local_to_global[b.n_owned:])
        assert np.array_equal(a.send_cells, b.send_cells) and a.recv == b.recv
        assert np.array_equal(np.sort(a.local_to_global[:a.n_owned]), np.sort(b.local_to_global[:b.n_owned]))
