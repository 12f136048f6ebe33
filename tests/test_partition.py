"""Graph partitioner (north_star: "partitioned ... by a graph partitioner"; the reference: DMPlex distribution,
seigen/elastic.py:404-414).  METIS on the dual graph vs recursive coordinate bisection on an unstructured (Delaunay)
mesh like the ones tests/tiling ran on (tests/tiling/launchers/executor.sh:99 ``domain$h.msh``): balanced, cut no
worse than RCB's, rank plans consistent."""
import ctypes
import os
import re

import numpy as np
import pytest

from seigen_b200 import partition_metis
from seigen_b200.layout import build_rank_plan, partition_cells
from seigen_b200.mesh import BOUNDARY, Mesh, RectangleMesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not partition_metis.available(), reason="libsg_metis.so not built")


def delaunay_mesh(n=4000, seed=3, lx=3.0, ly=1.0):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.uniform(size=(n, 2)) * [lx, ly]
    # an L-shaped domain with a hole: coordinate bisection has no reason to find good cuts here
    tri = Delaunay(pts)
    c = pts[tri.simplices].mean(axis=1)
    keep = ~((c[:, 0] > 0.5 * lx) & (c[:, 1] > 0.5 * ly)) & (np.hypot(c[:, 0] - 0.25 * lx, c[:, 1] - 0.5 * ly) > 0.15)
    cells = tri.simplices[keep]
    used = np.unique(cells)
    ren = np.full(n, -1)
    ren[used] = np.arange(len(used))
    return Mesh(pts[used], ren[cells].astype(np.int32))


def test_header_symbol_is_exported():
    header = open(os.path.join(ROOT, "include", "seigen_b200_partition.h")).read()
    declared = set(re.findall(r"\b(sg_[a-z_0-9]+)\s*\(", header))
    assert declared == {"sg_partition_graph"}
    lib = ctypes.CDLL(partition_metis.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name)


@pytest.mark.parametrize("k", [2, 4, 8])
def test_metis_cut_and_balance_on_unstructured_mesh(k):
    mesh = delaunay_mesh()
    E = mesh.num_cells()
    p_rcb = partition_cells(mesh, k, "rcb")
    p_met = partition_cells(mesh, k, "metis")
    assert p_met.shape == (E,) and set(np.unique(p_met)) == set(range(k))
    sizes = np.bincount(p_met, minlength=k)
    assert sizes.max() <= 1.03 * E / k + 1                      # METIS' default imbalance tolerance is 3 %
    cut_rcb = partition_metis.edge_cut(mesh.topology, p_rcb)
    cut_met = partition_metis.edge_cut(mesh.topology, p_met)
    assert cut_met <= cut_rcb, (cut_met, cut_rcb)
    # deterministic: every rank computes the partition for itself
    assert np.array_equal(p_met, partition_cells(mesh, k, "metis"))


def test_structured_mesh_cut_close_to_planes():
    mesh = RectangleMesh(64, 32, 2.0, 1.0)
    p = partition_cells(mesh, 4, "metis")
    cut = partition_metis.edge_cut(mesh.topology, p)
    cut_rcb = partition_metis.edge_cut(mesh.topology, partition_cells(mesh, 4, "rcb"))
    assert cut <= 1.5 * cut_rcb                                  # planes are optimal here; METIS must be close


def test_rank_plans_from_metis_partition_are_consistent():
    """What one rank sends is what the other expects, in the same order (the contract sg_peer_connect relies on)."""
    mesh = delaunay_mesh(1500, seed=5)
    k = 3
    mesh.partition_method = "metis"
    part = partition_cells(mesh, k, "metis")
    plans = [build_rank_plan(mesh, part, r, k) for r in range(k)]
    owned = np.concatenate([pl.local_to_global[:pl.n_owned] for pl in plans])
    assert np.array_equal(np.sort(owned), np.arange(mesh.num_cells()))
    for r, pl in enumerate(plans):
        for q, (first, count) in pl.recv.items():
            halo_global = pl.local_to_global[pl.n_owned + first: pl.n_owned + first + count]
            sent_global = plans[q].local_to_global[plans[q].send[r]]
            assert np.array_equal(halo_global, sent_global)
        # every cut-adjacent cell sits in the boundary block, every other owned cell does not
        nb_owner_remote = (pl.nbr >= pl.n_owned) & ((pl.code & BOUNDARY) == 0)
        assert nb_owner_remote[:pl.n_boundary].any(axis=1).all() and not nb_owner_remote[pl.n_boundary:].any()
