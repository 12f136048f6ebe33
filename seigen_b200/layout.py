"""Cell ordering, partitioning and halo plans for the device layout.

PyOP2 iterates cells in DMPlex order and finds facet neighbours through indirection maps; the
stage kernels instead want *tiles* of consecutive cells that are compact in space, so that most
facet neighbours of a tile's cells sit in the same shared-memory tile.  This module

* orders cells along a Hilbert curve through their centroids (``hilbert_order``);
* splits the mesh over ranks (recursive coordinate bisection or METIS, ``partition_cells``) -
  the role DMPlex distribution plays for the reference's MPI runs;
* builds, for one rank, the local numbering ``[cut-adjacent cells | other owned cells | halo cells
  grouped by owner]`` with its send lists (``build_rank_plan``) - the one-layer DG halo PyOP2
  exchanges around every par_loop.

Everything here is host-side NumPy and deterministic: every rank can rebuild every other rank's
plan from the global ``(coords, cells)`` arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .mesh import BOUNDARY, Mesh, Topology

__all__ = ["hilbert_key", "hilbert_order", "mesh_lattice", "partition_cells", "RankPlan", "build_rank_plan"]


def hilbert_key(points: np.ndarray, bits: int | None = None, bbox=None, lattice=None) -> np.ndarray:
    """Hilbert-curve index (uint64) of each point (n, d); Skilling's transpose algorithm, vectorised.  ``bbox`` =
    (lo, hi) fixes the box the curve fills (default: the points' own bounding box); ranks that pass the same box get
    the same key for the same point, whatever subset of the mesh they hold.

    ``lattice`` = per-axis spacing h (d,): the curve's grid is laid over the lattice ``lo + h * Z^d`` (each lattice
    cell = 2^sub grid cells per axis) instead of over the bounding box.  A run of 2^(d*k) lattice cells along the curve
    is then an *aligned* cube of lattice cells, so on a structured mesh (h = its spacing) a tile of cells is a
    square / cubic block of the mesh rather than a block shifted against it: 91 % instead of 87 % of the facets stay
    inside a 128-cell tile of the Marmousi grid, 66 % instead of 60 % inside a 32-cell tile of Kuhn tetrahedra."""
    pts = np.asarray(points, dtype=np.float64)
    n, d = pts.shape
    if bits is None:
        bits = 63 // d if d > 1 else 62
        bits = min(bits, 20)
    if bbox is None:
        lo = pts.min(axis=0)
        span = pts.max(axis=0) - lo
    else:
        lo = np.asarray(bbox[0], dtype=np.float64)
        span = np.asarray(bbox[1], dtype=np.float64) - lo
    span = np.where(span == 0, 1.0, span)
    scale = ((1 << bits) - 1) / span.max()
    if lattice is not None:
        h = np.asarray(lattice, dtype=np.float64)
        if h.shape == (d,) and np.all(h > 0) and np.all(np.isfinite(h)):
            ncell = int(np.ceil((span / h).max())) + 1                    # lattice cells along the longest axis
            lbits = max(1, int(np.ceil(np.log2(ncell))))
            if lbits <= bits:
                scale = float(1 << (bits - lbits)) / h                    # per axis: 2^sub grid cells per lattice cell
    X = np.clip(np.floor((pts - lo) * scale), 0, (1 << bits) - 1).astype(np.uint64).T.copy()      # (d, n)
    if d == 1:
        return X[0]
    M = np.uint64(1 << (bits - 1))
    Q = M
    one = np.uint64(1)
    while Q > one:
        P = Q - one
        for i in range(d):
            hit = (X[i] & Q) != 0
            t = (X[0] ^ X[i]) & P
            X0_new = np.where(hit, X[0] ^ P, X[0] ^ t)
            if i != 0:
                X[i] = np.where(hit, X[i], X[i] ^ t)
            X[0] = X0_new
        Q >>= one
    for i in range(1, d):
        X[i] ^= X[i - 1]
    t = np.zeros(n, dtype=np.uint64)
    Q = M
    while Q > one:
        t = np.where((X[d - 1] & Q) != 0, t ^ (Q - one), t)
        Q >>= one
    for i in range(d):
        X[i] ^= t
    key = np.zeros(n, dtype=np.uint64)
    for b in range(bits - 1, -1, -1):
        for i in range(d):
            key = (key << one) | ((X[i] >> np.uint64(b)) & one)
    return key


def mesh_lattice(mesh: Mesh) -> np.ndarray:
    """Per-axis spacing of the lattice the Hilbert grid is aligned with: the mean extent of a cell's bounding box
    (the spacing itself on the structured utility meshes, a mean cell size per axis otherwise).  Computed from the
    whole mesh, so every rank derives the same value."""
    h = np.zeros(mesh.dim)
    E = mesh.num_cells()
    for k in range(mesh.dim):
        xk = np.ascontiguousarray(mesh.coords[:, k])
        lo = hi = xk[mesh.cells[:, 0]]
        for v in range(1, mesh.cells.shape[1]):
            xv = xk[mesh.cells[:, v]]
            lo = np.minimum(lo, xv)
            hi = np.maximum(hi, xv)
        h[k] = float((hi - lo).sum()) / max(E, 1)
    return h


def hilbert_order(centroids: np.ndarray) -> np.ndarray:
    """Permutation ``order`` such that ``centroids[order]`` walks a Hilbert curve (stable)."""
    return np.argsort(hilbert_key(centroids), kind="stable")


def _rcb(cols, ids, nparts, out, first):
    """Recursive coordinate bisection on ``cols`` = one contiguous coordinate array per axis, restricted to ``ids``
    (ascending).  The split position is found with a selection (O(n)), not a sort; cells whose coordinate ties with
    the split value go left in index order, which is what a stable sort would do (structured meshes have thousands
    of equal centroid coordinates: the cut stays a clean plane)."""
    if nparts == 1:
        out[ids] = first
        return
    left_parts = nparts // 2
    sub = [c[ids] if ids is not None else c for c in cols]
    axis = int(np.argmax([float(x.max() - x.min()) for x in sub]))
    x = sub[axis]
    n = len(x)
    k = int(round(n * left_parts / nparts))
    if k <= 0 or k >= n:
        raise ValueError("more parts than cells")
    v = np.partition(x, k - 1)[k - 1]                 # k-th smallest coordinate
    left = x < v
    ties = np.flatnonzero(x == v)
    left[ties[:k - int(np.count_nonzero(left))]] = True
    idx = np.arange(n) if ids is None else ids
    _rcb(cols, idx[left], left_parts, out, first)
    _rcb(cols, idx[~left], nparts - left_parts, out, first + left_parts)


def partition_cells(mesh: Mesh, nparts: int, method: str = "rcb") -> np.ndarray:
    """Owner rank of every cell.  ``rcb``: recursive coordinate bisection (cut = planes for box meshes)."""
    E = mesh.num_cells()
    part = np.zeros(E, dtype=np.int32)
    if nparts <= 1:
        return part
    if method == "rcb":
        # centroid coordinates, one contiguous array per axis (column reductions over an (E, d) array are slow)
        nv = mesh.cells.shape[1]
        cols = []
        for k in range(mesh.dim):
            xk = np.ascontiguousarray(mesh.coords[:, k])
            acc = xk[mesh.cells[:, 0]]
            for v in range(1, nv):
                acc = acc + xk[mesh.cells[:, v]]
            cols.append(acc / nv)
        _rcb(cols, None, nparts, part, 0)
        return part
    if method == "metis":
        from .partition_metis import metis_partition
        return metis_partition(mesh.topology, nparts)
    raise ValueError("method must be 'rcb' or 'metis'")


@dataclass
class RankPlan:
    """Local numbering of one rank (all arrays index / hold GLOBAL cell ids unless noted)."""
    rank: int
    nranks: int
    local_to_global: np.ndarray            # (n_total,)  owned cells first, then halo
    n_owned: int
    n_boundary: int                        # owned cells [0, n_boundary) touch the cut
    nbr: np.ndarray                        # (n_owned, nf) LOCAL neighbour index
    code: np.ndarray                       # (n_owned, nf)
    jinv: np.ndarray                       # (n_owned, d, d)
    recv: dict = field(default_factory=dict)   # peer -> (first halo slot, count)   [slots relative to n_owned]
    send: dict = field(default_factory=dict)   # peer -> LOCAL owned cell ids, in the peer's halo order
    send_cells: np.ndarray | None = None       # concatenation of ``send`` in ascending peer order
    send_offsets: dict = field(default_factory=dict)   # peer -> (offset, count) into send_cells

    @property
    def n_total(self):
        return len(self.local_to_global)

    @property
    def n_halo(self):
        return self.n_total - self.n_owned


def _order_within_tiles(o_i, n_b, tile, tile_id, topo, cent, bbox, lattice):
    """Permutation of the interior cells ``o_i`` (positions n_b, n_b + 1, ... of the rank's numbering) that keeps every
    cell in its tile but puts the cells with a facet neighbour in ANOTHER tile first, grouped by that tile, by the local
    facet and the gluing code of the facet, and ordered inside a group by the Hilbert key of the point half-way between
    the two centroids -- a point both sides compute alike.  Consecutive lanes of a tile then read consecutive lanes of
    the neighbouring tile through the same facet slot: their 8-byte gathers share 32-byte L2 sectors (2.7 lanes per
    sector instead of 1.3 on the Marmousi grid at TILE 128, 1.4-1.7 instead of 1.15 for Kuhn tetrahedra), and the
    sectors are what the out-of-tile gathers cost."""
    n = len(o_i)
    if n == 0:
        return o_i
    my_tile = (n_b + np.arange(n)) // tile
    nb = topo.nbr[o_i]                                      # (n, nf) sub-mesh indices
    out = (tile_id[nb] != my_tile[:, None]) & (nb != o_i[:, None])
    has = out.any(axis=1)
    first = np.argmax(out, axis=1)                           # lowest local facet that leaves the tile
    rows = np.arange(n)
    other = nb[rows, first]
    grp = np.where(has, tile_id[other], np.iinfo(np.int64).max)
    fcode = np.where(has, topo.code[o_i][rows, first].astype(np.int64), 0)
    fidx = np.where(has, first, 0)
    fkey = np.zeros(n, dtype=np.uint64)
    if has.any():
        mid = 0.5 * (cent[o_i[has]] + cent[other[has]])
        fkey[has] = hilbert_key(mid, bbox=bbox, lattice=lattice)
    order = np.lexsort((rows, fkey, fcode, fidx, grp, my_tile))
    return o_i[order]


def build_rank_plan(mesh: Mesh, part: np.ndarray, rank: int, nranks: int, tile: int | None = None) -> RankPlan:
    """Local numbering of rank ``rank``.  With more than one rank only the sub-mesh made of the owned cells and
    the cells that share a vertex with them (a superset of the facet neighbours) gets its adjacency built, so
    the cost per rank follows the rank's share of the mesh, not the global mesh.  ``tile``: cells per tile of the
    kernels that will run on this numbering (``sg_tile_cells``); the interior cells are then also ordered inside
    their tiles (``_order_within_tiles``) -- an optimisation only, any order is correct."""
    part = np.asarray(part)
    E = mesh.num_cells()
    if nranks > 1:
        vflag = np.zeros(mesh.num_vertices(), dtype=bool)
        vflag[mesh.cells[part == rank].reshape(-1)] = True
        sub = np.flatnonzero(vflag[mesh.cells].any(axis=1))          # ascending global ids
        from .mesh import build_topology
        topo = build_topology(mesh.coords, mesh.cells[sub])
        cent = mesh.coords[mesh.cells[sub]].mean(axis=1)
    else:
        sub = np.arange(E)
        topo = mesh.topology
        cent = mesh.cell_centroids()
    spart = part[sub]
    nf = topo.nbr.shape[1]
    owned = np.flatnonzero(spart == rank)                             # sub-mesh indices from here on
    if len(owned) == 0:
        raise ValueError(f"rank {rank} owns no cells")
    nb = topo.nbr[owned]
    remote = spart[nb] != rank                                        # exterior facets point to self
    is_bnd = remote.any(axis=1)

    # one curve for all ranks (the mesh's bounding box): a rank's cut-adjacent cells and the halo copies its
    # neighbours hold of them are then sorted alike, so a tile's rows land on consecutive halo lanes (coalesced
    # NVLink stores in sg::halo_push)
    key = hilbert_key(cent, bbox=(mesh.coords.min(axis=0), mesh.coords.max(axis=0)), lattice=mesh_lattice(mesh))
    o_b = owned[is_bnd]
    o_i = owned[~is_bnd]
    o_b = o_b[np.argsort(key[o_b], kind="stable")]
    o_i = o_i[np.argsort(key[o_i], kind="stable")]
    if tile:
        # tile of every cell of the sub-mesh in the numbering built so far: cut-adjacent cells, interior cells, and one
        # pseudo-tile per owner for the halo cells
        tile_id = -1 - spart.astype(np.int64)
        tile_id[o_b] = np.arange(len(o_b)) // tile
        tile_id[o_i] = (len(o_b) + np.arange(len(o_i))) // tile
        o_i = _order_within_tiles(o_i, len(o_b), int(tile), tile_id, topo, cent,
                                  (mesh.coords.min(axis=0), mesh.coords.max(axis=0)), mesh_lattice(mesh))
    owned_sorted = np.concatenate([o_b, o_i])

    # halo: remote cells across a facet of an owned cell, grouped by owner, inside a group in the owner's own order of
    # its cut-adjacent cells = (Hilbert key, global id)  (sub-mesh indices ascend with global ids, so every rank
    # derives the same order)
    halo_ids = np.unique(nb[remote])
    halo_owner = spart[halo_ids]
    order = np.lexsort((halo_ids, key[halo_ids], halo_owner))
    halo_ids, halo_owner = halo_ids[order], halo_owner[order]
    recv = {}
    for q in np.unique(halo_owner):
        sel = np.flatnonzero(halo_owner == q)
        recv[int(q)] = (int(sel[0]), int(len(sel)))

    l2s = np.concatenate([owned_sorted, halo_ids]).astype(np.int64)
    s2l = np.full(len(sub), -1, dtype=np.int64)
    s2l[l2s] = np.arange(len(l2s))

    nbr_local = s2l[topo.nbr[owned_sorted]]
    assert (nbr_local >= 0).all()
    code = topo.code[owned_sorted].copy()
    jinv = np.ascontiguousarray(topo.jinv[owned_sorted])

    # send lists: my cells that peer q sees across its facets = my cut-adjacent cells with a neighbour owned by q,
    # in (Hilbert key, global id) order = my own local order = the order q's halo group uses
    send, send_offsets, chunks, off = {}, {}, [], 0
    pb = spart[topo.nbr[o_b]]
    for q in sorted(recv):
        mine = o_b[(pb == q).any(axis=1)]                              # o_b is sorted by (key, global id) already
        loc = s2l[mine]
        send[q] = loc
        send_offsets[q] = (off, len(loc))
        chunks.append(loc)
        off += len(loc)
    send_cells = np.concatenate(chunks).astype(np.int64) if chunks else np.zeros(0, dtype=np.int64)

    return RankPlan(rank=rank, nranks=nranks, local_to_global=sub[l2s].astype(np.int64), n_owned=len(owned_sorted),
                    n_boundary=len(o_b), nbr=np.ascontiguousarray(nbr_local.astype(np.int32)),
                    code=np.ascontiguousarray(code.astype(np.uint8)), jinv=jinv, recv=recv, send=send,
                    send_cells=send_cells, send_offsets=send_offsets)
