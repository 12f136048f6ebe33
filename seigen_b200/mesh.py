"""Simplicial meshes, facet adjacency and affine geometry for the DG stage kernels.

The mesh constructors reproduce the vertex/cell enumeration of the Firedrake
utility meshes the reference scripts call (``UnitSquareMesh`` /
``RectangleMesh`` in ``tests/eigenmode/eigenmode_2d.py:11`` and
``tests/explosive_source/explosive_source_lf4.py:10``, ``UnitCubeMesh`` in
``tests/eigenmode/eigenmode_3d.py:11``, ``IntervalMesh`` in
``tests/pulse/pulse_1d_lf4.py:10``) *before* DMPlex renumbering, which a DG
solver never observes except through the order of ``dat.data``.

What the kernels consume (``Topology``):

* ``nbr[E, nf]``   int32  face-neighbour cell (own index on a boundary facet)
* ``code[E, nf]``  uint8  ``f' * d! + s`` = neighbour's facet number and the vertex
  permutation that glues it to mine (index into ``RefElem.ftab``);
  bit 7 (``BOUNDARY``) marks an exterior facet
* ``jinv[E, d, d]`` float64  ``Jinv[r, k] = d(xi_r)/d(x_k)``

This replaces PyOP2's ``cell_node_map`` / ``interior_facet_node_map`` /
``exterior_facet_node_map`` indirections by one
facet-to-cell adjacency.
"""
from __future__ import annotations

import math

import numpy as np

from .refelem import RefElem, facet_perms, get_refelem

__all__ = ["Mesh", "Topology", "IntervalMesh", "UnitIntervalMesh", "RectangleMesh", "UnitSquareMesh",
           "BoxMesh", "UnitCubeMesh", "BOUNDARY", "build_topology", "perturb_vertices", "read_gmsh", "write_gmsh"]

BOUNDARY = 0x80


class _Comm:
    """Stand-in for ``mesh.comm`` (``seigen/elastic.py:85``): sums over torch.distributed ranks if initialised."""

    @property
    def rank(self):
        try:
            import torch.distributed as dist
            return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        except Exception:  # pragma: no cover
            return 0

    @property
    def size(self):
        try:
            import torch.distributed as dist
            return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        except Exception:  # pragma: no cover
            return 1

    def allreduce(self, value, op=None):
        return value


class Mesh:
    """A conforming simplicial mesh given by explicit ``(coords, cells)`` arrays, or read from a Gmsh ``.msh`` file
    (``Mesh(mesh_file)`` as in tests/tiling/explosive_source.py:531-532, for meshes made from
    tests/explosive_source/src/domain.geo)."""

    def __init__(self, coords, cells=None, name="mesh", dim=None):
        #: how the cells are split over ranks (layout.partition_cells): "rcb" (coordinate bisection: planes, optimal
        #: for the structured utility meshes) or "metis" (graph partitioner on the dual graph; the default for
        #: unstructured meshes read from a file when libsg_metis.so is built).  SG_PARTITION overrides both.
        self.partition_method = "rcb"
        if isinstance(coords, (str, bytes)) or hasattr(coords, "__fspath__"):
            name = str(coords)
            coords, cells = read_gmsh(coords, dim=dim)
            from . import partition_metis
            if partition_metis.available():
                self.partition_method = "metis"
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords[:, None]
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        if cells.shape[1] != coords.shape[1] + 1:
            raise ValueError("cells must have dim+1 vertices")
        self.coords = coords
        self.cells = cells
        self.dim = coords.shape[1]
        self.name = name
        self.comm = _Comm()
        self._topology = None

    # Firedrake-flavoured accessors the reference scripts touch
    def geometric_dimension(self):
        return self.dim

    def num_cells(self):
        return self.cells.shape[0]

    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def topology(self):
        if self._topology is None:
            self._topology = build_topology(self.coords, self.cells)
        return self._topology

    def init(self, *args, **kwargs):  # mesh.init() / mesh.topology.init(s_depth=..) are no-ops here
        return None

    def node_coords(self, elem: RefElem) -> np.ndarray:
        """Physical coordinates of every DG node: (E, nd, d)."""
        v = self.coords[self.cells]                       # (E, d+1, d)
        lam = elem.lattice.astype(np.float64) / elem.degree  # (nd, d+1) barycentric
        return np.einsum("av,evk->eak", lam, v)

    def cell_centroids(self) -> np.ndarray:
        return self.coords[self.cells].mean(axis=1)

    def min_cell_size(self) -> float:
        v = self.coords[self.cells]
        h = np.inf
        for a in range(self.dim + 1):
            for b in range(a + 1, self.dim + 1):
                h = min(h, float(np.sqrt(((v[:, a] - v[:, b]) ** 2).sum(-1)).min()))
        return h


class Topology:
    """Facet adjacency + affine geometry (see module docstring)."""

    def __init__(self, nbr, code, jinv, detj):
        self.nbr = nbr
        self.code = code
        self.jinv = jinv
        self.detj = detj

    def init(self, *args, **kwargs):
        return None

    @property
    def num_cells(self):
        return self.nbr.shape[0]

    def interior_facets(self):
        """(e, f, e', f') once per interior facet (e < e')."""
        E, nf = self.nbr.shape
        e = np.repeat(np.arange(E), nf)
        f = np.tile(np.arange(nf), E)
        n = self.nbr.reshape(-1)
        c = self.code.reshape(-1)
        keep = ((c & BOUNDARY) == 0) & (e < n)
        d = nf - 1
        return e[keep], f[keep], n[keep], (c[keep] // math.factorial(d))

    def exterior_facets(self):
        E, nf = self.nbr.shape
        e = np.repeat(np.arange(E), nf)
        f = np.tile(np.arange(nf), E)
        keep = (self.code.reshape(-1) & BOUNDARY) != 0
        return e[keep], f[keep]


def build_topology(coords: np.ndarray, cells: np.ndarray) -> Topology:
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells)
    E, nv = cells.shape
    d = nv - 1
    nf = nv
    perms = facet_perms(d)
    nperm = len(perms)
    # facet vertex tuples in own local order (facet f drops local vertex f)
    fverts = np.array([[v for v in range(nv) if v != f] for f in range(nf)])        # (nf, d)
    fg = cells[:, fverts].astype(np.int64)                                          # (E, nf, d) global ids
    flat = fg.reshape(E * nf, d)
    key = np.sort(flat, axis=1)
    order = np.lexsort(key.T[::-1])
    ks = key[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    # facets shared by more than two cells would be a non-manifold mesh
    if np.any(same[1:] & same[:-1]):
        raise ValueError("non-manifold mesh: a facet is shared by more than two cells")
    a = order[:-1][same]
    b = order[1:][same]

    nbr = np.repeat(np.arange(E, dtype=np.int32)[:, None], nf, axis=1)
    identity = perms.index(tuple(range(d)))
    code = (np.arange(nf, dtype=np.int64)[None, :] * nperm + identity).astype(np.uint8)
    code = np.repeat(code, E, axis=0) | np.uint8(BOUNDARY)

    # perm lookup: sigma encoded base d
    enc = {sum(s[j] * d ** j for j in range(d)): i for i, s in enumerate(perms)}
    lut = np.full(d ** d if d > 0 else 1, -1, dtype=np.int64)
    for k, v in enc.items():
        lut[k] = v

    def glue(me, other):
        """code for `me` looking at `other`: sigma[j] = position of my j-th facet vertex in other's tuple."""
        gm = flat[me]            # (n, d)
        go = flat[other]
        sigma = np.argmax(gm[:, :, None] == go[:, None, :], axis=2)   # (n, d)
        e_code = (sigma * (d ** np.arange(d))[None, :]).sum(axis=1)
        s = lut[e_code]
        assert (s >= 0).all()
        return ((other % nf) * nperm + s).astype(np.uint8)

    ea, fa = a // nf, a % nf
    eb, fb = b // nf, b % nf
    nbr[ea, fa] = eb
    nbr[eb, fb] = ea
    code[ea, fa] = glue(a, b)
    code[eb, fb] = glue(b, a)

    v = coords[cells]                                    # (E, d+1, d)
    J = np.swapaxes(v[:, 1:, :] - v[:, :1, :], 1, 2)     # J[e, k, r] = (v_{r+1} - v_0)_k
    detj = np.linalg.det(J)
    if np.any(np.abs(detj) < 1e-300):
        raise ValueError("degenerate cell (zero volume)")
    jinv = np.linalg.inv(J)                              # jinv[e, r, k]
    return Topology(nbr, code, np.ascontiguousarray(jinv), detj)


def read_gmsh(path, dim=None):
    """(coords, cells) of the highest-dimensional simplices in a Gmsh ASCII mesh (format 2.2 or 4.1).

    Lower-dimensional elements (the ``Physical Line`` tags of domain.geo) are ignored: every boundary is a free
    surface in ElasticLF4.  For planar meshes the constant z coordinate is dropped."""
    with open(path) as fh:
        lines = [ln.strip() for ln in fh]
    sect = {}
    i = 0
    while i < len(lines):
        if lines[i].startswith("$") and not lines[i].startswith("$End"):
            name = lines[i][1:]
            j = i + 1
            while not lines[j].startswith("$End"):
                j += 1
            sect[name] = lines[i + 1:j]
            i = j
        i += 1
    if "MeshFormat" not in sect:
        raise ValueError(f"{path}: not a Gmsh ASCII mesh")
    version, ftype = sect["MeshFormat"][0].split()[:2]
    if ftype != "0":
        raise ValueError(f"{path}: binary Gmsh files are not supported")
    simplex = {2: 3, 4: 4}                                  # Gmsh element type -> vertices (3-node triangle, 4-node tet)
    tags, pts, elems = [], [], {2: [], 4: []}
    if version.startswith("2"):
        n = int(sect["Nodes"][0])
        for ln in sect["Nodes"][1:n + 1]:
            t = ln.split()
            tags.append(int(t[0]))
            pts.append([float(x) for x in t[1:4]])
        m = int(sect["Elements"][0])
        for ln in sect["Elements"][1:m + 1]:
            t = ln.split()
            et, ntag = int(t[1]), int(t[2])
            if et in simplex:
                elems[et].append([int(x) for x in t[3 + ntag:3 + ntag + simplex[et]]])
    elif version.startswith("4"):
        body = sect["Nodes"]
        nblocks = int(body[0].split()[0])
        k = 1
        for _ in range(nblocks):
            nn = int(body[k].split()[3])
            k += 1
            tags += [int(x) for x in body[k:k + nn]]
            pts += [[float(x) for x in ln.split()[:3]] for ln in body[k + nn:k + 2 * nn]]
            k += 2 * nn
        body = sect["Elements"]
        nblocks = int(body[0].split()[0])
        k = 1
        for _ in range(nblocks):
            _, _, et, ne = (int(x) for x in body[k].split())
            k += 1
            if et in simplex:
                elems[et] += [[int(x) for x in ln.split()[1:1 + simplex[et]]] for ln in body[k:k + ne]]
            k += ne
    else:
        raise ValueError(f"{path}: unsupported Gmsh format version {version}")
    et = 4 if (elems[4] and dim != 2) else 2
    if not elems[et]:
        raise ValueError(f"{path}: no triangles or tetrahedra found")
    pts = np.array(pts, dtype=np.float64)
    lut = np.full(max(tags) + 1, -1, dtype=np.int64)
    lut[np.array(tags)] = np.arange(len(tags))
    cells = lut[np.array(elems[et], dtype=np.int64)]
    used = np.unique(cells)                                  # drop nodes no cell refers to
    renum = np.full(len(pts), -1, dtype=np.int64)
    renum[used] = np.arange(len(used))
    d = 3 if et == 4 else 2
    return pts[used][:, :d], renum[cells].astype(np.int32)


def write_gmsh(path, coords, cells):
    """Gmsh 2.2 ASCII writer (triangles / tetrahedra); the counterpart of read_gmsh for fixtures and exports."""
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells)
    et = {3: 2, 4: 4}[cells.shape[1]]
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(coords))
        for i, x in enumerate(coords):
            xyz = list(x) + [0.0] * (3 - len(x))
            fh.write("%d %.17g %.17g %.17g\n" % (i + 1, *xyz))
        fh.write("$EndNodes\n$Elements\n%d\n" % len(cells))
        for i, c in enumerate(cells):
            fh.write("%d %d 2 10 6 %s\n" % (i + 1, et, " ".join(str(int(v) + 1) for v in c)))
        fh.write("$EndElements\n")


# ----------------------------------------------------------------------------
# utility meshes (vertex / cell enumeration as in Firedrake's utility_meshes, before DMPlex reordering)
# ----------------------------------------------------------------------------
def IntervalMesh(ncells, length_or_left, right=None):
    if right is None:
        left, right = 0.0, float(length_or_left)
    else:
        left = float(length_or_left)
    ncells = int(ncells)
    x = np.linspace(left, right, ncells + 1)
    cells = np.stack([np.arange(ncells), np.arange(1, ncells + 1)], axis=1)
    return Mesh(x[:, None], cells, name="interval")


def UnitIntervalMesh(ncells):
    return IntervalMesh(ncells, 1.0)


def RectangleMesh(nx, ny, Lx, Ly, quadrilateral=False, reorder=None, diagonal="left"):
    if quadrilateral:
        raise NotImplementedError("only simplicial meshes are supported")
    nx, ny = int(nx), int(ny)
    xs = np.linspace(0.0, Lx, nx + 1)
    ys = np.linspace(0.0, Ly, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)        # vertex id = i*(ny+1) + j
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    i, j = i.reshape(-1), j.reshape(-1)
    v0 = i * (ny + 1) + j
    v1 = v0 + 1
    v2 = (i + 1) * (ny + 1) + j + 1
    v3 = (i + 1) * (ny + 1) + j
    quad = np.stack([v0, v1, v2, v3], axis=1)
    if diagonal == "left":
        idx = [0, 1, 3, 1, 2, 3]
    elif diagonal == "right":
        idx = [0, 1, 2, 0, 2, 3]
    else:
        raise ValueError("diagonal must be 'left' or 'right'")
    cells = quad[:, idx].reshape(-1, 3)
    return Mesh(coords, cells, name="rectangle")


def UnitSquareMesh(nx, ny, **kw):
    return RectangleMesh(nx, ny, 1.0, 1.0, **kw)


def BoxMesh(nx, ny, nz, Lx, Ly, Lz, reorder=None):
    nx, ny, nz = int(nx), int(ny), int(nz)
    xs = np.linspace(0.0, Lx, nx + 1)
    ys = np.linspace(0.0, Ly, ny + 1)
    zs = np.linspace(0.0, Lz, nz + 1)
    # vertex id = k*(nx+1)*(ny+1) + j*(nx+1) + i
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.reshape(-1), j.reshape(-1), k.reshape(-1)
    v0 = k * (nx + 1) * (ny + 1) + j * (nx + 1) + i
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v1 + (nx + 1)
    v4 = v0 + (nx + 1) * (ny + 1)
    v5 = v1 + (nx + 1) * (ny + 1)
    v6 = v2 + (nx + 1) * (ny + 1)
    v7 = v3 + (nx + 1) * (ny + 1)
    cube = np.stack([v0, v1, v2, v3, v4, v5, v6, v7], axis=1)
    # six tetrahedra around the main diagonal v0-v7
    idx = [0, 1, 3, 7, 0, 1, 7, 5, 0, 5, 7, 4, 0, 3, 2, 7, 0, 6, 4, 7, 0, 2, 6, 7]
    cells = cube[:, idx].reshape(-1, 4)
    return Mesh(coords, cells, name="box")


def UnitCubeMesh(nx, ny, nz, **kw):
    return BoxMesh(nx, ny, nz, 1.0, 1.0, 1.0, **kw)


def perturb_vertices(mesh: Mesh, amplitude: float, seed: int = 0) -> Mesh:
    """Randomly displaced copy (fraction of the shortest edge) - used by parity tests to break symmetry."""
    rng = np.random.default_rng(seed)
    h = mesh.min_cell_size()
    coords = mesh.coords + amplitude * h * rng.uniform(-1.0, 1.0, size=mesh.coords.shape)
    return Mesh(coords, mesh.cells.copy(), name=mesh.name + "_perturbed")
