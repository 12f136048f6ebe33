"""VTU / PVTU / PVD output for ``output=True`` (seigen/elastic.py:120-124, 221-232, 273, 310).

The reference streams ``File("velocity.pvd")`` / ``File("stress.pvd")``: one ``.vtu`` per ``write`` plus a ``.pvd``
collection (and a ``.pvtu`` per snapshot under MPI).  Here every snapshot carries the FULL P_p field: each cell is
cut into p^d linear sub-simplices whose corners are the cell's nd Lagrange nodes (points are duplicated per cell, as a
discontinuous field needs), so no nodal value is dropped and any VTK reader shows the piecewise-linear interpolant of
the degree-p data on the refined mesh.  Data arrays are inline base64 binary (``header_type="UInt64"``).
"""
from __future__ import annotations

import base64
import itertools
import os
from functools import lru_cache

import numpy as np

__all__ = ["subcells", "write_vtu", "write_pvtu", "write_pvd"]


@lru_cache(maxsize=None)
def _subcells_cached(d, p, lattice_key):
    lattice = np.array(lattice_key, dtype=np.int64)            # (nd, d+1) barycentric multi-indices, sum = p
    index = {tuple(int(x) for x in k[1:]): a for a, k in enumerate(lattice)}     # (k_1..k_d) -> node
    out = []
    if d == 1:
        order = sorted(range(len(lattice)), key=lambda a: lattice[a][1])
        return np.array([[order[i], order[i + 1]] for i in range(p)], dtype=np.int64)
    # Freudenthal / Kuhn triangulation of the ordered simplex p >= y_1 >= y_2 >= ... >= y_d >= 0 on the integer grid,
    # mapped back to lattice coordinates by x_i = y_i - y_{i+1} (x_d = y_d): p^d sub-simplices, all congruent pieces
    # of the reference cell.
    for base in itertools.product(range(p), repeat=d):
        for perm in itertools.permutations(range(d)):
            w = list(base)
            verts = [tuple(w)]
            for ax in perm:
                w[ax] += 1
                verts.append(tuple(w))
            if all(p >= v[0] and all(v[i] >= v[i + 1] for i in range(d - 1)) and v[-1] >= 0 for v in verts):
                nodes = []
                for v in verts:
                    x = tuple(v[i] - (v[i + 1] if i + 1 < d else 0) for i in range(d))
                    nodes.append(index[x])
                out.append(nodes)
    out = np.array(out, dtype=np.int64)
    assert len(out) == p ** d
    return out


def subcells(elem):
    """(p^d, d+1) node numbers (within a cell) of the linear sub-simplices of a degree-p Lagrange cell."""
    return _subcells_cached(elem.dim, elem.degree, tuple(tuple(int(x) for x in k) for k in elem.lattice))


def _b64(a):
    raw = np.ascontiguousarray(a).tobytes()
    return (base64.b64encode(np.uint64(len(raw)).tobytes()) + base64.b64encode(raw)).decode()


def _array(fh, name, a, ncomp=None):
    typ = {"float64": "Float64", "int64": "Int64", "uint8": "UInt8", "int32": "Int32"}[str(a.dtype)]
    nc = f' NumberOfComponents="{ncomp}"' if ncomp is not None else ""
    nm = f' Name="{name}"' if name else ""
    fh.write(f'<DataArray type="{typ}"{nm}{nc} format="binary">{_b64(a)}</DataArray>\n')


def _point_data(f):
    """(name, ncomp, values (npoints, ncomp)) in VTK conventions: vectors padded to 3, tensors to 3x3."""
    fs = f.function_space()
    d = fs.mesh().dim
    vals = np.asarray(f.dat.data)
    n = vals.shape[0]
    name = f.name() or "function"
    if fs.shape == ():
        return name, 1, vals.reshape(n, 1)
    if len(fs.shape) == 1:
        out = np.zeros((n, 3))
        out[:, :fs.shape[0]] = vals
        return name, 3, out
    out = np.zeros((n, 3, 3))
    out[:, :d, :d] = vals.reshape(n, d, d)
    return name, 9, out.reshape(n, 9)


def write_vtu(path, f):
    """One ``.vtu`` piece holding this rank's owned cells of Function ``f`` (all nd nodes of every cell)."""
    fs = f.function_space()
    mesh, el = fs.mesh(), fs.elem
    d = mesh.dim
    E = fs.plan.n_owned
    pts = np.zeros((E * el.nd, 3))
    pts[:, :d] = fs.node_coords()
    sub = subcells(el)                                                    # (nsub, d+1)
    conn = (np.arange(E, dtype=np.int64)[:, None, None] * el.nd + sub[None]).reshape(-1, d + 1)
    ncell = len(conn)
    offsets = (np.arange(1, ncell + 1, dtype=np.int64)) * (d + 1)
    ctype = np.full(ncell, {1: 3, 2: 5, 3: 10}[d], dtype=np.uint8)
    name, nc, vals = _point_data(f)
    with open(path, "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" '
                 'header_type="UInt64">\n<UnstructuredGrid>\n')
        fh.write(f'<Piece NumberOfPoints="{len(pts)}" NumberOfCells="{ncell}">\n<Points>\n')
        _array(fh, None, pts, 3)
        fh.write("</Points>\n<Cells>\n")
        _array(fh, "connectivity", conn.reshape(-1))
        _array(fh, "offsets", offsets)
        _array(fh, "types", ctype)
        kind = {1: "Scalars", 3: "Vectors", 9: "Tensors"}[nc]
        fh.write(f'</Cells>\n<PointData {kind}="{name}">\n')
        _array(fh, name, vals, nc)
        fh.write("</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    return name, nc


def write_pvtu(path, pieces, name, nc):
    with open(path, "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="PUnstructuredGrid" version="1.0" byte_order="LittleEndian" '
                 'header_type="UInt64">\n<PUnstructuredGrid GhostLevel="0">\n')
        fh.write('<PPoints><PDataArray type="Float64" NumberOfComponents="3"/></PPoints>\n')
        kind = {1: "Scalars", 3: "Vectors", 9: "Tensors"}[nc]
        fh.write(f'<PPointData {kind}="{name}"><PDataArray type="Float64" Name="{name}" NumberOfComponents="{nc}"/>'
                 '</PPointData>\n')
        for p in pieces:
            fh.write(f'<Piece Source="{os.path.basename(p)}"/>\n')
        fh.write("</PUnstructuredGrid>\n</VTKFile>\n")


def write_pvd(path, entries):
    """``entries``: [(time, file)] -- the collection ParaView opens (the reference's ``velocity.pvd``)."""
    with open(path, "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1" byte_order="LittleEndian">\n'
                 "<Collection>\n")
        for t, fn in entries:
            fh.write(f'<DataSet timestep="{t:.17g}" part="0" file="{os.path.basename(fn)}"/>\n')
        fh.write("</Collection>\n</VTKFile>\n")


def read_vtu_arrays(path):
    """Decode the arrays of a ``.vtu`` written by ``write_vtu`` (tests and round trips): {name: ndarray}."""
    import re
    text = open(path).read()
    out = {}
    types = {"Float64": np.float64, "Int64": np.int64, "UInt8": np.uint8, "Int32": np.int32}
    for m in re.finditer(r'<DataArray type="(\w+)"(?: Name="([^"]*)")?(?: NumberOfComponents="(\d+)")? '
                         r'format="binary">([^<]*)</DataArray>', text):
        typ, name, nc, payload = m.groups()
        head = base64.b64decode(payload[:12])
        nbytes = int(np.frombuffer(head, dtype=np.uint64)[0])
        raw = base64.b64decode(payload[12:])[:nbytes]
        a = np.frombuffer(raw, dtype=types[typ])
        if nc:
            a = a.reshape(-1, int(nc))
        out[name or "points"] = a
    return out
