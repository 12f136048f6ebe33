"""Minimal legacy-VTK writer for ``output=True`` (seigen/elastic.py:120-124, 221-232): cell-vertex values only."""
from __future__ import annotations

import numpy as np


def write_vtk(path, f):
    fs = f.function_space()
    mesh, el = fs.mesh(), fs.elem
    d = mesh.dim
    cells = mesh.cells[fs.cell_order]
    E = len(cells)
    nv = d + 1
    pts = mesh.coords[cells].reshape(-1, d)                    # discontinuous: every cell has its own vertices
    pts3 = np.zeros((len(pts), 3))
    pts3[:, :d] = pts
    vals = f.dat.data.reshape(E, el.nd, -1)[:, :nv, :].reshape(E * nv, -1)   # first d+1 nodes are the vertices
    ctype = {1: 3, 2: 5, 3: 10}[d]
    with open(path, "w") as fh:
        fh.write("# vtk DataFile Version 3.0\nseigen_b200\nASCII\nDATASET UNSTRUCTURED_GRID\n")
        fh.write(f"POINTS {len(pts3)} double\n")
        np.savetxt(fh, pts3, fmt="%.9g")
        fh.write(f"CELLS {E} {E * (nv + 1)}\n")
        conn = np.hstack([np.full((E, 1), nv), np.arange(E * nv).reshape(E, nv)])
        np.savetxt(fh, conn, fmt="%d")
        fh.write(f"CELL_TYPES {E}\n")
        np.savetxt(fh, np.full(E, ctype), fmt="%d")
        fh.write(f"POINT_DATA {len(pts3)}\n")
        name = f.name() or "f"
        for c in range(vals.shape[1]):
            fh.write(f"SCALARS {name}_{c} double 1\nLOOKUP_TABLE default\n")
            np.savetxt(fh, vals[:, c], fmt="%.9g")
