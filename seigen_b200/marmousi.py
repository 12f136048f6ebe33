"""Marmousi P-velocity model (the input data of ``seigen/marmousi.py`` + ``seigen/data/marmhard.dat``).

The reference module only interpolates the 384 x 122 grid (24 m spacing, 1500-5500 m/s) into DG1 and writes a
VTK file (marmousi.py:16-24); it never feeds ``ElasticLF4``.  Here the same grid, with the same index rule
(marmousi.py:9-11, including its ``data[i][-j]`` vertical flip: row j = 0 reads column 0, row j >= 1 reads
column 122 - j), provides per-cell Lame parameters for the heterogeneous benchmark configuration
(rho = 1, mu = lambda = Vp^2/3, sampled at cell centroids)."""
from __future__ import annotations

import os

import numpy as np

__all__ = ["load_vp_grid", "marmousi_vp", "marmousi_lame", "marmousi_lame_at", "LX", "LY"]

LX, LY = 9192.0, 2904.0
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "marmousi_vp.npz")


def load_vp_grid():
    z = np.load(_DATA)
    return z["vp"].astype(np.float64), float(z["spacing"])


def marmousi_vp(points):
    """Vp (m/s) at points (n, 2); x is wrapped with period LX so that tiled copies of the model can be built."""
    data, h = load_vp_grid()
    p = np.asarray(points, dtype=np.float64)
    x = np.mod(p[:, 0], LX)
    i = np.clip(np.floor(x / h).astype(np.int64), 0, data.shape[0] - 1)
    j = np.clip(np.floor(p[:, 1] / h).astype(np.int64), 0, data.shape[1] - 1)
    return data[i, -j]


def marmousi_lame_at(points):
    """(lambda, mu) at points (n, 2): Poisson solid with rho = 1, so that the P speed is Vp."""
    vp = marmousi_vp(points)
    mu = vp * vp / 3.0
    return mu.copy(), mu


def marmousi_lame(mesh):
    """Per-cell (lambda, mu) in the mesh's global cell order, sampled at the cell centroids."""
    return marmousi_lame_at(mesh.cell_centroids())
