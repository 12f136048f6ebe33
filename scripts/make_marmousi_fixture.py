"""Derive seigen_b200/data/marmousi_vp.npz from the reference's input data file (run in the build container only).

The reference ships the Marmousi P-velocity grid as 46 848 text lines (seigen/data/marmhard.dat, read by
seigen/marmousi.py:4-5 as a (384, 122) array, 24 m spacing).  The values are whole m/s, so the grid is stored
losslessly as int16 (94 KB -> ~60 KB compressed)."""
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/seigen/data/marmhard.dat"
vals = np.loadtxt(src).reshape(-1)[:384 * 122].reshape(384, 122)
assert np.all(vals == np.round(vals)) and vals.min() >= 0 and vals.max() < 32767
np.savez_compressed("seigen_b200/data/marmousi_vp.npz", vp=vals.astype(np.int16), spacing=np.float64(24.0))
print("wrote", vals.shape, vals.min(), vals.max())
