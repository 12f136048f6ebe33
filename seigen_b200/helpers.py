"""Scalar helpers of ``seigen/helpers.py`` (log :6-12, Vp :15-28, Vs :31-43, cfl_dt :46-54, get_dofs :57-67)."""
from __future__ import annotations

from math import sqrt

__all__ = ["log", "Vp", "Vs", "cfl_dt", "get_dofs"]


def _rank():
    try:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    except Exception:  # pragma: no cover
        return 0


#: where ``log`` writes; bench.py points it at stderr so that its stdout is exactly one JSON line
LOG_STREAM = None


def log(s):
    """Rank-0 print (helpers.py:6-12)."""
    if _rank() == 0:
        print(s, file=LOG_STREAM)


def Vp(mu, l, density):
    """P-wave velocity sqrt((lambda + 2 mu)/rho)."""
    return sqrt((l + 2 * mu) / density)


def Vs(mu, density):
    """S-wave velocity sqrt(mu/rho)."""
    return sqrt(mu / density)


def cfl_dt(dx, Vp, courant_number):
    """Time step permitted by the CFL condition: courant_number*dx/Vp."""
    return (courant_number * dx) / Vp


def get_dofs(mesh, p):
    """Total (stress, velocity) DoF counts over all ranks (helpers.py:57-67, with its missing imports fixed)."""
    from math import comb
    nd = comb(p + mesh.dim, mesh.dim)
    nodes = mesh.num_cells() * nd
    return nodes * mesh.dim * mesh.dim, nodes * mesh.dim
