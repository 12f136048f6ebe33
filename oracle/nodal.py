"""Second CPU oracle  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy statement of the quadrature-free nodal operator of SURVEY.md Appendix A, i.e. the
*algorithm the CUDA kernels implement*, driven by the tables the product's host code builds
(``RefElem`` matrices, facet adjacency, ``Jinv``).  It is passed those tables as plain arrays
(it does not import ``seigen_b200``), so that

* ``literal forms (elastic_oracle.py) == nodal operator`` proves the reformulation
  (``elastic.py:204-219`` followed by the inverse mass of ``:358-367, 376-381``), and
* the product's host-side tables are exercised on a machine without a GPU (``-m "not gpu"``
  tests), including the multi-rank halo logic under ``gloo``.

Same pinning status as ``elastic_oracle.py``: parity unpinned at the Firedrake boundary.
"""
from __future__ import annotations

import numpy as np

BOUNDARY = 0x80


class NodalOperator:
    def __init__(self, Dr, Lift, fnodes, ftab, nbr, code, jinv, n_owned=None):
        self.Dr = np.asarray(Dr)                    # (d, nd, nd)
        self.Lift = np.asarray(Lift)                # (nf, nd, nfp)
        self.fnodes = np.asarray(fnodes)            # (nf, nfp)
        self.ftab = np.asarray(ftab).reshape(-1, self.fnodes.shape[1])   # (nf*nperm, nfp)
        self.nbr = np.asarray(nbr)                  # (E, nf)  indices into the field's cell axis
        self.code = np.asarray(code)
        self.jinv = np.asarray(jinv)                # (E, d, d)
        self.E = self.nbr.shape[0] if n_owned is None else n_owned
        self.d = self.Dr.shape[0]
        self.nf = self.d + 1

    def ref_gradient(self, phi, boundary_coef):
        """R[e, r, a, ...] for nodal fields phi[e_total, a, ...] (first E cells are evaluated).

        jump = (phi_nbr - phi_own)/2 + boundary_coef * phi_own on exterior facets
        (boundary_coef = -1: numerical trace 0, the free surface of ``f``; 0: own trace, as in ``g``).
        """
        E, d = self.E, self.d
        own = phi[:E]
        R = np.einsum("rab,eb...->era...", self.Dr, own)
        ar = np.arange(E)
        for f in range(self.nf):
            c = self.code[:E, f]
            bnd = (c & BOUNDARY) != 0
            nodes_n = self.ftab[c & 0x7F]                                  # (E, nfp)
            vn = phi[self.nbr[:E, f][:, None], nodes_n]                   # (E, nfp, ...)
            vo = own[:, self.fnodes[f]]
            jump = 0.5 * (vn - vo)
            if boundary_coef != 0.0:
                jump[bnd] += boundary_coef * vo[bnd]
            lift = np.einsum("am,em...->ea...", self.Lift[f], jump)
            if f == 0:
                R += lift[:, None]
            else:
                R[:, f - 1] -= lift
        del ar
        return R

    def Dv(self, s):
        """(E_total, nd, d, d) -> (E, nd, d):  sum_j d~_j s_ij  with free-surface boundary trace."""
        R = self.ref_gradient(s, -1.0)                                     # (E, r, a, i, j)
        return np.einsum("erj,eraij->eai", self.jinv[:self.E], R)

    def Ds(self, u, lam, mu):
        """(E_total, nd, d) -> (E, nd, d, d):  lam*div*I + mu*(G + G^T),  G_ij = d~_j u_i."""
        R = self.ref_gradient(u, 0.0)                                      # (E, r, a, i)
        G = np.einsum("erj,erai->eaij", self.jinv[:self.E], R)
        lam = np.broadcast_to(np.asarray(lam, dtype=float), (self.E,))
        mu = np.broadcast_to(np.asarray(mu, dtype=float), (self.E,))
        div = np.einsum("eaii->ea", G)
        out = mu[:, None, None, None] * (G + np.swapaxes(G, 2, 3))
        out += (lam[:, None] * div)[:, :, None, None] * np.eye(self.d)[None, None]
        return out

    @staticmethod
    def absorb(W, sigma, u):
        """P(sigma, u)[e, a, i] = sum_{b,c} sigma[e, b] W[a, b, c] u[e, c, i]."""
        A = np.einsum("abc,eb->eac", W, sigma)
        return np.einsum("eac,eci->eai", A, u)
