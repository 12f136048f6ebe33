"""Reference-element tables for equispaced discontinuous Lagrange P_p on simplices.

Everything the fused stage kernels need about the reference cell is built here
once, in exact rational arithmetic (``fractions.Fraction``), and only then
rounded to float64: structural zeros of the derivative and lift matrices are
exact zeros, which lets the CUDA compiler delete those multiplies when the
tables are baked into ``csrc/tables_gen.h``.

The space is the one ``seigen/elastic.py:81-82`` asks Firedrake for
(``TensorFunctionSpace/VectorFunctionSpace(mesh, "DG", degree)``): every
component lives in the same scalar equispaced Lagrange space, nodes in FIAT's
entity order (vertices, edges, faces, interior).

Tables (d = dim, p = degree, nd nodes per cell, nfp nodes per facet, nf = d+1):

* ``nodes``    (nd, d)        reference coordinates, UFC simplex (0, e_1..e_d)
* ``lattice``  (nd, d+1)      integer barycentric multi-index (sum = p)
* ``M``, ``Minv`` (nd, nd)    reference mass matrix and inverse
* ``Dr``       (d, nd, nd)    strong derivative  Dr[r, a, b] = d(phi_b)/d(xi_r)(x_a)
* ``fnodes``   (nf, nfp)      nodes on facet f (facet f is opposite vertex f),
                              listed in the lattice order of the facet's own
                              vertices taken in increasing local index
* ``FM``       (nfp, nfp)     (d-1)-dimensional reference facet mass
* ``Lift``     (nf, nd, nfp)  Minv[:, fnodes[f]] @ FM
* ``ftab``     (nf, d!, nfp)  neighbour-side node matching: if the neighbour's
                              facet f' has its j-th vertex glued to my facet's
                              sigma[j]-th vertex ... see ``facet_match_table``.

With these, the central-flux DG derivative of SURVEY.md Appendix A is

    R_r = Dr[r] @ phi - Lift[r+1] @ jump_{r+1} + Lift[0] @ jump_0
    d~_k phi = sum_r Jinv[r, k] * R_r

because the physical facet mass is |F_f| (d-1)! FM, the cell mass |detJ| M, and
(d-1)! * |F_f| / |detJ| * n_f = -grad(lambda_f)  on any affine simplex, with
grad(lambda_{r+1}) = Jinv[r, :]  and  grad(lambda_0) = -sum_r Jinv[r, :].
"""
from __future__ import annotations

import itertools
import math
from fractions import Fraction
from functools import lru_cache

import numpy as np

__all__ = ["RefElem", "get_refelem", "facet_perms"]


# ----------------------------------------------------------------------------
# exact helpers
# ----------------------------------------------------------------------------
def _finv(A):
    """Exact inverse of a square Fraction matrix (Gauss-Jordan)."""
    n = len(A)
    M = [list(row) + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(A)]
    for c in range(n):
        piv = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[piv] = M[piv], M[c]
        inv = 1 / M[c][c]
        M[c] = [x * inv for x in M[c]]
        for r in range(n):
            if r != c and M[r][c] != 0:
                f = M[r][c]
                M[r] = [x - f * y for x, y in zip(M[r], M[c])]
    return [row[n:] for row in M]


def _fmatmul(A, B):
    Bt = list(zip(*B))
    return [[sum((a * b for a, b in zip(row, col) if a != 0 and b != 0), Fraction(0)) for col in Bt]
            for row in A]


def _monomials(d, p):
    """Exponent tuples of total degree <= p in d variables."""
    return [e for e in itertools.product(range(p + 1), repeat=d) if sum(e) <= p]


def _simplex_monomial_integral(expo):
    """int over the unit simplex of prod x_k^expo_k  =  prod(expo_k!) / (sum(expo) + d)!"""
    d = len(expo)
    num = 1
    for e in expo:
        num *= math.factorial(e)
    return Fraction(num, math.factorial(sum(expo) + d))


def _lattice_nodes(d, p):
    """FIAT entity-ordered equispaced lattice: integer barycentric multi-indices (k_0..k_d)."""
    if p == 0:
        raise ValueError("degree must be >= 1 (P0 has no facet nodes)")
    if d == 1:
        verts = [(p, 0), (0, p)]
        interior = [(p - i, i) for i in range(1, p)]
        return verts + interior

    def bary_from_vertices(vs, ks):
        """point = sum_j ks[j] * vertex vs[j] (in lattice units) as a (d+1) multi-index."""
        out = [0] * (d + 1)
        for v, k in zip(vs, ks):
            out[v] += k
        return tuple(out)

    nodes = []
    # vertices
    for v in range(d + 1):
        nodes.append(bary_from_vertices([v], [p]))
    # edges in UFC order
    if d == 2:
        edges = [(1, 2), (0, 2), (0, 1)]
    else:
        edges = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
    for (a, b) in edges:
        for i in range(1, p):  # from vertex a towards vertex b
            nodes.append(bary_from_vertices([a, b], [p - i, i]))
    # triangular faces (d == 3) in UFC order: face k is opposite vertex k
    if d == 3:
        faces = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
        for (a, b, c) in faces:
            for ii in range(1, p):
                for jj in range(1, p - ii):
                    nodes.append(bary_from_vertices([a, b, c], [p - ii - jj, jj, ii]))
    # cell interior
    if d == 2:
        for ii in range(1, p):
            for jj in range(1, p - ii):
                nodes.append((p - ii - jj, jj, ii))
    else:
        for kk in range(1, p):
            for ii in range(1, p - kk):
                for jj in range(1, p - kk - ii):
                    nodes.append((p - ii - jj - kk, jj, ii, kk))
    assert len(set(nodes)) == len(nodes) == math.comb(p + d, d)
    return nodes


def facet_perms(d):
    """All orderings of the d vertices of a facet, in a fixed (lexicographic) order."""
    return list(itertools.permutations(range(d)))


class RefElem:
    """Exact reference-element tables for (dim, degree); see module docstring."""

    def __init__(self, dim: int, degree: int):
        if dim not in (1, 2, 3):
            raise ValueError("dim must be 1, 2 or 3")
        if degree < 1:
            raise ValueError("degree must be >= 1")
        d, p = dim, degree
        self.dim, self.degree = d, p
        self.nd = math.comb(p + d, d)
        self.nfp = math.comb(p + d - 1, d - 1)
        self.nf = d + 1

        lat = _lattice_nodes(d, p)
        self.lattice = np.array(lat, dtype=np.int64)
        # reference coordinates xi_r = lambda_{r+1}
        X = [[Fraction(k[r + 1], p) for r in range(d)] for k in lat]
        self._X = X
        self.nodes = np.array([[float(x) for x in row] for row in X])
        self._node_of = {k: i for i, k in enumerate(lat)}

        mono = _monomials(d, p)
        assert len(mono) == self.nd
        # Vandermonde V[a][alpha] = x_a^alpha ; phi_b(x) = sum_alpha C[alpha][b] x^alpha
        V = [[self._mono_eval(e, x) for e in mono] for x in X]
        C = _finv(V)
        self._mono, self._C = mono, C
        # monomial Gram matrix
        G = [[_simplex_monomial_integral(tuple(a + b for a, b in zip(ea, eb))) for eb in mono] for ea in mono]
        Ct = [list(r) for r in zip(*C)]
        M = _fmatmul(_fmatmul(Ct, G), C)
        Minv = _finv(M)
        self._M, self._Minv = M, Minv
        self.M = self._tofloat(M)
        self.Minv = self._tofloat(Minv)

        # strong derivative matrices
        Dr = []
        for r in range(d):
            dV = [[self._mono_deriv_eval(e, r, x) for e in mono] for x in X]
            Dr.append(_fmatmul(dV, C))
        self._Dr = Dr
        self.Dr = np.array([self._tofloat(D) for D in Dr])

        # facet node lists, in the lattice order of the facet's own vertices (increasing local index)
        self.fverts = [tuple(v for v in range(d + 1) if v != f) for f in range(d + 1)]
        if d == 1:
            flat = [(1,)]  # a point
        else:
            flat = _lattice_nodes(d - 1, p)
        self._flat = flat
        fnodes = []
        for f in range(d + 1):
            row = []
            for kf in flat:
                k = [0] * (d + 1)
                for v, kk in zip(self.fverts[f], kf):
                    k[v] = kk if d > 1 else p
                row.append(self._node_of[tuple(k)])
            fnodes.append(row)
        self.fnodes = np.array(fnodes, dtype=np.int64)

        # facet mass (reference facet of measure 1/(d-1)!)
        if d == 1:
            FM = [[Fraction(1)]]
        else:
            fe = RefElem._cached(d - 1, p)
            FM = fe._M
        self._FM = FM
        self.FM = self._tofloat(FM)
        Lift = []
        for f in range(d + 1):
            cols = [[Minv[a][n] for n in fnodes[f]] for a in range(self.nd)]
            Lift.append(_fmatmul(cols, FM))
        self._Lift = Lift
        self.Lift = np.array([self._tofloat(L) for L in Lift])

        self.perms = facet_perms(d)
        self.ftab = self._facet_match_table()

    # -- small helpers ------------------------------------------------------
    @staticmethod
    @lru_cache(maxsize=None)
    def _cached(dim, degree):
        return RefElem(dim, degree)

    @staticmethod
    def _tofloat(A):
        return np.array([[float(x) for x in row] for row in A])

    @staticmethod
    def _mono_eval(e, x):
        out = Fraction(1)
        for ek, xk in zip(e, x):
            out *= xk ** ek
        return out

    @staticmethod
    def _mono_deriv_eval(e, r, x):
        if e[r] == 0:
            return Fraction(0)
        out = Fraction(e[r])
        for k, (ek, xk) in enumerate(zip(e, x)):
            out *= xk ** (ek - 1 if k == r else ek)
        return out

    # -- neighbour node matching ---------------------------------------------
    def _facet_match_table(self):
        """ftab[f', s, m]: local node of the NEIGHBOUR that coincides with my facet node m.

        My facet nodes are listed w.r.t. my facet vertices (w_0..w_{d-1}) in increasing
        local index.  ``perms[s] = sigma`` says: my facet vertex w_j is the neighbour's
        facet vertex number sigma[j] (numbered again by increasing neighbour-local
        index among the vertices of its facet f').  A lattice point with facet
        multi-index (k_0..k_{d-1}) on my side therefore has, on the neighbour's side,
        multi-index k' with k'[sigma[j]] = k[j].
        """
        d, p = self.dim, self.degree
        tab = np.zeros((self.nf, len(self.perms), self.nfp), dtype=np.int64)
        for fp in range(self.nf):
            for s, sigma in enumerate(self.perms):
                for m, kf in enumerate(self._flat):
                    kn = [0] * d
                    for j in range(d):
                        kn[sigma[j]] = kf[j] if d > 1 else p
                    k = [0] * (d + 1)
                    for v, kk in zip(self.fverts[fp], kn):
                        k[v] = kk
                    tab[fp, s, m] = self._node_of[tuple(k)]
        return tab

    # -- sponge: L2 projection of (sigma * u) ---------------------------------
    def absorption_tensor(self, sigma_degree: int) -> np.ndarray:
        """W[a, b, c] = sum_a' Minv[a, a'] * int phi_a' psi_b phi_c  (psi: P_q basis of sigma).

        After the inverse mass, the term ``-inner(w, absorption*u0)*dx`` of
        ``seigen/elastic.py:207-208`` is ``-(sum_b sigma_b W[:, b, :]) @ u``; |detJ| cancels.
        """
        q = sigma_degree
        key = ("W", q)
        cache = self.__dict__.setdefault("_cache", {})
        if key in cache:
            return cache[key]
        se = RefElem._cached(self.dim, q) if q != self.degree else self
        d = self.dim
        # triple products in monomial space
        monoP, CP = self._mono, self._C
        monoQ, CQ = se._mono, se._C
        nP, nQ = len(monoP), len(monoQ)
        I3 = {}

        def integ(e):
            if e not in I3:
                I3[e] = _simplex_monomial_integral(e)
            return I3[e]

        # T_mono[alpha][beta][gamma]
        T = np.zeros((self.nd, se.nd, self.nd), dtype=object)
        # transform index by index (small sizes): first contract gamma with CP, etc.
        Tm = [[[integ(tuple(a + b + c for a, b, c in zip(ea, eb, ec))) for ec in monoP] for eb in monoQ]
              for ea in monoP]
        # phi_a = sum_alpha CP[alpha][a] x^alpha
        CPt = [list(r) for r in zip(*CP)]   # [a][alpha]
        CQt = [list(r) for r in zip(*CQ)]   # [b][beta]
        # step 1: over gamma -> c
        S1 = [[[sum((Tm[al][be][ga] * CP[ga][c] for ga in range(nP) if CP[ga][c] != 0), Fraction(0))
                for c in range(nP)] for be in range(nQ)] for al in range(nP)]
        S2 = [[[sum((S1[al][be][c] * CQt[b][be] for be in range(nQ) if CQt[b][be] != 0), Fraction(0))
                for c in range(nP)] for b in range(nQ)] for al in range(nP)]
        S3 = [[[sum((S2[al][b][c] * CPt[a][al] for al in range(nP) if CPt[a][al] != 0), Fraction(0))
                for c in range(nP)] for b in range(nQ)] for a in range(nP)]
        Minv = self._Minv
        W = np.zeros((self.nd, se.nd, self.nd))
        for a in range(nP):
            for b in range(nQ):
                for c in range(nP):
                    W[a, b, c] = float(sum((Minv[a][ap] * S3[ap][b][c] for ap in range(nP)), Fraction(0)))
        del T, d
        cache[key] = W
        return W

    # -- evaluation (host-side interpolation / probing) -----------------------
    def tabulate(self, xi: np.ndarray) -> np.ndarray:
        """Basis values phi_b(xi) at reference points xi (n, d) -> (n, nd), float64."""
        xi = np.atleast_2d(np.asarray(xi, dtype=float))
        Vx = np.ones((xi.shape[0], self.nd))
        for col, e in enumerate(self._mono):
            for k, ek in enumerate(e):
                if ek:
                    Vx[:, col] *= xi[:, k] ** ek
        C = np.array([[float(x) for x in row] for row in self._C])
        return Vx @ C


def get_refelem(dim: int, degree: int) -> RefElem:
    return RefElem._cached(dim, degree)
