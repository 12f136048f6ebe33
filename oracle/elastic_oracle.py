"""CPU ORACLE for the ElasticLF4 explicit path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``seigen_b200``) never does.

What this is
------------
A literal NumPy restatement of what Firedrake/PyOP2 execute for
``/root/reference/seigen/elastic.py``:

* the UFL forms ``f`` (``elastic.py:204-209``) and ``g`` (``elastic.py:211-219``) are assembled
  term by term with Gauss quadrature over cells (``dx``), interior facets (``dS``: both
  restrictions gathered, results INC'ed into both cells) and exterior facets (``ds``),
  exactly the three loop types of SURVEY.md section 3.3;
* ``assemble(inner(w, u)*dx, inverse=True)`` (``elastic.py:376-381``) = per-cell inverse of the
  quadrature-assembled mass block;
* ``ExplicitElasticLF4.solve`` (``elastic.py:358-367``) = assemble RHS, multiply by the block inverse;
* the eight forms ``form_uh1 .. form_s1`` (``elastic.py:156-202`` with the explicit overrides
  ``:341-352``) and the time loop of ``ElasticLF4.run`` (``elastic.py:267-315``): same order of
  solves, same ``u0 <- u1`` / ``s0 <- s1`` assignments, same accumulated ``t += dt`` stop rule,
  source re-interpolated at the END time of each step (``:285-288``).

It deliberately shares no code with ``seigen_b200``: its own node enumeration, basis
tabulation, facet matching (dictionary of sorted vertex tuples), normals (from edge
vectors / cross products) and quadrature.  It never uses the quadrature-free nodal
operator the CUDA kernels implement.

Pinning status  --  **parity unpinned at the Firedrake boundary**
------------------------------------------------------------------
Firedrake / PyOP2 / TSFC / PETSc are not importable here and the reference repository stores
no DoF vectors, norms or hashes from a real run.  The oracle is pinned to what the
reference *does* hold for this path: the analytic eigenmode solutions of
``tests/eigenmode/eigenmode_2d.py:30-46`` / ``eigenmode_3d.py:30-50`` (errors and convergence
rates, ``tests/test_oracle_eigenmode.py``) and the external sensor trace
``tests/explosive_source/REF-C1`` (``tests/test_oracle_refc.py``, loose).  Because the scheme is
exact-integration DG on affine simplices, any correct implementation of the same forms agrees
with this one to round-off; the DoF *ordering* of a real Firedrake dump would have to be matched
by (cell vertices, node position).
"""
from __future__ import annotations

import itertools
import math

import numpy as np

__all__ = ["ElasticOracle", "lagrange_lattice", "step_count", "step_times"]


# ----------------------------------------------------------------------------------------
# reference element (independent restatement of FIAT's entity-ordered equispaced lattice)
# ----------------------------------------------------------------------------------------
def lagrange_lattice(dim: int, p: int) -> np.ndarray:
    """Reference coordinates (nd, dim) of equispaced P_p nodes: vertices, edges, faces, interior."""
    pts = [np.array(k[1:], dtype=float) / p
           for k in itertools.product(range(p + 1), repeat=dim + 1) if sum(k) == p]
    bary = lambda x: np.concatenate([[1.0 - x.sum()], x])
    verts = np.vstack([np.zeros(dim), np.eye(dim)])
    out = []

    def on_entity(x, vs):
        lam = bary(x)
        inside = all(lam[v] > 1e-12 for v in vs)
        rest = all(abs(lam[v]) < 1e-12 for v in range(dim + 1) if v not in vs)
        return inside and rest

    def entity_points(vs):
        cand = [x for x in pts if on_entity(x, vs)]
        # order: last entity vertex slowest, second vertex fastest (FIAT make_lattice order)
        def key(x):
            lam = bary(x)
            return tuple(round(lam[v] * p) for v in reversed(vs[1:]))
        return sorted(cand, key=key)

    for v in range(dim + 1):
        out += entity_points((v,))
    if dim >= 2:
        edges = [(1, 2), (0, 2), (0, 1)] if dim == 2 else [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
        for e in edges:
            out += entity_points(e)
    if dim == 3:
        for f in [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]:
            out += entity_points(f)
    out += entity_points(tuple(range(dim + 1)))
    out = np.array(out)
    assert out.shape[0] == math.comb(p + dim, dim)
    del verts
    return out


class _Lagrange:
    """Nodal basis evaluated through a numerically inverted Vandermonde matrix."""

    def __init__(self, dim, p):
        self.dim, self.p = dim, p
        self.nodes = lagrange_lattice(dim, p)
        self.nd = self.nodes.shape[0]
        self.expo = [e for e in itertools.product(range(p + 1), repeat=dim) if sum(e) <= p]
        self.C = np.linalg.inv(self._vander(self.nodes))

    def _vander(self, x):
        x = np.asarray(x, dtype=float)
        V = np.ones(x.shape[:-1] + (len(self.expo),))
        for c, e in enumerate(self.expo):
            for k, ek in enumerate(e):
                if ek:
                    V[..., c] *= x[..., k] ** ek
        return V

    def _dvander(self, x, r):
        x = np.asarray(x, dtype=float)
        V = np.zeros(x.shape[:-1] + (len(self.expo),))
        for c, e in enumerate(self.expo):
            if e[r] == 0:
                continue
            t = np.full(x.shape[:-1], float(e[r]))
            for k, ek in enumerate(e):
                pw = ek - 1 if k == r else ek
                if pw:
                    t = t * x[..., k] ** pw
            V[..., c] = t
        return V

    def tab(self, x):
        """phi_b(x): (..., nd)"""
        return self._vander(x) @ self.C

    def dtab(self, x):
        """d phi_b / d xi_r (x): (..., nd, dim)"""
        return np.stack([self._dvander(x, r) @ self.C for r in range(self.dim)], axis=-1)


def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def simplex_quadrature(dim: int, degree: int):
    """Collapsed Gauss rule on the unit simplex, exact for total degree <= ``degree``."""
    if dim == 0:
        return np.zeros((1, 0)), np.ones(1)
    n = (degree + dim) // 2 + 1
    g, w = _gauss01(n)
    if dim == 1:
        return g[:, None], w
    if dim == 2:
        U, V = np.meshgrid(g, g, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        x = U
        y = V * (1 - U)
        return np.stack([x.ravel(), y.ravel()], 1), (WU * WV * (1 - U)).ravel()
    U, V, W = np.meshgrid(g, g, g, indexing="ij")
    WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
    x = U
    y = V * (1 - U)
    z = W * (1 - U) * (1 - V)
    wt = WU * WV * WW * (1 - U) ** 2 * (1 - V)
    return np.stack([x.ravel(), y.ravel(), z.ravel()], 1), wt.ravel()


# ----------------------------------------------------------------------------------------
# time-loop bookkeeping of ElasticLF4.run  (elastic.py:279-280, 313)
# ----------------------------------------------------------------------------------------
def step_times(T: float, dt: float):
    """The values ``t`` takes in ``while t <= T + 1e-12: ...; t += dt`` starting from t = dt."""
    out = []
    t = dt
    while t <= T + 1e-12:
        out.append(t)
        t += dt
    return out


def step_count(T: float, dt: float) -> int:
    return len(step_times(T, dt))


# ----------------------------------------------------------------------------------------
# the oracle
# ----------------------------------------------------------------------------------------
class ElasticOracle:
    """Literal assembly of the ElasticLF4 explicit scheme on an explicit ``(coords, cells)`` mesh.

    Fields use the Firedrake ``dat.data`` layout: ``u[E*nd, d]`` / ``s[E*nd, d, d]`` (here kept
    reshaped as ``(E, nd, d)`` / ``(E, nd, d, d)``).
    """

    def __init__(self, coords, cells, degree, sigma_degree=None, lite=False):
        """``lite``: skip the per-cell tables only the NumPy assembly needs (``gphi``: E*nq*nd*d doubles, 10 GB for
        98 304 P3 tetrahedra); enough for the C restatement (``c_oracle.COracle``), which recomputes them per cell."""
        coords = np.asarray(coords, dtype=float)
        if coords.ndim == 1:
            coords = coords[:, None]
        cells = np.asarray(cells, dtype=np.int64)
        self.coords, self.cells = coords, cells
        self.dim = d = coords.shape[1]
        self.p = degree
        self.E = E = cells.shape[0]
        self.el = _Lagrange(d, degree)
        self.nd = self.el.nd
        self.sel = _Lagrange(d, sigma_degree) if sigma_degree is not None else None
        qdeg = 2 * degree + (sigma_degree or 0)
        self.xq, self.wq = simplex_quadrature(d, qdeg)
        self.phi = self.el.tab(self.xq)               # (nq, nd)
        self.dphi = self.el.dtab(self.xq)             # (nq, nd, d)   reference derivatives
        self.psi = self.sel.tab(self.xq) if self.sel is not None else None

        v = coords[cells]                             # (E, d+1, d)
        self.v0 = v[:, 0, :]
        self.J = np.swapaxes(v[:, 1:, :] - v[:, :1, :], 1, 2)   # J[e, k, r]
        self.detJ = np.abs(np.linalg.det(self.J))
        self.Jinv = np.linalg.inv(self.J)             # Jinv[e, r, k]
        # physical gradients of the basis at quadrature points: gphi[e, q, a, k]
        self.gphi = None if lite else np.einsum("qar,erk->eqak", self.dphi, self.Jinv)

        # mass blocks and their inverses  (assemble(inner(w,u)*dx, inverse=True))
        Mref = np.einsum("q,qa,qb->ab", self.wq, self.phi, self.phi)
        self.Mcell = self.detJ[:, None, None] * Mref[None]
        self.Minv_cell = np.linalg.inv(self.Mcell)

        self._build_facets()
        # parameters (plain attributes, as in the reference)
        self.density = 1.0
        self.l = None
        self.mu = None
        self.dt = None
        self.sigma = None           # (E, nd_sigma) nodal values of the absorption field, or None
        self.source = None          # callable t -> (E, nd, d, d) nodal source values, or None

    # -- facets ----------------------------------------------------------------------
    def _build_facets(self):
        d, E = self.dim, self.E
        cells, coords = self.cells, self.coords
        # facet matching through the sorted vertex tuple of every (cell, facet).  Small meshes: a dictionary, visited
        # in (cell, facet) order; large meshes: the same pairs in the same order from one sort (an interior facet is
        # listed where its SECOND cell meets it, as the dictionary walk does).
        if E <= 20000:
            seen = {}
            interior, exterior = [], []
            for e in range(E):
                for f in range(d + 1):
                    key = tuple(sorted(int(x) for i, x in enumerate(cells[e]) if i != f))
                    if key in seen:
                        interior.append(seen.pop(key) + (e, f))
                    else:
                        seen[key] = (e, f)
            exterior = sorted(seen.values())
            self.int_facets = np.array(interior, dtype=np.int64).reshape(-1, 4)     # e+, f+, e-, f-
            self.ext_facets = np.array(exterior, dtype=np.int64).reshape(-1, 2)
        else:
            nf = d + 1
            keys = np.stack([np.sort(np.delete(cells, f, axis=1), axis=1) for f in range(nf)], axis=1).reshape(E * nf, d)
            order = np.lexsort(keys.T[::-1])                  # groups equal tuples; stable: visit order inside a group
            ks = keys[order]
            same = np.all(ks[1:] == ks[:-1], axis=1)
            first, second = order[:-1][same], order[1:][same]
            by_second = np.argsort(second, kind="stable")
            first, second = first[by_second], second[by_second]
            self.int_facets = np.stack([first // nf, first % nf, second // nf, second % nf], axis=1).astype(np.int64)
            paired = np.zeros(E * nf, dtype=bool)
            paired[first] = True
            paired[second] = True
            ext = np.flatnonzero(~paired)
            self.ext_facets = np.stack([ext // nf, ext % nf], axis=1).astype(np.int64)

        fq, fw = simplex_quadrature(d - 1, 2 * self.p) if d > 1 else (np.zeros((1, 0)), np.ones(1))
        self.fw_ref = fw

        def facet_data(e, f):
            """physical quad points, outward unit normal and measure scale of facet f of cells e."""
            vid = np.array([[i for i in range(d + 1) if i != ff] for ff in range(d + 1)])[f]   # (n, d)
            vv = coords[cells[e[:, None], vid]]                    # (n, d, dim)
            opp = coords[cells[e, f]]                              # (n, dim)
            va = vv[:, 0, :]
            if d == 1:
                x = va[:, None, :]
                nrm = np.sign(va - opp)
                meas = np.ones(len(e))
            elif d == 2:
                t = vv[:, 1, :] - va
                x = va[:, None, :] + fq[None, :, 0:1] * t[:, None, :]
                length = np.linalg.norm(t, axis=1)
                nrm = np.stack([t[:, 1], -t[:, 0]], 1) / length[:, None]
                flip = np.einsum("nk,nk->n", nrm, opp - va) > 0
                nrm[flip] *= -1
                meas = length                                      # reference edge has measure 1
            else:
                t1 = vv[:, 1, :] - va
                t2 = vv[:, 2, :] - va
                x = va[:, None, :] + fq[None, :, 0:1] * t1[:, None, :] + fq[None, :, 1:2] * t2[:, None, :]
                c = np.cross(t1, t2)
                cn = np.linalg.norm(c, axis=1)
                nrm = c / cn[:, None]
                flip = np.einsum("nk,nk->n", nrm, opp - va) > 0
                nrm[flip] *= -1
                meas = cn                                          # = 2*area; reference facet measure 1/2
            return x, nrm, meas

        def tabulate_on(e, x):
            xi = np.einsum("nrk,nqk->nqr", self.Jinv[e], x - self.v0[e][:, None, :])
            return self.el.tab(xi)                                 # (n, nq, nd)

        if len(self.int_facets):
            ep, fp, em, fm = self.int_facets.T
            x, n_plus, meas = facet_data(ep, fp)
            self.iF = dict(ep=ep, em=em, n_plus=n_plus, n_minus=-n_plus, w=meas[:, None] * fw[None, :],
                           phi_p=tabulate_on(ep, x), phi_m=tabulate_on(em, x))
        else:
            self.iF = None
        e, f = self.ext_facets.T
        x, nrm, meas = facet_data(e, f)
        self.eF = dict(e=e, n=nrm, w=meas[:, None] * fw[None, :], phi=tabulate_on(e, x))

    # -- helpers -----------------------------------------------------------------------
    def zeros_u(self):
        return np.zeros((self.E, self.nd, self.dim))

    def zeros_s(self):
        return np.zeros((self.E, self.nd, self.dim, self.dim))

    def node_coords(self):
        lam0 = 1.0 - self.el.nodes.sum(1)
        lam = np.concatenate([lam0[:, None], self.el.nodes], axis=1)
        return np.einsum("av,evk->eak", lam, self.coords[self.cells])

    def sigma_node_coords(self):
        lam0 = 1.0 - self.sel.nodes.sum(1)
        lam = np.concatenate([lam0[:, None], self.sel.nodes], axis=1)
        return np.einsum("av,evk->eak", lam, self.coords[self.cells])

    def apply_inverse_mass(self, F):
        """``matrix.handle.mult(F_v, res)`` with the block-diagonal inverse (elastic.py:365-367)."""
        return np.einsum("eab,eb...->ea...", self.Minv_cell, F)

    def apply_mass(self, x):
        return np.einsum("eab,eb...->ea...", self.Mcell, x)

    # -- the two RHS forms ---------------------------------------------------------------
    def assemble_f(self, s, u0):
        """``assemble(f(w, s, u0, n, absorption))`` -> (E, nd, d).   elastic.py:204-209"""
        d = self.dim
        F = self.zeros_u()
        wdx = self.wq[None, :] * self.detJ[:, None]                          # (E, nq)
        sq = np.einsum("qb,ebij->eqij", self.phi, s)
        # - inner(grad(w), s0)*dx         grad(w)[i, j] = d w_i / d x_j
        F -= np.einsum("eq,eqaj,eqij->eai", wdx, self.gphi, sq)
        if self.iF is not None:
            iF = self.iF
            sp = np.einsum("nqb,nbij->nqij", iF["phi_p"], s[iF["ep"]])
            sm = np.einsum("nqb,nbij->nqij", iF["phi_m"], s[iF["em"]])
            avg = 0.5 * (sp + sm)
            # inner(avg(s0)*n('+'), w('+'))*dS + inner(avg(s0)*n('-'), w('-'))*dS
            Fp = np.einsum("nq,nqij,nj,nqa->nai", iF["w"], avg, iF["n_plus"], iF["phi_p"])
            Fm = np.einsum("nq,nqij,nj,nqa->nai", iF["w"], avg, iF["n_minus"], iF["phi_m"])
            np.add.at(F, iF["ep"], Fp)
            np.add.at(F, iF["em"], Fm)
        if self.sigma is not None:
            # - inner(w, absorption*u0)*dx
            sig_q = np.einsum("qb,eb->eq", self.psi, self.sigma)
            uq = np.einsum("qb,ebi->eqi", self.phi, u0)
            F -= np.einsum("eq,eq,eqi,qa->eai", wdx, sig_q, uq, self.phi)
        del d
        return F

    def assemble_g(self, u, src):
        """``assemble(g(v, u, I, n, l, mu, source))`` -> (E, nd, d, d).   elastic.py:211-219"""
        d = self.dim
        I = np.eye(d)
        lam, mu = self._cellwise(self.l), self._cellwise(self.mu)
        G = self.zeros_s()
        wdx = self.wq[None, :] * self.detJ[:, None]
        uq = np.einsum("qb,ebk->eqk", self.phi, u)
        # test function v = phi_a e_i (x) e_j:  tr(v) = phi_a delta_ij ; div(v)_m = d_j phi_a delta_im ;
        # div(v.T)_m = d_i phi_a delta_jm
        # - l*(v[i,j]*I[i,j]).dx(k)*u1[k]*dx
        t = np.einsum("eq,eqak,eqk->ea", wdx, self.gphi, uq)
        G -= (lam[:, None] * t)[:, :, None, None] * I[None, None]
        # - mu*inner(div(v), u1)*dx        -> G[a, i, j] -= mu * int d_j phi_a u_i
        G -= mu[:, None, None, None] * np.einsum("eq,eqaj,eqi->eaij", wdx, self.gphi, uq)
        # - mu*inner(div(v.T), u1)*dx      -> G[a, i, j] -= mu * int d_i phi_a u_j
        G -= mu[:, None, None, None] * np.einsum("eq,eqai,eqj->eaij", wdx, self.gphi, uq)
        if self.iF is not None:
            iF = self.iF
            ep, em = iF["ep"], iF["em"]
            up = np.einsum("nqb,nbk->nqk", iF["phi_p"], u[ep])
            um = np.einsum("nqb,nbk->nqk", iF["phi_m"], u[em])
            avg = 0.5 * (up + um)
            for side, e, nrm, ph in (("+", ep, iF["n_plus"], iF["phi_p"]), ("-", em, iF["n_minus"], iF["phi_m"])):
                # a test function living on one side only contributes through its own restriction:
                # jump(tr v, n)[k] -> phi_a delta_ij n_k ;  jump(v, n)[m] -> phi_a delta_im n_j ;
                # jump(v.T, n)[m] -> phi_a delta_jm n_i
                an = np.einsum("nq,nqk,nk,nqa->na", iF["w"], avg, nrm, ph)
                # + l*(jump(v[i,j], n[k])*I[i,j]*avg(u1[k]))*dS
                c1 = (lam[e][:, None] * an)[:, :, None, None] * I[None, None]
                # + mu*inner(avg(u1), jump(v, n))*dS    -> avg_i n_j
                c2 = mu[e][:, None, None, None] * np.einsum("nq,nqi,nj,nqa->naij", iF["w"], avg, nrm, ph)
                # + mu*inner(avg(u1), jump(v.T, n))*dS  -> avg_j n_i
                c3 = mu[e][:, None, None, None] * np.einsum("nq,nqj,ni,nqa->naij", iF["w"], avg, nrm, ph)
                np.add.at(G, e, c1 + c2 + c3)
                del side
        eF = self.eF
        e = eF["e"]
        ub = np.einsum("nqb,nbk->nqk", eF["phi"], u[e])
        # + l*(v[i,j]*I[i,j]*u1[k]*n[k])*ds
        un = np.einsum("nq,nqk,nk,nqa->na", eF["w"], ub, eF["n"], eF["phi"])
        c1 = (lam[e][:, None] * un)[:, :, None, None] * I[None, None]
        # + mu*inner(u1, dot(v, n))*ds     -> u_i n_j ;   + mu*inner(u1, dot(v.T, n))*ds -> u_j n_i
        c2 = mu[e][:, None, None, None] * np.einsum("nq,nqi,nj,nqa->naij", eF["w"], ub, eF["n"], eF["phi"])
        c3 = mu[e][:, None, None, None] * np.einsum("nq,nqj,ni,nqa->naij", eF["w"], ub, eF["n"], eF["phi"])
        np.add.at(G, e, c1 + c2 + c3)
        if src is not None:
            # + inner(v, source)*dx
            sq = np.einsum("qb,ebij->eqij", self.phi, src)
            G += np.einsum("eq,eqij,qa->eaij", wdx, sq, self.phi)
        return G

    def _cellwise(self, c):
        """Lame parameters: a float (every reference script) or one value per cell."""
        c = np.asarray(c, dtype=float)
        if c.ndim == 0:
            return np.full(self.E, float(c))
        assert c.shape == (self.E,)
        return c

    # -- solves (elastic.py:156-202, 341-352, 358-367) ---------------------------------
    def solve_f(self, s, u0):
        return self.apply_inverse_mass(self.assemble_f(s, u0))

    def solve_g(self, u, src):
        return self.apply_inverse_mass(self.assemble_g(u, src))

    def solve_u1(self, u0, uh1, uh2):
        dt = self.dt
        rhs = self.density * self.apply_mass(u0) + dt * self.apply_mass(uh1) + (dt ** 3 / 24.0) * self.apply_mass(uh2)
        return self.apply_inverse_mass(rhs)          # un-weighted inverse mass: SURVEY App. B-3

    def solve_s1(self, s0, sh1, sh2):
        dt = self.dt
        rhs = self.apply_mass(s0) + dt * self.apply_mass(sh1) + (dt ** 3 / 24.0) * self.apply_mass(sh2)
        return self.apply_inverse_mass(rhs)

    def step(self, u0, s0, t):
        """One pass of the loop body of ``run`` (elastic.py:283-304); returns (u1, s1) and the stage fields."""
        src = self.source(t) if self.source is not None else None
        uh1 = self.solve_f(s0, u0)
        stemp = self.solve_g(uh1, src)
        uh2 = self.solve_f(stemp, u0)
        u1 = self.solve_u1(u0, uh1, uh2)
        sh1 = self.solve_g(u1, src)
        utemp = self.solve_f(sh1, u1)
        sh2 = self.solve_g(utemp, src)
        s1 = self.solve_s1(s0, sh1, sh2)
        return u1, s1, dict(uh1=uh1, stemp=stemp, uh2=uh2, sh1=sh1, utemp=utemp, sh2=sh2)

    def run(self, u0, s0, T, callback=None):
        """``ElasticLF4.run(T)`` (elastic.py:267-315)."""
        u, s = np.array(u0, dtype=float), np.array(s0, dtype=float)
        for n, t in enumerate(step_times(T, self.dt)):
            u, s, _ = self.step(u, s, t)
            if callback is not None:
                callback(n, t, u, s)
        return u, s

    # -- error norms (true L2, by quadrature) -----------------------------------------------
    def l2_error(self, field, exact):
        """|| field_h - exact ||_L2 with ``exact(x) -> (..., comps)`` evaluated at quadrature points."""
        lam0 = 1.0 - self.xq.sum(1)
        lam = np.concatenate([lam0[:, None], self.xq], axis=1)
        xq = np.einsum("qv,evk->eqk", lam, self.coords[self.cells])
        fq = np.einsum("qb,eb...->eq...", self.phi, field)
        diff = fq - exact(xq)
        diff = diff.reshape(diff.shape[0], diff.shape[1], -1)
        wdx = self.wq[None, :] * self.detJ[:, None]
        return float(np.sqrt(np.einsum("eq,eqc,eqc->", wdx, diff, diff)))
