/* CPU ORACLE (C restatement) for the ElasticLF4 explicit path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Pinning status: parity unpinned at the Firedrake boundary (Firedrake/PyOP2/TSFC/PETSc are not installable
 * here and the reference stores no DoF vectors; see oracle/elastic_oracle.py and DESIGN.md).  This file is
 * checked against oracle/elastic_oracle.py (the literal NumPy restatement, itself pinned to the reference's
 * analytic eigenmode solutions and the REF-C1 trace) to 1e-13 in tests/test_oracle_c.py.
 *
 * Structure mirrors what PyOP2 executes for /root/reference/seigen/elastic.py, one sweep of memory per loop:
 *   solve(rhs, invmass, result)                      elastic.py:358-367
 *     F = 0; cell loop (dx); interior-facet loop (dS, INC into both cells); exterior-facet loop (ds, g only)
 *     result = blockdiag(Minv) * F                   elastic.py:365-367 / 476-484
 *   forms f, g                                        elastic.py:204-219
 *   stage forms and LF4 combination                   elastic.py:156-202, 341-352
 *   time loop                                         elastic.py:279-313
 * i.e. 25 field sweeps per time step (tests/tiling/utils.py:263-269), each kernel evaluating the UFL integrand
 * at quadrature points from tabulated basis functions like a TSFC-generated kernel would.
 *
 * Threading: cells are split into contiguous chunks, one per OpenMP thread (the analogue of one MPI rank's
 * owned cells); a facet shared by two chunks is evaluated by both threads, each incrementing only its own
 * cell (the analogue of PyOP2's redundant computation over the exec halo).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 3
#define MAXND 20

typedef struct {
  int d, nd, nq, nfq, nperm;
  int64_t E, nif, nef;
  const double* wq;    /* [nq] */
  const double* phi;   /* [nq][nd] */
  const double* dphi;  /* [nq][nd][d] reference derivatives */
  const double* fw;    /* [nfq] facet quadrature weights (reference facet) */
  const double* phif;  /* [d+1][nperm][nfq][nd] cell basis at facet quadrature points, facet vertices in order perm */
  const double* jinv;  /* [E][d][d]  Jinv[r][k] */
  const double* detj;  /* [E] |det J| */
  const double* minv;  /* [E][nd][nd] inverse of the assembled cell mass block */
  const double* mass;  /* [E][nd][nd] */
  const int32_t* ifac; /* [nif][6]: e+, f+, perm+, e-, f-, perm- */
  const double* inrm;  /* [nif][d] unit normal n('+') */
  const double* imeas; /* [nif] facet measure / reference facet measure */
  const int32_t* efac; /* [nef][3]: e, f, perm */
  const double* enrm;  /* [nef][d] */
  const double* emeas; /* [nef] */
  const double* lam;   /* [E] */
  const double* mu;    /* [E] */
  const double* sigq;  /* [E][nq] absorption at cell quadrature points, or NULL */
  double density;
} oracle_ctx;

static int nthreads_(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int oracle_num_threads(void) { return nthreads_(); }

/* Overrides OMP_NUM_THREADS for this process (torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline is
 * meant to use all the host cores it can, bench.py). */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static void chunk(int64_t n, int t, int nt, int64_t* lo, int64_t* hi) {
  *lo = n * t / nt;
  *hi = n * (t + 1) / nt;
}

/* result = blockdiag(minv) * F   (PETSc MatMult with the block-diagonal inverse mass, elastic.py:365-367) */
static void mass_apply(const oracle_ctx* c, const double* M, const double* F, double* out, int nc) {
  const int nd = c->nd;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < c->E; ++e) {
    const double* Me = M + e * nd * nd;
    const double* Fe = F + e * nd * nc;
    double* oe = out + e * nd * nc;
    for (int a = 0; a < nd; ++a)
      for (int k = 0; k < nc; ++k) {
        double acc = 0.0;
        for (int b = 0; b < nd; ++b) acc += Me[a * nd + b] * Fe[b * nc + k];
        oe[a * nc + k] = acc;
      }
  }
}

/* ---- f: velocity RHS  (elastic.py:204-209) ------------------------------------------------------------ */
static void f_cells(const oracle_ctx* c, const double* s, const double* u0, double* F) {
  const int d = c->d, nd = c->nd, nq = c->nq;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < c->E; ++e) {
    const double* se = s + e * nd * d * d;
    const double* ue = u0 + e * nd * d;
    const double* Ji = c->jinv + e * d * d;
    double* Fe = F + e * nd * d;
    for (int q = 0; q < nq; ++q) {
      const double w = c->wq[q] * c->detj[e];
      const double* ph = c->phi + q * nd;
      double sq[MAXD * MAXD] = {0}, uq[MAXD] = {0};
      for (int b = 0; b < nd; ++b) {
        for (int k = 0; k < d * d; ++k) sq[k] += ph[b] * se[b * d * d + k];
        if (c->sigq)
          for (int k = 0; k < d; ++k) uq[k] += ph[b] * ue[b * d + k];
      }
      const double sg = c->sigq ? c->sigq[e * nq + q] : 0.0;
      for (int a = 0; a < nd; ++a) {
        double g[MAXD];
        for (int j = 0; j < d; ++j) {
          double acc = 0.0;
          for (int r = 0; r < d; ++r) acc += c->dphi[(q * nd + a) * d + r] * Ji[r * d + j];
          g[j] = acc;
        }
        for (int i = 0; i < d; ++i) {
          double acc = 0.0;
          for (int j = 0; j < d; ++j) acc += g[j] * sq[i * d + j]; /* inner(grad(w), s) */
          Fe[a * d + i] -= w * acc;
          if (c->sigq) Fe[a * d + i] -= w * sg * uq[i] * ph[a]; /* inner(w, absorption*u0) */
        }
      }
    }
  }
}

static void f_int_facets(const oracle_ctx* c, const double* s, double* F) {
  const int d = c->d, nd = c->nd, nfq = c->nfq;
#pragma omp parallel
  {
    int64_t lo, hi;
    chunk(c->E, omp_get_thread_num(), omp_get_num_threads(), &lo, &hi);
    for (int64_t k = 0; k < c->nif; ++k) {
      const int32_t* fc = c->ifac + k * 6;
      const int64_t ep = fc[0], em = fc[3];
      const int own_p = ep >= lo && ep < hi, own_m = em >= lo && em < hi;
      if (!own_p && !own_m) continue;
      const double* php = c->phif + ((size_t)(fc[1] * c->nperm + fc[2]) * nfq) * nd;
      const double* phm = c->phif + ((size_t)(fc[4] * c->nperm + fc[5]) * nfq) * nd;
      const double* np = c->inrm + k * d;
      for (int q = 0; q < nfq; ++q) {
        const double w = c->fw[q] * c->imeas[k];
        double avg[MAXD * MAXD] = {0};
        for (int b = 0; b < nd; ++b)
          for (int m = 0; m < d * d; ++m)
            avg[m] += 0.5 * (php[q * nd + b] * s[(ep * nd + b) * d * d + m] + phm[q * nd + b] * s[(em * nd + b) * d * d + m]);
        double t[MAXD];
        for (int i = 0; i < d; ++i) {
          double acc = 0.0;
          for (int j = 0; j < d; ++j) acc += avg[i * d + j] * np[j];
          t[i] = acc;
        }
        for (int a = 0; a < nd; ++a)
          for (int i = 0; i < d; ++i) {
            if (own_p) F[(ep * nd + a) * d + i] += w * t[i] * php[q * nd + a];  /* inner(avg(s)*n('+'), w('+')) */
            if (own_m) F[(em * nd + a) * d + i] -= w * t[i] * phm[q * nd + a];  /* n('-') = -n('+') */
          }
      }
    }
  }
}

/* ---- g: stress RHS  (elastic.py:211-219) ---------------------------------------------------------------- */
static void g_cells(const oracle_ctx* c, const double* u, const double* src, double* G) {
  const int d = c->d, nd = c->nd, nq = c->nq;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < c->E; ++e) {
    const double* ue = u + e * nd * d;
    const double* Ji = c->jinv + e * d * d;
    const double lam = c->lam[e], mu = c->mu[e];
    double* Ge = G + e * nd * d * d;
    for (int q = 0; q < nq; ++q) {
      const double w = c->wq[q] * c->detj[e];
      const double* ph = c->phi + q * nd;
      double uq[MAXD] = {0}, sq[MAXD * MAXD] = {0};
      for (int b = 0; b < nd; ++b) {
        for (int k = 0; k < d; ++k) uq[k] += ph[b] * ue[b * d + k];
        if (src)
          for (int k = 0; k < d * d; ++k) sq[k] += ph[b] * src[(e * nd + b) * d * d + k];
      }
      for (int a = 0; a < nd; ++a) {
        double g[MAXD];
        double gu = 0.0;
        for (int j = 0; j < d; ++j) {
          double acc = 0.0;
          for (int r = 0; r < d; ++r) acc += c->dphi[(q * nd + a) * d + r] * Ji[r * d + j];
          g[j] = acc;
          gu += acc * uq[j];
        }
        for (int i = 0; i < d; ++i)
          for (int j = 0; j < d; ++j) {
            double v = 0.0;
            if (i == j) v -= lam * gu;            /* - l*(v[i,j]*I[i,j]).dx(k)*u[k] */
            v -= mu * g[j] * uq[i];               /* - mu*inner(div(v), u)   */
            v -= mu * g[i] * uq[j];               /* - mu*inner(div(v.T), u) */
            if (src) v += sq[i * d + j] * ph[a];  /* + inner(v, source)      */
            Ge[(a * d + i) * d + j] += w * v;
          }
      }
    }
  }
}

static void g_facet_add(int d, int nd, double* Gc, const double* ph_q, double w, const double* val, const double* n,
                        double lam, double mu) {
  /* test function on this cell: + l*tr-part*(val.n) + mu*(val_i n_j + val_j n_i) */
  double vn = 0.0;
  for (int k = 0; k < d; ++k) vn += val[k] * n[k];
  for (int a = 0; a < nd; ++a) {
    const double wa = w * ph_q[a];
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) {
        double v = mu * (val[i] * n[j] + val[j] * n[i]);
        if (i == j) v += lam * vn;
        Gc[(a * d + i) * d + j] += wa * v;
      }
  }
}

static void g_int_facets(const oracle_ctx* c, const double* u, double* G) {
  const int d = c->d, nd = c->nd, nfq = c->nfq;
#pragma omp parallel
  {
    int64_t lo, hi;
    chunk(c->E, omp_get_thread_num(), omp_get_num_threads(), &lo, &hi);
    for (int64_t k = 0; k < c->nif; ++k) {
      const int32_t* fc = c->ifac + k * 6;
      const int64_t ep = fc[0], em = fc[3];
      const int own_p = ep >= lo && ep < hi, own_m = em >= lo && em < hi;
      if (!own_p && !own_m) continue;
      const double* php = c->phif + ((size_t)(fc[1] * c->nperm + fc[2]) * nfq) * nd;
      const double* phm = c->phif + ((size_t)(fc[4] * c->nperm + fc[5]) * nfq) * nd;
      const double* np = c->inrm + k * d;
      double nm[MAXD];
      for (int j = 0; j < d; ++j) nm[j] = -np[j];
      for (int q = 0; q < nfq; ++q) {
        const double w = c->fw[q] * c->imeas[k];
        double avg[MAXD] = {0};
        for (int b = 0; b < nd; ++b)
          for (int m = 0; m < d; ++m)
            avg[m] += 0.5 * (php[q * nd + b] * u[(ep * nd + b) * d + m] + phm[q * nd + b] * u[(em * nd + b) * d + m]);
        if (own_p) g_facet_add(d, nd, G + ep * nd * d * d, php + q * nd, w, avg, np, c->lam[ep], c->mu[ep]);
        if (own_m) g_facet_add(d, nd, G + em * nd * d * d, phm + q * nd, w, avg, nm, c->lam[em], c->mu[em]);
      }
    }
  }
}

static void g_ext_facets(const oracle_ctx* c, const double* u, double* G) {
  const int d = c->d, nd = c->nd, nfq = c->nfq;
#pragma omp parallel
  {
    int64_t lo, hi;
    chunk(c->E, omp_get_thread_num(), omp_get_num_threads(), &lo, &hi);
    for (int64_t k = 0; k < c->nef; ++k) {
      const int32_t* fc = c->efac + k * 3;
      const int64_t e = fc[0];
      if (e < lo || e >= hi) continue;
      const double* ph = c->phif + ((size_t)(fc[1] * c->nperm + fc[2]) * nfq) * nd;
      for (int q = 0; q < nfq; ++q) {
        const double w = c->fw[q] * c->emeas[k];
        double ub[MAXD] = {0};
        for (int b = 0; b < nd; ++b)
          for (int m = 0; m < d; ++m) ub[m] += ph[q * nd + b] * u[(e * nd + b) * d + m];
        g_facet_add(d, nd, G + e * nd * d * d, ph + q * nd, w, ub, c->enrm + k * d, c->lam[e], c->mu[e]);
      }
    }
  }
}

/* ---- solves ------------------------------------------------------------------------------------------------ */
void oracle_solve_f(const oracle_ctx* c, const double* s, const double* u0, double* work, double* out) {
  const size_t n = (size_t)c->E * c->nd * c->d;
  memset(work, 0, n * sizeof(double));                 /* assemble() zeroes a fresh Function */
  f_cells(c, s, u0, work);
  if (c->nif) f_int_facets(c, s, work);
  mass_apply(c, c->minv, work, out, c->d);
}

void oracle_solve_g(const oracle_ctx* c, const double* u, const double* src, double* work, double* out) {
  const size_t n = (size_t)c->E * c->nd * c->d * c->d;
  memset(work, 0, n * sizeof(double));
  g_cells(c, u, src, work);
  if (c->nif) g_int_facets(c, u, work);
  g_ext_facets(c, u, work);
  mass_apply(c, c->minv, work, out, c->d * c->d);
}

/* result = Minv * (c0*M*a + c1*M*b + c2*M*cc)   -- form_u1 / form_s1 (elastic.py:341-352): the reference
 * assembles the mass-weighted right-hand side with one cell loop and multiplies by the inverse. */
static void solve_axpy(const oracle_ctx* c, double c0, const double* a, double c1, const double* b, double c2,
                       const double* cc, double* work, double* out, int nc) {
  const int nd = c->nd;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < c->E; ++e) {
    const double* Me = c->mass + e * nd * nd;
    for (int x = 0; x < nd; ++x)
      for (int k = 0; k < nc; ++k) {
        double acc = 0.0;
        for (int y = 0; y < nd; ++y) {
          const size_t o = ((size_t)e * nd + y) * nc + k;
          acc += Me[x * nd + y] * (c0 * a[o] + c1 * b[o] + c2 * cc[o]);
        }
        work[((size_t)e * nd + x) * nc + k] = acc;
      }
  }
  mass_apply(c, c->minv, work, out, nc);
}

/* One pass of the loop body of ElasticLF4.run (elastic.py:283-304).  u, s are updated in place (u0 <- u1,
 * s0 <- s1).  scratch: uh1, uh2 [U-size]; stemp, sh1 [S-size] ... supplied by the caller as bufU[3], bufS[3]. */
void oracle_step(const oracle_ctx* c, double* u, double* s, const double* src, double dt, double* bufU0,
                 double* bufU1, double* bufU2, double* bufS0, double* bufS1, double* bufS2) {
  const double c3 = dt * dt * dt / 24.0;
  const size_t nU = (size_t)c->E * c->nd * c->d, nS = nU * c->d;
  double* uh1 = bufU0; double* uh2 = bufU1; double* wU = bufU2;
  double* stemp = bufS0; double* sh2 = bufS1; double* wS = bufS2;
  oracle_solve_f(c, s, u, wU, uh1);                       /* :292 */
  oracle_solve_g(c, uh1, src, wS, stemp);                 /* :293 */
  oracle_solve_f(c, stemp, u, wU, uh2);                   /* :294 */
  solve_axpy(c, c->density, u, dt, uh1, c3, uh2, wU, uh1, c->d);   /* :295 (u1 lands in uh1's buffer) */
  memcpy(u, uh1, nU * sizeof(double));                    /* :296 u0.assign(u1) */
  double* sh1 = stemp;
  oracle_solve_g(c, u, src, wS, sh1);                     /* :300 */
  double* utemp = uh2;
  oracle_solve_f(c, sh1, u, wU, utemp);                   /* :301 */
  oracle_solve_g(c, utemp, src, wS, sh2);                 /* :302 */
  solve_axpy(c, 1.0, s, dt, sh1, c3, sh2, wS, sh2, c->d * c->d);   /* :303 */
  memcpy(s, sh2, nS * sizeof(double));                    /* :304 s0.assign(s1) */
}
