/* CPU ORACLE, fused variant (C/OpenMP) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * BASELINE.md section 3 names two CPU baselines: the PyOP2-structured restatement (oracle/elastic_c.c: 25 sweeps per
 * time step, quadrature kernels, scatter-add, block inverse mass -- what the reference executes for
 * /root/reference/seigen/elastic.py:358-367) and THIS one: the same six-pass algorithm the GPU runs (SURVEY.md 8a
 * K1..K6), i.e. the quadrature-free nodal operator of SURVEY.md Appendix A with the LF4 combinations
 * (elastic.py:341-352) fused into passes 3 and 6.  It is the C statement of oracle/nodal.py and is checked against the
 * literal oracle (oracle/elastic_oracle.py) to 1e-12 in tests/test_oracle_fused.py.  Same pinning status: parity
 * unpinned at the Firedrake boundary.  Only tests/ and bench.py's CPU legs may load it.
 *
 * Fields use the boundary layout of the reference: u[cell][node][i], s[cell][node][i][j] (Firedrake dat.data).
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 3
#define MAXND 35
#define MAXNFP 10
#define SG_BOUNDARY 0x80

typedef struct {
  int d, nd, nfp, nperm;
  int64_t E;
  const double* Dr;      /* [d][nd][nd]   strong derivative, inverse mass folded in (reference element) */
  const double* Lift;    /* [d+1][nd][nfp] */
  const int32_t* fnodes; /* [d+1][nfp]    own node of facet node m */
  const int32_t* ftab;   /* [(d+1)*nperm][nfp] neighbour-side node of my facet node m */
  const int32_t* nbr;    /* [E][d+1] */
  const uint8_t* code;   /* [E][d+1]  f'*nperm + s, | SG_BOUNDARY on exterior facets */
  const double* jinv;    /* [E][d][d] */
  const double* lam;     /* [E] */
  const double* mu;      /* [E] */
  const int32_t* absidx; /* [E] row of absmat or -1; NULL = no sponge */
  const double* absmat;  /* [nabs][nd][nd]  A = Minv * int phi_a sigma phi_c */
  double density;
} fused_ctx;

void fused_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int fused_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline void facet_dir(const fused_ctx* c, const double* Ji, int f, double* gf) {
  const int d = c->d;
  for (int j = 0; j < d; ++j) {
    if (f == 0) {
      double a = 0.0;
      for (int r = 0; r < d; ++r) a += Ji[r * d + j];
      gf[j] = a;
    } else {
      gf[j] = -Ji[(f - 1) * d + j];
    }
  }
}

/* out_i = c0*ax0_i + c1*ax1_i + c2*(Dv(s)_i - (A uabs)_i)     (ax0 == NULL: out_i = Dv(s)_i - (A uabs)_i)
 * Dv(s)_i = sum_j d~_j s_ij, central flux, numerical trace 0 on exterior facets (elastic.py:204-209). */
void fused_pass_f(const fused_ctx* c, const double* s, const double* uabs, const double* ax0, const double* ax1,
                  double c0, double c1, double c2, double* out) {
  const int d = c->d, nd = c->nd, nfp = c->nfp, nf = c->d + 1, dd = d * d;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < c->E; ++e) {
    const double* S = s + e * nd * dd;
    const double* Ji = c->jinv + e * dd;
    double acc[MAXD][MAXND];
    for (int i = 0; i < d; ++i) {
      double t[MAXD][MAXND];
      for (int r = 0; r < d; ++r)
        for (int b = 0; b < nd; ++b) {
          double a = 0.0;
          for (int j = 0; j < d; ++j) a += Ji[r * d + j] * S[(b * d + i) * d + j];
          t[r][b] = a;
        }
      for (int a = 0; a < nd; ++a) {
        double v = 0.0;
        for (int r = 0; r < d; ++r) {
          const double* D = c->Dr + ((size_t)r * nd + a) * nd;
          for (int b = 0; b < nd; ++b) v += D[b] * t[r][b];
        }
        acc[i][a] = v;
      }
    }
    for (int f = 0; f < nf; ++f) {
      const unsigned cd = c->code[e * nf + f];
      const int bnd = (cd & SG_BOUNDARY) != 0;
      const double cn = bnd ? 0.0 : 0.5, co = bnd ? 1.0 : 0.5;
      const double* Sn = s + (int64_t)c->nbr[e * nf + f] * nd * dd;
      const int32_t* row = c->ftab + (size_t)(cd & 0x7f) * nfp;
      double gf[MAXD];
      facet_dir(c, Ji, f, gf);
      for (int m = 0; m < nfp; ++m) {
        const int on = c->fnodes[f * nfp + m], nn = row[m];
        for (int i = 0; i < d; ++i) {
          double q = 0.0;
          for (int j = 0; j < d; ++j) q += gf[j] * (cn * Sn[(nn * d + i) * d + j] - co * S[(on * d + i) * d + j]);
          const double* L = c->Lift + (size_t)f * nd * nfp + m;
          for (int a = 0; a < nd; ++a) acc[i][a] += L[a * nfp] * q;
        }
      }
    }
    if (c->absidx && c->absidx[e] >= 0) {
      const double* A = c->absmat + (size_t)c->absidx[e] * nd * nd;
      const double* U = uabs + e * nd * d;
      double ua[MAXND][MAXD];
      for (int b = 0; b < nd; ++b)
        for (int i = 0; i < d; ++i) ua[b][i] = U[b * d + i];
      for (int i = 0; i < d; ++i)
        for (int a = 0; a < nd; ++a) {
          double v = 0.0;
          for (int b = 0; b < nd; ++b) v += A[a * nd + b] * ua[b][i];
          acc[i][a] -= v;
        }
    }
    double* O = out + e * nd * d;
    if (ax0) {
      const double* A0 = ax0 + e * nd * d;
      const double* A1 = ax1 + e * nd * d;
      for (int a = 0; a < nd; ++a)
        for (int i = 0; i < d; ++i) O[a * d + i] = c0 * A0[a * d + i] + c1 * A1[a * d + i] + c2 * acc[i][a];
    } else {
      for (int a = 0; a < nd; ++a)
        for (int i = 0; i < d; ++i) O[a * d + i] = acc[i][a];
    }
  }
}

/* out_ij = c0*ax0_ij + c1*ax1_ij + c2*(lam delta_ij div + mu (G_ij + G_ji) + src_ij),  G_ij = d~_j u_i, own trace on
 * exterior facets (elastic.py:211-219).  ax0 == NULL: plain. */
void fused_pass_g(const fused_ctx* c, const double* u, const double* src, const double* ax0, const double* ax1,
                  double c0, double c1, double c2, double* out) {
  const int d = c->d, nd = c->nd, nfp = c->nfp, nf = c->d + 1, dd = d * d;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < c->E; ++e) {
    const double* U = u + e * nd * d;
    const double* Ji = c->jinv + e * dd;
    double G[MAXD][MAXD][MAXND]; /* G[i][j][a] */
    for (int i = 0; i < d; ++i) {
      double R[MAXD][MAXND];
      for (int r = 0; r < d; ++r)
        for (int a = 0; a < nd; ++a) {
          const double* D = c->Dr + ((size_t)r * nd + a) * nd;
          double v = 0.0;
          for (int b = 0; b < nd; ++b) v += D[b] * U[b * d + i];
          R[r][a] = v;
        }
      for (int f = 0; f < nf; ++f) {
        const unsigned cd = c->code[e * nf + f];
        if (cd & SG_BOUNDARY) continue; /* jump = 0 */
        const double* Un = u + (int64_t)c->nbr[e * nf + f] * nd * d;
        const int32_t* row = c->ftab + (size_t)(cd & 0x7f) * nfp;
        for (int m = 0; m < nfp; ++m) {
          const double jump = 0.5 * (Un[row[m] * d + i] - U[c->fnodes[f * nfp + m] * d + i]);
          const double* L = c->Lift + (size_t)f * nd * nfp + m;
          if (f == 0) {
            for (int a = 0; a < nd; ++a) {
              const double v = L[a * nfp] * jump;
              for (int r = 0; r < d; ++r) R[r][a] += v;
            }
          } else {
            for (int a = 0; a < nd; ++a) R[f - 1][a] -= L[a * nfp] * jump;
          }
        }
      }
      for (int j = 0; j < d; ++j)
        for (int a = 0; a < nd; ++a) {
          double v = 0.0;
          for (int r = 0; r < d; ++r) v += Ji[r * d + j] * R[r][a];
          G[i][j][a] = v;
        }
    }
    const double lam = c->lam[e], mu = c->mu[e];
    double* O = out + e * nd * dd;
    for (int a = 0; a < nd; ++a) {
      double div = 0.0;
      for (int k = 0; k < d; ++k) div += G[k][k][a];
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          double v = mu * (G[i][j][a] + G[j][i][a]);
          if (i == j) v += lam * div;
          const size_t o = (size_t)(a * d + i) * d + j;
          if (src) v += src[e * nd * dd + o];
          O[o] = ax0 ? c0 * ax0[e * nd * dd + o] + c1 * ax1[e * nd * dd + o] + c2 * v : v;
        }
    }
  }
}

/* One LF4 time step, six passes; u, s updated in place; uh, sh: scratch fields. */
void fused_step(const fused_ctx* c, double* u, double* s, const double* src, double dt, double* uh, double* sh) {
  const double c3 = dt * dt * dt / 24.0;
  fused_pass_f(c, s, u, NULL, NULL, 0, 0, 0, uh);              /* K1 uh1   = Dv(s0) - P(sigma, u0)          :292 */
  fused_pass_g(c, uh, src, NULL, NULL, 0, 0, 0, sh);           /* K2 stemp = Ds(uh1) + src                  :293 */
  fused_pass_f(c, sh, u, u, uh, c->density, dt, c3, u);        /* K3 u1    = rho u0 + dt uh1 + c3 uh2       :294-296 */
  fused_pass_g(c, u, src, NULL, NULL, 0, 0, 0, sh);            /* K4 sh1   = Ds(u1) + src                   :300 */
  fused_pass_f(c, sh, u, NULL, NULL, 0, 0, 0, uh);             /* K5 utemp = Dv(sh1) - P(sigma, u1)         :301 */
  fused_pass_g(c, uh, src, s, sh, 1.0, dt, c3, s);             /* K6 s1    = s0 + dt sh1 + c3 sh2           :302-304 */
}
