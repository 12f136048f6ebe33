// Fused stage kernels of the ElasticLF4 explicit update for sm_100a (FP64 on CUDA cores).
//
// One launch = one of the six field-producing passes of a time step:
//   F-type  (K1, K3, K5):  velocity RHS  f  of seigen/elastic.py:204-209  after the inverse mass
//   G-type  (K2, K4, K6):  stress   RHS  g  of seigen/elastic.py:211-219  after the inverse mass
// optionally fused with the LF4 combination of elastic.py:341-352 (AXPY variants K3, K6).
// Each pass reads its input field once (own cells through one TMA bulk copy per tile, facet
// neighbours from the same shared-memory tile or, across tile borders, from L2) and writes its output once.
//
// Device layout of every field ("tile-blocked SoA"):  value(cell e, row k) lives at
//     base[((e / TILE) * K + k) * TILE + (e % TILE)],   k = comp * ND + node,
// so a tile of TILE cells is one contiguous K*TILE*8-byte block (a single cp.async.bulk), a warp
// reading row k of its cells reads 256 contiguous bytes, and a neighbour's row is TILE doubles apart
// in shared and in global memory alike -- which lets one generic-address load serve both cases.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

template <int D, int P> struct ElemOps;   // specialised by the generated headers
#include "gen/elem_d1p1.cuh"
#include "gen/elem_d1p2.cuh"
#include "gen/elem_d1p3.cuh"
#include "gen/elem_d2p1.cuh"
#include "gen/elem_d2p2.cuh"
#include "gen/elem_d2p3.cuh"
#include "gen/elem_d2p4.cuh"
#include "gen/elem_d3p1.cuh"
#include "gen/elem_d3p2.cuh"
#include "gen/elem_d3p3.cuh"

namespace sg {

struct StageParams {
  const double* in;       // input field (S for F-type, U for G-type), owned + halo tiles
  double* out;            // output field
  const double* ax0;      // AXPY: c0 * ax0 (u0 / s0), own cells only (may alias out)
  const double* ax1;      // AXPY: c1 * ax1 (uh1 / sh1), own cells only
  const double* absu;     // F-type: velocity the sponge multiplies (u0 / u1); unused if absidx == nullptr
  const double* geo;      // [tile][D*D][TILE]  Jinv per cell, or nullptr when geometry classes are used
  const uint16_t* geoidx; // [tile][TILE] class of each cell (affine-congruent cells share one Jinv)
  const double* geotab;   // [nclass][D*D]
  int32_t nclass;
  const int32_t* nbr;     // [tile][NF][TILE]   neighbour cell (device index)
  const uint8_t* code;    // [tile][NF][TILE]
  const int32_t* absidx;  // [tile][TILE] row of absmat or -1; nullptr = no sponge anywhere
  const double* absmat;   // [ND*ND][nabs_pad]
  int64_t nabs_pad;
  const double* mat;      // per cell [tile][2][TILE] (lambda, mu) or nullptr
  double lam_c, mu_c;
  double c0, c1, c2;      // AXPY: out = c0*ax0 + c1*ax1 + c2*rhs
  // nodal source values added where g is evaluated (elastic.py:165, 183, 195, 285-288), G-type only
  const int32_t* src_start;  // [tiles + 1] first source entry of each tile, or nullptr
  const int32_t* src_off;    // [nsrc] offset inside the tile's K*TILE block
  const double* amp;         // [nsteps][nsrc]
  const int64_t* step;       // device-side step counter
  int64_t nsteps, nsrc;
  double src_scale;
  int32_t tile0;          // first tile of this launch
  int32_t ntiles;         // tiles of this launch
  unsigned int* sched;    // [4] dynamic tile scheduler of this launch: next ticket, CTAs finished, boundary tiles
                          // pushed (all 0 at launch)
  // Halo exchange fused into the pass (one process per GPU, peers' fields mapped through CUDA IPC).  Tiles
  // [0, push_tiles) hold the cut-adjacent cells: they are handed out first, the CTA that computes one waits for the
  // peers' rows of the previous exchange before it reads a halo cell, and afterwards stores the rows of its
  // cut-adjacent cells straight into the peers' halo tiles; whoever finishes the last of these tiles publishes the
  // new epoch to every peer.  push_tiles = 0: no exchange in this launch.
  int32_t push_tiles;
  int32_t npeers;
  int32_t K_out;                        // rows per cell of the output field
  const int32_t* push_start;            // [push_tiles + 1] first entry of each tile
  const int32_t* push_lane;             // [n] lane of the cell inside its tile
  const int32_t* push_peer;             // [n] index into rfield / rflag
  const int64_t* push_dst;              // [n] device cell index in the peer's copy of the field
  double* const* rfield;                // [npeers] the peers' copy of the output field
  unsigned long long* ctl;              // my control words (SG_CTL_*)
  unsigned long long* const* rflag;     // [npeers] my flag slot in each peer's control words
  long long timeout_cycles;
  int32_t dbg;                          // measurement only (SG_EXCHANGE_DEBUG): 1 skip the wait, 2 skip the remote stores,
                                        // 4 skip the system fence -- results are wrong, timings show what each part costs
  int32_t pdl_late;                     // 1: let the next kernel start being scheduled when this CTA is done with its
                                        // tiles instead of when it starts (programmatic dependent launch)
  int64_t* bump;                        // last pass of a step inside the step graph: the CTA that finishes last
                                        // advances the device-side step counter (nullptr otherwise)
};

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (global -> shared), sm_90+ PTX
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// L2 prefetch of a contiguous block (no shared memory, no completion to wait for)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory plan of one pipeline stage (= everything one tile needs, all of it brought in by bulk copies):
//   input tile | [ax0 | ax1] | nbr | code | geometry (class ids or Jinv) | [lambda, mu] | [absidx]
// and of the CTA-wide area behind the NS stages:  mbarriers + slot tile ids | facet node table | geometry class
// table | [X].  The same function sizes the launch on the host and lays the buffers out on the device.
// ---------------------------------------------------------------------------------------------
constexpr int GEO_SMEM_CLASSES = 64;

struct StagePlan {
  uint32_t in, ax0, ax1, nbr, code, geo, mat, abs, stage_bytes;   // offsets inside a stage
  uint32_t bars, ftab, gtab, x, total;                            // offsets from the start of shared memory
  uint32_t in_b, ax_b, nbr_b, code_b, geo_b, mat_b, abs_b;        // copy sizes (0 = not copied)
};

__host__ __device__ constexpr uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }
__host__ __device__ constexpr uint32_t up128(uint32_t x) { return (x + 127u) & ~127u; }

// Stored stress components.  SYM = false: all D*D of them, as the reference's TensorFunctionSpace does
// (elastic.py:81).  SYM = true: the upper triangle only -- Ds(u) is symmetric by construction (elastic.py:211-219),
// so when s0 and the source are symmetric every stress field of the step is, bit for bit, and (i,j) / (j,i) need
// not both travel through HBM (the library checks the premise, sg_set_state / sg_set_source -> SG_EASYM).
template <int D, bool SYM> __host__ __device__ constexpr int ncs() { return SYM ? D * (D + 1) / 2 : D * D; }
template <int D, bool SYM> __host__ __device__ constexpr int scomp(int i, int j) {
  if (!SYM) return i * D + j;
  const int a = i < j ? i : j, b = i < j ? j : i;
  return a * D - a * (a - 1) / 2 + (b - a);   // row-major upper triangle
}

template <int D, int ND, int TILE, int NS, bool GTYPE, bool AXS, bool XS = GTYPE, bool SYM = false>
__host__ __device__ inline StagePlan make_plan(bool classes, bool per_cell_mat, bool sponge, int ftab_size) {
  constexpr uint32_t KS = ncs<D, SYM>() * ND, KX = D * D * ND, KU = D * ND, NF = D + 1;
  constexpr uint32_t KIN = GTYPE ? KU : KS, KAX = GTYPE ? KS : KU;
  StagePlan p{};
  uint32_t o = 0;
  p.in_b = KIN * TILE * 8;
  p.in = o; o += p.in_b;
  p.ax_b = AXS ? KAX * TILE * 8 : 0;
  p.ax0 = o; o += p.ax_b;
  p.ax1 = o; o += p.ax_b;
  p.nbr_b = NF * TILE * 4;
  p.nbr = o; o += p.nbr_b;
  p.code_b = up16(NF * TILE);
  p.code = o; o += p.code_b;
  p.geo_b = classes ? up16(TILE * 2) : (uint32_t)(D * D * TILE * 8);
  p.geo = o; o += p.geo_b;
  p.mat_b = (GTYPE && per_cell_mat) ? 2 * TILE * 8 : 0;
  p.mat = o; o += p.mat_b;
  p.abs_b = (!GTYPE && sponge) ? up16(TILE * 4) : 0;
  p.abs = o; o += p.abs_b;
  p.stage_bytes = up128(o);
  o = p.stage_bytes * NS;
  p.bars = o; o += up16(8 * NS + 4 * NS);   // mbarriers, then the tile id of each slot
  p.ftab = o; o += up16((uint32_t)ftab_size);
  p.gtab = o; o += classes ? (uint32_t)(GEO_SMEM_CLASSES * D * D * 8) : 0;
  o = up128(o);
  p.x = o; o += (GTYPE && XS) ? KX * TILE * 8 : 0;
  p.total = o;
  return p;
}

// ---------------------------------------------------------------------------------------------
// contexts handed to the generated contractions
// ---------------------------------------------------------------------------------------------
template <int D, int ND, int NFP, int TILE> struct FaceGeom {
  double ji[D][D];
  int nb[D + 1];
  unsigned cd[D + 1];
};

// F-type: row i of  Dv(s)_i = sum_j d~_j s_ij   (free-surface trace on exterior facets: s^ = 0)
template <int D, int ND, int NFP, int TILE, int KS, bool SYM = false> struct FCtx {
  const double* own;          // &sIn[lane]
  const double* tileS;        // sIn (shared)
  const double* gIn;          // global input field
  const unsigned char* sft;   // neighbour node table (shared)
  const FaceGeom<D, ND, NFP, TILE>* g;
  int tile;
  int coff[D];                // coff[j] = row of s_ij (node 0) * TILE, for this thread's i
  // per-facet state
  const double* nbp;
  const unsigned char* row;
  double gn[D], go[D];        // facet direction times the weight of the neighbour's / the own trace
  // qall of the largest all-rows kernel (3D P2, 255 registers) forms the weighted jump of each stored component first:
  // NC values live instead of 2 NC, 3 flops more per facet node, K1 130 -> 124 us (profiles/r02_experiment_qall_forms.log);
  // everywhere else the folded weights are as fast or faster
  static constexpr bool QALL_DELTA = (D == 3 && D * ND >= 30);
  double gf[D], cn1, co1;     // unfolded facet direction and weights (QALL_DELTA)

  __device__ __forceinline__ void t(int b, double* t) const {
    double s[D];
#pragma unroll
    for (int j = 0; j < D; ++j) s[j] = own[coff[j] + b * TILE];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double a = g->ji[r][0] * s[0];
#pragma unroll
      for (int j = 1; j < D; ++j) a = fma(g->ji[r][j], s[j], a);
      t[r] = a;
    }
  }
  __device__ __forceinline__ void face(int f) {
    const int n = g->nb[f];
    const unsigned c = g->cd[f];
    const bool bnd = (c & 0x80u) != 0;
    // numerical trace: (s_nbr + s_own)/2 on interior facets, 0 on exterior ones (free surface); the facet term is
    // gf . (trace - s_own) = gn . s_nbr + go . s_own with the two weights folded into the facet direction once per facet
    const double cn = bnd ? 0.0 : 0.5, co = bnd ? -1.0 : -0.5;
    row = sft + (c & 0x7fu) * NFP;
    const int nt = n / TILE, nl = n % TILE;
    nbp = (nt == tile) ? (tileS + nl) : (gIn + (size_t)nt * (KS * TILE) + nl);
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double gfj;
      if (f == 0) {
        double a = g->ji[0][j];
#pragma unroll
        for (int r = 1; r < D; ++r) a += g->ji[r][j];
        gfj = a;
      } else {
        gfj = -g->ji[f - 1][j];
      }
      gn[j] = cn * gfj;
      go[j] = co * gfj;
      if (QALL_DELTA) gf[j] = gfj;
    }
    if (QALL_DELTA) {
      cn1 = cn;
      co1 = co;
    }
  }
  __device__ __forceinline__ double q(int on, int m) const {
    const int nn = row[m];
    double acc = go[0] * own[coff[0] + on * TILE];
#pragma unroll
    for (int j = 1; j < D; ++j) acc = fma(go[j], own[coff[j] + on * TILE], acc);
#pragma unroll
    for (int j = 0; j < D; ++j) acc = fma(gn[j], nbp[coff[j] + nn * TILE], acc);
    return acc;
  }
  // all D rows at once (one thread per cell): each stored stress component is read once
  __device__ __forceinline__ void tall(int b, double (*t)[D]) const {
    constexpr int NC = ncs<D, SYM>();
    double s[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) s[k] = own[(k * ND + b) * TILE];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int r = 0; r < D; ++r) {
        double a = g->ji[r][0] * s[scomp<D, SYM>(i, 0)];
#pragma unroll
        for (int j = 1; j < D; ++j) a = fma(g->ji[r][j], s[scomp<D, SYM>(i, j)], a);
        t[i][r] = a;
      }
  }
  __device__ __forceinline__ void qall(int on, int m, double* q) const {
    constexpr int NC = ncs<D, SYM>();
    const int nn = row[m];
    if (QALL_DELTA) {
      double dlt[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) dlt[k] = fma(cn1, nbp[(k * ND + nn) * TILE], co1 * own[(k * ND + on) * TILE]);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = gf[0] * dlt[scomp<D, SYM>(i, 0)];
#pragma unroll
        for (int j = 1; j < D; ++j) acc = fma(gf[j], dlt[scomp<D, SYM>(i, j)], acc);
        q[i] = acc;
      }
      return;
    }
    double vo[NC], vn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      vo[k] = own[(k * ND + on) * TILE];
      vn[k] = nbp[(k * ND + nn) * TILE];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double acc = go[0] * vo[scomp<D, SYM>(i, 0)];
#pragma unroll
      for (int j = 1; j < D; ++j) acc = fma(go[j], vo[scomp<D, SYM>(i, j)], acc);
#pragma unroll
      for (int j = 0; j < D; ++j) acc = fma(gn[j], vn[scomp<D, SYM>(i, j)], acc);
      q[i] = acc;
    }
  }
};

// G-type: reference gradients of u_i (own trace on exterior facets: jump = 0)
template <int D, int ND, int NFP, int TILE> struct GCtx {
  const double* own;          // &sIn[(i*ND)*TILE + lane]
  const double* tileU;
  const double* gIn;
  const unsigned char* sft;
  const FaceGeom<D, ND, NFP, TILE>* g;
  int tile;
  int rowoff;                 // i*ND*TILE
  const double* nbp;
  const unsigned char* row;
  double cn;

  __device__ __forceinline__ double v(int b) const { return own[b * TILE]; }
  __device__ __forceinline__ void face(int f) {
    const int n = g->nb[f];
    const unsigned c = g->cd[f];
    cn = (c & 0x80u) ? 0.0 : 0.5;
    row = sft + (c & 0x7fu) * NFP;
    const int nt = n / TILE, nl = n % TILE;
    const double* base = (nt == tile) ? (tileU + nl) : (gIn + (size_t)nt * (D * ND * TILE) + nl);
    nbp = base + rowoff;
  }
  __device__ __forceinline__ double jump(int on, int m) const {
    const int nn = row[m];
    return cn * (nbp[nn * TILE] - own[on * TILE]);
  }
};

// ---------------------------------------------------------------------------------------------
// producer side: one thread starts every bulk copy of a tile; they all complete on the stage's mbarrier
// ---------------------------------------------------------------------------------------------
template <int D, int ND, int TILE, bool GTYPE, bool SYM>
__device__ __forceinline__ void issue_tile(const StageParams& p, const StagePlan& pl, unsigned char* stage,
                                           uint64_t* bar, int tile) {
  constexpr int KS = ncs<D, SYM>() * ND, KU = D * ND, NF = D + 1;
  constexpr int KIN = GTYPE ? KU : KS, KAX = GTYPE ? KS : KU;
  const uint32_t total = pl.in_b + 2 * pl.ax_b + pl.nbr_b + pl.code_b + pl.geo_b + pl.mat_b + pl.abs_b;
  mbar_expect_tx(bar, total);
  bulk_g2s(stage + pl.in, p.in + (size_t)tile * (KIN * TILE), pl.in_b, bar);
  if (pl.ax_b) {
    bulk_g2s(stage + pl.ax0, p.ax0 + (size_t)tile * (KAX * TILE), pl.ax_b, bar);
    bulk_g2s(stage + pl.ax1, p.ax1 + (size_t)tile * (KAX * TILE), pl.ax_b, bar);
  } else if (p.ax0 != nullptr) {
    // AXPY operands read straight from global by the compute threads (large elements: no shared memory left to
    // stage them): start them towards L2 now, a pipeline stage ahead of their use
    bulk_prefetch_l2(p.ax0 + (size_t)tile * (KAX * TILE), KAX * TILE * 8);
    bulk_prefetch_l2(p.ax1 + (size_t)tile * (KAX * TILE), KAX * TILE * 8);
  }
  bulk_g2s(stage + pl.nbr, p.nbr + (size_t)tile * (NF * TILE), pl.nbr_b, bar);
  bulk_g2s(stage + pl.code, p.code + (size_t)tile * pl.code_b, pl.code_b, bar);
  if (p.geoidx != nullptr)
    bulk_g2s(stage + pl.geo, p.geoidx + (size_t)tile * (pl.geo_b / 2), pl.geo_b, bar);
  else
    bulk_g2s(stage + pl.geo, p.geo + (size_t)tile * (D * D * TILE), pl.geo_b, bar);
  if (pl.mat_b) bulk_g2s(stage + pl.mat, p.mat + (size_t)tile * (2 * TILE), pl.mat_b, bar);
  if (pl.abs_b) bulk_g2s(stage + pl.abs, p.absidx + (size_t)tile * (pl.abs_b / 4), pl.abs_b, bar);
}

// Programmatic dependent launch (sm_90+): the stage kernels of a step are chained with programmatic edges, so the
// launch and the CTA prologue of pass k+1 (barrier init, table copies) overlap the tail of pass k.  pdl_launch lets the
// next kernel in the stream start being scheduled once every CTA of this grid has started; pdl_wait blocks until the
// previous kernel has completed and its writes are visible (a no-op for a kernel launched without the attribute).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Dynamic tile scheduler.  CTA b starts with tile b; further tiles are tickets gridDim.x + atomicAdd(next).  A
// ticket is requested one iteration before its bulk copies are issued, so the atomic's latency is never waited
// for.  The last CTA to finish zeroes both words for the next launch on the stream.
__device__ __forceinline__ int sched_next(unsigned int* sched) { return (int)(gridDim.x + atomicAdd(sched, 1u)); }
__device__ __forceinline__ void sched_done(unsigned int* sched, int64_t* bump = nullptr) {
  __threadfence();
  if (atomicAdd(sched + 1, 1u) == gridDim.x - 1) {
    sched[0] = 0u;
    sched[1] = 0u;
    if (bump != nullptr) *bump += 1;
    __threadfence();
  }
}

template <int D, int ND, int NFP, int TILE>
__device__ __forceinline__ void load_geom(const StageParams& p, const StagePlan& pl, const unsigned char* stage,
                                          const double* gtab, int lane, FaceGeom<D, ND, NFP, TILE>& g) {
  constexpr int NF = D + 1;
  if (p.geoidx != nullptr) {
    const int cls = reinterpret_cast<const uint16_t*>(stage + pl.geo)[lane];
    const double* gt = (p.nclass <= GEO_SMEM_CLASSES) ? gtab + cls * (D * D) : p.geotab + (size_t)cls * (D * D);
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = 0; k < D; ++k) g.ji[r][k] = gt[r * D + k];
  } else {
    const double* geo = reinterpret_cast<const double*>(stage + pl.geo) + lane;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = 0; k < D; ++k) g.ji[r][k] = geo[(r * D + k) * TILE];
  }
  const int32_t* nb = reinterpret_cast<const int32_t*>(stage + pl.nbr) + lane;
  const uint8_t* cd = stage + pl.code + lane;
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    g.nb[f] = nb[f * TILE];
    g.cd[f] = cd[f * TILE];
  }
}

// CTA-wide setup shared by both kernels: barriers, facet node table, geometry class table
template <int NS, int NTHREADS>
__device__ __forceinline__ void cta_setup(const StageParams& p, const StagePlan& pl, unsigned char* smem,
                                          const unsigned char* ftab, int ftab_size, int dd) {
  const int tid = threadIdx.x;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bars);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(bars + s, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < ftab_size; i += NTHREADS) smem[pl.ftab + i] = ftab[i];
  if (p.geoidx != nullptr && p.nclass <= GEO_SMEM_CLASSES) {
    double* gt = reinterpret_cast<double*>(smem + pl.gtab);
    for (int i = tid; i < p.nclass * dd; i += NTHREADS) gt[i] = p.geotab[i];
  }
  __syncthreads();
}

// Producer prologue / per-iteration step (thread 0 only).  slot_tile[s] = tile held or being filled by slot s.
#define SG_PIPE_PROLOGUE(GT)                                                                               \
  int* slot_tile = reinterpret_cast<int*>(smem + pl.bars + 8 * NS);                                        \
  int t_ahead = 0;                                                                                         \
  if (tid == 0) {                                                                                          \
    int t = blockIdx.x;                                                                                    \
    _Pragma("unroll") for (int s = 0; s < NS - 1; ++s) {                                                   \
      slot_tile[s] = t;                                                                                    \
      if (t < p.ntiles) issue_tile<D, ND, TILE, GT, SYM>(p, pl, smem + s * pl.stage_bytes, bars + s, p.tile0 + t); \
      t = sched_next(p.sched);                                                                             \
    }                                                                                                      \
    t_ahead = t;                                                                                           \
  }                                                                                                        \
  __syncthreads();   /* slot_tile[0 .. NS-2] visible to every warp */

#define SG_PIPE_ADVANCE(GT)                                                                                \
  if (tid == 0) {                                                                                          \
    const int sn = (it + NS - 1) % NS;                                                                     \
    slot_tile[sn] = t_ahead;                                                                               \
    if (t_ahead < p.ntiles) {                                                                              \
      issue_tile<D, ND, TILE, GT, SYM>(p, pl, smem + sn * pl.stage_bytes, bars + sn, p.tile0 + t_ahead);        \
      t_ahead = sched_next(p.sched);                                                                       \
    }                                                                                                      \
  }                                                                                                        \
  if (NS == 1) __syncthreads();

// ---------------------------------------------------------------------------------------------
// Halo exchange inside a pass (StageParams::push_*).  ctl layout (uint64): [0, 16) flags written by my peers,
// [16] exchanges I have published, [18] error word (1 = a wait timed out).  Every rank issues the same sequence of
// exchanges, so "the peers' rows of the previous exchange have landed" == flag[p] >= number of exchanges I have
// published myself.
// ---------------------------------------------------------------------------------------------
constexpr int SG_CTL_SENT = 16, SG_CTL_ERROR = 18, SG_CTL_WORDS = 32;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Spin (bounded) until flag slot `slot` of ctl has reached `epoch`; raises the error word instead of hanging.
__device__ __forceinline__ void wait_flag(unsigned long long* ctl, int slot, unsigned long long epoch,
                                          long long timeout_cycles) {
  const unsigned long long* f = ctl + slot;
  if (ld_acquire_sys(f) >= epoch) return;
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < epoch) {
    if (clock64() - t0 > timeout_cycles) {
      ctl[SG_CTL_ERROR] = 1;
      break;
    }
    __nanosleep(32);
  }
}

// Before a CTA reads halo cells: the peers' rows of the last exchange must have landed.  Called by all threads.
__device__ __forceinline__ void halo_wait(const StageParams& p) {
  if ((int)threadIdx.x < p.npeers && !(p.dbg & 1)) {
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(p.ctl + SG_CTL_SENT);
    wait_flag(p.ctl, threadIdx.x, epoch, p.timeout_cycles);
  }
  __syncthreads();
}

// After a CTA has stored boundary tile `tile` (and synchronised): copy the rows of its cut-adjacent cells into the
// peers' halo tiles (entries sorted by peer and remote cell: consecutive threads write consecutive remote lanes),
// then count the tile; the CTA that completes the last one publishes the epoch.  Called by all threads.
template <int TILE, int NT>
__device__ __forceinline__ void halo_push(const StageParams& p, int tile) {
  const int lo = p.push_start[tile], n = p.push_start[tile + 1] - lo;
  const int K = p.K_out;
  const double* src = p.out + (size_t)tile * K * TILE;
  // one work item = (cell entry, batch of HB rows): the rows of a batch are loaded together and then stored together,
  // so an item costs one round trip to L2 instead of HB (this loop sits on the exchange's critical path: the
  // neighbour's next pass waits for these rows)
  constexpr int HB = 8;
  const int nb = (K + HB - 1) / HB;
  for (int idx = threadIdx.x; idx < n * nb; idx += NT) {
    const int c = lo + idx % n, k0 = (idx / n) * HB;
    const int64_t r = p.push_dst[c];
    double* dst = p.rfield[p.push_peer[c]] + ((r / TILE) * K + k0) * TILE + r % TILE;
    const double* s = src + k0 * TILE + p.push_lane[c];
    double v[HB];
#pragma unroll
    for (int k = 0; k < HB; ++k)
      if (k0 + k < K) v[k] = s[k * TILE];
#pragma unroll
    for (int k = 0; k < HB; ++k)
      if (k0 + k < K && !(p.dbg & 2)) dst[k * TILE] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!(p.dbg & 4)) __threadfence_system();                 // this CTA's remote rows before the count / the flag
    if (atomicAdd(p.sched + 2, 1u) == (unsigned)p.push_tiles - 1u) {
      p.sched[2] = 0u;
      // every other boundary CTA fenced its rows before it counted; the release store below orders this thread's
      // observation of the count (and with it their rows) before the flag
      const unsigned long long epoch = p.ctl[SG_CTL_SENT] + 1;
      p.ctl[SG_CTL_SENT] = epoch;
      for (int q = 0; q < p.npeers; ++q) st_release_sys(p.rflag[q], epoch);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// F-type pass:   out_i = Dv(in)_i - A_cell * absu_i            (K1, K5)
//                out_i = c0*ax0_i + c1*ax1_i + c2*(Dv(in)_i - A_cell*absu_i)   (K3, AXPY)
// Persistent CTAs with a dynamic tile scheduler; the bulk copies of the next NS-1 tiles are in flight while a
// tile is computed.
// ---------------------------------------------------------------------------------------------
template <int D, int P, int TILE, int SPLIT, int MINB, int NS, bool AXPY, bool AXS, bool SYM>
__global__ void __launch_bounds__(TILE* SPLIT, MINB) stage_f_kernel(const StageParams p) {
  using E = ElemOps<D, P>;
  constexpr int ND = E::ND, NFP = E::NFP, KU = D * ND, KS = ncs<D, SYM>() * ND, IPT = D / SPLIT, NT = TILE * SPLIT;
  static_assert(D % SPLIT == 0, "SPLIT must divide D");
  extern __shared__ __align__(128) unsigned char smem[];
  const StagePlan pl = make_plan<D, ND, TILE, NS, false, (AXPY && AXS), false, SYM>(
      p.geoidx != nullptr, false, p.absidx != nullptr, E::FTAB_SIZE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bars);
  const unsigned char* sft = smem + pl.ftab;
  const double* gtab = reinterpret_cast<const double*>(smem + pl.gtab);
  if (!p.pdl_late) pdl_launch_dependents();
  cta_setup<NS, NT>(p, pl, smem, E::ftab(), E::FTAB_SIZE, D * D);
  pdl_wait();   // everything above reads tables that no kernel writes; from here on the previous pass's output is read

  const int tid = threadIdx.x, lane = tid % TILE, ig = tid / TILE;
  SG_PIPE_PROLOGUE(false)
  for (int it = 0;; ++it) {
    const int slot = it % NS;
    SG_PIPE_ADVANCE(false)
    const int t = slot_tile[slot];
    if (t >= p.ntiles) break;
    const int tile = p.tile0 + t;
    unsigned char* stage = smem + slot * pl.stage_bytes;
    const double* sIn = reinterpret_cast<const double*>(stage + pl.in);
    const bool btile = tile < p.push_tiles;
    if (btile) halo_wait(p);
    mbar_wait(bars + slot, (it / NS) & 1);

    FaceGeom<D, ND, NFP, TILE> g;
    load_geom<D, ND, NFP, TILE>(p, pl, stage, gtab, lane, g);
    int aidx = -1;
    if (pl.abs_b) aidx = reinterpret_cast<const int32_t*>(stage + pl.abs)[lane];

    FCtx<D, ND, NFP, TILE, KS, SYM> c;
    c.tileS = sIn;
    c.gIn = p.in;
    c.sft = sft;
    c.g = &g;
    c.tile = tile;
    c.own = sIn + lane;
    // One thread per cell with small elements: all D rows are computed together (volF_all / liftF_all), so that a
    // stress component two rows need (s_ij = s_ji with symmetric storage) is gathered once -- 9 instead of 12
    // neighbour loads per facet node pair in 2D, 6 instead of 9 in 3D.
    constexpr bool ROWS_FIRST = (SPLIT == 1) && (D * ND <= 30);
    double accs[ROWS_FIRST ? IPT : 1][ND];
    auto compute = [&](int ii, double* acc) {
      const int i = ig * IPT + ii;
#pragma unroll
      for (int j = 0; j < D; ++j) c.coff[j] = scomp<D, SYM>(i, j) * (ND * TILE);
#pragma unroll
      for (int a = 0; a < ND; ++a) acc[a] = 0.0;
      E::volF(c, acc);
      E::liftF(c, acc);
    };
    auto finish = [&](int ii, double* acc) {
      const int i = ig * IPT + ii;
      const size_t orow = ((size_t)tile * KU + i * ND) * TILE + lane;
      if (aidx >= 0) {
        // sponge: - Minv * int phi_a (sigma u_i)   (elastic.py:207-208), A precomputed per sponge cell
        double ua[ND];
#pragma unroll
        for (int b = 0; b < ND; ++b) ua[b] = p.absu[orow + (size_t)b * TILE];
        const double* A = p.absmat + aidx;
#pragma unroll
        for (int a = 0; a < ND; ++a)
#pragma unroll
          for (int b = 0; b < ND; ++b) acc[a] = fma(-A[(size_t)(a * ND + b) * p.nabs_pad], ua[b], acc[a]);
      }
      if (AXPY && AXS) {
        const double* a0 = reinterpret_cast<const double*>(stage + pl.ax0) + i * ND * TILE + lane;
        const double* a1 = reinterpret_cast<const double*>(stage + pl.ax1) + i * ND * TILE + lane;
#pragma unroll
        for (int a = 0; a < ND; ++a)
          p.out[orow + (size_t)a * TILE] = fma(p.c0, a0[a * TILE], fma(p.c1, a1[a * TILE], p.c2 * acc[a]));
      } else if (AXPY) {
        // operands straight from global memory (L2-prefetched by issue_tile).  out aliases ax0 (u is updated in
        // place), so the compiler may not move a load above an earlier store: load a batch first, then store it
        constexpr int CH = 10;
#pragma unroll
        for (int a0 = 0; a0 < ND; a0 += CH) {
          double x0[CH], x1[CH];
#pragma unroll
          for (int k = 0; k < CH; ++k)
            if (a0 + k < ND) {
              x0[k] = p.ax0[orow + (size_t)(a0 + k) * TILE];
              x1[k] = p.ax1[orow + (size_t)(a0 + k) * TILE];
            }
#pragma unroll
          for (int k = 0; k < CH; ++k)
            if (a0 + k < ND) p.out[orow + (size_t)(a0 + k) * TILE] = fma(p.c0, x0[k], fma(p.c1, x1[k], p.c2 * acc[a0 + k]));
        }
      } else {
#pragma unroll
        for (int a = 0; a < ND; ++a) p.out[orow + (size_t)a * TILE] = acc[a];
      }
    };
    if (ROWS_FIRST) {
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii)
#pragma unroll
        for (int a = 0; a < ND; ++a) accs[ii][a] = 0.0;
      E::volF_all(c, accs);
      E::liftF_all(c, accs);
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) finish(ii, accs[ii]);
    } else {
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        compute(ii, accs[0]);
        finish(ii, accs[0]);
      }
    }
    __syncthreads();   // the stage may be refilled from the next iteration on
    if (btile) halo_push<TILE, NT>(p, tile);
  }
  if (p.pdl_late) pdl_launch_dependents();
  if (tid == 0) sched_done(p.sched, p.bump);
}

// ---------------------------------------------------------------------------------------------
// G-type pass:   out_ij = lam*delta_ij*div + mu*(G_ij + G_ji) [+ src],  G_ij = d~_j in_i      (K2, K4)
//                out_ij = c0*ax0_ij + c1*ax1_ij + c2*(...)                                   (K6, AXPY)
// XREG (only with SPLIT == 1): the gradients G_ij stay in registers instead of the shared X buffer.
// ---------------------------------------------------------------------------------------------
template <int D, int P, int TILE, int SPLIT, int MINB, int NS, bool AXPY, bool AXS, bool XREG, bool SYM>
__global__ void __launch_bounds__(TILE* SPLIT, MINB) stage_g_kernel(const StageParams p) {
  using E = ElemOps<D, P>;
  constexpr int ND = E::ND, NFP = E::NFP, KS = ncs<D, SYM>() * ND, KX = D * D * ND, IPT = D / SPLIT, NT = TILE * SPLIT;
  static_assert(!XREG || SPLIT == 1, "register-resident gradients need one thread per cell");
  extern __shared__ __align__(128) unsigned char smem[];
  const StagePlan pl = make_plan<D, ND, TILE, NS, true, (AXPY && AXS), !XREG, SYM>(
      p.geoidx != nullptr, p.mat != nullptr, false, E::FTAB_SIZE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.bars);
  const unsigned char* sft = smem + pl.ftab;
  const double* gtab = reinterpret_cast<const double*>(smem + pl.gtab);
  double* sX = reinterpret_cast<double*>(smem + pl.x);
  if (!p.pdl_late) pdl_launch_dependents();
  cta_setup<NS, NT>(p, pl, smem, E::ftab(), E::FTAB_SIZE, D * D);
  pdl_wait();   // (see stage_f_kernel)

  const int tid = threadIdx.x, lane = tid % TILE, ig = tid / TILE;
  SG_PIPE_PROLOGUE(true)
  for (int it = 0;; ++it) {
    const int slot = it % NS;
    SG_PIPE_ADVANCE(true)
    const int t = slot_tile[slot];
    if (t >= p.ntiles) break;
    const int tile = p.tile0 + t;
    // source entries of this tile: requested now, looked at after the tile has been computed and stored
    int s_lo = 0, s_hi = 0;
    if (p.src_start != nullptr) {
      s_lo = p.src_start[tile];
      s_hi = p.src_start[tile + 1];
    }
    unsigned char* stage = smem + slot * pl.stage_bytes;
    const double* sIn = reinterpret_cast<const double*>(stage + pl.in);
    const bool btile = tile < p.push_tiles;
    if (btile) halo_wait(p);
    mbar_wait(bars + slot, (it / NS) & 1);

    FaceGeom<D, ND, NFP, TILE> g;
    load_geom<D, ND, NFP, TILE>(p, pl, stage, gtab, lane, g);

    GCtx<D, ND, NFP, TILE> c;
    c.tileU = sIn;
    c.gIn = p.in;
    c.sft = sft;
    c.g = &g;
    c.tile = tile;
    double XR[XREG ? KX : 1];
#pragma unroll
    for (int ii = 0; ii < IPT; ++ii) {
      const int i = ig * IPT + ii;
      c.rowoff = i * ND * TILE;
      c.own = sIn + c.rowoff + lane;
      double R[D * ND];
#pragma unroll
      for (int a = 0; a < D * ND; ++a) R[a] = 0.0;
      E::volG(c, R);
      E::liftG(c, R);
      double* X = sX + (size_t)(i * D) * ND * TILE + lane;
#pragma unroll
      for (int j = 0; j < D; ++j)
#pragma unroll
        for (int a = 0; a < ND; ++a) {
          double v = g.ji[0][j] * R[a];
#pragma unroll
          for (int r = 1; r < D; ++r) v = fma(g.ji[r][j], R[r * ND + a], v);
          if (XREG)
            XR[(i * D + j) * ND + a] = v;
          else
            X[(j * ND + a) * TILE] = v;
        }
    }
    if (SPLIT > 1) __syncthreads();

    double lam = p.lam_c, mu = p.mu_c;
    if (pl.mat_b) {
      const double* m = reinterpret_cast<const double*>(stage + pl.mat) + lane;
      lam = m[0];
      mu = m[TILE];
    }
    const double* X = sX + lane;
    double* out = p.out + (size_t)tile * (KS * TILE);
#pragma unroll
    for (int ii = 0; ii < IPT; ++ii) {
      const int i = ig * IPT + ii;
      // AXPY operands straight from global memory (large elements; L2-prefetched by issue_tile): out aliases ax0
      // (s is updated in place), so the compiler may not move a load above an earlier store -- load the operands of
      // CH nodes first, then compute and store them
      constexpr bool DIRECT = AXPY && !AXS;
      constexpr int CH = 5;
      double x0[DIRECT ? CH : 1][D], x1[DIRECT ? CH : 1][D];
#pragma unroll
      for (int a = 0; a < ND; ++a) {
        if (DIRECT && a % CH == 0) {
#pragma unroll
          for (int k = 0; k < CH; ++k)
#pragma unroll
            for (int jj = 0; jj < D; ++jj) {
              const int j = SYM ? (i + jj) % D : jj;
              if (a + k >= ND || (SYM && !(D == 2 ? (jj == 0 || i == 0) : (jj < 2)))) continue;
              const size_t og = (size_t)tile * (KS * TILE) + (scomp<D, SYM>(i, j) * ND + a + k) * TILE + lane;
              x0[DIRECT ? k : 0][jj] = p.ax0[og];
              x1[DIRECT ? k : 0][jj] = p.ax1[og];
            }
        }
        double div = XREG ? XR[a] : X[(0 * ND + a) * TILE];
#pragma unroll
        for (int k = 1; k < D; ++k) div += XREG ? XR[(k * D + k) * ND + a] : X[((k * D + k) * ND + a) * TILE];
        const double ld = lam * div;
#pragma unroll
        for (int jj = 0; jj < D; ++jj) {
          // SYM: row i produces (i,i) and (i,i+1 mod D) [D = 3] / row 0 produces (0,0), (0,1), row 1 (1,1) [D = 2]
          const int j = SYM ? (i + jj) % D : jj;
          if (SYM && !(D == 2 ? (jj == 0 || i == 0) : (jj < 2))) continue;
          double v = XREG ? mu * (XR[(i * D + j) * ND + a] + XR[(j * D + i) * ND + a])
                          : mu * (X[((i * D + j) * ND + a) * TILE] + X[((j * D + i) * ND + a) * TILE]);
          if (j == i) v += ld;
          const int o = (scomp<D, SYM>(i, j) * ND + a) * TILE + lane;
          if (AXPY && AXS) {
            const double* a0 = reinterpret_cast<const double*>(stage + pl.ax0);
            const double* a1 = reinterpret_cast<const double*>(stage + pl.ax1);
            v = fma(p.c0, a0[o], fma(p.c1, a1[o], p.c2 * v));
          } else if (AXPY) {
            v = fma(p.c0, x0[DIRECT ? a % CH : 0][jj], fma(p.c1, x1[DIRECT ? a % CH : 0][jj], p.c2 * v));
          }
          out[o] = v;
        }
      }
    }
    __syncthreads();   // stage and X free again; this tile's stores are ordered before the source update below
    if (s_hi > s_lo) {
      const int64_t st = *p.step;
      if (st >= 0 && st < p.nsteps)
        for (int k = s_lo + tid; k < s_hi; k += NT) out[p.src_off[k]] += p.src_scale * p.amp[st * p.nsrc + k];
    }
    if (btile) {
      __syncthreads();   // source values are part of what travels
      halo_push<TILE, NT>(p, tile);
    }
  }
  if (p.pdl_late) pdl_launch_dependents();
  if (tid == 0) sched_done(p.sched, p.bump);
}

#ifndef SG_STAGE_KERNELS_ONLY   // the instantiation units (sg_inst_*.cu) only need the stage kernels
// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------
// boundary (AoS, cell-major: [cell][node][comp]) <-> device (tile-blocked SoA: [tile][comp*ND+node][lane]).
// symd = 0: the device stores the ncomp boundary components as they are.  symd = D > 0 (stress fields of a
// symmetric-storage solver): the boundary has D*D components, the device the upper triangle; on the way in the
// two halves are compared and *asym is raised if they differ, on the way out (j,i) is filled from (i,j).
template <bool TO_DEVICE>
__global__ void relayout_kernel(double* __restrict__ dev, double* __restrict__ host_order, int64_t ncell,
                                int64_t n_owned, int64_t n_owned_pad, int nd, int ncomp, int tile, int symd,
                                unsigned int* __restrict__ asym) {
  const int KH = nd * ncomp;                                        // boundary doubles per cell
  const int K = symd ? nd * (symd * (symd + 1) / 2) : KH;           // device rows per cell
  const int64_t total = ncell * KH;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = idx / KH;
    const int r = (int)(idx % KH);
    const int node = r / ncomp, comp = r % ncomp;
    int dcomp = comp;
    bool lower = false;
    if (symd) {
      const int i = comp / symd, j = comp % symd;
      const int a = i < j ? i : j, b = i < j ? j : i;
      dcomp = a * symd - a * (a - 1) / 2 + (b - a);
      lower = i > j;
    }
    const int64_t k = dcomp * nd + node;
    const int64_t e = cell < n_owned ? cell : cell - n_owned + n_owned_pad;
    const int64_t d = ((e / tile) * K + k) * tile + e % tile;
    if (TO_DEVICE) {
      const double v = host_order[idx];
      if (!lower) {
        dev[d] = v;
      } else {
        const double w = host_order[idx - comp + (comp % symd) * symd + comp / symd];   // the (j,i) entry
        if (v != w && !(v != v && w != w)) *asym = 1u;
      }
    } else {
      host_order[idx] = dev[d];
    }
  }
}

// device-side step counter: indexes the source table so that the step graph replays without host arguments
__global__ void bump_step_kernel(int64_t* step) { *step += 1; }

// receivers (the sensors of tests/explosive_source/uy.py:36-43): velocity at fixed points, sampled after every step.
// rec[step][r][comp] = sum_a w[r][a] * u[cell_r, a, comp];  one thread per (receiver, component)
__global__ void receivers_kernel(const double* __restrict__ u, const int64_t* __restrict__ cell,
                                 const double* __restrict__ w, double* __restrict__ rec,
                                 const int64_t* __restrict__ step, int64_t max_steps, int n, int nd, int d, int tile) {
  const int64_t st = *step;
  if (st < 0 || st >= max_steps) return;
  const int K = nd * d;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * d; idx += gridDim.x * blockDim.x) {
    const int r = idx / d, comp = idx % d;
    const int64_t e = cell[r];
    const double* base = u + ((e / tile) * K + comp * nd) * tile + e % tile;
    double acc = 0.0;
    for (int a = 0; a < nd; ++a) acc = fma(w[r * nd + a], base[(int64_t)a * tile], acc);
    rec[(st * n + r) * d + comp] = acc;
  }
}
__global__ void set_step_kernel(int64_t* step, int64_t v) { *step = v; }

// halo pack / unpack: whole cells, K rows each; buffer is [cell][k]
__global__ void pack_kernel(const double* __restrict__ field, const int64_t* __restrict__ cells, int64_t n, int K,
                            int tile, double* __restrict__ dst) {
  const int64_t total = n * K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / K;
    const int k = (int)(idx % K);
    const int64_t e = cells[c];
    dst[idx] = field[((e / tile) * K + k) * tile + e % tile];
  }
}
__global__ void unpack_kernel(double* __restrict__ field, int64_t e0, int64_t n, int K, int tile,
                              const double* __restrict__ src) {
  const int64_t total = n * K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx / K;
    const int k = (int)(idx % K);
    const int64_t e = e0 + c;
    field[((e / tile) * K + k) * tile + e % tile] = src[idx];
  }
}

// ---------------------------------------------------------------------------------------------
// peer-memory halo exchange (one process per GPU, buffers mapped through CUDA IPC over NVLink)
// ---------------------------------------------------------------------------------------------
// Writes field rows of my cut-adjacent cells straight into the halo tiles of the owning peers' copy of the same
// field.  idx = k * n + c: consecutive threads write consecutive halo cells of one row (coalesced NVLink stores).
__global__ void push_kernel(const double* __restrict__ field, const int64_t* __restrict__ cells,
                            const int64_t* __restrict__ dst_cell, const int32_t* __restrict__ peer_of,
                            double* const* __restrict__ rfield, int64_t n, int K, int tile) {
  const int64_t total = n * K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx % n;
    const int k = (int)(idx / n);
    const int64_t e = cells[c], r = dst_cell[c];
    double* dst = rfield[peer_of[c]];
    dst[((r / tile) * K + k) * tile + r % tile] = field[((e / tile) * K + k) * tile + e % tile];
  }
}

// Stand-alone exchange (sg_exchange: after sg_set_state, and the two-stream schedule kept for comparison).
// After push_kernel (stream order): make the pushed rows visible system-wide, then publish the new epoch in every
// peer's flag slot for me.
__global__ void signal_kernel(unsigned long long* ctl, unsigned long long* const* __restrict__ rflag, int npeers) {
  __shared__ unsigned long long epoch;
  if (threadIdx.x == 0) {
    epoch = ctl[SG_CTL_SENT] + 1;
    ctl[SG_CTL_SENT] = epoch;
  }
  __syncthreads();
  __threadfence_system();
  if ((int)threadIdx.x < npeers) st_release_sys(rflag[threadIdx.x], epoch);
}

// Spins until every peer has published as many exchanges as this rank has (bounded: sets the error word instead of
// hanging if a peer never arrives).
__global__ void wait_kernel(unsigned long long* ctl, int npeers, long long timeout_cycles) {
  if ((int)threadIdx.x < npeers) wait_flag(ctl, threadIdx.x, ctl[SG_CTL_SENT], timeout_cycles);
  __threadfence_system();
}

#endif  // SG_STAGE_KERNELS_ONLY

}  // namespace sg
