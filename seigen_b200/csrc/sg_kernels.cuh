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
  const int32_t* nbr;     // [tile][NF][TILE]   neighbour cell (device index)
  const uint8_t* code;    // [tile][NF][TILE]
  const int32_t* absidx;  // [tile][TILE] row of absmat or -1; nullptr = no sponge anywhere
  const double* absmat;   // [ND*ND][nabs_pad]
  int64_t nabs_pad;
  const double* lam;      // per cell [tile][TILE] or nullptr
  const double* mu;
  double lam_c, mu_c;
  double c0, c1, c2;      // AXPY: out = c0*ax0 + c1*ax1 + c2*rhs
  int32_t tile0;          // first tile of this launch
};

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (global -> shared), sm_90+ PTX
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
__device__ __forceinline__ void prefetch_l2(const void* src, uint32_t bytes) {
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

template <int TILE> __device__ __forceinline__ int tile_of(int e) { return e / TILE; }

// ---------------------------------------------------------------------------------------------
// contexts handed to the generated contractions
// ---------------------------------------------------------------------------------------------
template <int D, int ND, int NFP, int TILE> struct FaceGeom {
  double ji[D][D];
  int nb[D + 1];
  unsigned cd[D + 1];
};

// F-type: row i of  Dv(s)_i = sum_j d~_j s_ij   (free-surface trace on exterior facets: s^ = 0)
template <int D, int ND, int NFP, int TILE> struct FCtx {
  const double* own;          // &sIn[(i*D*ND)*TILE + lane]
  const double* tileS;        // sIn (shared)
  const double* gIn;          // global input field
  const unsigned char* sft;   // neighbour node table (shared)
  const FaceGeom<D, ND, NFP, TILE>* g;
  int tile;
  int rowoff;                 // i*D*ND*TILE
  // per-facet state
  const double* nbp;
  const unsigned char* row;
  double gf[D], cn, co;

  __device__ __forceinline__ void t(int b, double* t) const {
    double s[D];
#pragma unroll
    for (int j = 0; j < D; ++j) s[j] = own[(j * ND + b) * TILE];
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
    cn = bnd ? 0.0 : 0.5;
    co = bnd ? 1.0 : 0.5;
    row = sft + (c & 0x7fu) * NFP;
    const int nt = n / TILE, nl = n % TILE;
    const double* base = (nt == tile) ? (tileS + nl) : (gIn + (size_t)nt * (D * D * ND * TILE) + nl);
    nbp = base + rowoff;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      if (f == 0) {
        double a = g->ji[0][j];
#pragma unroll
        for (int r = 1; r < D; ++r) a += g->ji[r][j];
        gf[j] = a;
      } else {
        gf[j] = -g->ji[f - 1][j];
      }
    }
  }
  __device__ __forceinline__ double q(int on, int m) const {
    const int nn = row[m];
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const double o = own[(j * ND + on) * TILE];
      const double v = nbp[(j * ND + nn) * TILE];
      acc = fma(gf[j], cn * v - co * o, acc);
    }
    return acc;
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

template <int D, int P, int TILE> struct SmemLayout {
  using E = ElemOps<D, P>;
  static constexpr int KS = D * D * E::ND, KU = D * E::ND;
  static constexpr size_t tail = 16 + ((E::FTAB_SIZE + 15) / 16) * 16;   // mbarrier + ftab
  static constexpr size_t f_bytes = (size_t)KS * TILE * 8 + tail;
  static constexpr size_t g_bytes = (size_t)(KU + KS) * TILE * 8 + tail;
};

// ---------------------------------------------------------------------------------------------
// common prologue: start the bulk copy of the input tile, fetch per-cell geometry meanwhile
// ---------------------------------------------------------------------------------------------
template <int D, int ND, int NFP, int TILE, int KIN, int KAX, int NTHREADS, int FTAB_SIZE>
__device__ __forceinline__ void stage_prologue(const StageParams& p, int tile, int lane, double* sIn,
                                               uint64_t* bar, unsigned char* sft, const unsigned char* ftab,
                                               FaceGeom<D, ND, NFP, TILE>& g) {
  constexpr int NF = D + 1;
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(bar, 1);
  for (int i = tid; i < FTAB_SIZE; i += NTHREADS) sft[i] = ftab[i];
  __syncthreads();
  if (tid == 0) {
    constexpr uint32_t BYTES = KIN * TILE * 8;
    mbar_expect_tx(bar, BYTES);
    bulk_g2s(sIn, p.in + (size_t)tile * (KIN * TILE), BYTES, bar);
    if (KAX > 0) {   // LF4 combination operands: start their DRAM reads now, consume them from L2 in the epilogue
      prefetch_l2(p.ax0 + (size_t)tile * (KAX * TILE), KAX * TILE * 8);
      prefetch_l2(p.ax1 + (size_t)tile * (KAX * TILE), KAX * TILE * 8);
    }
  }
  if (p.geoidx != nullptr) {
    const double* gt = p.geotab + (size_t)p.geoidx[(size_t)tile * TILE + lane] * (D * D);
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = 0; k < D; ++k) g.ji[r][k] = __ldg(gt + r * D + k);
  } else {
    const double* geo = p.geo + (size_t)tile * (D * D * TILE) + lane;
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int k = 0; k < D; ++k) g.ji[r][k] = geo[(r * D + k) * TILE];
  }
  const int32_t* nb = p.nbr + (size_t)tile * (NF * TILE) + lane;
  const uint8_t* cd = p.code + (size_t)tile * (NF * TILE) + lane;
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    g.nb[f] = nb[f * TILE];
    g.cd[f] = cd[f * TILE];
  }
  mbar_wait(bar, 0);
}

// ---------------------------------------------------------------------------------------------
// F-type pass:   out_i = Dv(in)_i - A_cell * absu_i            (K1, K5)
//                out_i = c0*ax0_i + c1*ax1_i + c2*(Dv(in)_i - A_cell*absu_i)   (K3, AXPY)
// ---------------------------------------------------------------------------------------------
template <int D, int P, int TILE, int SPLIT, int MINB, bool AXPY>
__global__ void __launch_bounds__(TILE* SPLIT, MINB) stage_f_kernel(const StageParams p) {
  using E = ElemOps<D, P>;
  constexpr int ND = E::ND, NFP = E::NFP, KS = D * D * ND, KU = D * ND, IPT = D / SPLIT;
  static_assert(D % SPLIT == 0, "SPLIT must divide D");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sIn = reinterpret_cast<double*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sIn + KS * TILE);
  unsigned char* sft = reinterpret_cast<unsigned char*>(bar + 2);

  const int lane = threadIdx.x % TILE, ig = threadIdx.x / TILE;
  const int tile = p.tile0 + blockIdx.x;
  FaceGeom<D, ND, NFP, TILE> g;
  stage_prologue<D, ND, NFP, TILE, KS, (AXPY ? KU : 0), TILE * SPLIT, E::FTAB_SIZE>(p, tile, lane, sIn, bar, sft, E::ftab(), g);

  int aidx = -1;
  if (p.absidx != nullptr) aidx = p.absidx[(size_t)tile * TILE + lane];

  FCtx<D, ND, NFP, TILE> c;
  c.tileS = sIn;
  c.gIn = p.in;
  c.sft = sft;
  c.g = &g;
  c.tile = tile;
#pragma unroll
  for (int ii = 0; ii < IPT; ++ii) {
    const int i = ig * IPT + ii;
    c.rowoff = i * D * ND * TILE;
    c.own = sIn + c.rowoff + lane;
    double acc[ND];
#pragma unroll
    for (int a = 0; a < ND; ++a) acc[a] = 0.0;
    E::volF(c, acc);
    E::liftF(c, acc);
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
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      double v = acc[a];
      if (AXPY) v = fma(p.c0, p.ax0[orow + (size_t)a * TILE], fma(p.c1, p.ax1[orow + (size_t)a * TILE], p.c2 * v));
      p.out[orow + (size_t)a * TILE] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// G-type pass:   out_ij = lam*delta_ij*div + mu*(G_ij + G_ji),  G_ij = d~_j in_i      (K2, K4)
//                out_ij = c0*ax0_ij + c1*ax1_ij + c2*(...)                           (K6, AXPY)
// ---------------------------------------------------------------------------------------------
template <int D, int P, int TILE, int SPLIT, int MINB, bool AXPY>
__global__ void __launch_bounds__(TILE* SPLIT, MINB) stage_g_kernel(const StageParams p) {
  using E = ElemOps<D, P>;
  constexpr int ND = E::ND, NFP = E::NFP, KS = D * D * ND, KU = D * ND, IPT = D / SPLIT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sIn = reinterpret_cast<double*>(smem_raw);
  double* sX = sIn + KU * TILE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sX + KS * TILE);
  unsigned char* sft = reinterpret_cast<unsigned char*>(bar + 2);

  const int lane = threadIdx.x % TILE, ig = threadIdx.x / TILE;
  const int tile = p.tile0 + blockIdx.x;
  FaceGeom<D, ND, NFP, TILE> g;
  stage_prologue<D, ND, NFP, TILE, KU, (AXPY ? KS : 0), TILE * SPLIT, E::FTAB_SIZE>(p, tile, lane, sIn, bar, sft, E::ftab(), g);

  GCtx<D, ND, NFP, TILE> c;
  c.tileU = sIn;
  c.gIn = p.in;
  c.sft = sft;
  c.g = &g;
  c.tile = tile;
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
        X[(j * ND + a) * TILE] = v;
      }
  }
  if (SPLIT > 1) __syncthreads();

  double lam = p.lam_c, mu = p.mu_c;
  if (p.lam != nullptr) {
    lam = p.lam[(size_t)tile * TILE + lane];
    mu = p.mu[(size_t)tile * TILE + lane];
  }
  const double* X = sX + lane;
#pragma unroll
  for (int ii = 0; ii < IPT; ++ii) {
    const int i = ig * IPT + ii;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      double div = X[(0 * ND + a) * TILE];
#pragma unroll
      for (int k = 1; k < D; ++k) div += X[((k * D + k) * ND + a) * TILE];
      const double ld = lam * div;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double v = mu * (X[((i * D + j) * ND + a) * TILE] + X[((j * D + i) * ND + a) * TILE]);
        if (j == i) v += ld;
        const size_t o = ((size_t)tile * KS + (i * D + j) * ND + a) * TILE + lane;
        if (AXPY) v = fma(p.c0, p.ax0[o], fma(p.c1, p.ax1[o], p.c2 * v));
        p.out[o] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------
// boundary (AoS, cell-major: [cell][node][comp]) <-> device (tile-blocked SoA: [tile][comp*ND+node][lane])
template <bool TO_DEVICE>
__global__ void relayout_kernel(double* __restrict__ dev, double* __restrict__ host_order, int64_t ncell,
                                int64_t n_owned, int64_t n_owned_pad, int nd, int ncomp, int tile) {
  const int K = nd * ncomp;
  const int64_t total = ncell * K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = idx / K;
    const int r = (int)(idx % K);
    const int node = r / ncomp, comp = r % ncomp;
    const int64_t k = comp * nd + node;
    const int64_t e = cell < n_owned ? cell : cell - n_owned + n_owned_pad;
    const int64_t d = ((e / tile) * K + k) * tile + e % tile;
    if (TO_DEVICE)
      dev[d] = host_order[idx];
    else
      host_order[idx] = dev[d];
  }
}

// stress += scale * amp[step][k] at the listed device addresses (elastic.py:149-154, 285-288)
__global__ void add_source_kernel(double* __restrict__ s, const int64_t* __restrict__ addr,
                                  const double* __restrict__ amp, const int64_t* __restrict__ step, int64_t nsteps,
                                  int64_t nsrc, double scale, int64_t addr_lo, int64_t addr_hi) {
  const int64_t st = *step;
  if (st < 0 || st >= nsteps) return;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nsrc; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = addr[k];
    if (a >= addr_lo && a < addr_hi) s[a] += scale * amp[st * nsrc + k];
  }
}
__global__ void bump_step_kernel(int64_t* step) { *step += 1; }
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

// ctl layout (uint64): [0, 16) flags written by my peers, [16] exchanges I have signalled, [17] exchanges I have
// waited for, [18] error word (1 = a wait timed out)
constexpr int SG_CTL_SENT = 16, SG_CTL_WAITED = 17, SG_CTL_ERROR = 18, SG_CTL_WORDS = 32;

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
  if ((int)threadIdx.x < npeers) {
    unsigned long long* f = rflag[threadIdx.x];
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
  }
}

// Spins until every peer has published the epoch this rank is about to consume (bounded: sets the error word
// instead of hanging if a peer never arrives).
__global__ void wait_kernel(unsigned long long* ctl, int npeers, long long timeout_cycles) {
  __shared__ unsigned long long epoch;
  if (threadIdx.x == 0) {
    epoch = ctl[SG_CTL_WAITED] + 1;
    ctl[SG_CTL_WAITED] = epoch;
  }
  __syncthreads();
  if ((int)threadIdx.x < npeers) {
    const unsigned long long* f = ctl + threadIdx.x;
    const long long t0 = clock64();
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > timeout_cycles) {
        ctl[SG_CTL_ERROR] = 1;
        break;
      }
      __nanosleep(200);
    } while (true);
  }
  __threadfence_system();
}

}  // namespace sg
