// C ABI (include/seigen_b200.h) over the fused sm_100a stage kernels.  No torch types, no CPU fallback.
#include "../../include/seigen_b200.h"
#include "sg_kernels.cuh"
#include "sg_variants.h"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

// NVTX ranges around the host-side phases (visible in Nsight Systems / filterable in Nsight Compute); the role of the
// reference's timed_region('timestepping' / 'solver setup' / 'i/o') markers (seigen/elastic.py:76, 247, 278)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define SG_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      return fail(SG_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
    }                                                                                          \
  } while (0)

const std::vector<Variant>& variants() {
  static const std::vector<Variant> v = [] {
    std::vector<Variant> t;
    sg_variants_1d(t);
    sg_variants_2d_low(t);
    sg_variants_2d_high(t);
    sg_variants_3d_p1(t);
    sg_variants_3d_p2(t);
    sg_variants_3d_p3(t);
    return t;
  }();
  return v;
}

int env_int(const char* name) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : 0;
}

const Variant* find_variant(int dim, int degree) {
  const int tile = env_int("SG_TILE"), split = env_int("SG_SPLIT"), minb = env_int("SG_MINB"), ns = env_int("SG_NS");
  const int minba = env_int("SG_MINBA");
  const char* ex = std::getenv("SG_XREG");
  const char* ea = std::getenv("SG_AXS");
  const int xreg = ex ? std::atoi(ex) : -1, axs = ea ? std::atoi(ea) : -1;
  const Variant* first = nullptr;
  for (const Variant& v : variants()) {
    if (v.dim != dim || v.degree != degree) continue;
    if (!first) first = &v;
    if ((tile == 0 || v.tile == tile) && (split == 0 || v.split == split) && (minb == 0 || v.minb == minb) &&
        (minba == 0 || v.minba == minba) && (ns == 0 || v.ns_plain * 10 + v.ns_axpy == ns) &&
        (xreg < 0 || v.xreg == xreg) && (axs < 0 || v.axs == axs))
      return &v;
  }
  return first;
}

template <class T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    if (p && count == n) return cudaSuccess;   // same size: keep the buffer (pointers baked into CUDA graphs stay valid)
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, count * sizeof(T));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace

struct sg_solver {
  const Variant* var = nullptr;
  int dim = 0, degree = 0, nd = 0, nf = 0, tile = 0, device = 0;
  int KU = 0, KS = 0;               // device rows per cell of a velocity / stress field
  int sym = 0, ncs = 0, KS_full = 0; // symmetric stress storage; stored stress components; nd*dim*dim
  DevBuf<unsigned int> asym;         // raised by the relayout kernel when a stress field handed in is not symmetric
  int64_t n_owned = 0, n_total = 0, n_owned_pad = 0, n_halo = 0, n_dev = 0, n_boundary = 0;
  int tiles_owned = 0, tiles_total = 0, tiles_boundary = 0;
  DevBuf<double> u, s, uh, sh;          // state + scratch, tile-blocked
  DevBuf<double> geo, geotab, mat, absmat, amp;
  DevBuf<uint16_t> geoidx;
  int64_t n_geo_classes = 0;
  DevBuf<int32_t> nbr, absidx;
  DevBuf<uint8_t> code;
  DevBuf<int64_t> step_dev, send_cells;
  DevBuf<int32_t> src_start, src_off;
  DevBuf<int64_t> rec_cell;               // receivers: owning cell, basis weights, samples [max_steps][nrec][dim]
  DevBuf<double> rec_w, rec_data;
  int64_t nrec = 0, rec_steps = 0;
  DevBuf<unsigned int> sched;             // [3 parts][4] dynamic tile scheduler words (sg::sched_next, sg::halo_push)
  int nsm = 148;
  int occ[4] = {0, 0, 0, 0};            // resident CTAs per SM of f_plain, f_axpy, g_plain, g_axpy
  uint32_t occ_smem[4] = {0, 0, 0, 0};  // the shared-memory size those were computed for
  int64_t nabs_pad = 0, nsrc = 0, src_steps = 0, nsend = 0;
  bool have_material = false, per_cell = false;
  double density = 1.0, lam_c = 0.0, mu_c = 0.0;
  cudaStream_t stream = nullptr, comm = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr;
  // peer-memory halo exchange (CUDA IPC): see sg_peer_connect
  int npeers = 0;
  std::vector<void*> ipc_opened;
  DevBuf<unsigned long long> ctl;            // flags + epoch counters (sg::SG_CTL_*)
  DevBuf<int64_t> send_dst;                  // [nsend] device cell index in the destination rank's fields
  DevBuf<int32_t> send_peer;                 // [nsend] index into the peer tables
  DevBuf<double*> rfield;                    // [4][npeers] peers' u, s, uh, sh
  DevBuf<unsigned long long*> rflag;         // [npeers] my flag slot in each peer's ctl
  // the same send list regrouped by boundary tile for the exchange fused into the stage kernels (sg::halo_push)
  DevBuf<int32_t> push_start, push_lane, push_peer;
  DevBuf<int64_t> push_dst;
  int push_tiles = 0;
  long long timeout_cycles = (long long)40e9;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_int[6] = {}, ev_bnd[6] = {};
  // CUDA graph of one time step
  cudaGraphExec_t graph = nullptr;
  double graph_dt = 0.0;
  uint64_t config_version = 0, graph_version = ~0ull;

  int64_t dev_index(int64_t c) const { return c < n_owned ? c : c - n_owned + n_owned_pad; }
};

namespace {

int grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

void drop_graph(sg_solver* h) {
  if (h->graph) cudaGraphExecDestroy(h->graph);
  h->graph = nullptr;
}

sg::StageParams base_params(sg_solver* h) {
  sg::StageParams p;
  std::memset(&p, 0, sizeof(p));
  p.geo = h->geo.p;
  p.geoidx = h->geoidx.p;
  p.geotab = h->geotab.p;
  p.nclass = (int32_t)h->n_geo_classes;
  p.nbr = h->nbr.p;
  p.code = h->code.p;
  p.absidx = h->nabs_pad > 0 ? h->absidx.p : nullptr;
  p.absmat = h->absmat.p;
  p.nabs_pad = h->nabs_pad;
  p.mat = h->per_cell ? h->mat.p : nullptr;
  p.lam_c = h->lam_c;
  p.mu_c = h->mu_c;
  return p;
}

int field_ncomp(sg_solver* h, int which);
const int STAGE_OUTPUT[7] = {-1, SG_FIELD_UH, SG_FIELD_SH, SG_FIELD_U, SG_FIELD_SH, SG_FIELD_UH, SG_FIELD_S};

// in_step: the launch is one of the six of a captured time step (chained with programmatic dependent launch; the
// last one advances the step counter)
int launch_stage(sg_solver* h, int stage, int part, double dt, cudaStream_t st, bool push = false,
                 bool in_step = false) {
  int t0 = 0, nt = h->tiles_owned;
  if (part == SG_PART_BOUNDARY) {
    nt = h->tiles_boundary;
  } else if (part == SG_PART_INTERIOR) {
    t0 = h->tiles_boundary;
    nt = h->tiles_owned - h->tiles_boundary;
  } else if (part != SG_PART_ALL) {
    return fail(SG_EINVAL, "sg_stage: bad part");
  }
  const Variant* v = h->var;
  sg::StageParams p = base_params(h);
  p.tile0 = t0;
  p.ntiles = nt;
  p.sched = h->sched.p + 4 * part;
  const double c3 = dt * dt * dt / 24.0;
  const bool classes = h->geoidx.p != nullptr, sponge = p.absidx != nullptr, mat = p.mat != nullptr;
  const void* fn = nullptr;
  sg::StagePlan pl{};
  int* occ = nullptr;
  bool gtype = false;
  switch (stage) {
    case 1:  // uh1 = Dv(s0) - P(sigma, u0)                         elastic.py:157-161, 292
      p.in = h->s.p; p.out = h->uh.p; p.absu = h->u.p;
      fn = v->f_plain[h->sym]; pl = v->plan_f[h->sym](classes, mat, sponge); occ = &h->occ[0];
      break;
    case 2:  // stemp = Ds(uh1) + src                               elastic.py:163-167, 293
      p.in = h->uh.p; p.out = h->sh.p; p.src_scale = 1.0; gtype = true;
      fn = v->g_plain[h->sym]; pl = v->plan_g[h->sym](classes, mat, sponge); occ = &h->occ[2];
      break;
    case 3:  // u1 = rho*u0 + dt*uh1 + dt^3/24*(Dv(stemp) - P(sigma, u0))   elastic.py:169-173, 341-345, 294-296
      p.in = h->sh.p; p.out = h->u.p; p.ax0 = h->u.p; p.ax1 = h->uh.p; p.absu = h->u.p;
      p.c0 = h->density; p.c1 = dt; p.c2 = c3;
      fn = v->f_axpy[h->sym]; pl = v->plan_f_axpy[h->sym](classes, mat, sponge); occ = &h->occ[1];
      break;
    case 4:  // sh1 = Ds(u1) + src                                  elastic.py:181-185, 300
      p.in = h->u.p; p.out = h->sh.p; p.src_scale = 1.0; gtype = true;
      fn = v->g_plain[h->sym]; pl = v->plan_g[h->sym](classes, mat, sponge); occ = &h->occ[2];
      break;
    case 5:  // utemp = Dv(sh1) - P(sigma, u1)                      elastic.py:187-191, 301
      p.in = h->sh.p; p.out = h->uh.p; p.absu = h->u.p;
      fn = v->f_plain[h->sym]; pl = v->plan_f[h->sym](classes, mat, sponge); occ = &h->occ[0];
      break;
    case 6:  // s1 = s0 + dt*sh1 + dt^3/24*(Ds(utemp) + src)        elastic.py:193-197, 348-352, 302-304
      p.in = h->uh.p; p.out = h->s.p; p.ax0 = h->s.p; p.ax1 = h->sh.p;
      p.c0 = 1.0; p.c1 = dt; p.c2 = c3; p.src_scale = c3; gtype = true;
      fn = v->g_axpy[h->sym]; pl = v->plan_g_axpy[h->sym](classes, mat, sponge); occ = &h->occ[3];
      break;
    default:
      return fail(SG_EINVAL, "sg_stage: stage must be 1..6");
  }
  // source: added wherever g is evaluated (elastic.py:165, 183, 195), by the CTA that produced the tile
  if (gtype && h->nsrc > 0) {
    p.src_start = h->src_start.p;
    p.src_off = h->src_off.p;
    p.amp = h->amp.p;
    p.step = h->step_dev.p;
    p.nsteps = h->src_steps;
    p.nsrc = h->nsrc;
  }
  // the device-side step counter (source table index, receivers) advances with the last pass of a step; with
  // receivers a separate kernel does it after they have been sampled
  if (in_step && stage == 6 && h->nsrc > 0 && h->nrec == 0) p.bump = h->step_dev.p;
  // the trigger at the end of a CTA's tile loop measured as good as or better than at its start for every element
  // (profiles/r02_pdl_modes.log: 3D P3 54.3 vs 52.1 G, 2D P4 87.1 vs 84.2 G); SG_PDL_EARLY=1 selects the early one
  p.pdl_late = env_int("SG_PDL_EARLY") ? 0 : 1;
  if (push && h->npeers > 0 && h->push_tiles > 0) {
    // halo exchange of this pass's output fused into the kernel (sg::halo_wait / sg::halo_push)
    const int which = STAGE_OUTPUT[stage];
    p.push_tiles = h->push_tiles;
    p.npeers = h->npeers;
    p.K_out = field_ncomp(h, which) * h->nd;
    p.push_start = h->push_start.p;
    p.push_lane = h->push_lane.p;
    p.push_peer = h->push_peer.p;
    p.push_dst = h->push_dst.p;
    p.rfield = h->rfield.p + (size_t)which * h->npeers;
    p.ctl = h->ctl.p;
    p.rflag = h->rflag.p;
    p.timeout_cycles = h->timeout_cycles;
    p.dbg = env_int("SG_EXCHANGE_DEBUG");
  }
  if (nt > 0) {
    const int nthreads = v->threads[occ - h->occ];
    if (h->occ_smem[occ - h->occ] != pl.total) {   // (re)size the kernel's shared memory and its persistent grid
      SG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
      SG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      int nb = 0;
      SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, nthreads, pl.total));
      if (nb < 1) return fail(SG_ECUDA, "stage kernel does not fit on an SM (shared memory plan too large)");
      {
        // Shared memory and L1 share the SM's 256 KB.  The facet gathers of out-of-tile neighbours go through L1, so
        // give shared memory only what the resident CTAs need (smallest hardware configuration that keeps the
        // occupancy computed above) and leave the rest to L1: 3D P3 K1 152 -> 127 us, 2D P3 +3 %
        // (profiles/r02_experiment_l1_prefetch_and_carveout.log).  SG_CARVEOUT=<percent> overrides, 100 = round 1.
        static const int config_kb[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
        const size_t need = (size_t)nb * (pl.total + 1024);      // 1 KB per CTA is reserved by the system
        int pick = 228;
        for (int kb : config_kb)
          if ((size_t)kb * 1024 >= need) {
            pick = kb;
            break;
          }
        int carve = (pick * 100 + 227) / 228;
        if (env_int("SG_CARVEOUT") > 0) carve = env_int("SG_CARVEOUT");
        if (carve < 100) {
          SG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
          int nb2 = 0;
          SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, fn, nthreads, pl.total));
          if (nb2 < nb && env_int("SG_CARVEOUT") == 0)      // the hint would cost occupancy: keep the maximum
            SG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        }
      }
      *occ = nb;
      h->occ_smem[occ - h->occ] = pl.total;
    }
    int grid = h->nsm * *occ;
    const int cap = env_int("SG_GRID_PER_SM");
    if (cap > 0) grid = h->nsm * cap;
    // experiment for the next round (off by default, unmeasured): leave room for the boundary kernel's CTAs next to
    // the persistent interior grid, so that boundary(k+1) + exchange overlap interior(k+1) instead of trailing it
    // (3D P3 at N = 2 loses 42 us per pass to that chain: profiles/README.md)
    if (part == SG_PART_INTERIOR && h->npeers > 0 && env_int("SG_INTERIOR_RESERVE") > 0) {
      const int reserve = std::min(h->tiles_boundary, grid / 4);
      grid -= reserve;
    }
    if (grid > nt) grid = nt;
    if (env_int("SG_GRID_MAX") > 0) grid = std::min(grid, env_int("SG_GRID_MAX"));   // tests: many tiles per CTA on a small mesh
    void* args[] = {(void*)&p};
    if (in_step && env_int("SG_NO_PDL") == 0) {
      // programmatic dependent launch: this kernel's launch + CTA prologue overlap the previous pass's tail
      // (sg::pdl_wait in the kernel orders every read of the previous output)
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr.val.programmaticStreamSerializationAllowed = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(nthreads);
      cfg.dynamicSmemBytes = pl.total;
      cfg.stream = st;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      SG_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    } else if (part == SG_PART_BOUNDARY && env_int("SG_LAUNCH_PRIORITY") > 0) {
      // experiment for the next round (off by default, unmeasured): give the boundary kernel node an explicit
      // priority inside the captured graph instead of relying on the comm stream's
      int least = 0, greatest = 0;
      SG_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributePriority;
      attr.val.priority = greatest;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(nthreads);
      cfg.dynamicSmemBytes = pl.total;
      cfg.stream = st;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      SG_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    } else {
      SG_CUDA(cudaLaunchKernel(fn, dim3(grid), dim3(nthreads), args, pl.total, st));
    }
  }
  return SG_OK;
}

int relayout(sg_solver* h, double* dev, double* host_order, int ncomp, bool to_device, cudaStream_t st,
             int64_t ncell) {
  // ncomp: components at the boundary; stress fields of a symmetric-storage solver are packed on the device
  const int64_t total = ncell * h->nd * ncomp;
  if (total == 0) return SG_OK;
  const int symd = (h->sym && ncomp == h->dim * h->dim) ? h->dim : 0;
  if (to_device)
    sg::relayout_kernel<true><<<grid_for(total), 256, 0, st>>>(dev, host_order, ncell, h->n_owned, h->n_owned_pad,
                                                               h->nd, ncomp, h->tile, symd, h->asym.p);
  else
    sg::relayout_kernel<false><<<grid_for(total), 256, 0, st>>>(dev, host_order, ncell, h->n_owned, h->n_owned_pad,
                                                                h->nd, ncomp, h->tile, symd, h->asym.p);
  SG_CUDA(cudaGetLastError());
  return SG_OK;
}

DevBuf<double>* field_buf(sg_solver* h, int which);
int field_ncomp_boundary(sg_solver* h, int which);

// push my cut-adjacent cells of field `which` into the peers' halo tiles, publish, then wait for the peers' rows
int enqueue_exchange(sg_solver* h, int which, cudaStream_t st) {
  if (h->npeers == 0) return SG_OK;
  const int K = field_ncomp(h, which) * h->nd;
  // (a single push+signal+wait kernel was tried and was slower: profiles/r01_experiment_fused_exchange.log)
  sg::push_kernel<<<grid_for(h->nsend * K), 256, 0, st>>>(field_buf(h, which)->p, h->send_cells.p, h->send_dst.p,
                                                          h->send_peer.p, h->rfield.p + (size_t)which * h->npeers,
                                                          h->nsend, K, h->tile);
  SG_CUDA(cudaGetLastError());
  sg::signal_kernel<<<1, 32, 0, st>>>(h->ctl.p, h->rflag.p, h->npeers);
  SG_CUDA(cudaGetLastError());
  sg::wait_kernel<<<1, 32, 0, st>>>(h->ctl.p, h->npeers, h->timeout_cycles);
  SG_CUDA(cudaGetLastError());
  return SG_OK;
}

// One time step on a rank with peers, two-stream schedule (SG_PEER_SCHED=split; the default is the fused schedule
// below): two chains that only meet where the data says they must.
//   compute stream:  interior(1) -> interior(2) -> ... -> interior(6)
//   comm stream:     boundary(1) -> push/signal/wait -> boundary(2) -> push/signal/wait -> ...
// interior(k+1) needs boundary(k) (it reads cut-adjacent neighbours) but never the halo, so the exchange latency is
// off its critical path; boundary(k+1) needs the halo rows (wait) and interior(k).  Each kernel of pass k+1 waits
// for both kernels of pass k, which also orders every write-after-read on the recycled scratch fields.
int enqueue_step_peers(sg_solver* h, double dt) {
  cudaStream_t st = h->stream, cm = h->comm;
  SG_CUDA(cudaEventRecord(h->ev_fork, st));
  SG_CUDA(cudaStreamWaitEvent(cm, h->ev_fork, 0));
  for (int k = 1; k <= 6; ++k) {
    if (k > 1) {
      SG_CUDA(cudaStreamWaitEvent(cm, h->ev_int[k - 2], 0));   // boundary(k) after interior(k-1)
      SG_CUDA(cudaStreamWaitEvent(st, h->ev_bnd[k - 2], 0));   // interior(k) after boundary(k-1)
    }
    int rc = launch_stage(h, k, SG_PART_BOUNDARY, dt, cm);
    if (rc) return rc;
    SG_CUDA(cudaEventRecord(h->ev_bnd[k - 1], cm));
    rc = enqueue_exchange(h, STAGE_OUTPUT[k], cm);
    if (rc) return rc;
    rc = launch_stage(h, k, SG_PART_INTERIOR, dt, st);
    if (rc) return rc;
    SG_CUDA(cudaEventRecord(h->ev_int[k - 1], st));
  }
  SG_CUDA(cudaEventRecord(h->ev_join, cm));
  SG_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
  return SG_OK;
}

// One time step on a rank with peers, fused schedule: six launches over ALL owned tiles, exactly the single-GPU
// step.  The tiles of the cut-adjacent cells come first in the tile order, so the persistent CTAs take them first;
// each of them waits (device-side flag) for the peers' rows of the previous pass, computes, stores its cells' rows
// into the peers' halo tiles over NVLink and the last one publishes the epoch -- all while the other CTAs are
// already working through the interior tiles.  No second stream, no events, no extra kernels: the exchange costs
// nothing on the critical path unless a peer is late.
int enqueue_step_fused(sg_solver* h, double dt) {
  for (int k = 1; k <= 6; ++k) {
    int rc = launch_stage(h, k, SG_PART_ALL, dt, h->stream, true, true);
    if (rc) return rc;
  }
  return SG_OK;
}

DevBuf<double>* field_buf(sg_solver* h, int which) {
  switch (which) {
    case SG_FIELD_U: return &h->u;
    case SG_FIELD_S: return &h->s;
    case SG_FIELD_UH: return &h->uh;
    case SG_FIELD_SH: return &h->sh;
  }
  return nullptr;
}
// components stored on the device / crossing the boundary
int field_ncomp(sg_solver* h, int which) {
  return (which == SG_FIELD_U || which == SG_FIELD_UH) ? h->dim : h->ncs;
}
int field_ncomp_boundary(sg_solver* h, int which) {
  return (which == SG_FIELD_U || which == SG_FIELD_UH) ? h->dim : h->dim * h->dim;
}

}  // namespace

extern "C" {

const char* sg_last_error(void) { return g_err.c_str(); }
int sg_version(void) { return 1; }

int sg_nodes_per_cell(int dim, int degree) {
  const Variant* v = find_variant(dim, degree);
  return v ? v->nd : SG_EINVAL;
}
int sg_tile_cells(int dim, int degree) {
  const Variant* v = find_variant(dim, degree);
  return v ? v->tile : SG_EINVAL;
}

int sg_create(sg_solver** out, const sg_mesh_desc* d) {
  NvtxRange nvtx_range("sg_create");
  if (!out || !d) return fail(SG_EINVAL, "sg_create: null argument");
  *out = nullptr;
  const Variant* v = find_variant(d->dim, d->degree);
  if (!v) return fail(SG_EINVAL, "sg_create: unsupported (dim, degree); supported: 1D P1-P3, 2D P1-P4, 3D P1-P3");
  if (d->n_owned <= 0 || d->n_total < d->n_owned || !d->nbr || !d->code || !d->jinv)
    return fail(SG_EINVAL, "sg_create: bad mesh description");
  if (d->n_boundary < 0 || d->n_boundary > d->n_owned) return fail(SG_EINVAL, "sg_create: bad n_boundary");
  int ndev = 0;
  SG_CUDA(cudaGetDeviceCount(&ndev));
  if (d->device < 0 || d->device >= ndev) return fail(SG_EINVAL, "sg_create: no such CUDA device");
  SG_CUDA(cudaSetDevice(d->device));

  sg_solver* h = new sg_solver();
  h->var = v;
  h->dim = v->dim;
  h->degree = v->degree;
  h->nd = v->nd;
  h->nf = v->dim + 1;
  h->tile = v->tile;
  h->device = d->device;
  h->sym = d->symmetric_stress ? 1 : 0;
  h->ncs = h->sym ? v->dim * (v->dim + 1) / 2 : v->dim * v->dim;
  h->KU = v->dim * v->nd;
  h->KS = h->ncs * v->nd;
  h->KS_full = v->dim * v->dim * v->nd;
  h->n_owned = d->n_owned;
  h->n_total = d->n_total;
  h->n_halo = d->n_total - d->n_owned;
  h->n_boundary = d->n_boundary;
  const int T = v->tile;
  h->tiles_owned = (int)((h->n_owned + T - 1) / T);
  h->n_owned_pad = (int64_t)h->tiles_owned * T;
  h->tiles_total = h->tiles_owned + (int)((h->n_halo + T - 1) / T);
  h->n_dev = (int64_t)h->tiles_total * T;
  h->tiles_boundary = (int)((h->n_boundary + T - 1) / T);
  if (h->n_dev >= (int64_t)1 << 31) {
    delete h;
    return fail(SG_EINVAL, "sg_create: too many cells for 32-bit neighbour indices");
  }

  auto cleanup = [&](int code) {
    sg_destroy(h);
    return code;
  };
#define SG_CUDA_H(expr)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      g_err = std::string(#expr) + ": " + cudaGetErrorString(_e);                          \
      return cleanup(SG_ECUDA);                                                            \
    }                                                                                      \
  } while (0)

  SG_CUDA_H(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    // the comm stream carries the boundary tiles and the halo exchange of every pass: highest priority, so that when
    // boundary(k+1) and interior(k+1) become ready together the boundary CTAs are dispatched first and the
    // push/signal/wait chain overlaps the interior kernel instead of trailing it (kernel nodes captured from this
    // stream keep the priority inside the step graph)
    int least = 0, greatest = 0;
    SG_CUDA_H(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const int prio = env_int("SG_COMM_PRIORITY_OFF") ? least : greatest;
    SG_CUDA_H(cudaStreamCreateWithPriority(&h->comm, cudaStreamNonBlocking, prio));
  }
  SG_CUDA_H(cudaEventCreate(&h->ev0));
  SG_CUDA_H(cudaEventCreate(&h->ev1));
  SG_CUDA_H(cudaEventCreateWithFlags(&h->ev_sync, cudaEventDisableTiming));
  SG_CUDA_H(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  SG_CUDA_H(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  for (int k = 0; k < 6; ++k) {
    SG_CUDA_H(cudaEventCreateWithFlags(&h->ev_int[k], cudaEventDisableTiming));
    SG_CUDA_H(cudaEventCreateWithFlags(&h->ev_bnd[k], cudaEventDisableTiming));
  }
  SG_CUDA_H(h->sched.alloc(12));
  SG_CUDA_H(cudaMemsetAsync(h->sched.p, 0, 12 * sizeof(unsigned int), h->stream));
  if (const char* e = std::getenv("SG_PEER_TIMEOUT_S")) {
    const double sec = std::atof(e);
    if (sec > 0) h->timeout_cycles = (long long)(sec * 1.9e9);   // clock64 ticks at ~1.9 GHz
  }
  SG_CUDA_H(h->ctl.alloc(sg::SG_CTL_WORDS));
  SG_CUDA_H(cudaMemsetAsync(h->ctl.p, 0, sg::SG_CTL_WORDS * 8, h->stream));

  SG_CUDA_H(h->asym.alloc(1));
  SG_CUDA_H(cudaMemsetAsync(h->asym.p, 0, sizeof(unsigned int), h->stream));
  // stress buffers are sized for full storage either way: they double as staging buffers for the boundary layout
  const size_t nU = (size_t)h->n_dev * h->KU, nS = (size_t)h->n_dev * h->KS_full;
  SG_CUDA_H(h->u.alloc(nU));
  SG_CUDA_H(h->uh.alloc(nU));
  SG_CUDA_H(h->s.alloc(nS));
  SG_CUDA_H(h->sh.alloc(nS));
  SG_CUDA_H(cudaMemsetAsync(h->u.p, 0, nU * 8, h->stream));
  SG_CUDA_H(cudaMemsetAsync(h->uh.p, 0, nU * 8, h->stream));
  SG_CUDA_H(cudaMemsetAsync(h->s.p, 0, nS * 8, h->stream));
  SG_CUDA_H(cudaMemsetAsync(h->sh.p, 0, nS * 8, h->stream));
  SG_CUDA_H(h->step_dev.alloc(1));
  SG_CUDA_H(cudaMemsetAsync(h->step_dev.p, 0, 8, h->stream));

  // adjacency + geometry of owned tiles, re-laid tile-blocked on the host (one-off)
  const int nf = h->nf, dd = h->dim * h->dim;
  const size_t npad = (size_t)h->n_owned_pad;
  std::vector<int32_t> nbr(npad * nf);
  std::vector<uint8_t> code(npad * nf);
  std::vector<double> geo(npad * dd, 0.0);
  for (size_t e = 0; e < npad; ++e) {
    const size_t t = e / T, l = e % T;
    for (int f = 0; f < nf; ++f) {
      int32_t n = (int32_t)e;
      uint8_t c = (uint8_t)SG_BOUNDARY;   // padding lanes: exterior facet onto itself (row 0 of the node table)
      if ((int64_t)e < h->n_owned) {
        const int64_t nn = d->nbr[e * nf + f];
        if (nn < 0 || nn >= h->n_total) {
          g_err = "sg_create: neighbour index out of range";
          return cleanup(SG_EINVAL);
        }
        n = (int32_t)h->dev_index(nn);
        c = d->code[e * nf + f];
      }
      nbr[(t * nf + f) * T + l] = n;
      code[(t * nf + f) * T + l] = c;
    }
    if ((int64_t)e < h->n_owned)
      for (int k = 0; k < dd; ++k) geo[(t * dd + k) * T + l] = d->jinv[e * dd + k];
  }
  SG_CUDA_H(h->nbr.alloc(nbr.size()));
  SG_CUDA_H(h->code.alloc(code.size()));
  SG_CUDA_H(cudaMemcpy(h->nbr.p, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice));
  SG_CUDA_H(cudaMemcpy(h->code.p, code.data(), code.size(), cudaMemcpyHostToDevice));

  // Geometry classes: on (piecewise) uniform meshes thousands of cells are translates of one another and share
  // Jinv up to the round-off of the vertex coordinates.  Such cells get a 2-byte class id instead of D*D doubles
  // (32-72 B of HBM traffic per cell per pass).  Classes are merged only within SG_GEOM_TOL relative.
  bool use_classes = false;
  if (d->geom_classes) {
    // key = (binary exponent of the largest entry, the d*d entries in units of 2^-36 of it): fixed-size, hashed with
    // FNV-1a -- no per-cell heap allocation (this loop runs over every owned cell)
    struct GeoKey {
      int64_t v[10];
      bool operator==(const GeoKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
    };
    struct GeoKeyHash {
      size_t operator()(const GeoKey& k) const {
        uint64_t x = 1469598103934665603ull;
        for (int i = 0; i < 10; ++i) {
          x ^= (uint64_t)k.v[i];
          x *= 1099511628211ull;
        }
        return (size_t)x;
      }
    };
    std::unordered_map<GeoKey, uint16_t, GeoKeyHash> seen;
    std::vector<double> tab;
    std::vector<uint16_t> gi(npad, 0);
    const size_t max_classes = 4096;
    use_classes = true;
    for (int64_t e = 0; e < h->n_owned && use_classes; ++e) {
      const double* J = d->jinv + (size_t)e * dd;
      double nrm = 0.0;
      for (int k = 0; k < dd; ++k) nrm = std::fmax(nrm, std::fabs(J[k]));
      int ex = 0;
      std::frexp(nrm, &ex);
      const double q = std::ldexp(1.0, ex - 36);   // quantum: 2^-36 of the largest entry (~1.5e-11 relative)
      GeoKey ks{};
      ks.v[0] = ex;
      for (int k = 0; k < dd; ++k) ks.v[1 + k] = (int64_t)std::llround(J[k] / q);
      auto it = seen.find(ks);
      uint16_t cls;
      if (it == seen.end()) {
        if (seen.size() >= max_classes) {
          use_classes = false;
          break;
        }
        cls = (uint16_t)seen.size();
        seen.emplace(ks, cls);
        tab.insert(tab.end(), J, J + dd);
      } else {
        cls = it->second;
      }
      const size_t t = (size_t)e / T, l = (size_t)e % T;
      gi[t * T + l] = cls;
    }
    if (use_classes) {
      // padding lanes need a zero Jinv: give them their own class
      if (npad > (size_t)h->n_owned) {
        const uint16_t zc = (uint16_t)(tab.size() / dd);
        tab.insert(tab.end(), dd, 0.0);
        for (size_t e = (size_t)h->n_owned; e < npad; ++e) gi[e] = zc;
      }
      h->n_geo_classes = (int64_t)(tab.size() / dd);
      SG_CUDA_H(h->geoidx.alloc(gi.size()));
      SG_CUDA_H(h->geotab.alloc(tab.size()));
      SG_CUDA_H(cudaMemcpy(h->geoidx.p, gi.data(), gi.size() * 2, cudaMemcpyHostToDevice));
      SG_CUDA_H(cudaMemcpy(h->geotab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    }
  }
  if (!use_classes) {
    SG_CUDA_H(h->geo.alloc(geo.size()));
    SG_CUDA_H(cudaMemcpy(h->geo.p, geo.data(), geo.size() * 8, cudaMemcpyHostToDevice));
  }

  {
    cudaDeviceProp prop;
    SG_CUDA_H(cudaGetDeviceProperties(&prop, d->device));
    h->nsm = prop.multiProcessorCount;
  }
  SG_CUDA_H(cudaStreamSynchronize(h->stream));
#undef SG_CUDA_H
  *out = h;
  return SG_OK;
}

void sg_destroy(sg_solver* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm) cudaStreamSynchronize(h->comm);
  drop_graph(h);
  h->u.release(); h->s.release(); h->uh.release(); h->sh.release();
  h->geo.release(); h->geotab.release(); h->geoidx.release(); h->mat.release(); h->absmat.release(); h->amp.release();
  h->nbr.release(); h->absidx.release(); h->code.release();
  h->src_start.release(); h->src_off.release(); h->step_dev.release(); h->send_cells.release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_sync) cudaEventDestroy(h->ev_sync);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (int k = 0; k < 6; ++k) {
    if (h->ev_int[k]) cudaEventDestroy(h->ev_int[k]);
    if (h->ev_bnd[k]) cudaEventDestroy(h->ev_bnd[k]);
  }
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  h->rec_cell.release(); h->rec_w.release(); h->rec_data.release();
  h->asym.release();
  h->sched.release(); h->ctl.release(); h->send_dst.release(); h->send_peer.release(); h->rfield.release(); h->rflag.release();
  h->push_start.release(); h->push_lane.release(); h->push_peer.release(); h->push_dst.release();
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->comm) cudaStreamDestroy(h->comm);
  delete h;
}

int sg_set_material(sg_solver* h, double density, double lam, double mu, const double* lam_cell,
                    const double* mu_cell) {
  if (!h) return fail(SG_EINVAL, "null solver");
  if ((lam_cell == nullptr) != (mu_cell == nullptr))
    return fail(SG_EINVAL, "sg_set_material: give both per-cell arrays or neither");
  SG_CUDA(cudaSetDevice(h->device));
  const bool per_cell = lam_cell != nullptr;
  const double* old_mat = h->mat.p;
  bool changed = !h->have_material || per_cell != h->per_cell || density != h->density ||
                 (!per_cell && (lam != h->lam_c || mu != h->mu_c));
  if (!changed && !per_cell) return SG_OK;          // same scalars again: nothing queued work could observe
  SG_CUDA(cudaStreamSynchronize(h->stream));
  h->density = density;
  h->lam_c = lam;
  h->mu_c = mu;
  h->per_cell = per_cell;
  if (h->per_cell) {
    // [tile][2][TILE]: lambda row then mu row of each tile, one bulk copy per tile
    const int T = h->tile;
    std::vector<double> m((size_t)h->n_owned_pad * 2, 0.0);
    for (int64_t e = 0; e < h->n_owned; ++e) {
      const size_t t = (size_t)e / T, l = (size_t)e % T;
      m[(t * 2 + 0) * T + l] = lam_cell[e];
      m[(t * 2 + 1) * T + l] = mu_cell[e];
    }
    SG_CUDA(h->mat.alloc(m.size()));
    SG_CUDA(cudaMemcpy(h->mat.p, m.data(), m.size() * 8, cudaMemcpyHostToDevice));
    changed = changed || h->mat.p != old_mat;
  }
  h->have_material = true;
  if (changed) h->config_version++;   // kernel arguments are baked into the step graph
  return SG_OK;
}

int sg_set_absorption(sg_solver* h, int64_t n, const int64_t* cell, const double* mats) {
  if (!h) return fail(SG_EINVAL, "null solver");
  if (n < 0 || (n > 0 && (!cell || !mats))) return fail(SG_EINVAL, "sg_set_absorption: bad arguments");
  SG_CUDA(cudaSetDevice(h->device));
  if (n == 0 && h->nabs_pad == 0) return SG_OK;
  SG_CUDA(cudaStreamSynchronize(h->stream));
  h->config_version++;
  if (n == 0) {
    h->nabs_pad = 0;
    h->absidx.release();
    h->absmat.release();
    return SG_OK;
  }
  const int nd2 = h->nd * h->nd;
  const int64_t npad = (n + 31) / 32 * 32;
  std::vector<int32_t> idx((size_t)h->n_owned_pad, -1);
  std::vector<double> m((size_t)npad * nd2, 0.0);
  for (int64_t k = 0; k < n; ++k) {
    if (cell[k] < 0 || cell[k] >= h->n_owned) return fail(SG_EINVAL, "sg_set_absorption: cell out of range");
    idx[(size_t)cell[k]] = (int32_t)k;
    for (int q = 0; q < nd2; ++q) m[(size_t)q * npad + k] = mats[(size_t)k * nd2 + q];
  }
  SG_CUDA(h->absidx.alloc(idx.size()));
  SG_CUDA(h->absmat.alloc(m.size()));
  SG_CUDA(cudaMemcpy(h->absidx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemcpy(h->absmat.p, m.data(), m.size() * 8, cudaMemcpyHostToDevice));
  h->nabs_pad = npad;
  return SG_OK;
}

int sg_set_source(sg_solver* h, int64_t nsrc, const int64_t* sdof, int64_t nsteps, const double* amp) {
  NvtxRange nvtx_range("sg_set_source");
  if (!h) return fail(SG_EINVAL, "null solver");
  if (nsrc < 0 || nsteps < 0 || (nsrc > 0 && (!sdof || (nsteps > 0 && !amp))))
    return fail(SG_EINVAL, "sg_set_source: bad arguments");
  SG_CUDA(cudaSetDevice(h->device));
  if ((nsrc == 0 || nsteps == 0) && h->nsrc == 0) return SG_OK;
  SG_CUDA(cudaStreamSynchronize(h->stream));
  h->config_version++;
  h->nsrc = 0;
  h->src_steps = 0;
  if (nsrc == 0 || nsteps == 0) return SG_OK;
  const int dd = h->dim * h->dim, D = h->dim;
  // entries sorted by tile: the G-type CTA that produces a tile adds that tile's source values (src_start/src_off)
  std::vector<std::pair<int64_t, int64_t>> key((size_t)nsrc);   // (tile * tile_elems + offset, original column)
  const int64_t tile_elems = (int64_t)h->KS * h->tile;
  for (int64_t k = 0; k < nsrc; ++k) {
    const int64_t dof = sdof[k];
    const int64_t cell = dof / ((int64_t)h->nd * dd);
    if (dof < 0 || cell >= h->n_owned) return fail(SG_EINVAL, "sg_set_source: dof outside owned cells");
    const int r = (int)(dof % ((int64_t)h->nd * dd));
    const int node = r / dd;
    int comp = r % dd;
    if (h->sym) {
      const int i = comp / D, j = comp % D, a = i < j ? i : j, b = i < j ? j : i;
      comp = a * D - a * (a - 1) / 2 + (b - a);
    }
    key[(size_t)k] = {(cell / h->tile) * tile_elems + (int64_t)(comp * h->nd + node) * h->tile + cell % h->tile, k};
  }
  std::sort(key.begin(), key.end());
  if (h->sym) {
    // (i,j) and (j,i) land on the same packed entry: they must carry the same values at every step (a zero entry
    // may be left out by the caller), then one of them is kept
    std::vector<std::pair<int64_t, int64_t>> uniq;
    for (int64_t k = 0; k < nsrc;) {
      int64_t e = k + 1;
      while (e < nsrc && key[(size_t)e].first == key[(size_t)k].first) ++e;
      const int r0 = (int)(sdof[key[(size_t)k].second] % ((int64_t)h->nd * dd)) % dd;
      const bool offdiag = r0 / D != r0 % D;
      bool ok = (e - k) == (offdiag ? 2 : 1);
      if (ok && offdiag) {
        const int64_t c0 = key[(size_t)k].second, c1 = key[(size_t)k + 1].second;
        ok = sdof[c0] != sdof[c1];
        for (int64_t n = 0; n < nsteps && ok; ++n) ok = amp[(size_t)(n * nsrc + c0)] == amp[(size_t)(n * nsrc + c1)];
      } else if (!ok && offdiag && e - k == 1) {
        // a lone off-diagonal entry is symmetric only if it is zero throughout
        ok = true;
        for (int64_t n = 0; n < nsteps && ok; ++n) ok = amp[(size_t)(n * nsrc + key[(size_t)k].second)] == 0.0;
      }
      if (!ok)
        return fail(e - k > 2 ? SG_EINVAL : SG_EASYM,
                    "sg_set_source: the source is not symmetric (or lists a dof twice) but the solver was created "
                    "with symmetric_stress = 1");
      uniq.push_back(key[(size_t)k]);
      k = e;
    }
    key.swap(uniq);
  } else {
    for (int64_t k = 1; k < nsrc; ++k)
      if (key[(size_t)k].first == key[(size_t)k - 1].first) return fail(SG_EINVAL, "sg_set_source: duplicate dof");
  }
  const int64_t ncol = nsrc;          // columns of the caller's table
  nsrc = (int64_t)key.size();         // entries kept on the device
  std::vector<int32_t> start((size_t)h->tiles_owned + 1, 0), off((size_t)nsrc);
  std::vector<double> a((size_t)nsrc * nsteps);
  for (int64_t k = 0; k < nsrc; ++k) {
    const int64_t t = key[(size_t)k].first / tile_elems;
    start[(size_t)t + 1]++;
    off[(size_t)k] = (int32_t)(key[(size_t)k].first % tile_elems);
    for (int64_t n = 0; n < nsteps; ++n) a[(size_t)(n * nsrc + k)] = amp[(size_t)(n * ncol + key[(size_t)k].second)];
  }
  for (size_t t = 0; t < (size_t)h->tiles_owned; ++t) start[t + 1] += start[t];
  SG_CUDA(h->src_start.alloc(start.size()));
  SG_CUDA(h->src_off.alloc(off.size()));
  SG_CUDA(h->amp.alloc(a.size()));
  SG_CUDA(cudaMemcpy(h->src_start.p, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemcpy(h->src_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemcpy(h->amp.p, a.data(), a.size() * 8, cudaMemcpyHostToDevice));
  h->nsrc = nsrc;
  h->src_steps = nsteps;
  return SG_OK;
}

int sg_set_state_async(sg_solver* h, const double* u, const double* s) {
  NvtxRange nvtx_range("sg_set_state");
  if (!h) return fail(SG_EINVAL, "null solver");
  SG_CUDA(cudaSetDevice(h->device));
  // the scratch fields double as staging buffers: they are dead between time steps
  if (u) {
    SG_CUDA(cudaMemcpyAsync(h->uh.p, u, (size_t)h->n_owned * h->KU * 8, cudaMemcpyHostToDevice, h->stream));
    int rc = relayout(h, h->u.p, h->uh.p, h->dim, true, h->stream, h->n_owned);
    if (rc) return rc;
  }
  if (s) {
    SG_CUDA(cudaMemcpyAsync(h->sh.p, s, (size_t)h->n_owned * h->KS_full * 8, cudaMemcpyHostToDevice, h->stream));
    int rc = relayout(h, h->s.p, h->sh.p, h->dim * h->dim, true, h->stream, h->n_owned);
    if (rc) return rc;
  }
  return SG_OK;
}

int sg_set_state_finish(sg_solver* h) {
  if (!h) return fail(SG_EINVAL, "null solver");
  SG_CUDA(cudaSetDevice(h->device));
  unsigned int asym = 0;
  if (h->sym) {
    SG_CUDA(cudaMemcpyAsync(&asym, h->asym.p, sizeof(asym), cudaMemcpyDeviceToHost, h->stream));
    SG_CUDA(cudaMemsetAsync(h->asym.p, 0, sizeof(asym), h->stream));
  }
  SG_CUDA(cudaStreamSynchronize(h->stream));
  if (asym)
    return fail(SG_EASYM, "sg_set_state: the stress handed in is not symmetric but the solver was created with "
                          "symmetric_stress = 1; create it with symmetric_stress = 0");
  return SG_OK;
}

int sg_set_state(sg_solver* h, const double* u, const double* s) {
  const int rc = sg_set_state_async(h, u, s);
  return rc ? rc : sg_set_state_finish(h);
}

int sg_get_state(sg_solver* h, double* u, double* s) {
  NvtxRange nvtx_range("sg_get_state");
  if (!h) return fail(SG_EINVAL, "null solver");
  SG_CUDA(cudaSetDevice(h->device));
  SG_CUDA(cudaStreamSynchronize(h->comm));
  // owned cells only; the scratch fields double as staging buffers (dead between time steps)
  if (u) {
    int rc = relayout(h, h->u.p, h->uh.p, h->dim, false, h->stream, h->n_owned);
    if (rc) return rc;
    SG_CUDA(cudaMemcpyAsync(u, h->uh.p, (size_t)h->n_owned * h->KU * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (s) {
    int rc = relayout(h, h->s.p, h->sh.p, h->dim * h->dim, false, h->stream, h->n_owned);
    if (rc) return rc;
    SG_CUDA(cudaMemcpyAsync(s, h->sh.p, (size_t)h->n_owned * h->KS_full * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  SG_CUDA(cudaStreamSynchronize(h->stream));
  return SG_OK;
}

int sg_get_field(sg_solver* h, int which, double* out) {
  if (!h || !out) return fail(SG_EINVAL, "sg_get_field: null argument");
  DevBuf<double>* f = field_buf(h, which);
  if (!f) return fail(SG_EINVAL, "sg_get_field: bad field id");
  SG_CUDA(cudaSetDevice(h->device));
  SG_CUDA(cudaStreamSynchronize(h->comm));
  const int nc = field_ncomp_boundary(h, which);
  DevBuf<double> tmp;   // test/diagnostic path: a private staging buffer keeps all four fields intact
  SG_CUDA(tmp.alloc((size_t)h->n_total * h->nd * nc));
  int rc = relayout(h, f->p, tmp.p, nc, false, h->stream, h->n_total);
  if (rc == SG_OK) {
    cudaError_t e = cudaMemcpyAsync(out, tmp.p, tmp.n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) rc = fail(SG_ECUDA, cudaGetErrorString(e));
  }
  tmp.release();
  return rc;
}

int sg_stage(sg_solver* h, int stage, int part, double dt, int64_t step) {
  if (!h) return fail(SG_EINVAL, "null solver");
  if (!h->have_material) return fail(SG_ESTATE, "sg_stage: call sg_set_material first");
  SG_CUDA(cudaSetDevice(h->device));
  if (h->nsrc > 0) {
    sg::set_step_kernel<<<1, 1, 0, h->stream>>>(h->step_dev.p, step);
    SG_CUDA(cudaGetLastError());
  }
  return launch_stage(h, stage, part, dt, h->stream);
}

int sg_step(sg_solver* h, int64_t nsteps, double dt, int64_t first_step) {
  NvtxRange nvtx_range("sg_step");
  if (!h) return fail(SG_EINVAL, "null solver");
  if (!h->have_material) return fail(SG_ESTATE, "sg_step: call sg_set_material first");
  if (nsteps < 0) return fail(SG_EINVAL, "sg_step: nsteps < 0");
  SG_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  if (!h->graph || h->graph_dt != dt || h->graph_version != h->config_version) {
    drop_graph(h);
    cudaGraph_t g = nullptr;
    SG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = SG_OK;
    if (h->npeers > 0 && env_int("SG_PEER_SCHED_SPLIT") > 0)
      rc = enqueue_step_peers(h, dt);
    else if (h->npeers > 0)
      rc = enqueue_step_fused(h, dt);
    else
      for (int k = 1; k <= 6 && rc == SG_OK; ++k) rc = launch_stage(h, k, SG_PART_ALL, dt, st, false, true);
    if (rc == SG_OK && h->nrec > 0)
      sg::receivers_kernel<<<1, 128, 0, st>>>(h->u.p, h->rec_cell.p, h->rec_w.p, h->rec_data.p, h->step_dev.p,
                                              h->rec_steps, (int)h->nrec, h->nd, h->dim, h->tile);
    const bool split_sched = h->npeers > 0 && env_int("SG_PEER_SCHED_SPLIT") > 0;
    if (rc == SG_OK && (h->nrec > 0 || (h->nsrc > 0 && split_sched))) sg::bump_step_kernel<<<1, 1, 0, st>>>(h->step_dev.p);
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != SG_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (e != cudaSuccess) return fail(SG_ECUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
      h->graph = nullptr;
      return fail(SG_ECUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    }
    h->graph_dt = dt;
    h->graph_version = h->config_version;
  }
  if (h->nsrc > 0 || h->nrec > 0) {
    sg::set_step_kernel<<<1, 1, 0, st>>>(h->step_dev.p, first_step);
    SG_CUDA(cudaGetLastError());
  }
  SG_CUDA(cudaEventRecord(h->ev0, st));
  for (int64_t n = 0; n < nsteps; ++n) SG_CUDA(cudaGraphLaunch(h->graph, st));
  SG_CUDA(cudaEventRecord(h->ev1, st));
  return SG_OK;
}

int sg_time_stage(sg_solver* h, int stage, int part, double dt, int reps, double* ms_avg) {
  if (!h || !ms_avg || reps <= 0) return fail(SG_EINVAL, "sg_time_stage: bad arguments");
  if (!h->have_material) return fail(SG_ESTATE, "sg_time_stage: call sg_set_material first");
  SG_CUDA(cudaSetDevice(h->device));
  int rc = launch_stage(h, stage, part, dt, h->stream);   // warm-up launch, untimed
  if (rc) return rc;
  SG_CUDA(cudaEventRecord(h->ev0, h->stream));
  for (int r = 0; r < reps; ++r) {
    rc = launch_stage(h, stage, part, dt, h->stream);
    if (rc) return rc;
  }
  SG_CUDA(cudaEventRecord(h->ev1, h->stream));
  SG_CUDA(cudaEventSynchronize(h->ev1));
  float f = 0.f;
  SG_CUDA(cudaEventElapsedTime(&f, h->ev0, h->ev1));
  *ms_avg = (double)f / reps;
  return SG_OK;
}

int sg_synchronize(sg_solver* h) {
  if (!h) return fail(SG_EINVAL, "null solver");
  SG_CUDA(cudaSetDevice(h->device));
  SG_CUDA(cudaStreamSynchronize(h->stream));
  SG_CUDA(cudaStreamSynchronize(h->comm));
  return SG_OK;
}

int sg_set_receivers(sg_solver* h, int64_t n, const int64_t* cell, const double* weights, int64_t max_steps) {
  if (!h) return fail(SG_EINVAL, "null solver");
  if (n < 0 || max_steps < 0 || (n > 0 && (!cell || !weights))) return fail(SG_EINVAL, "sg_set_receivers: bad arguments");
  SG_CUDA(cudaSetDevice(h->device));
  if (n == 0 && h->nrec == 0) return SG_OK;
  SG_CUDA(cudaStreamSynchronize(h->stream));
  h->config_version++;
  h->nrec = 0;
  h->rec_steps = 0;
  if (n == 0 || max_steps == 0) return SG_OK;
  std::vector<int64_t> c((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    if (cell[k] < 0 || cell[k] >= h->n_owned) return fail(SG_EINVAL, "sg_set_receivers: cell is not owned");
    c[(size_t)k] = cell[k];
  }
  SG_CUDA(h->rec_cell.alloc((size_t)n));
  SG_CUDA(h->rec_w.alloc((size_t)n * h->nd));
  SG_CUDA(h->rec_data.alloc((size_t)max_steps * n * h->dim));
  SG_CUDA(cudaMemcpy(h->rec_cell.p, c.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemcpy(h->rec_w.p, weights, (size_t)n * h->nd * 8, cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemset(h->rec_data.p, 0, (size_t)max_steps * n * h->dim * 8));
  h->nrec = n;
  h->rec_steps = max_steps;
  return SG_OK;
}

int sg_get_receivers(sg_solver* h, int64_t first_step, int64_t nsteps, double* out) {
  if (!h || !out) return fail(SG_EINVAL, "sg_get_receivers: null argument");
  if (first_step < 0 || nsteps < 0 || first_step + nsteps > h->rec_steps)
    return fail(SG_EINVAL, "sg_get_receivers: step range outside the recorded window");
  SG_CUDA(cudaSetDevice(h->device));
  const size_t row = (size_t)h->nrec * h->dim;
  SG_CUDA(cudaMemcpyAsync(out, h->rec_data.p + (size_t)first_step * row, (size_t)nsteps * row * 8,
                          cudaMemcpyDeviceToHost, h->stream));
  SG_CUDA(cudaStreamSynchronize(h->stream));
  return SG_OK;
}

int sg_record_receivers(sg_solver* h, int64_t step) {
  if (!h) return fail(SG_EINVAL, "null solver");
  if (h->nrec == 0) return SG_OK;
  SG_CUDA(cudaSetDevice(h->device));
  sg::set_step_kernel<<<1, 1, 0, h->stream>>>(h->step_dev.p, step);
  SG_CUDA(cudaGetLastError());
  sg::receivers_kernel<<<1, 128, 0, h->stream>>>(h->u.p, h->rec_cell.p, h->rec_w.p, h->rec_data.p, h->step_dev.p,
                                                 h->rec_steps, (int)h->nrec, h->nd, h->dim, h->tile);
  SG_CUDA(cudaGetLastError());
  return SG_OK;
}

int sg_mark(sg_solver* h, int which) {
  if (!h || (which != 0 && which != 1)) return fail(SG_EINVAL, "sg_mark: bad arguments");
  SG_CUDA(cudaSetDevice(h->device));
  SG_CUDA(cudaEventRecord(which ? h->ev1 : h->ev0, h->stream));
  return SG_OK;
}

int sg_last_step_ms(sg_solver* h, double* ms) {
  if (!h || !ms) return fail(SG_EINVAL, "sg_last_step_ms: null argument");
  SG_CUDA(cudaSetDevice(h->device));
  SG_CUDA(cudaEventSynchronize(h->ev1));
  float f = 0.f;
  SG_CUDA(cudaEventElapsedTime(&f, h->ev0, h->ev1));
  *ms = f;
  return SG_OK;
}

int sg_set_halo_plan(sg_solver* h, int64_t nsend, const int64_t* send_cells) {
  if (!h) return fail(SG_EINVAL, "null solver");
  if (nsend < 0 || (nsend > 0 && !send_cells)) return fail(SG_EINVAL, "sg_set_halo_plan: bad arguments");
  SG_CUDA(cudaSetDevice(h->device));
  for (int64_t k = 0; k < nsend; ++k)
    if (send_cells[k] < 0 || send_cells[k] >= h->n_owned)
      return fail(SG_EINVAL, "sg_set_halo_plan: send cell is not owned");
  SG_CUDA(h->send_cells.alloc((size_t)nsend));
  if (nsend) SG_CUDA(cudaMemcpy(h->send_cells.p, send_cells, (size_t)nsend * 8, cudaMemcpyHostToDevice));
  h->nsend = nsend;
  return SG_OK;
}

int sg_pack(sg_solver* h, int which, double* dst, int on_comm_stream) {
  if (!h) return fail(SG_EINVAL, "null solver");
  DevBuf<double>* f = field_buf(h, which);
  if (!f || (!dst && h->nsend)) return fail(SG_EINVAL, "sg_pack: bad arguments");
  if (h->nsend == 0) return SG_OK;
  SG_CUDA(cudaSetDevice(h->device));
  const int K = field_ncomp(h, which) * h->nd;
  cudaStream_t st = on_comm_stream ? h->comm : h->stream;
  sg::pack_kernel<<<grid_for(h->nsend * K), 256, 0, st>>>(f->p, h->send_cells.p, h->nsend, K, h->tile, dst);
  SG_CUDA(cudaGetLastError());
  return SG_OK;
}

int sg_unpack(sg_solver* h, int which, const double* src, int64_t first, int64_t count, int on_comm_stream) {
  if (!h) return fail(SG_EINVAL, "null solver");
  DevBuf<double>* f = field_buf(h, which);
  if (!f || first < 0 || count < 0 || first + count > h->n_halo || (!src && count))
    return fail(SG_EINVAL, "sg_unpack: bad arguments");
  if (count == 0) return SG_OK;
  SG_CUDA(cudaSetDevice(h->device));
  const int K = field_ncomp(h, which) * h->nd;
  cudaStream_t st = on_comm_stream ? h->comm : h->stream;
  sg::unpack_kernel<<<grid_for(count * K), 256, 0, st>>>(f->p, h->n_owned_pad + first, count, K, h->tile, src);
  SG_CUDA(cudaGetLastError());
  return SG_OK;
}

int sg_comm_wait_compute(sg_solver* h) {
  if (!h) return fail(SG_EINVAL, "null solver");
  SG_CUDA(cudaEventRecord(h->ev_sync, h->stream));
  SG_CUDA(cudaStreamWaitEvent(h->comm, h->ev_sync, 0));
  return SG_OK;
}
int sg_compute_wait_comm(sg_solver* h) {
  if (!h) return fail(SG_EINVAL, "null solver");
  SG_CUDA(cudaEventRecord(h->ev_sync, h->comm));
  SG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_sync, 0));
  return SG_OK;
}
void* sg_stream(sg_solver* h, int comm) { return h ? (void*)(comm ? h->comm : h->stream) : nullptr; }
void* sg_field_ptr(sg_solver* h, int which) {
  if (!h) return nullptr;
  DevBuf<double>* f = field_buf(h, which);
  return f ? (void*)f->p : nullptr;
}

int sg_ipc_export(sg_solver* h, unsigned char* out) {
  if (!h || !out) return fail(SG_EINVAL, "sg_ipc_export: null argument");
  SG_CUDA(cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == SG_IPC_HANDLE_BYTES, "handle size");
  void* bufs[5] = {h->u.p, h->s.p, h->uh.p, h->sh.p, h->ctl.p};
  for (int i = 0; i < 5; ++i) {
    cudaIpcMemHandle_t mh;
    SG_CUDA(cudaIpcGetMemHandle(&mh, bufs[i]));
    std::memcpy(out + (size_t)i * SG_IPC_HANDLE_BYTES, &mh, SG_IPC_HANDLE_BYTES);
  }
  return SG_OK;
}

int sg_peer_connect(sg_solver* h, int32_t npeers, const sg_peer_desc* peers) {
  NvtxRange nvtx_range("sg_peer_connect");
  if (!h || npeers < 0 || npeers > 16 || (npeers > 0 && !peers)) return fail(SG_EINVAL, "sg_peer_connect: bad arguments");
  SG_CUDA(cudaSetDevice(h->device));
  SG_CUDA(cudaStreamSynchronize(h->stream));
  SG_CUDA(cudaStreamSynchronize(h->comm));
  h->config_version++;
  h->npeers = 0;
  drop_graph(h);                                   // the step graph holds the peers' pointers
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  h->ipc_opened.clear();
  if (npeers == 0) return SG_OK;
  std::vector<double*> rf((size_t)4 * npeers);
  std::vector<unsigned long long*> fl((size_t)npeers);
  std::vector<int64_t> dst((size_t)h->nsend, -1);
  std::vector<int32_t> who((size_t)h->nsend, -1);
  for (int i = 0; i < npeers; ++i) {
    const sg_peer_desc& d = peers[i];
    if (d.send_offset < 0 || d.send_count < 0 || d.send_offset + d.send_count > h->nsend || d.flag_slot < 0 ||
        d.flag_slot >= 16 || d.remote_first_cell < 0)
      return fail(SG_EINVAL, "sg_peer_connect: bad peer descriptor");
    void* ptrs[5];
    for (int b = 0; b < 5; ++b) {
      cudaIpcMemHandle_t mh;
      std::memcpy(&mh, d.handles[b], SG_IPC_HANDLE_BYTES);
      cudaError_t e = cudaIpcOpenMemHandle(&ptrs[b], mh, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return fail(SG_ECUDA, std::string("cudaIpcOpenMemHandle (peer rank ") + std::to_string(d.rank) + "): " +
                                  cudaGetErrorString(e));
      h->ipc_opened.push_back(ptrs[b]);
    }
    for (int b = 0; b < 4; ++b) rf[(size_t)b * npeers + i] = (double*)ptrs[b];
    fl[(size_t)i] = (unsigned long long*)ptrs[4] + d.flag_slot;
    for (int64_t k = 0; k < d.send_count; ++k) {
      dst[(size_t)(d.send_offset + k)] = d.remote_first_cell + k;
      who[(size_t)(d.send_offset + k)] = i;
    }
  }
  for (int64_t k = 0; k < h->nsend; ++k)
    if (who[(size_t)k] < 0) return fail(SG_EINVAL, "sg_peer_connect: a send cell has no destination");
  // the send list regrouped by (tile, peer, remote cell) for the fused exchange; needs the send cells on the host
  {
    std::vector<int64_t> cells((size_t)h->nsend);
    if (h->nsend) SG_CUDA(cudaMemcpy(cells.data(), h->send_cells.p, (size_t)h->nsend * 8, cudaMemcpyDeviceToHost));
    const int T = h->tile;
    h->push_tiles = h->tiles_boundary;
    std::vector<int64_t> order((size_t)h->nsend);
    for (int64_t k = 0; k < h->nsend; ++k) {
      order[(size_t)k] = k;
      if (cells[(size_t)k] / T >= h->tiles_boundary)
        return fail(SG_EINVAL, "sg_peer_connect: a send cell lies outside the boundary tiles (n_boundary too small)");
    }
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      const int64_t ta = cells[(size_t)a] / T, tb = cells[(size_t)b] / T;
      if (ta != tb) return ta < tb;
      if (who[(size_t)a] != who[(size_t)b]) return who[(size_t)a] < who[(size_t)b];
      return dst[(size_t)a] < dst[(size_t)b];
    });
    std::vector<int32_t> pstart((size_t)h->push_tiles + 1, 0), plane((size_t)h->nsend), ppeer((size_t)h->nsend);
    std::vector<int64_t> pdst((size_t)h->nsend);
    for (int64_t i = 0; i < h->nsend; ++i) {
      const int64_t k = order[(size_t)i];
      pstart[(size_t)(cells[(size_t)k] / T) + 1]++;
      plane[(size_t)i] = (int32_t)(cells[(size_t)k] % T);
      ppeer[(size_t)i] = who[(size_t)k];
      pdst[(size_t)i] = dst[(size_t)k];
    }
    for (int t = 0; t < h->push_tiles; ++t) pstart[(size_t)t + 1] += pstart[(size_t)t];
    SG_CUDA(h->push_start.alloc(pstart.size()));
    SG_CUDA(h->push_lane.alloc(std::max<size_t>(plane.size(), 1)));
    SG_CUDA(h->push_peer.alloc(std::max<size_t>(ppeer.size(), 1)));
    SG_CUDA(h->push_dst.alloc(std::max<size_t>(pdst.size(), 1)));
    SG_CUDA(cudaMemcpy(h->push_start.p, pstart.data(), pstart.size() * 4, cudaMemcpyHostToDevice));
    if (h->nsend) {
      SG_CUDA(cudaMemcpy(h->push_lane.p, plane.data(), plane.size() * 4, cudaMemcpyHostToDevice));
      SG_CUDA(cudaMemcpy(h->push_peer.p, ppeer.data(), ppeer.size() * 4, cudaMemcpyHostToDevice));
      SG_CUDA(cudaMemcpy(h->push_dst.p, pdst.data(), pdst.size() * 8, cudaMemcpyHostToDevice));
    }
  }
  SG_CUDA(h->rfield.alloc(rf.size()));
  SG_CUDA(h->rflag.alloc(fl.size()));
  SG_CUDA(h->send_dst.alloc(dst.size()));
  SG_CUDA(h->send_peer.alloc(who.size()));
  SG_CUDA(cudaMemcpy(h->rfield.p, rf.data(), rf.size() * sizeof(double*), cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemcpy(h->rflag.p, fl.data(), fl.size() * sizeof(void*), cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemcpy(h->send_dst.p, dst.data(), dst.size() * 8, cudaMemcpyHostToDevice));
  SG_CUDA(cudaMemcpy(h->send_peer.p, who.data(), who.size() * 4, cudaMemcpyHostToDevice));
  h->npeers = npeers;
  return SG_OK;
}

int sg_exchange(sg_solver* h, int which) {
  if (!h || !field_buf(h, which)) return fail(SG_EINVAL, "sg_exchange: bad arguments");
  SG_CUDA(cudaSetDevice(h->device));
  return enqueue_exchange(h, which, h->stream);
}

int sg_peer_error(sg_solver* h, int64_t* err) {
  if (!h || !err) return fail(SG_EINVAL, "sg_peer_error: null argument");
  SG_CUDA(cudaSetDevice(h->device));
  SG_CUDA(cudaStreamSynchronize(h->stream));
  SG_CUDA(cudaStreamSynchronize(h->comm));
  unsigned long long v = 0;
  SG_CUDA(cudaMemcpy(&v, h->ctl.p + sg::SG_CTL_ERROR, 8, cudaMemcpyDeviceToHost));
  *err = (int64_t)v;
  return SG_OK;
}

void* sg_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0) return nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
    g_err = "sg_host_alloc: cudaHostAlloc failed";
    return nullptr;
  }
  return p;
}
void sg_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
