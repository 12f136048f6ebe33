/* seigen_b200 -- C ABI of the B200-native ElasticLF4 explicit velocity-stress update.
 *
 * This is the boundary a host binding (ctypes in seigen_b200/capi.py) talks to.
 * It replaces, for the explicit ElasticLF4 path of devitocodes/seigen only, what the reference
 * obtains from Firedrake / PyOP2 / PETSc at run time.  Citations are relative to the reference
 * repository (seigen/elastic.py unless another file is named).
 *
 * Conventions
 *   - every function returns SG_OK (0) or a negative SG_E* code; sg_last_error() returns a
 *     thread-local, human readable message for the last failure on the calling thread;
 *   - the caller owns every host array; the library copies what it needs before returning and
 *     keeps no host pointer;
 *   - an sg_solver owns its device memory, stream, events and CUDA graphs.  One solver must not be
 *     used from two threads at once; distinct solvers are independent;
 *   - fields cross the boundary in Firedrake's dat.data layout: velocity u[cell*nd + node][i],
 *     stress s[cell*nd + node][i][j], C-contiguous float64.  Inside the library
 *     they live in a tile-blocked SoA layout;
 *   - there is no CPU fallback: every entry point needs a CUDA device.
 */
#ifndef SEIGEN_B200_H
#define SEIGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_OK 0
#define SG_EINVAL (-1)   /* bad argument / unsupported (dim, degree) */
#define SG_ECUDA (-2)    /* a CUDA runtime call failed */
#define SG_ESTATE (-3)   /* call made in the wrong state (e.g. stepping before materials are set) */
#define SG_EASYM (-4)    /* asymmetric stress / source handed to a solver created with symmetric_stress = 1 */

#define SG_BOUNDARY 0x80 /* bit 7 of a facet code marks an exterior facet */

typedef struct sg_solver sg_solver;

/* Mesh partition handed to sg_create.  It is the facet-to-cell adjacency that stands in for
 * PyOP2's cell_node_map / interior_facet_node_map / exterior_facet_node_map (the maps every
 * par_loop of ExplicitElasticLF4.solve, elastic.py:358-367, is driven by) plus the affine
 * geometry TSFC would recompute from the coordinate field in every kernel. */
typedef struct sg_mesh_desc {
  int32_t dim;          /* 2 (triangles) or 3 (tetrahedra) */
  int32_t degree;       /* DG polynomial degree: 1..4 in 2D, 1..3 in 3D (elastic.py:81-82) */
  int64_t n_owned;      /* cells this solver updates */
  int64_t n_total;      /* n_owned + halo cells; halo cells (index >= n_owned) are only read */
  const int32_t* nbr;   /* [n_owned][dim+1] cell across facet f (own index on an exterior facet) */
  const uint8_t* code;  /* [n_owned][dim+1] f' * dim! + s (neighbour's facet number and gluing
                           permutation, seigen_b200/refelem.py ftab), | SG_BOUNDARY if exterior */
  const double* jinv;   /* [n_owned][dim][dim] Jinv[r][k] = d(xi_r)/d(x_k) */
  int32_t device;       /* CUDA device ordinal */
  int32_t n_boundary;   /* cells [0, n_boundary) touch a partition cut: they are updated first so the
                           halo exchange of their values overlaps the rest (0 on a single GPU) */
  int32_t geom_classes; /* 1: cells whose Jinv agree to 2^-36 relative (translates of one another on uniform
                           meshes) share one geometry record, saving dim*dim*8 B of HBM traffic per cell per
                           pass; falls back to per-cell geometry above 4096 classes.  0: always per cell */
  int32_t symmetric_stress; /* 1: keep only the upper triangle of every stress field on the device.  The stress RHS
                           g (elastic.py:211-219) is symmetric by construction, so if s0 and the source are symmetric
                           every stress of the run is, bit for bit, and the step moves 1/6 (2D) to 1/4 (3D) fewer
                           bytes.  The boundary still carries all dim*dim components (TensorFunctionSpace,
                           elastic.py:81); sg_set_state / sg_set_source return SG_EASYM if the premise fails and
                           the caller re-creates the solver with 0 (what seigen_b200/elastic.py does).  0: store all
                           dim*dim components, no premise */
} sg_mesh_desc;

/* Which device field sg_get_field / sg_field_ptr address. */
#define SG_FIELD_U 0   /* u0 / u1      (elastic.py:98, 102) */
#define SG_FIELD_S 1   /* s0 / s1      (elastic.py:93, 97)  */
#define SG_FIELD_UH 2  /* uh1 / utemp  (elastic.py:99-100)  -- scratch velocity */
#define SG_FIELD_SH 3  /* stemp / sh1  (elastic.py:95, 94)  -- scratch stress   */

/* Which cells a stage launch covers (multi-GPU overlap). */
#define SG_PART_ALL 0
#define SG_PART_BOUNDARY 1
#define SG_PART_INTERIOR 2

const char* sg_last_error(void);
int sg_version(void);

/* ElasticLF4.__init__ (elastic.py:66-124): allocates u0/s0 and the two scratch fields on the device. */
int sg_create(sg_solver** out, const sg_mesh_desc* desc);
void sg_destroy(sg_solver* h);

/* Plain attributes elastic.density / .l / .mu (tests/eigenmode/eigenmode_2d.py:17-20).  lam_cell / mu_cell
 * are optional per-cell values [n_owned]; NULL = use the scalars. */
int sg_set_material(sg_solver* h, double density, double lam, double mu,
                    const double* lam_cell, const double* mu_cell);

/* elastic.absorption (elastic.py:136-141, term elastic.py:207-208).  For each listed cell the caller passes
 * the nd x nd matrix A = sum_b sigma_b * Minv * T[:, b, :] so that P(sigma, u)_i = A u_i.  n = 0 clears. */
int sg_set_absorption(sg_solver* h, int64_t n, const int64_t* cell, const double* mats);

/* elastic.source (elastic.py:149-154, 285-288): nodal source values.  sdof[k] indexes the stress field in
 * boundary layout, amp[step][k] is the value interpolated at the END time of step `step` (0-based).
 * nsrc = 0 clears. */
int sg_set_source(sg_solver* h, int64_t nsrc, const int64_t* sdof, int64_t nsteps, const double* amp);

/* u0.assign / s0.assign (tests/explosive_source/explosive_source_lf4.py:48-52); arrays are [n_owned*nd][d] and
 * [n_owned*nd][d][d] (owned cells; halo cells are filled by the halo exchange).  Either pointer may be NULL (field
 * left unchanged). */
int sg_set_state(sg_solver* h, const double* u, const double* s);
/* The same in two halves, so that host-side setup (material / source tables) can run while the state crosses PCIe:
 * _async queues the copies (the host arrays must stay untouched until _finish), _finish waits and reports SG_EASYM. */
int sg_set_state_async(sg_solver* h, const double* u, const double* s);
int sg_set_state_finish(sg_solver* h);
/* u1.dat.data / s1.dat.data after run (elastic.py:315), owned cells.  Synchronises with all queued work. */
int sg_get_state(sg_solver* h, double* u, double* s);
int sg_get_field(sg_solver* h, int which, double* out);

/* Receivers: the sensors tests/explosive_source/uy.py:36-43 probes in the VTU output ((45,149), (90,149), (140,149)).
 * Receiver k sits in owned cell cell[k] with basis values weights[k][nd] at its position; after every time step
 * sg_step records u at the receivers on the device.  sg_get_receivers copies steps [first_step, first_step+nsteps)
 * as out[step][k][dim].  n = 0 clears. */
int sg_set_receivers(sg_solver* h, int64_t n, const int64_t* cell, const double* weights, int64_t max_steps);
int sg_get_receivers(sg_solver* h, int64_t first_step, int64_t nsteps, double* out);

/* For drivers that advance stage by stage with sg_stage (library transport, SG_HALO=nccl): record the receivers for
 * 0-based step `step` now, on the compute stream -- what sg_step does itself after every step. */
int sg_record_receivers(sg_solver* h, int64_t step);

/* The loop body of ElasticLF4.run (elastic.py:283-304) `nsteps` times, starting at 0-based step index
 * `first_step` (only used to index the source table).  Work is queued on the solver's stream; the call
 * returns without waiting (sg_get_state / sg_synchronize wait). */
int sg_step(sg_solver* h, int64_t nsteps, double dt, int64_t first_step);
int sg_synchronize(sg_solver* h);
/* Device time in milliseconds between the start and the end of the last sg_step call (CUDA events on the
 * solver's stream); synchronises. */
int sg_last_step_ms(sg_solver* h, double* ms);
/* Records the start (which = 0) / end (which = 1) event of that interval on the compute stream by hand: for
 * drivers that advance stage by stage with sg_stage (the role of PyOP2's timed_region('timestepping'),
 * elastic.py:278). */
int sg_mark(sg_solver* h, int which);

/* One of the six fused passes (K1 uh1, K2 stemp, K3 u1, K4 sh1, K5 utemp, K6 s1), on all cells or on
 * the boundary / interior part.  Used by per-stage parity tests and by the multi-GPU driver, which exchanges
 * halos between stages.  `step` indexes the source table. */
int sg_stage(sg_solver* h, int stage, int part, double dt, int64_t step);

/* Measurement aid: `reps` back-to-back launches of one stage's kernel on the solver's stream between two CUDA
 * events; *ms_avg = elapsed / reps.  The state is advanced as a side effect. */
int sg_time_stage(sg_solver* h, int stage, int part, double dt, int reps, double* ms_avg);

/* Multi-GPU plumbing.  Halo exchange moves whole cells in device (tile-blocked) layout:
 * sg_set_halo_plan registers the owned cells whose values neighbours need, grouped by destination;
 * sg_pack gathers field `which` of those cells into a contiguous device buffer (cell-major, K doubles per
 * cell); sg_unpack scatters a received buffer into halo cells [n_owned + first, n_owned + first + count).
 * Buffers are device pointers owned by the caller (e.g. torch tensors used with torch.distributed). */
int sg_set_halo_plan(sg_solver* h, int64_t nsend, const int64_t* send_cells);
int sg_pack(sg_solver* h, int which, double* dst, int on_comm_stream);
int sg_unpack(sg_solver* h, int which, const double* src, int64_t first, int64_t count, int on_comm_stream);
/* Stream ordering helpers for the overlap schedule: the solver has a compute and a comm stream. */
int sg_comm_wait_compute(sg_solver* h);   /* comm stream waits for everything queued on compute */
int sg_compute_wait_comm(sg_solver* h);   /* compute stream waits for everything queued on comm */
void* sg_stream(sg_solver* h, int comm);  /* cudaStream_t, for torch.cuda.ExternalStream */
void* sg_field_ptr(sg_solver* h, int which);

/* Peer-memory halo exchange (one process per GPU on one NVSwitch domain).  Instead of packing into a buffer
 * that a library sends, each rank writes the rows of its cut-adjacent cells straight into the halo tiles of the
 * neighbouring ranks' fields through CUDA-IPC mappings (NVLink stores), then publishes an epoch flag in the
 * neighbour's memory; the consumer spins on its own flag.  With peers connected, sg_step replays ONE CUDA graph per
 * time step whose six stage kernels carry the six exchanges themselves: the CTAs that compute the tiles of the
 * cut-adjacent cells (handed out first) wait for the peers' rows of the previous pass, store their own rows into the
 * peers' halo tiles and publish the epoch, while the other CTAs are already on the interior tiles -- no host work,
 * no library call, no extra kernel inside a step.  (SG_PEER_SCHED_SPLIT=1 selects the earlier two-stream schedule:
 * boundary kernel -> push/signal/wait kernels on a comm stream next to the interior kernel.)
 * SG_PEER_TIMEOUT_S bounds every device-side wait (default ~20 s).
 *   sg_ipc_export    5 handles of SG_IPC_HANDLE_BYTES bytes: u, s, uh, sh, control words
 *   sg_peer_connect  the handles of every neighbouring rank + where my cells land in its fields
 *   sg_exchange      one immediate exchange of field `which` on the compute stream (after sg_set_state)
 *   sg_peer_error    synchronises; *err != 0 if a wait timed out (a peer never published its rows)
 * Every rank must issue the same sequence of exchanges.  The halo cells of a field are written by exchanges only
 * (sg_set_state ignores the halo part of its arguments). */
#define SG_IPC_HANDLE_BYTES 64
typedef struct sg_peer_desc {
  int32_t rank;               /* the neighbour (informational) */
  int32_t flag_slot;          /* my index among that neighbour's peers: the control word I publish into */
  int64_t send_offset;        /* my cells for this neighbour: send_cells[send_offset, send_offset + send_count) */
  int64_t send_count;
  int64_t remote_first_cell;  /* device cell index (tile-padded) of the first of them in the neighbour's fields */
  unsigned char handles[5][SG_IPC_HANDLE_BYTES];
} sg_peer_desc;
int sg_ipc_export(sg_solver* h, unsigned char* out);
int sg_peer_connect(sg_solver* h, int32_t npeers, const sg_peer_desc* peers);
int sg_exchange(sg_solver* h, int which);
int sg_peer_error(sg_solver* h, int64_t* err);

/* Sizes, so that a binding can allocate: nd nodes per cell, tile (cells per device tile). */
int sg_nodes_per_cell(int dim, int degree);
int sg_tile_cells(int dim, int degree);

/* Page-locked host buffers for sg_set_state / sg_get_state (optional; any host pointer works). */
void* sg_host_alloc(int64_t bytes);
void sg_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SEIGEN_B200_H */
