/* Graph partitioner entry point of seigen_b200 (include/seigen_b200.h: sg_partition_graph).
 *
 * The reference gets its mesh distribution from PETSc DMPlex (seigen/elastic.py:404-414 -> mesh.topology.init), which
 * hands the dual graph to a graph partitioner.  Here the dual graph (cells, shared facets) built by
 * seigen_b200/mesh.py goes to METIS 5 -- the static library NVIDIA ships with the CUDA toolkit for cuSOLVER
 * (targets/x86_64-linux/lib/libmetis_static.a; no header is shipped, so the two entry points are declared here with
 * the ABI that build uses: idx_t = int64_t; real_t arguments are only ever passed as NULL = METIS defaults).
 */
#include <stddef.h>
#include <stdint.h>

typedef int64_t idx_t;
int METIS_PartGraphKway(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt, idx_t* vsize, idx_t* adjwgt,
                        idx_t* nparts, void* tpwgts, void* ubvec, idx_t* options, idx_t* edgecut, idx_t* part);
int METIS_PartGraphRecursive(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt, idx_t* vsize,
                             idx_t* adjwgt, idx_t* nparts, void* tpwgts, void* ubvec, idx_t* options, idx_t* edgecut,
                             idx_t* part);

/* CSR graph (xadj[n+1], adjncy[xadj[n]]) -> part[n] in [0, nparts), *edgecut = cut edges.  Returns 0 on success,
 * -1 on bad arguments, -2 if METIS reports an error. */
int sg_partition_graph(int64_t n, const int64_t* xadj, const int64_t* adjncy, int64_t nparts, int recursive,
                       int64_t* part, int64_t* edgecut) {
  if (n <= 0 || nparts <= 0 || !xadj || !adjncy || !part || !edgecut) return -1;
  if (nparts == 1) {
    for (int64_t i = 0; i < n; ++i) part[i] = 0;
    *edgecut = 0;
    return 0;
  }
  idx_t nv = n, ncon = 1, np = nparts, ec = 0;
  const int rc = recursive
                     ? METIS_PartGraphRecursive(&nv, &ncon, (idx_t*)xadj, (idx_t*)adjncy, NULL, NULL, NULL, &np, NULL,
                                                NULL, NULL, &ec, part)
                     : METIS_PartGraphKway(&nv, &ncon, (idx_t*)xadj, (idx_t*)adjncy, NULL, NULL, NULL, &np, NULL, NULL,
                                           NULL, &ec, part);
  *edgecut = ec;
  return rc == 1 ? 0 : -2; /* METIS_OK == 1 */
}
