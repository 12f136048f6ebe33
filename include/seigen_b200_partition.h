/* seigen_b200 -- graph partitioner entry point (libsg_metis.so, built from seigen_b200/csrc/sg_metis.c).
 *
 * Replaces, for the ElasticLF4 path, the mesh distribution the reference obtains from PETSc DMPlex
 * (seigen/elastic.py:404-414, mesh.topology.init(s_depth=...)): the dual graph of the mesh (cells,
 * shared facets) is cut into `nparts` balanced parts minimising the number of cut facets, i.e. the halo volume every
 * LF4 pass exchanges.  Host-side, no CUDA device needed.  Binding: seigen_b200/partition_metis.py (ctypes).
 */
#ifndef SEIGEN_B200_PARTITION_H
#define SEIGEN_B200_PARTITION_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* CSR graph: xadj[n + 1], adjncy[xadj[n]] (every edge listed from both ends).  part[n] receives values in
 * [0, nparts); *edgecut the number of cut edges.  recursive = 0: multilevel k-way, 1: multilevel recursive bisection.
 * Returns 0, -1 (bad arguments) or -2 (partitioner error).  Deterministic for a given graph. */
int sg_partition_graph(int64_t n, const int64_t* xadj, const int64_t* adjncy, int64_t nparts, int recursive,
                       int64_t* part, int64_t* edgecut);
#ifdef __cplusplus
}
#endif
#endif
