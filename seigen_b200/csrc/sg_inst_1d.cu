// Stage-kernel instantiations of the 1-D elements (the reference's tests/pulse/pulse_1d_lf4.py: IntervalMesh, DG P1,
// dimension=1).  One scalar velocity and one stress component per node; a facet is a point (NFP = 1).  These meshes
// are tiny (400 cells in the shipped script), so there is one untuned variant per degree.
//                D  P  TILE SPLIT MINB MINBA NSP NSA AXS   XREG
#define SG_STAGE_KERNELS_ONLY
#include "sg_variants.h"

void sg_variants_1d(std::vector<Variant>& v) {
  v.push_back(make_variant<1, 1, 128, 1, 4, 4, 2, 2, true, true>());
  v.push_back(make_variant<1, 2, 128, 1, 4, 4, 2, 2, true, true>());
  v.push_back(make_variant<1, 3, 128, 1, 4, 4, 2, 2, true, true>());
}
