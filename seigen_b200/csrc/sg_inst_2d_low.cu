// Stage-kernel instantiations of one group of elements (see sg_variants.h).  First entry of each (dim, degree) is the
// default; the others are tuning candidates selectable with SG_TILE / SG_SPLIT / SG_MINB / SG_NS (scripts/perf_probe.py).
//                D  P  TILE SPLIT MINB MINBA NSP NSA AXS   XREG
#define SG_STAGE_KERNELS_ONLY
#include "sg_variants.h"

void sg_variants_2d_low(std::vector<Variant>& v) {
  v.push_back(make_variant<2, 1, 128, 1, 4, 2, 2, 2, true, true>());
  v.push_back(make_variant<2, 1, 64, 1, 8, 4, 2, 2, true, true>());
  v.push_back(make_variant<2, 2, 128, 1, 4, 2, 2, 2, true, true>());
  v.push_back(make_variant<2, 2, 64, 1, 8, 3, 2, 2, true, true>());
  v.push_back(make_variant<2, 2, 256, 1, 2, 1, 2, 2, true, true>());
  v.push_back(make_variant<2, 2, 32, 1, 16, 6, 2, 2, true, true>());
}
