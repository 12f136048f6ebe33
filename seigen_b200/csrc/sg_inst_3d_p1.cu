// Stage-kernel instantiations of one group of elements (see sg_variants.h).  First entry of each (dim, degree) is the
// default; the others are tuning candidates selectable with SG_TILE / SG_SPLIT / SG_MINB / SG_NS (scripts/perf_probe.py).
//                D  P  TILE SPLIT MINB MINBA NSP NSA AXS   XREG
#define SG_STAGE_KERNELS_ONLY
#include "sg_variants.h"

void sg_variants_3d_p1(std::vector<Variant>& v) {
  {
    // 3D P1, TILE 128: K3 with a two-stage pipeline and 2 CTAs per SM, K6 with a single stage and 3 CTAs per SM
    const Variant a = make_variant<3, 1, 128, 1, 2, 2, 2, 2, true, true>();
    const Variant b = make_variant<3, 1, 128, 1, 2, 3, 2, 1, true, true>();
    v.push_back(compose_variant(a, a, a, b, 11));
    v.push_back(a);
    v.push_back(b);
  }
  v.push_back(make_variant<3, 1, 64, 1, 4, 4, 2, 2, true, true>());
  v.push_back(make_variant<3, 1, 32, 1, 8, 4, 2, 2, true, true>());
}
