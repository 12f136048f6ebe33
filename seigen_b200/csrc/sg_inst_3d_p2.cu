// Stage-kernel instantiations of one group of elements (see sg_variants.h).  First entry of each (dim, degree) is the
// default; the others are tuning candidates selectable with SG_TILE / SG_SPLIT / SG_MINB / SG_NS (scripts/perf_probe.py).
//                D  P  TILE SPLIT MINB MINBA NSP NSA AXS   XREG
#define SG_STAGE_KERNELS_ONLY
#include "sg_variants.h"

void sg_variants_3d_p2(std::vector<Variant>& v) {
  {
    // 3D P2, TILE 64: F-type plain pass with one thread per cell (all rows together: each stress component of a
    // neighbour gathered once), everything else with one tensor row per thread; AXPY operands straight from L2
    const Variant rows = make_variant<3, 2, 64, 3, 2, 2, 2, 2, false, false>();
    const Variant one = make_variant<3, 2, 64, 1, 4, 2, 2, 2, false, false>();
    v.push_back(compose_variant(one, rows, rows, rows, 13));
    v.push_back(rows);
    v.push_back(one);
  }
  v.push_back(make_variant<3, 2, 32, 3, 3, 3, 2, 2, true, false>());
}
