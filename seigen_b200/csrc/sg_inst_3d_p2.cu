// Stage-kernel instantiations of one group of elements (see sg_variants.h).  First entry of each (dim, degree) is the
// default; the others are tuning candidates selectable with SG_TILE / SG_SPLIT / SG_MINB / SG_NS (scripts/perf_probe.py).
//                D  P  TILE SPLIT MINB MINBA NSP NSA AXS   XREG
#define SG_STAGE_KERNELS_ONLY
#include "sg_variants.h"

void sg_variants_3d_p2(std::vector<Variant>& v) {
  {
    // 3D P2, TILE 64: one tensor row per thread in every pass, AXPY operands straight from L2.  (With tiles aligned
    // to the mesh lattice K1/K5 take 107 us this way and 125 us with one thread per cell and all rows together --
    // the order was the other way round, 124 against 110 us, on the unaligned Hilbert tiles of the first sweeps;
    // profiles/r02_tune_axpy_operands.log.)
    const Variant rows = make_variant<3, 2, 64, 3, 2, 2, 2, 2, false, false>();
    const Variant one = make_variant<3, 2, 64, 1, 4, 2, 2, 2, false, false>();
    v.push_back(rows);
    v.push_back(compose_variant(one, rows, rows, rows, 13));
    v.push_back(one);
  }
  v.push_back(make_variant<3, 2, 32, 3, 3, 3, 2, 2, true, false>());
}
