// Stage-kernel instantiations of one group of elements (see sg_variants.h).  First entry of each (dim, degree) is the
// default; the others are tuning candidates selectable with SG_TILE / SG_SPLIT / SG_MINB / SG_NS (scripts/perf_probe.py).
//                D  P  TILE SPLIT MINB MINBA NSP NSA AXS   XREG
#define SG_STAGE_KERNELS_ONLY
#include "sg_variants.h"

void sg_variants_2d_high(std::vector<Variant>& v) {
  // 2D P3, TILE 64: AXPY passes single-stage with 3 CTAs per SM, operands staged through shared memory (K3 94.9 us,
  // K6 111.1 us; two stages + operands straight from L2: 98.8 / 123.2 us; profiles/r02_tune_axpy_operands.log)
  v.push_back(make_variant<2, 3, 64, 1, 4, 3, 2, 1, true, true>());
  v.push_back(make_variant<2, 3, 64, 1, 4, 2, 2, 2, false, true>());
  v.push_back(make_variant<2, 3, 64, 1, 4, 2, 2, 2, true, true>());
  v.push_back(make_variant<2, 3, 32, 1, 8, 4, 2, 2, true, true>());
  {
    // 2D P4, TILE 32: F-type plain and G-type plain passes with one thread per cell (gradients in registers), the
    // AXPY passes with one tensor row per thread (K3 staged through shared memory, K6 straight from L2)
    const Variant one = make_variant<2, 4, 32, 1, 4, 3, 2, 2, false, true>();
    const Variant rows_staged = make_variant<2, 4, 32, 2, 3, 3, 2, 2, true, false>();
    const Variant rows_direct = make_variant<2, 4, 32, 2, 3, 3, 2, 2, false, false>();
    v.push_back(compose_variant(one, rows_staged, one, rows_direct, 12));
    v.push_back(rows_direct);
    v.push_back(one);
    v.push_back(rows_staged);
  }
}
