// Stage-kernel instantiations of one group of elements (see sg_variants.h).  First entry of each (dim, degree) is the
// default; the others are tuning candidates selectable with SG_TILE / SG_SPLIT / SG_MINB / SG_NS (scripts/perf_probe.py).
//                D  P  TILE SPLIT MINB MINBA NSP NSA AXS   XREG
#define SG_STAGE_KERNELS_ONLY
#include "sg_variants.h"

void sg_variants_2d_low(std::vector<Variant>& v) {
  v.push_back(make_variant<2, 1, 128, 1, 4, 2, 2, 2, true, true>());
  v.push_back(make_variant<2, 1, 64, 1, 8, 4, 2, 2, true, true>());
  // 2D P2, TILE 128: the AXPY passes with a single-stage pipeline and 4 CTAs per SM (operands staged through shared
  // memory): K3 105.6 us, K6 121.1 us = 0.99 of the copy peak on the bytes it moves; two stages and 2 CTAs per SM gave
  // 115 / 127 us, operands straight from L2 106 / 135 us   (profiles/r02_tune_axpy_operands.log)
  v.push_back(make_variant<2, 2, 128, 1, 4, 4, 2, 1, true, true>());
  v.push_back(make_variant<2, 2, 128, 1, 4, 2, 2, 2, true, true>());
  v.push_back(make_variant<2, 2, 128, 1, 4, 3, 2, 2, false, true>());
  v.push_back(make_variant<2, 2, 64, 1, 8, 3, 2, 2, true, true>());
  v.push_back(make_variant<2, 2, 256, 1, 2, 1, 2, 2, true, true>());
  v.push_back(make_variant<2, 2, 32, 1, 16, 6, 2, 2, true, true>());
}
