// Kernel variants: one entry per (dim, degree, tile, split, ...) combination with its four stage kernels and the
// shared-memory plans they were compiled for.  Included by sg_api.cu (the table) and by the sg_inst_*.cu translation
// units, which instantiate the kernels of one group of elements each so that the library builds in parallel.
#pragma once
#include <vector>

#include "sg_kernels.cuh"

struct Variant {
  int dim, degree, nd, nfp, tile, split, minb, minba, ns_plain, ns_axpy, axs, xreg;
  // [0]: full stress storage (D*D components), [1]: symmetric storage (upper triangle)
  sg::StagePlan (*plan_f[2])(bool classes, bool mat, bool sponge);
  sg::StagePlan (*plan_f_axpy[2])(bool classes, bool mat, bool sponge);
  sg::StagePlan (*plan_g[2])(bool classes, bool mat, bool sponge);
  sg::StagePlan (*plan_g_axpy[2])(bool classes, bool mat, bool sponge);
  const void* f_plain[2];
  const void* f_axpy[2];
  const void* g_plain[2];
  const void* g_axpy[2];
  // threads per CTA of f_plain, f_axpy, g_plain, g_axpy (TILE * that kernel's SPLIT): the four kernels of a variant
  // share the tile size (= the field layout) and nothing else, so a variant may take each kernel from a different
  // instantiation (compose_variant)
  int threads[4];
};

// TILE cells per tile, SPLIT threads per cell, MINB / MINBA CTAs per SM the compiler must allow for the plain / AXPY
// kernels (register cap), NSP / NSA pipeline depth of the plain / AXPY kernels, AXS: stage the AXPY operands through
// shared memory (bulk copies) instead of reading them from L2, XREG: G-type gradients in registers (SPLIT == 1)
template <int D, int P, int TILE, int SPLIT, int MINB, int MINBA, int NSP, int NSA, bool AXS, bool XREG, bool SYM>
inline void fill_variant(Variant& v) {
  using E = ElemOps<D, P>;
  constexpr int m = SYM ? 1 : 0;
  v.plan_f[m] = [](bool c, bool mt, bool sp) { return sg::make_plan<D, E::ND, TILE, NSP, false, false, false, SYM>(c, mt, sp, E::FTAB_SIZE); };
  v.plan_f_axpy[m] = [](bool c, bool mt, bool sp) { return sg::make_plan<D, E::ND, TILE, NSA, false, AXS, false, SYM>(c, mt, sp, E::FTAB_SIZE); };
  v.plan_g[m] = [](bool c, bool mt, bool sp) { return sg::make_plan<D, E::ND, TILE, NSP, true, false, !XREG, SYM>(c, mt, sp, E::FTAB_SIZE); };
  v.plan_g_axpy[m] = [](bool c, bool mt, bool sp) { return sg::make_plan<D, E::ND, TILE, NSA, true, AXS, !XREG, SYM>(c, mt, sp, E::FTAB_SIZE); };
  v.f_plain[m] = (const void*)&sg::stage_f_kernel<D, P, TILE, SPLIT, MINB, NSP, false, false, SYM>;
  v.f_axpy[m] = (const void*)&sg::stage_f_kernel<D, P, TILE, SPLIT, MINBA, NSA, true, AXS, SYM>;
  v.g_plain[m] = (const void*)&sg::stage_g_kernel<D, P, TILE, SPLIT, MINB, NSP, false, false, XREG, SYM>;
  v.g_axpy[m] = (const void*)&sg::stage_g_kernel<D, P, TILE, SPLIT, MINBA, NSA, true, AXS, XREG, SYM>;
}

template <int D, int P, int TILE, int SPLIT, int MINB, int MINBA, int NSP, int NSA, bool AXS, bool XREG>
inline Variant make_variant() {
  using E = ElemOps<D, P>;
  Variant v;
  v.dim = D;
  v.degree = P;
  v.nd = E::ND;
  v.nfp = E::NFP;
  v.tile = TILE;
  v.split = SPLIT;
  v.minb = MINB;
  v.minba = MINBA;
  v.ns_plain = NSP;
  v.ns_axpy = NSA;
  v.axs = AXS;
  v.xreg = XREG;
  for (int k = 0; k < 4; ++k) v.threads[k] = TILE * SPLIT;
  fill_variant<D, P, TILE, SPLIT, MINB, MINBA, NSP, NSA, AXS, XREG, false>(v);
  fill_variant<D, P, TILE, SPLIT, MINB, MINBA, NSP, NSA, AXS, XREG, true>(v);
  return v;
}


// A variant whose four kernels come from four instantiations with the same tile size (chosen per pass by
// scripts/tune_stages.py).  `id` distinguishes it from its sources in SG_SPLIT-style selection: it reports split = id.
inline Variant compose_variant(const Variant& f_plain, const Variant& f_axpy, const Variant& g_plain,
                               const Variant& g_axpy, int id) {
  Variant v = f_plain;
  for (int m = 0; m < 2; ++m) {
    v.plan_f_axpy[m] = f_axpy.plan_f_axpy[m];
    v.f_axpy[m] = f_axpy.f_axpy[m];
    v.plan_g[m] = g_plain.plan_g[m];
    v.g_plain[m] = g_plain.g_plain[m];
    v.plan_g_axpy[m] = g_axpy.plan_g_axpy[m];
    v.g_axpy[m] = g_axpy.g_axpy[m];
  }
  v.threads[1] = f_axpy.threads[1];
  v.threads[2] = g_plain.threads[2];
  v.threads[3] = g_axpy.threads[3];
  v.split = id;
  return v;
}

// defined in sg_inst_*.cu; each appends its variants (the first entry of a (dim, degree) is the default)
void sg_variants_1d(std::vector<Variant>& v);        // 1D P1-P3 (tests/pulse/pulse_1d_lf4.py)
void sg_variants_2d_low(std::vector<Variant>& v);    // 2D P1, P2
void sg_variants_2d_high(std::vector<Variant>& v);   // 2D P3, P4
void sg_variants_3d_p1(std::vector<Variant>& v);
void sg_variants_3d_p2(std::vector<Variant>& v);
void sg_variants_3d_p3(std::vector<Variant>& v);
