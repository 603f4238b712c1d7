// TEST INFRASTRUCTURE ONLY.
// extern "C" driver around the REFERENCE's own matching-boundary kernel
// (kernel::bc::MatchBoundaries_kernel, src/kernels/fields_bcs.hpp:42-560), compiled in place
// from $(REF)/src against the serial mini-Kokkos in ref_shim/ (no reference source is copied).
// SRPIC, Minkowski 1D/2D/3D. The pgen side (MatchFields functor) is a polynomial field setter
// defined HERE (ours), whose fp32 expressions tests/bcs_cases.py repeats term by term.
#include "oracle.h"

#include "enums.h"
#include "global.h"

#include "arch/kokkos_aliases.h"
#include "utils/numeric.h"

#include "metrics/minkowski.h"

#include "kernels/fields_bcs.hpp"

#include <stdexcept>
#include <vector>

using namespace ntt;

namespace {
  // value = a0 + a1 * x[0] + a2 * x[1] + a3 * x[2] (left to right, fp32), one row per component
  // in the order ex1, ex2, ex3, bx1, bx2, bx3
  template <Dimension D>
  struct PolySetter {
    float a[6][4];

    Inline auto eval(int c, const coord_t<D>& x) const -> real_t {
      real_t v = a[c][0];
      for (int d = 0; d < (int)D; ++d) v = v + a[c][1 + d] * x[d];
      return v;
    }
    Inline auto ex1(const coord_t<D>& x) const -> real_t { return eval(0, x); }
    Inline auto ex2(const coord_t<D>& x) const -> real_t { return eval(1, x); }
    Inline auto ex3(const coord_t<D>& x) const -> real_t { return eval(2, x); }
    Inline auto bx1(const coord_t<D>& x) const -> real_t { return eval(3, x); }
    Inline auto bx2(const coord_t<D>& x) const -> real_t { return eval(4, x); }
    Inline auto bx3(const coord_t<D>& x) const -> real_t { return eval(5, x); }
  };

  // the same polynomial, magnetic components only (the kernel skips what a setter lacks)
  template <Dimension D>
  struct PolySetterB {
    float a[6][4];

    Inline auto eval(int c, const coord_t<D>& x) const -> real_t {
      real_t v = a[c][0];
      for (int d = 0; d < (int)D; ++d) v = v + a[c][1 + d] * x[d];
      return v;
    }
    Inline auto bx1(const coord_t<D>& x) const -> real_t { return eval(3, x); }
    Inline auto bx2(const coord_t<D>& x) const -> real_t { return eval(4, x); }
    Inline auto bx3(const coord_t<D>& x) const -> real_t { return eval(5, x); }
  };

  template <Dimension D>
  auto make_metric(const orc_grid_t* g, float dx, const float* xmin) -> metric::Minkowski<D> {
    std::vector<ncells_t> res;
    boundaries_t<real_t>  ext;
    for (int a = 0; a < (int)D; ++a) {
      res.push_back((ncells_t)g->n[a]);
      ext.push_back({ xmin[a], xmin[a] + dx * (real_t)g->n[a] });
    }
    return metric::Minkowski<D>(res, ext);
  }

  template <Dimension D>
  auto wrap6(const orc_grid_t* g, float* p) -> ndfield_t<D, 6> {
    const std::size_t G2 = 2 * (std::size_t)g->ng;
    if constexpr (D == Dim::_1D) {
      return ndfield_t<D, 6>(p, g->n[0] + G2);
    } else if constexpr (D == Dim::_2D) {
      return ndfield_t<D, 6>(p, g->n[0] + G2, g->n[1] + G2);
    } else {
      return ndfield_t<D, 6>(p, g->n[0] + G2, g->n[1] + G2, g->n[2] + G2);
    }
  }

  template <Dimension D, class FS, in O>
  void match_run(const orc_grid_t* g, float* em, const FS& fs, float dx, const float* xmin,
                 float xg_edge, float ds, int tags, const int* rmin, const int* rmax) {
    using M = metric::Minkowski<D>;
    auto                 fld    = wrap6<D>(g, em);
    const auto           metric = make_metric<D>(g, dx, xmin);
    boundaries_t<FldsBC> bnd;
    for (int a = 0; a < (int)D; ++a) bnd.push_back({ FldsBC::MATCH, FldsBC::MATCH });
    kernel::bc::MatchBoundaries_kernel<SimEngine::SRPIC, M, FS, O> k(fld, fs, metric, xg_edge, ds,
                                                                     (BCTags)tags, bnd);
    if constexpr (D == Dim::_1D) {
      for (ncells_t i = rmin[0]; i < (ncells_t)rmax[0]; ++i) k(i);
    } else if constexpr (D == Dim::_2D) {
      for (ncells_t i = rmin[0]; i < (ncells_t)rmax[0]; ++i)
        for (ncells_t j = rmin[1]; j < (ncells_t)rmax[1]; ++j) k(i, j);
    } else {
      for (ncells_t i = rmin[0]; i < (ncells_t)rmax[0]; ++i)
        for (ncells_t j = rmin[1]; j < (ncells_t)rmax[1]; ++j)
          for (ncells_t l = rmin[2]; l < (ncells_t)rmax[2]; ++l) k(i, j, l);
    }
  }

  template <Dimension D, class FS>
  void match_dir(const orc_grid_t* g, float* em, const FS& fs, float dx, const float* xmin, int o,
                 float xg_edge, float ds, int tags, const int* rmin, const int* rmax) {
    if (o == 0) {
      match_run<D, FS, in::x1>(g, em, fs, dx, xmin, xg_edge, ds, tags, rmin, rmax);
    } else if constexpr (D != Dim::_1D) {
      if (o == 1) {
        match_run<D, FS, in::x2>(g, em, fs, dx, xmin, xg_edge, ds, tags, rmin, rmax);
      } else if constexpr (D == Dim::_3D) {
        match_run<D, FS, in::x3>(g, em, fs, dx, xmin, xg_edge, ds, tags, rmin, rmax);
      }
    }
  }

  template <Dimension D>
  void match_dim(const orc_grid_t* g, float* em, const float* coef, int b_only, float dx,
                 const float* xmin, int o, float xg_edge, float ds, int tags, const int* rmin,
                 const int* rmax) {
    if (o < 0 || o >= (int)D) throw std::runtime_error("ref match: bad direction");
    if (b_only) {
      PolySetterB<D> fs;
      for (int c = 0; c < 6; ++c)
        for (int q = 0; q < 4; ++q) fs.a[c][q] = coef[4 * c + q];
      match_dir<D>(g, em, fs, dx, xmin, o, xg_edge, ds, tags, rmin, rmax);
    } else {
      PolySetter<D> fs;
      for (int c = 0; c < 6; ++c)
        for (int q = 0; q < 4; ++q) fs.a[c][q] = coef[4 * c + q];
      match_dir<D>(g, em, fs, dx, xmin, o, xg_edge, ds, tags, rmin, rmax);
    }
  }
} // namespace

namespace {
  // kernel::bc::ConductorBoundaries_kernel<Dim::_2D, o, P> over the range
  // srpic::PerfectConductorFieldsIn builds (src/engines/srpic/fields_bcs.h:384-470)
  template <in O, bool P>
  void conductor_run(const orc_grid_t* g, float* em, int tags) {
    auto           fld = wrap6<Dim::_2D>(g, em);
    const int      o   = (O == in::x1) ? 0 : 1;
    const ncells_t G   = (ncells_t)g->ng;
    const ncells_t i_edge = P ? (G + (ncells_t)g->n[o]) : G; // i_max / i_min
    kernel::bc::ConductorBoundaries_kernel<Dim::_2D, O, P> k(fld, i_edge, (BCTags)tags);
    ncells_t lo[2] = { 0, 0 }, hi[2] = { (ncells_t)g->n[0] + 2 * G, (ncells_t)g->n[1] + 2 * G };
    hi[o] = P ? G : G + 1;
    for (ncells_t i = lo[0]; i < hi[0]; ++i)
      for (ncells_t j = lo[1]; j < hi[1]; ++j) k(i, j);
  }

  // kernel::bc::AxisBoundaries_kernel<Dim::_2D, P> over n_all(x1) (fields_bcs.h:205-232)
  template <bool P>
  void axis_run(const orc_grid_t* g, float* em, int tags) {
    auto           fld = wrap6<Dim::_2D>(g, em);
    const ncells_t G   = (ncells_t)g->ng;
    const ncells_t i_edge = P ? (G + (ncells_t)g->n[1]) : G;
    kernel::bc::AxisBoundaries_kernel<Dim::_2D, P> k(fld, i_edge, (BCTags)tags);
    for (ncells_t i = 0; i < (ncells_t)g->n[0] + 2 * G; ++i) k(i);
  }
} // namespace

extern "C" {
void ref_conductor_fields(const orc_grid_t* g, float* em, int o, int sign, int tags) {
  if (g->dim != 2) throw std::runtime_error("ref conductor: 2D only");
  if (o == 0) {
    if (sign > 0) conductor_run<in::x1, true>(g, em, tags); else conductor_run<in::x1, false>(g, em, tags);
  } else {
    if (sign > 0) conductor_run<in::x2, true>(g, em, tags); else conductor_run<in::x2, false>(g, em, tags);
  }
}
void ref_axis_fields(const orc_grid_t* g, float* em, int sign, int tags) {
  if (sign > 0) axis_run<true>(g, em, tags); else axis_run<false>(g, em, tags);
}
int ref_bc_tag_e() { return (int)BC::E; }
int ref_bc_tag_b() { return (int)BC::B; }
void ref_match_fields(const orc_grid_t* g, float* em, const float* coef, int b_only, float dx,
                      const float* xmin, int o, float xg_edge, float ds, int tags, const int* rmin,
                      const int* rmax) {
  switch (g->dim) {
    case 1: match_dim<Dim::_1D>(g, em, coef, b_only, dx, xmin, o, xg_edge, ds, tags, rmin, rmax); break;
    case 2: match_dim<Dim::_2D>(g, em, coef, b_only, dx, xmin, o, xg_edge, ds, tags, rmin, rmax); break;
    default: match_dim<Dim::_3D>(g, em, coef, b_only, dx, xmin, o, xg_edge, ds, tags, rmin, rmax); break;
  }
}
}
