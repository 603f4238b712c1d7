// TEST INFRASTRUCTURE ONLY.
// extern "C" driver around the REFERENCE's own reduced-statistics kernels
// (src/kernels/reduced_stats.hpp), compiled in place from $(REF)/src against the serial
// mini-Kokkos in ref_shim/ (no reference source is copied). Minkowski 1D/2D/3D, and the 2D
// curvilinear SRPIC metrics (spherical, qspherical).
// The functors are called cell by cell / particle by particle in serial order with the
// reference's own accumulator type (real_t), i.e. what Kokkos::parallel_reduce does on the
// Serial backend (src/framework/domain/metadomain_stats.cpp:88-183).
#include "oracle.h"

#include "enums.h"
#include "global.h"

#include "arch/kokkos_aliases.h"
#include "utils/numeric.h"

#include "metrics/minkowski.h"
#include "metrics/qspherical.h"
#include "metrics/spherical.h"

#include "framework/containers/particles.h"
#include "kernels/reduced_stats.hpp"

#include <map>
#include <new>
#include <string>
#include <stdexcept>
#include <vector>

using namespace ntt;

namespace {
  template <Dimension D>
  auto make_metric(const orc_grid_t* g, float dx) -> metric::Minkowski<D> {
    std::vector<ncells_t> res;
    boundaries_t<real_t>  ext;
    for (int a = 0; a < (int)D; ++a) {
      res.push_back((ncells_t)g->n[a]);
      ext.push_back({ ZERO, dx * (real_t)g->n[a] });
    }
    return metric::Minkowski<D>(res, ext);
  }

  template <Dimension D, unsigned short N>
  auto wrap(const orc_grid_t* g, float* p) -> ndfield_t<D, N> {
    const std::size_t G2 = 2 * (std::size_t)g->ng;
    if constexpr (D == Dim::_1D) {
      return ndfield_t<D, N>(p, g->n[0] + G2);
    } else if constexpr (D == Dim::_2D) {
      return ndfield_t<D, N>(p, g->n[0] + G2, g->n[1] + G2);
    } else {
      return ndfield_t<D, N>(p, g->n[0] + G2, g->n[1] + G2, g->n[2] + G2);
    }
  }

  template <class M, StatsID::type F, unsigned I>
  float fields_metric(const M& metric, const orc_grid_t* g, float* em, float* cur) {
    constexpr Dimension D = M::Dim;
    kernel::ReducedFields_kernel<SimEngine::SRPIC, M, F, I> k(wrap<D, 6>(g, em), wrap<D, 3>(g, cur), metric);
    real_t          buff = ZERO;
    const ncells_t  G    = (ncells_t)g->ng;
    if constexpr (D == Dim::_1D) {
      for (ncells_t i = G; i < g->n[0] + G; ++i) k(i, buff);
    } else if constexpr (D == Dim::_2D) {
      // Kokkos LayoutRight host iteration order of an MDRangePolicy: last index fastest
      for (ncells_t i = G; i < g->n[0] + G; ++i)
        for (ncells_t j = G; j < g->n[1] + G; ++j) k(i, j, buff);
    } else {
      for (ncells_t i = G; i < g->n[0] + G; ++i)
        for (ncells_t j = G; j < g->n[1] + G; ++j)
          for (ncells_t l = G; l < g->n[2] + G; ++l) k(i, j, l, buff);
    }
    return buff;
  }

  template <Dimension D, StatsID::type F, unsigned I>
  float fields_one(const orc_grid_t* g, float* em, float* cur, float dx) {
    return fields_metric<metric::Minkowski<D>, F, I>(make_metric<D>(g, dx), g, em, cur);
  }

  template <class M, StatsID::type F>
  float fields_comp_metric(const M& m, const orc_grid_t* g, float* em, float* cur, int comp) {
    switch (comp) {
      case 1: return fields_metric<M, F, 1>(m, g, em, cur);
      case 2: return fields_metric<M, F, 2>(m, g, em, cur);
      case 3: return fields_metric<M, F, 3>(m, g, em, cur);
      default: throw std::runtime_error("ref stats: component must be 1..3");
    }
  }

  template <class M>
  float fields_any_metric(const M& m, const orc_grid_t* g, float* em, float* cur, int what, int comp) {
    switch (what) {
      case 0: return fields_comp_metric<M, StatsID::B2>(m, g, em, cur, comp);
      case 1: return fields_comp_metric<M, StatsID::E2>(m, g, em, cur, comp);
      case 2: return fields_comp_metric<M, StatsID::ExB>(m, g, em, cur, comp);
      case 3: return fields_metric<M, StatsID::JdotE, 0>(m, g, em, cur);
      default: throw std::runtime_error("ref stats: unknown field statistic");
    }
  }

  template <Dimension D, StatsID::type F>
  float fields_comp(const orc_grid_t* g, float* em, float* cur, float dx, int comp) {
    switch (comp) {
      case 1: return fields_one<D, F, 1>(g, em, cur, dx);
      case 2: return fields_one<D, F, 2>(g, em, cur, dx);
      case 3: return fields_one<D, F, 3>(g, em, cur, dx);
      default: throw std::runtime_error("ref stats: component must be 1..3");
    }
  }

  template <Dimension D>
  float fields_dim(const orc_grid_t* g, float* em, float* cur, float dx, int what, int comp) {
    switch (what) {
      case 0: return fields_comp<D, StatsID::B2>(g, em, cur, dx, comp);
      case 1: return fields_comp<D, StatsID::E2>(g, em, cur, dx, comp);
      case 2: return fields_comp<D, StatsID::ExB>(g, em, cur, dx, comp);
      case 3: return fields_one<D, StatsID::JdotE, 0>(g, em, cur, dx);
      default: throw std::runtime_error("ref stats: unknown field statistic");
    }
  }
  template <class M, StatsID::type P>
  float moments_metric(const M& metric, const orc_prtls_t* p, uint32_t n, float mass, float charge,
                       int use_weights, int c1, int c2) {
    constexpr Dimension              D = M::Dim;
    Particles<D, M::CoordType>       prtls;
    // the kernel reads mass() / charge() of the ParticleSpecies base, whose members are const and
    // whose allocating constructor lives in particles.cpp (not built here): re-construct the base
    // subobject of the empty container in place
    ParticleSpecies* base = static_cast<ParticleSpecies*>(&prtls);
    base->~ParticleSpecies();
    new (base) ParticleSpecies(1u, "s", mass, charge, (npart_t)n, 0u, 0u, ParticlePusher::BORIS, false,
                               RadiativeDrag::NONE, EmissionType::NONE, 0, 0);
    ParticleArrays& a = prtls;
    a.i1     = array_t<int*>(p->i1, n);
    a.i2     = array_t<int*>(p->i2, n);
    a.i3     = array_t<int*>(p->i3, n);
    a.dx1    = array_t<prtldx_t*>(p->dx1, n);
    a.dx2    = array_t<prtldx_t*>(p->dx2, n);
    a.dx3    = array_t<prtldx_t*>(p->dx3, n);
    a.ux1    = array_t<real_t*>(p->ux1, n);
    a.ux2    = array_t<real_t*>(p->ux2, n);
    a.ux3    = array_t<real_t*>(p->ux3, n);
    a.weight = array_t<real_t*>(p->weight, n);
    a.tag    = array_t<short*>(p->tag, n);
    if constexpr (M::CoordType != Coord::Cartesian) a.phi = array_t<real_t*>(p->phi, n);
    std::vector<uint8_t> comps;
    if (P == StatsID::T) comps = { (uint8_t)c1, (uint8_t)c2 };
    kernel::ReducedParticleMoments_kernel<SimEngine::SRPIC, M, P> k(comps, prtls, use_weights != 0, metric);
    real_t buff = ZERO;
    for (npart_t q = 0; q < n; ++q) k(q, buff);
    return buff;
  }

  template <Dimension D, StatsID::type P>
  float moments_one(const orc_grid_t* g, const orc_prtls_t* p, uint32_t n, float mass, float charge,
                    int use_weights, float dx, int c1, int c2) {
    return moments_metric<metric::Minkowski<D>, P>(make_metric<D>(g, dx), p, n, mass, charge, use_weights, c1, c2);
  }

  template <class M>
  float moments_any_metric(const M& m, const orc_prtls_t* p, uint32_t n, float mass, float charge,
                           int use_weights, int what, int c1, int c2) {
    switch (what) {
      case 0: return moments_metric<M, StatsID::Npart>(m, p, n, mass, charge, use_weights, c1, c2);
      case 1: return moments_metric<M, StatsID::N>(m, p, n, mass, charge, use_weights, c1, c2);
      case 2: return moments_metric<M, StatsID::Rho>(m, p, n, mass, charge, use_weights, c1, c2);
      case 3: return moments_metric<M, StatsID::Charge>(m, p, n, mass, charge, use_weights, c1, c2);
      case 4: return moments_metric<M, StatsID::T>(m, p, n, mass, charge, use_weights, c1, c2);
      default: throw std::runtime_error("ref stats: unknown particle statistic");
    }
  }

  // 2D curvilinear SR metrics: kind 1 = spherical, 2 = qspherical; ext = x1min, x1max, x2min, x2max, r0, h
  template <class M>
  M make_curv(const orc_grid_t* g, const float* ext) {
    std::vector<ncells_t> res { (ncells_t)g->n[0], (ncells_t)g->n[1] };
    boundaries_t<real_t>  e { { ext[0], ext[1] }, { ext[2], ext[3] } };
    std::map<std::string, real_t> prm { { "r0", ext[4] }, { "h", ext[5] }, { "a", ZERO } };
    return M(res, e, prm);
  }

  template <Dimension D>
  float moments_dim(const orc_grid_t* g, const orc_prtls_t* p, uint32_t n, float mass, float charge,
                    int use_weights, float dx, int what, int c1, int c2) {
    switch (what) {
      case 0: return moments_one<D, StatsID::Npart>(g, p, n, mass, charge, use_weights, dx, c1, c2);
      case 1: return moments_one<D, StatsID::N>(g, p, n, mass, charge, use_weights, dx, c1, c2);
      case 2: return moments_one<D, StatsID::Rho>(g, p, n, mass, charge, use_weights, dx, c1, c2);
      case 3: return moments_one<D, StatsID::Charge>(g, p, n, mass, charge, use_weights, dx, c1, c2);
      case 4: return moments_one<D, StatsID::T>(g, p, n, mass, charge, use_weights, dx, c1, c2);
      default: throw std::runtime_error("ref stats: unknown particle statistic");
    }
  }
} // namespace

extern "C" {
float ref_stats_particles(const orc_grid_t* g, const orc_prtls_t* p, uint32_t n, float mass,
                          float charge, int use_weights, float dx, int what, int c1, int c2) {
  switch (g->dim) {
    case 1: return moments_dim<Dim::_1D>(g, p, n, mass, charge, use_weights, dx, what, c1, c2);
    case 2: return moments_dim<Dim::_2D>(g, p, n, mass, charge, use_weights, dx, what, c1, c2);
    default: return moments_dim<Dim::_3D>(g, p, n, mass, charge, use_weights, dx, what, c1, c2);
  }
}

float ref_stats_fields_curv(int kind, const orc_grid_t* g, const float* ext, float* em, float* cur, int what,
                            int comp) {
  if (g->dim != 2) throw std::runtime_error("ref stats: curvilinear metrics are 2D");
  if (kind == 1) return fields_any_metric(make_curv<metric::Spherical<Dim::_2D>>(g, ext), g, em, cur, what, comp);
  return fields_any_metric(make_curv<metric::QSpherical<Dim::_2D>>(g, ext), g, em, cur, what, comp);
}

float ref_stats_particles_curv(int kind, const orc_grid_t* g, const float* ext, const orc_prtls_t* p, uint32_t n,
                               float mass, float charge, int use_weights, int what, int c1, int c2) {
  if (g->dim != 2) throw std::runtime_error("ref stats: curvilinear metrics are 2D");
  if (kind == 1) {
    return moments_any_metric(make_curv<metric::Spherical<Dim::_2D>>(g, ext), p, n, mass, charge, use_weights,
                              what, c1, c2);
  }
  return moments_any_metric(make_curv<metric::QSpherical<Dim::_2D>>(g, ext), p, n, mass, charge, use_weights, what,
                            c1, c2);
}

float ref_stats_fields(const orc_grid_t* g, float* em, float* cur, float dx, int what, int comp) {
  switch (g->dim) {
    case 1: return fields_dim<Dim::_1D>(g, em, cur, dx, what, comp);
    case 2: return fields_dim<Dim::_2D>(g, em, cur, dx, what, comp);
    default: return fields_dim<Dim::_3D>(g, em, cur, dx, what, comp);
  }
}
}
