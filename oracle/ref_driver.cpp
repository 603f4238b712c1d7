// TEST INFRASTRUCTURE ONLY.
// Thin extern "C" driver around the REFERENCE's own kernel headers, compiled in
// place from $(REF)/src against the serial mini-Kokkos in ref_shim/ (no reference
// source is copied into this repo). One library per compile-time SHAPE_ORDER:
// oracle/_ref/libref_o<O>.so. Signatures mirror oracle.h (orc_* -> ref_*), so the
// tests can run oracle and reference side by side on identical buffers.
#include "oracle.h"

#include "enums.h"
#include "global.h"

#include "arch/kokkos_aliases.h"
#include "utils/numeric.h"

#include "metrics/minkowski.h"

#include "kernels/ampere_mink.hpp"
#include "kernels/currents_deposit.hpp"
#include "kernels/digital_filter.hpp"
#include "kernels/faraday_mink.hpp"
#include "kernels/pushers/sr.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

using namespace ntt;

namespace {

  template <Dimension D>
  auto make_metric(const orc_grid_t* g, float dx, const float* xmin) -> metric::Minkowski<D> {
    std::vector<ncells_t> res;
    boundaries_t<real_t>  ext;
    for (int a = 0; a < (int)D; ++a) {
      res.push_back((ncells_t)g->n[a]);
      const real_t lo = xmin ? xmin[a] : ZERO;
      ext.push_back({ lo, lo + dx * (real_t)g->n[a] });
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

  void check_ng(const orc_grid_t* g) {
    if ((uint32_t)g->ng != N_GHOSTS) {
      throw std::runtime_error("ref: grid ng != compile-time N_GHOSTS");
    }
  }

  // number of worker threads for the *_mt entry points (1 = the serial reference order)
  int g_threads = 1;

  template <class F>
  void parallel_chunks(std::size_t n, const F& fn) {
    const int nt = std::max(1, std::min<int>(g_threads, (int)std::max<std::size_t>(n, 1)));
    if (nt == 1) {
      fn(0, n, 0);
      return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
      const std::size_t lo = n * t / nt, hi = n * (t + 1) / nt;
      th.emplace_back([=, &fn]() { fn(lo, hi, t); });
    }
    for (auto& x : th) x.join();
  }

  // active cells; the slowest active dimension is split across the worker threads (cells are
  // independent in every field kernel, like a Kokkos MDRangePolicy)
  template <class K, Dimension D>
  void loop_active(const orc_grid_t* g, const K& k) {
    const ncells_t G = N_GHOSTS;
    if constexpr (D == Dim::_1D) {
      parallel_chunks(g->n[0], [&](std::size_t lo, std::size_t hi, int) {
        for (ncells_t i = G + lo; i < G + hi; ++i) k(i);
      });
    } else if constexpr (D == Dim::_2D) {
      parallel_chunks(g->n[1], [&](std::size_t lo, std::size_t hi, int) {
        for (ncells_t j = G + lo; j < G + hi; ++j)
          for (ncells_t i = G; i < g->n[0] + G; ++i) k(i, j);
      });
    } else {
      parallel_chunks(g->n[2], [&](std::size_t lo, std::size_t hi, int) {
        for (ncells_t l = G + lo; l < G + hi; ++l)
          for (ncells_t j = G; j < g->n[1] + G; ++j)
            for (ncells_t i = G; i < g->n[0] + G; ++i) k(i, j, l);
      });
    }
  }

  PrtlBC to_pbc(int b) {
    switch (b) {
      case ORC_PBC_PERIODIC: return PrtlBC::PERIODIC;
      case ORC_PBC_ABSORB: return PrtlBC::ABSORB;
      case ORC_PBC_REFLECT: return PrtlBC::REFLECT;
      case ORC_PBC_AXIS: return PrtlBC::AXIS;
      default: return PrtlBC::SYNC;
    }
  }

  FldsBC to_fbc(int b) {
    switch (b) {
      case ORC_FBC_PERIODIC: return FldsBC::PERIODIC;
      case ORC_FBC_CONDUCTOR: return FldsBC::CONDUCTOR;
      case ORC_FBC_AXIS: return FldsBC::AXIS;
      default: return FldsBC::SYNC;
    }
  }

  template <Dimension D>
  void faraday(const orc_grid_t* g, float* em, float c1, float c2, const float* st) {
    auto EB = wrap<D, 6>(g, em);
    if (st) {
      loop_active<decltype(kernel::mink::Faraday_kernel<D>(EB, c1, c2)), D>(
        g,
        kernel::mink::Faraday_kernel<D>(EB, c1, c2, st[0], st[1], st[2], st[3], st[4], st[5],
                                        st[6], st[7], st[8]));
    } else {
      loop_active<decltype(kernel::mink::Faraday_kernel<D>(EB, c1, c2)), D>(
        g, kernel::mink::Faraday_kernel<D>(EB, c1, c2));
    }
  }

  template <Dimension D>
  void ampere(const orc_grid_t* g, float* em, float c1, float c2) {
    auto EB = wrap<D, 6>(g, em);
    loop_active<kernel::mink::Ampere_kernel<D>, D>(g, kernel::mink::Ampere_kernel<D>(EB, c1, c2));
  }

  template <Dimension D>
  void currents_ampere(const orc_grid_t* g, float* em, float* cur, float coeff, float ppc0) {
    auto E = wrap<D, 6>(g, em);
    auto J = wrap<D, 3>(g, cur);
    loop_active<kernel::mink::CurrentsAmpere_kernel<D>, D>(
      g, kernel::mink::CurrentsAmpere_kernel<D>(E, J, coeff, ppc0));
  }

  template <Dimension D>
  void filter_pass(const orc_grid_t* g, float* cur, const float* buff, const int* fbc) {
    auto                 A = wrap<D, 3>(g, cur);
    auto                 B = wrap<D, 3>(g, const_cast<float*>(buff));
    ncells_t             size[(int)D];
    boundaries_t<FldsBC> bnd;
    for (int a = 0; a < (int)D; ++a) {
      size[a] = (ncells_t)g->n[a];
      bnd.push_back({ to_fbc(fbc[2 * a]), to_fbc(fbc[2 * a + 1]) });
    }
    using K = kernel::DigitalFilter_kernel<D, Coord::Cartesian>;
    loop_active<K, D>(g, K(A, B, size, bnd));
  }

  void fill_arrays(ParticleArrays& a, const orc_prtls_t* p, uint32_t n) {
    a.i1       = array_t<int*>(p->i1, n);
    a.i2       = array_t<int*>(p->i2, n);
    a.i3       = array_t<int*>(p->i3, n);
    a.dx1      = array_t<prtldx_t*>(p->dx1, n);
    a.dx2      = array_t<prtldx_t*>(p->dx2, n);
    a.dx3      = array_t<prtldx_t*>(p->dx3, n);
    a.ux1      = array_t<real_t*>(p->ux1, n);
    a.ux2      = array_t<real_t*>(p->ux2, n);
    a.ux3      = array_t<real_t*>(p->ux3, n);
    a.weight   = array_t<real_t*>(p->weight, n);
    a.i1_prev  = array_t<int*>(p->i1_prev, n);
    a.i2_prev  = array_t<int*>(p->i2_prev, n);
    a.i3_prev  = array_t<int*>(p->i3_prev, n);
    a.dx1_prev = array_t<prtldx_t*>(p->dx1_prev, n);
    a.dx2_prev = array_t<prtldx_t*>(p->dx2_prev, n);
    a.dx3_prev = array_t<prtldx_t*>(p->dx3_prev, n);
    a.tag      = array_t<short*>(p->tag, n);
    a.phi      = array_t<real_t*>(p->phi, n);
  }

  template <Dimension D>
  void push(const orc_grid_t* g, const orc_pusher_t* c, const orc_prtls_t* p, uint32_t n,
            const float* em) {
    using M = metric::Minkowski<D>;
    auto                 metric = make_metric<D>(g, c->dx, c->xmin);
    boundaries_t<PrtlBC> bnd;
    for (int a = 0; a < (int)D; ++a) {
      bnd.push_back({ to_pbc(c->pbc[2 * a]), to_pbc(c->pbc[2 * a + 1]) });
    }
    kernel::sr::PusherContext ctx { (spidx_t)1,
                                    (ParticlePusherFlags)c->pusher_flags,
                                    (RadiativeDragFlags)c->drag_flags,
                                    c->mass,
                                    c->charge,
                                    c->time,
                                    c->dt,
                                    c->omegaB0,
                                    g->n[0],
                                    g->n[1],
                                    g->n[2] };
    ctx.gca.larmor_max         = c->gca_larmor_max;
    ctx.gca.e_ovr_b_sqr_max    = c->gca_e_ovr_b_sqr_max;
    ctx.synchrotron_drag.coeff = c->sync_coeff;
    ctx.compton_drag.coeff     = c->compton_coeff;
    ctx.atmosphere = kernel::sr::PusherAtmosphereContext(c->atm_gx1, c->atm_gx2, c->atm_gx3,
                                                         c->atm_x_surf, c->atm_ds);
    kernel::sr::PusherBoundaries<D> pb { bnd };
    ParticleArrays                  arr { 1u };
    fill_arrays(arr, p, n);
    auto                      EBw = wrap<D, 6>(g, const_cast<float*>(em));
    randacc_ndfield_t<D, 6>   EB(EBw);
    if (c->has_atmosphere) {
      using P = kernel::sr::PusherPolicy<M, ::traits::emission::NoPolicy_t,
                                         ::traits::custom_prtl_update::NoPolicy_t,
                                         ::traits::extfields::NoPolicy_t, true>;
      kernel::sr::Pusher_kernel<M, P> k(ctx, pb, arr, EB, metric, P {});
      parallel_chunks(n, [&](std::size_t lo, std::size_t hi, int) {
        for (uint32_t q = lo; q < hi; ++q) k(q);
      });
    } else {
      kernel::sr::Pusher_kernel<M> k(ctx, pb, arr, EB, metric);
      parallel_chunks(n, [&](std::size_t lo, std::size_t hi, int) {
        for (uint32_t q = lo; q < hi; ++q) k(q);
      });
    }
  }

  template <Dimension D>
  void deposit(const orc_grid_t* g, const orc_prtls_t* p, uint32_t n, float charge, float dt,
               float dx, float* cur) {
    using M     = metric::Minkowski<D>;
    auto metric = make_metric<D>(g, dx, nullptr);
    auto J      = wrap<D, 3>(g, cur);
    auto Js     = Kokkos::Experimental::create_scatter_view(J);
    ParticleArrays a { 1u };
    fill_arrays(a, p, n);
    if (g_threads <= 1) {
      kernel::DepositCurrents_kernel<SimEngine::SRPIC, M, SHAPE_ORDER> k(
        Js, a.i1, a.i2, a.i3, a.i1_prev, a.i2_prev, a.i3_prev, a.dx1, a.dx2, a.dx3, a.dx1_prev,
        a.dx2_prev, a.dx3_prev, a.ux1, a.ux2, a.ux3, a.phi, a.weight, a.tag, metric, charge, dt);
      for (uint32_t q = 0; q < n; ++q) k(q);
      return;
    }
    // threaded: one private copy of J per worker, reduced afterwards -- the duplication
    // strategy Kokkos' ScatterView uses on the OpenMP backend
    const std::size_t tot = J.size();
    const int         nt  = g_threads;
    std::vector<std::vector<float>> priv(nt, std::vector<float>(tot, 0.0f));
    parallel_chunks(n, [&](std::size_t lo, std::size_t hi, int t) {
      auto Jt  = wrap<D, 3>(g, priv[t].data());
      auto Jts = Kokkos::Experimental::create_scatter_view(Jt);
      kernel::DepositCurrents_kernel<SimEngine::SRPIC, M, SHAPE_ORDER> k(
        Jts, a.i1, a.i2, a.i3, a.i1_prev, a.i2_prev, a.i3_prev, a.dx1, a.dx2, a.dx3, a.dx1_prev,
        a.dx2_prev, a.dx3_prev, a.ux1, a.ux2, a.ux3, a.phi, a.weight, a.tag, metric, charge, dt);
      for (uint32_t q = lo; q < hi; ++q) k(q);
    });
    parallel_chunks(tot, [&](std::size_t lo, std::size_t hi, int) {
      for (int t = 0; t < nt; ++t)
        for (std::size_t x = lo; x < hi; ++x) cur[x] += priv[t][x];
    });
  }

} // namespace

#define BY_DIM(fn, ...)                                                                        \
  do {                                                                                         \
    check_ng(g);                                                                               \
    if (g->dim == 1) fn<Dim::_1D>(__VA_ARGS__);                                                \
    else if (g->dim == 2) fn<Dim::_2D>(__VA_ARGS__);                                           \
    else fn<Dim::_3D>(__VA_ARGS__);                                                            \
  } while (0)

extern "C" {
int  ref_shape_order() { return SHAPE_ORDER; }
void ref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int  ref_nghosts() { return (int)N_GHOSTS; }
void ref_faraday_mink(const orc_grid_t* g, float* em, float c1, float c2, const float* st) {
  BY_DIM(faraday, g, em, c1, c2, st);
}
void ref_ampere_mink(const orc_grid_t* g, float* em, float c1, float c2) {
  BY_DIM(ampere, g, em, c1, c2);
}
void ref_currents_ampere_mink(const orc_grid_t* g, float* em, float* cur, float coeff, float ppc0) {
  BY_DIM(currents_ampere, g, em, cur, coeff, ppc0);
}
void ref_filter_pass(const orc_grid_t* g, float* cur, const float* buff, const int* fbc) {
  BY_DIM(filter_pass, g, cur, buff, fbc);
}
void ref_push_sr_mink(const orc_grid_t* g, int order, const orc_pusher_t* ctx,
                      const orc_prtls_t* p, uint32_t npart, const float* em) {
  if (order != SHAPE_ORDER) throw std::runtime_error("ref: order != compile-time SHAPE_ORDER");
  BY_DIM(push, g, ctx, p, npart, em);
}
void ref_deposit_mink(const orc_grid_t* g, int order, const orc_prtls_t* p, uint32_t npart,
                      float charge, float dt, float dx, float* cur) {
  if (order != SHAPE_ORDER) throw std::runtime_error("ref: order != compile-time SHAPE_ORDER");
  BY_DIM(deposit, g, p, npart, charge, dt, dx, cur);
}
}
