// TEST INFRASTRUCTURE ONLY.
// extern "C" driver around the REFERENCE's own curvilinear-SR and GR kernel headers and metric
// classes, compiled in place from $(REF)/src against the serial mini-Kokkos in ref_shim/ (no
// reference source is copied into this repo). One library per compile-time SHAPE_ORDER:
// oracle/_ref/libref_curv_o<O>.so. This IS the checker for the curvilinear / GR rows: the
// golden vectors under tests/golden/ are produced from it (tests/golden/make_curv_golden.py).
#include "oracle.h"

#include "enums.h"
#include "global.h"

#include "arch/kokkos_aliases.h"
#include "utils/numeric.h"

#include "metrics/kerr_schild.h"
#include "metrics/kerr_schild_0.h"
#include "metrics/qkerr_schild.h"
#include "metrics/qspherical.h"
#include "metrics/spherical.h"

#include "kernels/ampere_gr.hpp"
#include "kernels/ampere_sr.hpp"
#include "kernels/aux_fields_gr.hpp"
#include "kernels/currents_deposit.hpp"
#include "kernels/digital_filter.hpp"
#include "kernels/faraday_gr.hpp"
#include "kernels/faraday_sr.hpp"
#include "kernels/pushers/gr.hpp"
#include "kernels/pushers/sr.hpp"

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

using namespace ntt;

extern "C" {
// metric kinds as in include/entity_b200.h (EB200_METRIC_*)
typedef struct {
  int   kind; // 1 spherical, 2 qspherical, 3 kerr_schild, 4 qkerr_schild, 5 kerr_schild_0
  int   n1, n2;
  float x1min, x1max, x2min, x2max;
  float r0, h, a;
} refc_metric_t;

typedef struct {
  int   pusher_flags;
  float mass, charge, dt, omegaB0, epsilon;
  int   niter;
  int   pbc[6];
} refc_pusher_gr_t;
}

namespace {
  constexpr auto D2 = Dim::_2D;

  template <class M>
  M make(const refc_metric_t* m) {
    std::vector<ncells_t> res { (ncells_t)m->n1, (ncells_t)m->n2 };
    boundaries_t<real_t>  ext { { m->x1min, m->x1max }, { m->x2min, m->x2max } };
    std::map<std::string, real_t> prm { { "r0", m->r0 }, { "h", m->h }, { "a", m->a } };
    return M(res, ext, prm);
  }

  template <unsigned short N>
  auto wrap(const orc_grid_t* g, float* p) -> ndfield_t<D2, N> {
    const std::size_t G2 = 2 * (std::size_t)g->ng;
    return ndfield_t<D2, N>(p, g->n[0] + G2, g->n[1] + G2);
  }

  void check_ng(const orc_grid_t* g) {
    if ((uint32_t)g->ng != N_GHOSTS || g->dim != 2) {
      throw std::runtime_error("refc: 2D grids with ng == N_GHOSTS only");
    }
  }

  PrtlBC to_pbc(int b, bool gr) {
    switch (b) {
      case ORC_PBC_PERIODIC: return PrtlBC::PERIODIC;
      case ORC_PBC_ABSORB: return gr ? PrtlBC::HORIZON : PrtlBC::ABSORB;
      case ORC_PBC_REFLECT: return PrtlBC::REFLECT;
      case ORC_PBC_AXIS: return PrtlBC::AXIS;
      default: return PrtlBC::SYNC;
    }
  }

  boundaries_t<FldsBC> to_fbcs(const int* fbc) {
    auto cv = [](int b) {
      switch (b) {
        case ORC_FBC_PERIODIC: return FldsBC::PERIODIC;
        case ORC_FBC_CONDUCTOR: return FldsBC::CONDUCTOR;
        case ORC_FBC_AXIS: return FldsBC::AXIS;
        default: return FldsBC::SYNC;
      }
    };
    return { { cv(fbc[0]), cv(fbc[1]) }, { cv(fbc[2]), cv(fbc[3]) } };
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

  // rows i2 in [j0, j1), columns i1 in [i0, i1): serial, row-ascending (the order a Serial
  // Kokkos backend with LayoutRight iteration would NOT necessarily use; it only matters for
  // the in-place GR Ampere axis rows, see curv.cu)
  template <class K>
  void loop2(int i0, int i1, int j0, int j1, const K& k) {
    for (int j = j0; j < j1; ++j)
      for (int i = i0; i < i1; ++i) k((cellidx_t)i, (cellidx_t)j);
  }

  /* ------------------------------------------------------------------ metrics */
  template <class M>
  void eval_sr(const M& m, int nq, const float* x1, const float* x2, float* out) {
    for (int q = 0; q < nq; ++q) {
      const coord_t<D2> x { x1[q], x2[q] };
      float*            o = out + 16 * q;
      o[0] = m.template h_<1, 1>(x);
      o[1] = m.template h_<2, 2>(x);
      o[2] = m.template h_<3, 3>(x);
      o[3] = m.template sqrt_h_<1, 1>(x);
      o[4] = m.template sqrt_h_<2, 2>(x);
      o[5] = m.template sqrt_h_<3, 3>(x);
      o[6] = m.sqrt_det_h(x);
      o[7] = m.polar_area(x1[q]);
      o[8] = m.template convert<1, Crd::Cd, Crd::Ph>(x1[q]);
      o[9] = m.template convert<2, Crd::Cd, Crd::Ph>(x2[q]);
      o[10] = m.template convert<1, Crd::Ph, Crd::Cd>(o[8]);
      o[11] = m.template convert<2, Crd::Ph, Crd::Cd>(o[9]);
      for (int k = 12; k < 16; ++k) o[k] = 0.0f;
    }
  }

  template <class M>
  void eval_gr(const M& m, int nq, const float* x1, const float* x2, float* out) {
    for (int q = 0; q < nq; ++q) {
      const coord_t<D2> x { x1[q], x2[q] };
      float*            o = out + 32 * q;
      o[0]  = m.template h_<1, 1>(x);
      o[1]  = m.template h_<2, 2>(x);
      o[2]  = m.template h_<3, 3>(x);
      o[3]  = m.template h_<1, 3>(x);
      o[4]  = m.template h<1, 1>(x);
      o[5]  = m.template h<2, 2>(x);
      o[6]  = m.template h<3, 3>(x);
      o[7]  = m.template h<1, 3>(x);
      o[8]  = m.alpha(x);
      o[9]  = m.beta1(x);
      o[10] = m.sqrt_det_h(x);
      o[11] = m.sqrt_det_h_tilde(x);
      o[12] = m.polar_area(x1[q]);
      o[13] = m.dr_alpha(x);
      o[14] = m.dt_alpha(x);
      o[15] = m.dr_beta1(x);
      o[16] = m.dt_beta1(x);
      o[17] = m.dr_h11(x);
      o[18] = m.dr_h22(x);
      o[19] = m.dr_h33(x);
      o[20] = m.dr_h13(x);
      o[21] = m.dt_h11(x);
      o[22] = m.dt_h22(x);
      o[23] = m.dt_h33(x);
      o[24] = m.dt_h13(x);
      o[25] = m.template convert<2, Crd::Cd, Crd::Ph>(x2[q]);
      o[26] = m.template convert<2, Crd::Ph, Crd::Cd>(o[25]);
      for (int k = 27; k < 32; ++k) o[k] = 0.0f;
    }
  }

  /* --------------------------------------------------------------- SR kernels */
  template <class M>
  void push_sr(const M& metric, const orc_grid_t* g, const orc_pusher_t* c, const orc_prtls_t* p,
               uint32_t n, const float* em) {
    boundaries_t<PrtlBC> bnd { { to_pbc(c->pbc[0], false), to_pbc(c->pbc[1], false) },
                               { to_pbc(c->pbc[2], false), to_pbc(c->pbc[3], false) } };
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
    kernel::sr::PusherBoundaries<D2> pb { bnd };
    ParticleArrays                   arr { 1u };
    fill_arrays(arr, p, n);
    auto                     EBw = wrap<6>(g, const_cast<float*>(em));
    randacc_ndfield_t<D2, 6> EB(EBw);
    if (c->has_atmosphere) {
      using P = kernel::sr::PusherPolicy<M, ::traits::emission::NoPolicy_t,
                                         ::traits::custom_prtl_update::NoPolicy_t,
                                         ::traits::extfields::NoPolicy_t, true>;
      kernel::sr::Pusher_kernel<M, P> k(ctx, pb, arr, EB, metric, P {});
      for (uint32_t q = 0; q < n; ++q) k(q);
    } else {
      kernel::sr::Pusher_kernel<M> k(ctx, pb, arr, EB, metric);
      for (uint32_t q = 0; q < n; ++q) k(q);
    }
  }

  template <SimEngine::type S, class M>
  void deposit(const M& metric, const orc_grid_t* g, const orc_prtls_t* p, uint32_t n,
               float charge, float dt, float* cur) {
    auto J  = wrap<3>(g, cur);
    auto Js = Kokkos::Experimental::create_scatter_view(J);
    ParticleArrays a { 1u };
    fill_arrays(a, p, n);
    kernel::DepositCurrents_kernel<S, M, SHAPE_ORDER> k(
      Js, a.i1, a.i2, a.i3, a.i1_prev, a.i2_prev, a.i3_prev, a.dx1, a.dx2, a.dx3, a.dx1_prev,
      a.dx2_prev, a.dx3_prev, a.ux1, a.ux2, a.ux3, a.phi, a.weight, a.tag, metric, charge, dt);
    for (uint32_t q = 0; q < n; ++q) k(q);
  }

  template <class M>
  void fields_sr(const M& metric, int which, const orc_grid_t* g, float* em, float* cur,
                 float coeff, float inv_n0, const int* fbc) {
    const int  G = (int)N_GHOSTS, n1 = g->n[0], n2 = g->n[1];
    const auto b = to_fbcs(fbc);
    const int  rows = n2 + (fbc[3] == ORC_FBC_AXIS ? 1 : 0);
    auto       EB   = wrap<6>(g, em);
    if (which == 0) {
      loop2(G, G + n1, G, G + n2, kernel::sr::Faraday_kernel<M>(EB, metric, coeff, b));
    } else if (which == 1) {
      loop2(G, G + n1, G, G + rows,
            kernel::sr::Ampere_kernel<M>(EB, metric, coeff, (ncells_t)n2, b));
    } else {
      auto J = wrap<3>(g, cur);
      loop2(G, G + n1, G, G + rows,
            kernel::sr::CurrentsAmpere_kernel<M>(EB, J, metric, coeff, inv_n0, (ncells_t)n2, b));
    }
  }

  /* --------------------------------------------------------------- GR kernels */
  template <class M>
  void push_gr(const M& metric, const orc_grid_t* g, const refc_pusher_gr_t* c,
               const orc_prtls_t* p, uint32_t n, const float* em, const float* em0) {
    boundaries_t<PrtlBC> bnd { { to_pbc(c->pbc[0], true), to_pbc(c->pbc[1], true) },
                               { to_pbc(c->pbc[2], true), to_pbc(c->pbc[3], true) } };
    kernel::gr::PusherContext ctx { c->mass, c->charge, c->dt, c->omegaB0, c->epsilon,
                                    (unsigned short)c->niter, g->n[0], g->n[1], g->n[2] };
    kernel::gr::PusherBoundaries<D2> pb { bnd };
    ParticleArrays                   arr { 1u };
    fill_arrays(arr, p, n);
    auto DB  = wrap<6>(g, const_cast<float*>(em));
    auto DB0 = wrap<6>(g, const_cast<float*>(em0));
    kernel::gr::Pusher_kernel<M> k(ctx, pb, arr, DB, DB0, metric);
    if (c->pusher_flags == 1) {
      for (uint32_t q = 0; q < n; ++q) k(kernel::gr::Massless_t {}, q);
    } else {
      for (uint32_t q = 0; q < n; ++q) k(kernel::gr::Massive_t {}, q);
    }
  }

  template <class M>
  void fields_gr(const M& metric, int which, const orc_grid_t* g, float* a, float* b_, float* c,
                 float coeff, const int* fbc) {
    const int  G = (int)N_GHOSTS, n1 = g->n[0], n2 = g->n[1];
    const auto bc = to_fbcs(fbc);
    const int  rows_axis = n2 + (fbc[3] == ORC_FBC_AXIS ? 1 : 0);
    switch (which) {
      case 0: // aux E: a = D, b = B, c = E out
        loop2(G - 1, G + n1, G, G + rows_axis,
              kernel::gr::ComputeAuxE_kernel<M>(wrap<6>(g, a), wrap<6>(g, b_), wrap<6>(g, c), metric));
        break;
      case 1: // aux H
        loop2(G - 1, G + n1, G, G + rows_axis,
              kernel::gr::ComputeAuxH_kernel<M>(wrap<6>(g, a), wrap<6>(g, b_), wrap<6>(g, c), metric));
        break;
      case 2: // faraday: a = Bin, b = Bout, c = E
        loop2(G, G + n1, G, G + n2,
              kernel::gr::Faraday_kernel<M>(wrap<6>(g, a), wrap<6>(g, b_), wrap<6>(g, c), metric,
                                            coeff, (ncells_t)n2, bc));
        break;
      case 3: // ampere: a = Din, b = Dout, c = H
        loop2(G, G + n1, G, G + n2 + 1,
              kernel::gr::Ampere_kernel<M>(wrap<6>(g, a), wrap<6>(g, b_), wrap<6>(g, c), metric,
                                           coeff, (ncells_t)n2, bc));
        break;
      case 4: // currents ampere: a = D, b = J
        loop2(G, G + n1, G, G + n2 + 1,
              kernel::gr::CurrentsAmpere_kernel<M>(wrap<6>(g, a), wrap<3>(g, b_), metric, coeff,
                                                   (ncells_t)n2, bc));
        break;
      default: throw std::runtime_error("refc: bad GR field kernel id");
    }
  }

} // namespace

#define SR_METRIC(m, CALL)                                                                     \
  do {                                                                                         \
    if ((m)->kind == 1) {                                                                      \
      auto M_ = make<metric::Spherical<D2>>(m);                                                \
      CALL;                                                                                    \
    } else if ((m)->kind == 2) {                                                               \
      auto M_ = make<metric::QSpherical<D2>>(m);                                               \
      CALL;                                                                                    \
    } else                                                                                     \
      throw std::runtime_error("refc: not an SR curvilinear metric");                          \
  } while (0)

#define GR_METRIC(m, CALL)                                                                     \
  do {                                                                                         \
    if ((m)->kind == 3) {                                                                      \
      auto M_ = make<metric::KerrSchild<D2>>(m);                                               \
      CALL;                                                                                    \
    } else if ((m)->kind == 4) {                                                               \
      auto M_ = make<metric::QKerrSchild<D2>>(m);                                              \
      CALL;                                                                                    \
    } else if ((m)->kind == 5) {                                                               \
      auto M_ = make<metric::KerrSchild0<D2>>(m);                                              \
      CALL;                                                                                    \
    } else                                                                                     \
      throw std::runtime_error("refc: not a GR metric");                                       \
  } while (0)

extern "C" {
int refc_shape_order() { return SHAPE_ORDER; }
int refc_nghosts() { return (int)N_GHOSTS; }

// out: [nq][16] for SR metrics, [nq][32] for GR metrics (layout: eval_sr / eval_gr above)
void refc_metric_eval(const refc_metric_t* m, int nq, const float* x1, const float* x2, float* out) {
  if (m->kind <= 2) {
    SR_METRIC(m, eval_sr(M_, nq, x1, x2, out));
  } else {
    GR_METRIC(m, eval_gr(M_, nq, x1, x2, out));
  }
}

float refc_metric_dxmin(const refc_metric_t* m) {
  float r = 0.0f;
  if (m->kind <= 2) {
    SR_METRIC(m, r = M_.dxMin());
  } else {
    GR_METRIC(m, r = M_.dxMin());
  }
  return r;
}

void refc_push_sr(const refc_metric_t* m, const orc_grid_t* g, const orc_pusher_t* c,
                  const orc_prtls_t* p, uint32_t n, const float* em) {
  check_ng(g);
  SR_METRIC(m, push_sr(M_, g, c, p, n, em));
}

void refc_deposit(const refc_metric_t* m, const orc_grid_t* g, const orc_prtls_t* p, uint32_t n,
                  float charge, float dt, float* cur) {
  check_ng(g);
  if (m->kind <= 2) {
    SR_METRIC(m, deposit<SimEngine::SRPIC>(M_, g, p, n, charge, dt, cur));
  } else {
    GR_METRIC(m, deposit<SimEngine::GRPIC>(M_, g, p, n, charge, dt, cur));
  }
}

// which: 0 Faraday, 1 Ampere, 2 CurrentsAmpere
void refc_fields_sr(const refc_metric_t* m, int which, const orc_grid_t* g, float* em, float* cur,
                    float coeff, float inv_n0, const int* fbc) {
  check_ng(g);
  SR_METRIC(m, fields_sr(M_, which, g, em, cur, coeff, inv_n0, fbc));
}

void refc_filter_sph(const orc_grid_t* g, float* cur, const float* buff, const int* fbc) {
  check_ng(g);
  auto     A = wrap<3>(g, cur);
  auto     B = wrap<3>(g, const_cast<float*>(buff));
  ncells_t size[2] { (ncells_t)g->n[0], (ncells_t)g->n[1] };
  const int G = (int)N_GHOSTS;
  const int rows = g->n[1] + (fbc[3] == ORC_FBC_AXIS ? 1 : 0);
  using K = kernel::DigitalFilter_kernel<D2, Coord::Spherical>;
  loop2(G, G + g->n[0], G, G + rows, K(A, B, size, to_fbcs(fbc)));
}

void refc_push_gr(const refc_metric_t* m, const orc_grid_t* g, const refc_pusher_gr_t* c,
                  const orc_prtls_t* p, uint32_t n, const float* em, const float* em0) {
  check_ng(g);
  GR_METRIC(m, push_gr(M_, g, c, p, n, em, em0));
}

// which: 0 aux E, 1 aux H, 2 Faraday, 3 Ampere, 4 CurrentsAmpere (argument roles in fields_gr)
void refc_fields_gr(const refc_metric_t* m, int which, const orc_grid_t* g, float* a, float* b,
                    float* c, float coeff, const int* fbc) {
  check_ng(g);
  GR_METRIC(m, fields_gr(M_, which, g, a, b, c, coeff, fbc));
}

// TimeAverageDB_kernel(em, em0): em0 = (em0 + em) / 2; TimeAverageJ_kernel(cur, cur0): cur = ...
void refc_time_average_db(const orc_grid_t* g, float* em, float* em0) {
  check_ng(g);
  const int G = (int)N_GHOSTS;
  loop2(G, G + g->n[0], G, G + g->n[1],
        kernel::gr::TimeAverageDB_kernel<D2>(wrap<6>(g, em), wrap<6>(g, em0)));
}

void refc_time_average_j(const orc_grid_t* g, float* cur, float* cur0) {
  check_ng(g);
  const int G = (int)N_GHOSTS;
  loop2(G, G + g->n[0], G, G + g->n[1],
        kernel::gr::TimeAverageJ_kernel<D2>(wrap<3>(g, cur), wrap<3>(g, cur0)));
}
}
