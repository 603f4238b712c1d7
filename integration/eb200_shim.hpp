// The shim an Entity maintainer adds, as compilable code: the SRPIC Minkowski
// dispatchers srpic::{Faraday, Ampere, CurrentsAmpere, ParticlePush, CurrentsDeposit} hand their
// Kokkos views to libentity_b200.so instead of launching the Kokkos kernels. Nothing else of the
// reference changes: engines, framework, pgens, boundaries, exchanges, sort and the time loop are
// the reference's own. integration/apply_shim.py inserts the calls into a SCRATCH COPY of the
// reference (never into /root/reference); oracle/build_entity_xc.sh cuda_shim builds it.
//
// Views of a Kokkos-CUDA build are LayoutLeft device arrays: fields (n1 + 2G, n2 + 2G, ncomp) with
// i1 fastest and the component slowest -- the "component planes" layout of include/entity_b200.h --
// and one 1D view per particle array.
#pragma once
#include "../include/entity_b200.h"

#include <Kokkos_Core.hpp>

#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

namespace eb200shim {
  inline void check(eb200_ctx_t* c, int rc, const char* what) {
    if (rc != 0) {
      throw std::runtime_error(std::string("entity_b200: ") + what + ": " + eb200_last_error(c));
    }
  }

  // what srpic::ParticlePush tells srpic::CurrentsDeposit within one step
  struct StepState {
    bool fused = false;
  };
  inline StepState& state() {
    static StepState s;
    return s;
  }

  // srpic::CurrentsFilter through eb200_filter: opt-in at run time (EB200_SHIM_FILTER=1)
  inline bool filter_enabled() {
    static const bool on = std::getenv("EB200_SHIM_FILTER") != nullptr;
    return on;
  }

  // kernels are enqueued on the stream of Kokkos' default execution space instance
  inline void* stream() { return (void*)Kokkos::Cuda().cuda_stream(); }

  // one context per local domain, created on first use
  template <class DOM>
  eb200_ctx_t* ctx(DOM& dom) {
    static std::map<unsigned int, eb200_ctx_t*> ctxs;
    auto it = ctxs.find((unsigned int)dom.index());
    if (it != ctxs.end()) return it->second;
    constexpr int D = (int)DOM::D;
    eb200_config_t cfg {};
    cudaGetDevice(&cfg.device);
    cfg.strict_fp   = 0;
    cfg.grid.dim    = D;
    cfg.grid.ng     = (int)N_GHOSTS;
    cfg.shape_order = (int)SHAPE_ORDER;
    cfg.metric      = EB200_METRIC_MINKOWSKI;
    const auto n    = dom.mesh.n_active();
    const auto ext  = dom.mesh.extent();
    for (int a = 0; a < 3; ++a) cfg.grid.n[a] = a < D ? (int)n[a] : 1;
    cfg.metric_params[0] = (float)math::sqrt(dom.mesh.metric.template h_<1, 1>({}));
    for (int a = 0; a < D; ++a) cfg.metric_params[1 + a] = (float)ext[a].first;
    cfg.maxnpart = 0;
    for (auto& sp : dom.species) cfg.maxnpart = std::max<uint32_t>(cfg.maxnpart, (uint32_t)sp.maxnpart());
    eb200_ctx_t* c = nullptr;
    check(nullptr, eb200_init(&cfg, &c), "eb200_init");
    return ctxs[(unsigned int)dom.index()] = c;
  }

  template <class SP>
  eb200_prtls_t prtls(SP& sp) {
    eb200_prtls_t p {};
    p.i1 = sp.i1.data(), p.i2 = sp.i2.data(), p.i3 = sp.i3.data();
    p.dx1 = sp.dx1.data(), p.dx2 = sp.dx2.data(), p.dx3 = sp.dx3.data();
    p.ux1 = sp.ux1.data(), p.ux2 = sp.ux2.data(), p.ux3 = sp.ux3.data();
    p.weight = sp.weight.data();
    p.i1_prev = sp.i1_prev.data(), p.i2_prev = sp.i2_prev.data(), p.i3_prev = sp.i3_prev.data();
    p.dx1_prev = sp.dx1_prev.data(), p.dx2_prev = sp.dx2_prev.data(), p.dx3_prev = sp.dx3_prev.data();
    p.tag = sp.tag.data();
    p.phi = sp.phi.data();
    // payload planes are not touched by the pusher or the deposit
    return p;
  }

  // PusherBoundaries (src/kernels/pushers/context.h:124-178) as EB200_PBC_*
  template <class BCS>
  void particle_bcs(const BCS& b, int dim, int* pbc) {
    auto code = [](const ntt::PrtlBC& x) {
      if (x == ntt::PrtlBC::PERIODIC) return (int)EB200_PBC_PERIODIC;
      if (x == ntt::PrtlBC::ABSORB || x == ntt::PrtlBC::ATMOSPHERE) return (int)EB200_PBC_ABSORB;
      if (x == ntt::PrtlBC::REFLECT) return (int)EB200_PBC_REFLECT;
      if (x == ntt::PrtlBC::AXIS) return (int)EB200_PBC_AXIS;
      return (int)EB200_PBC_NONE;
    };
    for (int a = 0; a < 6; ++a) pbc[a] = EB200_PBC_NONE;
    for (int a = 0; a < dim; ++a) {
      pbc[2 * a]     = code(b[a].first);
      pbc[2 * a + 1] = code(b[a].second);
    }
  }

  // Mesh::flds_bc() as EB200_FBC_* for the filter: periodic faces wrap, conductor faces mirror
  // (digital_filter.hpp:99-388), every other kind leaves its ghost layer alone
  template <class BCS>
  void field_bcs(const BCS& b, int dim, int* fbc) {
    auto code = [](const ntt::FldsBC& x) {
      if (x == ntt::FldsBC::PERIODIC) return (int)EB200_FBC_PERIODIC;
      if (x == ntt::FldsBC::CONDUCTOR) return (int)EB200_FBC_CONDUCTOR;
      return (int)EB200_FBC_NONE;
    };
    for (int a = 0; a < 6; ++a) fbc[a] = EB200_FBC_NONE;
    for (int a = 0; a < dim; ++a) {
      fbc[2 * a]     = code(b[a].first);
      fbc[2 * a + 1] = code(b[a].second);
    }
  }
} // namespace eb200shim
