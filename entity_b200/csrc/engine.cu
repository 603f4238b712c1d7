// entity_b200 -- host-side mirror of the reference's SRPIC engine layer for one Minkowski
// domain. The functions carry the names and argument meaning of the dispatchers they replace
// (src/engines/srpic/{fieldsolvers.h,particle_pusher.h,currents.h}, src/engines/srpic/srpic.hpp),
// compute the same coefficients in the same precision, and enqueue the sm_100a kernels through
// the C ABI instead of Kokkos::parallel_for. No arithmetic on field or particle data happens here.
#include "launch.h"

#include <cmath>
#include <cstring>
#include <limits>

namespace eb200 {
  namespace srpic {

    struct Domain {
      eb200_ctx_t*                ctx;
      const eb200_srpic_params_t* prm;
      eb200_grid_t                grid;
      float                       dx;
      float                       xmin[3];
      float *                     em, *cur, *buff;
      eb200_species_t*            species;
      int                         nspecies;
      eb200_stream_t              stream;
    };

#define TRY(expr)                                                                              \
  do {                                                                                         \
    int rc_ = (expr);                                                                          \
    if (rc_ != EB200_OK) return rc_;                                                           \
  } while (0)

    // srpic::Faraday, fieldsolvers.h:36-98 (Cartesian branch)
    int Faraday(Domain& dom, float fraction) {
      const float dt = dom.prm->dt;
      const float dT = fraction * dom.prm->correction * dt;
      const float dx = std::sqrt(dom.dx * dom.dx); // sqrt(h_<1,1>)
      float       coeff1, coeff2;
      if (dom.grid.dim == 2) {
        coeff1 = dT / (dx * dx);
        coeff2 = dT;
      } else {
        coeff1 = dT / dx;
        coeff2 = 0.0f;
      }
      return eb200_faraday(dom.ctx, dom.em, coeff1, coeff2, dom.prm->stencil, dom.stream);
    }

    // srpic::Ampere, fieldsolvers.h:101-139
    int Ampere(Domain& dom, float fraction) {
      const float dt = dom.prm->dt;
      const float dT = fraction * dom.prm->correction * dt;
      const float dx = std::sqrt(dom.dx * dom.dx);
      float       coeff1, coeff2;
      if (dom.grid.dim == 2) {
        coeff1 = dT / (dx * dx);
        coeff2 = dT;
      } else {
        coeff1 = dT / dx;
        coeff2 = 0.0f;
      }
      return eb200_ampere(dom.ctx, dom.em, coeff1, coeff2, dom.stream);
    }

    // srpic::CurrentsAmpere, fieldsolvers.h:142-199 (Cartesian, no external current)
    int CurrentsAmpere(Domain& dom) {
      const eb200_srpic_params_t& p     = *dom.prm;
      const float                 coeff = -p.dt * p.q0 / (p.B0 * p.V0);
      return eb200_currents_ampere(dom.ctx, dom.em, dom.cur, coeff, p.ppc0, dom.stream);
    }

    static eb200_pusher_t pusher_context(const Domain& dom, const eb200_species_t& sp, double time) {
      // particle_pusher.h:92-141
      const eb200_srpic_params_t& p = *dom.prm;
      eb200_pusher_t              c;
      std::memset(&c, 0, sizeof(c));
      c.pusher_flags = sp.pusher_flags;
      c.drag_flags   = sp.drag_flags;
      c.mass         = sp.mass;
      c.charge       = sp.charge;
      c.time         = time;
      c.dt           = p.dt;
      c.omegaB0      = p.omegaB0;
      if (sp.pusher_flags & EB200_PUSHER_GCA) {
        c.gca_larmor_max      = p.gca_larmor_max;
        c.gca_e_ovr_b_sqr_max = p.gca_e_ovr_b_max * p.gca_e_ovr_b_max; // context.h:34-36
      }
      if (sp.drag_flags & EB200_DRAG_SYNCHROTRON) { // context.h:44-47
        const float gm = p.sync_gamma_rad * sp.mass;
        c.sync_coeff   = 0.1f * p.dt * p.omegaB0 / (gm * gm);
      }
      if (sp.drag_flags & EB200_DRAG_COMPTON) {
        const float gm  = p.compton_gamma_rad * sp.mass;
        c.compton_coeff = 0.1f * p.dt * p.omegaB0 / (gm * gm);
      }
      for (int a = 0; a < 6; ++a) c.pbc[a] = p.pbc[a];
      c.tag_outgoing = 0;
      for (int a = 0; a < 6; ++a) {
        if (a < 2 * dom.grid.dim && p.pbc[a] == EB200_PBC_NONE) c.tag_outgoing = 1;
      }
      c.dx = dom.dx;
      for (int a = 0; a < 3; ++a) c.xmin[a] = dom.xmin[a];
      return c;
    }

    static bool almost_zero(float x) { return std::fabs(x) <= std::numeric_limits<float>::epsilon(); }

    // srpic::ParticlePush, particle_pusher.h:36-185
    int ParticlePush(Domain& dom, double time) {
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.pusher_flags == EB200_PUSHER_NONE || sp.npart == 0) continue;
        const eb200_pusher_t c = pusher_context(dom, sp, time);
        TRY(eb200_push_sr(dom.ctx, &c, &sp.arrays, sp.npart, dom.em, dom.stream));
      }
      return EB200_OK;
    }

    // srpic::CurrentsDeposit, currents.h:64-87
    int CurrentsDeposit(Domain& dom) {
      TRY(eb200_zero_currents(dom.ctx, dom.cur, dom.stream));
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.pusher_flags == EB200_PUSHER_NONE || sp.npart == 0 || almost_zero(sp.charge)) continue;
        TRY(eb200_deposit(dom.ctx, &sp.arrays, sp.npart, sp.charge, dom.prm->dt, dom.cur,
                          dom.prm->deposit_mode, dom.stream));
      }
      return EB200_OK;
    }

    // ParticlePush + CurrentsDeposit in one pass per species (same result up to the order of
    // the additions into J)
    int ParticlePushAndDeposit(Domain& dom, double time) {
      TRY(eb200_zero_currents(dom.ctx, dom.cur, dom.stream));
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.pusher_flags == EB200_PUSHER_NONE || sp.npart == 0) continue;
        const eb200_pusher_t c = pusher_context(dom, sp, time);
        if (almost_zero(sp.charge)) {
          TRY(eb200_push_sr(dom.ctx, &c, &sp.arrays, sp.npart, dom.em, dom.stream));
        } else {
          TRY(eb200_push_deposit_sr(dom.ctx, &c, &sp.arrays, sp.npart, dom.em, dom.cur,
                                    dom.stream));
        }
      }
      return EB200_OK;
    }

    // srpic::CurrentsFilter, currents.h:89-119
    int CurrentsFilter(Domain& dom) {
      return eb200_filter(dom.ctx, dom.cur, dom.buff, dom.prm->nfilter, dom.prm->fbc, dom.stream);
    }

    // Metadomain::SortParticles, metadomain_sort.cpp:17-37
    int SortParticles(Domain& dom, uint32_t step) {
      const int ci = dom.prm->clear_interval, si = dom.prm->sort_interval;
      const bool clear = (ci > 0) && (step % (uint32_t)ci == 0u) && (step > 0u);
      const bool sort  = (si > 0) && (step % (uint32_t)si == 0u);
      if (!clear && !sort) return EB200_OK;
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.npart == 0) continue;
        uint32_t n = sp.npart;
        TRY(eb200_sort_particles(dom.ctx, &sp.arrays, &n, clear ? 1 : 0, dom.stream));
        sp.npart = n;
      }
      return EB200_OK;
    }

    // SRPICEngine::step_forward, srpic.hpp:65-188
    int step_forward(Domain& dom, uint32_t step, double time) {
      const eb200_srpic_params_t& p = *dom.prm;
      if (step == 0) {
        TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 0, 6, p.fbc, dom.stream));
      }
      if (p.fieldsolver_enabled) {
        TRY(Faraday(dom, 0.5f));
        TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 3, 6, p.fbc, dom.stream));
      }
      if (p.deposit_enabled && p.fuse_push_deposit) {
        TRY(ParticlePushAndDeposit(dom, time));
      } else {
        TRY(ParticlePush(dom, time));
        if (p.deposit_enabled) {
          TRY(CurrentsDeposit(dom));
        }
      }
      if (p.deposit_enabled) {
        TRY(eb200_sync_currents(dom.ctx, dom.cur, dom.buff, p.fbc, dom.stream));
        TRY(eb200_comm_fields(dom.ctx, dom.cur, 3, 0, 3, p.fbc, dom.stream));
        TRY(CurrentsFilter(dom));
      }
      // CommunicateParticles: a single periodic domain has no neighbour to migrate to
      if (p.fieldsolver_enabled) {
        TRY(Faraday(dom, 0.5f));
        TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 3, 6, p.fbc, dom.stream));
        TRY(Ampere(dom, 1.0f));
        if (p.deposit_enabled) {
          TRY(CurrentsAmpere(dom));
        }
        TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 0, 3, p.fbc, dom.stream));
        TRY(eb200_comm_fields(dom.ctx, dom.cur, 3, 0, 3, p.fbc, dom.stream));
      }
      TRY(SortParticles(dom, step));
      return EB200_OK;
    }

  } // namespace srpic
} // namespace eb200

// accessors into the opaque context (capi.cu)
extern "C" int eb200_ctx_grid(const eb200_ctx_t* ctx, eb200_grid_t* grid, float* dx, float* xmin3);

extern "C" int eb200_srpic_step(eb200_ctx_t* ctx, const eb200_srpic_params_t* prm, float* em,
                                float* cur, float* buff, eb200_species_t* species, int nspecies,
                                uint32_t step, double time, eb200_stream_t stream) {
  if (!ctx || !prm || !em || !cur || !buff || (nspecies > 0 && !species)) return EB200_ERR_ARG;
  eb200::srpic::Domain dom;
  dom.ctx = ctx;
  dom.prm = prm;
  if (eb200_ctx_grid(ctx, &dom.grid, &dom.dx, dom.xmin) != EB200_OK) return EB200_ERR_ARG;
  dom.em       = em;
  dom.cur      = cur;
  dom.buff     = buff;
  dom.species  = species;
  dom.nspecies = nspecies;
  dom.stream   = stream;
  return eb200::srpic::step_forward(dom, step, time);
}
