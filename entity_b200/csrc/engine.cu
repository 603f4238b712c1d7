// entity_b200 -- host-side mirror of the reference's SRPIC engine layer for one Minkowski
// domain. The functions carry the names and argument meaning of the dispatchers they replace
// (src/engines/srpic/{fieldsolvers.h,particle_pusher.h,currents.h}, src/engines/srpic/srpic.hpp),
// compute the same coefficients in the same precision, and enqueue the sm_100a kernels through
// the C ABI instead of Kokkos::parallel_for. No arithmetic on field or particle data happens here.
#include "launch.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

// ---- phase profiler: event pairs on the caller's stream, owned by the engine state -------
namespace eb200 {
  struct Profiler {
    bool on = false;
    struct Rec {
      int         phase;
      cudaEvent_t a, b;
    };
    std::vector<Rec>         recs;
    std::vector<cudaEvent_t> pool;

    cudaEvent_t get() {
      if (!pool.empty()) {
        cudaEvent_t e = pool.back();
        pool.pop_back();
        return e;
      }
      cudaEvent_t e;
      cudaEventCreate(&e);
      return e;
    }
  };

  struct HostMirror {
    std::vector<void*>  ptrs;
    std::vector<size_t> sizes;
    // streamed host step: copy streams (non-blocking) and a pool of timing-free events
    cudaStream_t             main = nullptr, up = nullptr, down = nullptr;
    std::vector<cudaEvent_t> events;
    size_t                   next_event = 0;

    bool streams() {
      if (main) return true;
      if (cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking) != cudaSuccess) return false;
      if (cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking) != cudaSuccess) return false;
      if (cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking) != cudaSuccess) return false;
      return true;
    }
    cudaEvent_t event() {
      if (next_event == events.size()) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        events.push_back(e);
      }
      return events[next_event++];
    }
  };

  // One step over HOST particle arrays, streamed: while the fused push+deposit works on chunk k
  // of a species, chunk k+1 is on its way up and chunk k-1 on its way down (PCIe is full
  // duplex). Only what the step reads goes up (tag, i, dx, u, weight: i_prev/dx_prev are
  // overwritten by the pusher before anything reads them, sr.hpp:137-153) and only what it
  // writes comes down (i, dx, u, i_prev, dx_prev, tag: weight is not modified).
  struct HostStreamer {
    HostMirror*      m;
    eb200_species_t* host; // host pointers, same order as the device species of the Domain
    size_t           chunk;
    uint64_t         up_bytes = 0, down_bytes = 0;
  };

  struct EngineState {
    Profiler   prof;
    HostMirror mirror;
    // MATCH faces applied by srpic::FieldBoundaries inside the step (eb200_srpic_set_match)
    std::vector<eb200_match_face_t> match;
    const float*                    match_target = nullptr;
    int                             match_mask   = 0;
    // MATCH / ATMOSPHERE faces of a curvilinear domain (eb200_srpic_set_field_bcs)
    std::vector<eb200_field_bc_t> curv_bcs;
    // the pgen's ext_current as a mode table (eb200_srpic_set_ext_current)
    bool                has_ext = false;
    eb200_ext_current_t ext {};
    // srpic::ParticleInjector for an ATMOSPHERE face (eb200_srpic_set_atmosphere_injector)
    bool               has_atm = false;
    eb200_atmosphere_t atm {};
    // emission policies per emitting species (eb200_srpic_set_emission)
    struct Emit {
      bool             on = false;
      int              photon_species = -1;
      eb200_emission_t policy {};
    };
    std::vector<Emit> emission;
  };

  struct PhaseScope {
    Profiler*    p;
    cudaStream_t st;
    size_t       idx;
    bool         live;

    PhaseScope(Profiler* prof, int phase, cudaStream_t s) : p { prof }, st { s }, idx { 0 }, live { false } {
      if (p && p->on) {
        Profiler::Rec r { phase, p->get(), p->get() };
        cudaEventRecord(r.a, st);
        p->recs.push_back(r);
        idx  = p->recs.size() - 1;
        live = true;
      }
    }

    ~PhaseScope() {
      if (live) cudaEventRecord(p->recs[idx].b, st);
    }
  };
} // namespace eb200

extern "C" eb200::EngineState* eb200_ctx_engine_state(eb200_ctx_t* ctx);
extern "C" int                 eb200_ctx_metric(const eb200_ctx_t* ctx);
extern "C" int                 eb200_ctx_has_comm(const eb200_ctx_t* ctx);
extern "C" int                 eb200_ctx_sort_flags(const eb200_ctx_t* ctx);
extern "C" int                 eb200_ctx_lean_prev(const eb200_ctx_t* ctx);

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
      Profiler*                   prof;
      HostStreamer*               host = nullptr; // non-null: particle arrays stream from/to the host
      const EngineState*          eng  = nullptr; // MATCH faces, if any
      bool                        curv = false;   // spherical / qspherical metric
      uint32_t                    step = 0;       // of this step_forward (streams of the emission draws)
    };

#define PHASE(dom, which) PhaseScope phase_scope_((dom).prof, (which), (cudaStream_t)(dom).stream)

#define TRY(expr)                                                                              \
  do {                                                                                         \
    int rc_ = (expr);                                                                          \
    if (rc_ != EB200_OK) return rc_;                                                           \
  } while (0)

    // srpic::Faraday, fieldsolvers.h:36-98 (Cartesian branch)
    int Faraday(Domain& dom, float fraction) {
      const float dt = dom.prm->dt;
      const float dT = fraction * dom.prm->correction * dt;
      if (dom.curv) { // kernel::sr::Faraday_kernel over rangeActiveCells (fieldsolvers.h:89-97)
        return eb200_faraday_sr(dom.ctx, dom.em, dT, dom.prm->fbc, dom.stream);
      }
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
      if (dom.curv) { // kernel::sr::Ampere_kernel over RangeWithAxisBCs (fieldsolvers.h:128-137)
        return eb200_ampere_sr(dom.ctx, dom.em, dT, dom.prm->fbc, dom.stream);
      }
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
      if (dom.curv) { // fieldsolvers.h:183-198
        const float coeff = -p.dt * p.q0 * p.n0 / p.B0;
        return eb200_currents_ampere_sr(dom.ctx, dom.em, dom.cur, coeff, 1.0f / p.n0, p.fbc,
                                        dom.stream);
      }
      const float                 coeff = -p.dt * p.q0 / (p.B0 * p.V0);
      if (dom.eng && dom.eng->has_ext) {
        return eb200_currents_ampere_ext(dom.ctx, dom.em, dom.cur, coeff, p.ppc0, &dom.eng->ext,
                                         dom.stream);
      }
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
      if (p.has_atmosphere) { // particle_pusher.h:45-80, 121-127
        c.has_atmosphere = 1;
        c.atm_gx1 = p.atm_g[0], c.atm_gx2 = p.atm_g[1], c.atm_gx3 = p.atm_g[2];
        c.atm_x_surf = p.atm_x_surf;
        c.atm_ds     = p.atm_ds;
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

    // the pusher of a species with an emission policy (particle_pusher.h:128-183): the photons
    // are appended to the emitted species, whose npart grows before its own turn in the loop
    static const EngineState::Emit* emission_of(const Domain& dom, int s) {
      if (!dom.eng || s >= (int)dom.eng->emission.size() || !dom.eng->emission[s].on) return nullptr;
      return &dom.eng->emission[s];
    }

    static int PushWithEmission(Domain& dom, int s, const eb200_pusher_t& c, const EngineState::Emit& e) {
      if (e.photon_species < 0 || e.photon_species >= dom.nspecies || e.photon_species == s) {
        return EB200_ERR_ARG;
      }
      eb200_species_t& sp = dom.species[s];
      eb200_species_t& ph = dom.species[e.photon_species];
      eb200_emission_t pol = e.policy;
      pol.photons          = ph.arrays;
      pol.photon_npart     = ph.npart;
      pol.photon_maxnpart  = ph.maxnpart;
      pol.step             = dom.step;
      pol.call             = (uint32_t)s;
      const int rc = eb200_push_sr_emission(dom.ctx, &c, &sp.arrays, sp.npart, dom.em, &pol, dom.stream);
      ph.npart = pol.photon_npart;
      return rc;
    }

    // srpic::ParticlePush, particle_pusher.h:36-185
    int ParticlePush(Domain& dom, double time) {
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.pusher_flags == EB200_PUSHER_NONE || sp.npart == 0) continue;
        const eb200_pusher_t c = pusher_context(dom, sp, time);
        if (const EngineState::Emit* e = emission_of(dom, s)) {
          TRY(PushWithEmission(dom, s, c, *e));
          continue;
        }
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
    // ParticleArrays slots (eb200_prtls_t order) the SR step reads / writes
    static const int    kSlotsIn[]   = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 16 };
    static const int    kSlotsOut[]  = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 14, 15, 16 };
    static const size_t kSlotElem[17] = { 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2 };

    static int StreamedPushAndDeposit(Domain& dom, double time) {
      HostStreamer& H  = *dom.host;
      cudaStream_t  st = (cudaStream_t)dom.stream;
      const int     mode = dom.prm->deposit_mode == EB200_DEPOSIT_AGGREGATED ? EB200_DEPOSIT_AGGREGATED
                                                                             : EB200_DEPOSIT_ATOMIC;
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.pusher_flags == EB200_PUSHER_NONE || sp.npart == 0) continue;
        const eb200_pusher_t c  = pusher_context(dom, sp, time);
        void**               hp = (void**)&H.host[s].arrays;
        void**               dp = (void**)&sp.arrays;
        for (size_t lo = 0; lo < sp.npart; lo += H.chunk) {
          const size_t n = std::min<size_t>(H.chunk, sp.npart - lo);
          for (int k : kSlotsIn) {
            if (!hp[k]) continue;
            const size_t e = kSlotElem[k];
            if (cudaMemcpyAsync((char*)dp[k] + lo * e, (const char*)hp[k] + lo * e, n * e,
                                cudaMemcpyHostToDevice, H.m->up) != cudaSuccess) return EB200_ERR_CUDA;
            H.up_bytes += n * e;
          }
          cudaEvent_t arrived = H.m->event();
          cudaEventRecord(arrived, H.m->up);
          cudaStreamWaitEvent(st, arrived, 0);
          eb200_prtls_t part = sp.arrays;
          void**        pp   = (void**)&part;
          for (int k = 0; k < 17; ++k) {
            if (pp[k]) pp[k] = (char*)pp[k] + lo * kSlotElem[k];
          }
          if (almost_zero(sp.charge)) {
            TRY(eb200_push_sr(dom.ctx, &c, &part, (uint32_t)n, dom.em, dom.stream));
          } else {
            TRY(eb200_push_deposit_sr(dom.ctx, &c, &part, (uint32_t)n, dom.em, dom.cur, mode, dom.stream));
          }
          cudaEvent_t pushed = H.m->event();
          cudaEventRecord(pushed, st);
          cudaStreamWaitEvent(H.m->down, pushed, 0);
          const bool lean = eb200_ctx_lean_prev(dom.ctx) != 0;
          for (int k : kSlotsOut) {
            if (!hp[k]) continue;
            if (lean && k >= 10 && k <= 15) continue; // i*_prev / dx*_prev: not stored, not read
            const size_t e = kSlotElem[k];
            if (cudaMemcpyAsync((char*)hp[k] + lo * e, (const char*)dp[k] + lo * e, n * e,
                                cudaMemcpyDeviceToHost, H.m->down) != cudaSuccess) return EB200_ERR_CUDA;
            H.down_bytes += n * e;
          }
        }
      }
      return EB200_OK;
    }

    static int PushAndDepositSpecies(Domain& dom, double time);

    int ParticlePushAndDeposit(Domain& dom, double time) {
      TRY(eb200_zero_currents(dom.ctx, dom.cur, dom.stream));
      TRY(eb200_pack_fields_hold(dom.ctx, dom.em, dom.stream));
      const int rc = dom.host ? StreamedPushAndDeposit(dom, time) : PushAndDepositSpecies(dom, time);
      const int rc2 = eb200_pack_fields_release(dom.ctx);
      return rc != EB200_OK ? rc : rc2;
    }

    static int PushAndDepositSpecies(Domain& dom, double time) {
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.pusher_flags == EB200_PUSHER_NONE || sp.npart == 0) continue;
        const eb200_pusher_t c = pusher_context(dom, sp, time);
        const int mode = dom.prm->deposit_mode == EB200_DEPOSIT_AGGREGATED ? EB200_DEPOSIT_AGGREGATED
                                                                         : EB200_DEPOSIT_ATOMIC;
        if (const EngineState::Emit* e = emission_of(dom, s)) {
          // no fused kernel carries the emission hook: the pusher with the policy, then the deposit
          TRY(PushWithEmission(dom, s, c, *e));
          if (!almost_zero(sp.charge)) {
            TRY(eb200_deposit(dom.ctx, &sp.arrays, sp.npart, sp.charge, dom.prm->dt, dom.cur, mode, dom.stream));
          }
          continue;
        }
        if (almost_zero(sp.charge)) {
          TRY(eb200_push_sr(dom.ctx, &c, &sp.arrays, sp.npart, dom.em, dom.stream));
        } else {
          TRY(eb200_push_deposit_sr(dom.ctx, &c, &sp.arrays, sp.npart, dom.em, dom.cur, mode, dom.stream));
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
        // flags: counting / radix sort by build; i*_prev / dx*_prev left out when the caller
        // declared them unobserved (eb200_set_lean_prev)
        TRY(eb200_sort_particles(dom.ctx, &sp.arrays, &n, (clear ? 1 : 0) | eb200_ctx_sort_flags(dom.ctx),
                                 dom.stream));
        sp.npart = n;
      }
      return EB200_OK;
    }

    // srpic::FieldBoundaries (fields_bcs.h:600-672) for the MATCH faces registered with the
    // context; the other kinds (CONDUCTOR inside the filter / field kernels, PERIODIC and SYNC
    // inside the exchanges) are handled where the data moves
    int FieldBoundaries(Domain& dom, int tags) {
      if (dom.curv) {
        // dir::Directions<Dim::_2D>::orth = (-x1, -x2, +x2, +x1), fields_bcs.h:619-668
        PHASE(dom, EB200_PHASE_FIELDSOLVER);
        static const int orth[4] = { 0, 2, 3, 1 }; // face = 2 * dim + (sign > 0)
        for (int q = 0; q < 4; ++q) {
          const int face = orth[q], o = face >> 1, sign = (face & 1) ? +1 : -1;
          if (dom.prm->fbc[face] == EB200_FBC_AXIS) {
            TRY(eb200_axis_fields(dom.ctx, dom.em, sign, tags, dom.stream));
            continue;
          }
          if (!dom.eng) continue;
          for (const eb200_field_bc_t& b : dom.eng->curv_bcs) {
            if (b.o != o || (b.sign > 0) != (sign > 0)) continue;
            if (b.kind == EB200_FBC_MATCH) {
              TRY(eb200_match_fields_curv(dom.ctx, dom.em, b.target, b.o, b.xg_edge, b.ds, tags, b.mask,
                                          b.range_min, b.range_max, dom.prm->fbc, dom.stream));
            } else if (b.kind == EB200_FBC_ATMOSPHERE) {
              TRY(eb200_enforce_fields(dom.ctx, dom.em, b.target, b.o, b.sign, b.i_edge, tags, b.mask,
                                       b.range_min, b.range_max, dom.stream));
            }
          }
        }
        return EB200_OK;
      }
      if (!dom.eng || dom.eng->match.empty()) return EB200_OK;
      PHASE(dom, EB200_PHASE_FIELDSOLVER);
      for (const eb200_match_face_t& f : dom.eng->match) {
        TRY(eb200_match_fields(dom.ctx, dom.em, dom.eng->match_target, f.o, f.xg_edge, f.ds, tags,
                               dom.eng->match_mask, f.range_min, f.range_max, dom.stream));
      }
      return EB200_OK;
    }

    // SRPICEngine::step_forward, srpic.hpp:65-188
    // srpic::ParticleInjector (particles_bcs.h:155-166): the atmosphere face registered with
    // the context; buff component 0 is the density plane (the reference uses bckp)
    int ParticleInjector(Domain& dom, uint32_t step) {
      EngineState* e = eb200_ctx_engine_state(dom.ctx);
      if (!e->has_atm) return EB200_OK;
      return eb200_atmosphere_particles(dom.ctx, &e->atm, dom.species, dom.nspecies, dom.buff, 0,
                                        step, dom.stream);
    }

    int step_forward(Domain& dom, uint32_t step, double time) {
      const eb200_srpic_params_t& p = *dom.prm;
      dom.step = step;
      if (step == 0) {
        {
          PHASE(dom, EB200_PHASE_COMM);
          TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 0, 6, p.fbc, dom.stream));
        }
        TRY(FieldBoundaries(dom, EB200_BC_B | EB200_BC_E));
        TRY(ParticleInjector(dom, step)); // srpic.hpp:81
      }
      if (p.fieldsolver_enabled) {
        {
          PHASE(dom, EB200_PHASE_FIELDSOLVER);
          TRY(Faraday(dom, 0.5f));
        }
        {
          PHASE(dom, EB200_PHASE_COMM);
          TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 3, 6, p.fbc, dom.stream));
        }
        TRY(FieldBoundaries(dom, EB200_BC_B));
      }
      {
        PHASE(dom, EB200_PHASE_PUSH_DEPOSIT);
        if (p.deposit_enabled && p.fuse_push_deposit && !dom.curv) {
          TRY(ParticlePushAndDeposit(dom, time));
        } else {
          TRY(ParticlePush(dom, time));
          if (p.deposit_enabled) {
            TRY(CurrentsDeposit(dom));
          }
        }
      }
      if (p.deposit_enabled) {
        {
          PHASE(dom, EB200_PHASE_COMM);
          TRY(eb200_sync_currents(dom.ctx, dom.cur, dom.buff, p.fbc, dom.stream));
          TRY(eb200_comm_fields(dom.ctx, dom.cur, 3, 0, 3, p.fbc, dom.stream));
        }
        PHASE(dom, EB200_PHASE_FILTER);
        TRY(CurrentsFilter(dom));
      }
      // CommunicateParticles (srpic.hpp:130-132): a single self-periodic domain has no
      // neighbour to migrate to; species without a pusher never carry a send tag
      if (eb200_ctx_has_comm(dom.ctx)) {
        PHASE(dom, EB200_PHASE_MIGRATION);
        TRY(eb200_comm_particles(dom.ctx, dom.species, dom.nspecies, dom.stream));
      }
      if (p.fieldsolver_enabled) {
        {
          PHASE(dom, EB200_PHASE_FIELDSOLVER);
          TRY(Faraday(dom, 0.5f));
        }
        {
          PHASE(dom, EB200_PHASE_COMM);
          TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 3, 6, p.fbc, dom.stream));
        }
        TRY(FieldBoundaries(dom, EB200_BC_B));
        {
          PHASE(dom, EB200_PHASE_FIELDSOLVER);
          TRY(Ampere(dom, 1.0f));
          if (p.deposit_enabled) {
            TRY(CurrentsAmpere(dom));
          }
        }
        {
          PHASE(dom, EB200_PHASE_COMM);
          TRY(eb200_comm_fields(dom.ctx, dom.em, 6, 0, 3, p.fbc, dom.stream));
          TRY(eb200_comm_fields(dom.ctx, dom.cur, 3, 0, 3, p.fbc, dom.stream));
        }
        TRY(FieldBoundaries(dom, EB200_BC_E));
      }
      TRY(ParticleInjector(dom, step)); // srpic.hpp:179-183
      {
        PHASE(dom, EB200_PHASE_SORT);
        TRY(SortParticles(dom, step));
      }
      return EB200_OK;
    }

  } // namespace srpic
} // namespace eb200

// ============================================================================ GRPIC
extern "C" int eb200_ctx_grid(const eb200_ctx_t* ctx, eb200_grid_t* grid, float* dx, float* xmin3);

namespace eb200 {
  namespace grpic {
    static bool almost_zero(float x) { return std::fabs(x) <= std::numeric_limits<float>::epsilon(); }

    struct Domain {
      eb200_ctx_t*                ctx;
      const eb200_grpic_params_t* prm;
      eb200_grid_t                grid;
      float *                     em, *em0, *cur, *cur0, *aux, *buff;
      const float*                match_target;
      eb200_species_t*            species;
      int                         nspecies;
      eb200_stream_t              stream;
      Profiler*                   prof;
    };

    enum class gr_getE { D0_B, D_B0 };
    enum class gr_getH { D_B0, D0_B0 };
    enum class gr_faraday { aux, main };
    enum class gr_ampere { init, aux, main };
    enum class gr_bc { main, aux, curr };

    // fieldsolvers.h:79-128
    int ComputeAuxE(Domain& dom, gr_getE g) {
      const float* D = (g == gr_getE::D0_B) ? dom.em0 : dom.em;
      const float* B = (g == gr_getE::D0_B) ? dom.em : dom.em0;
      return eb200_gr_aux_e(dom.ctx, D, B, dom.aux, dom.prm->fbc, dom.stream);
    }

    int ComputeAuxH(Domain& dom, gr_getH g) {
      const float* D = (g == gr_getH::D_B0) ? dom.em : dom.em0;
      return eb200_gr_aux_h(dom.ctx, D, dom.em0, dom.aux, dom.prm->fbc, dom.stream);
    }

    // fieldsolvers.h:130-166
    int Faraday(Domain& dom, gr_faraday g, float fraction) {
      const float dT = fraction * dom.prm->correction * dom.prm->dt;
      const float* Bin = (g == gr_faraday::aux) ? dom.em0 : dom.em;
      return eb200_faraday_gr(dom.ctx, Bin, dom.em0, dom.aux, dT, dom.prm->fbc, dom.stream);
    }

    // fieldsolvers.h:168-222
    int Ampere(Domain& dom, gr_ampere g, float fraction) {
      const float  dT   = fraction * dom.prm->correction * dom.prm->dt;
      const float* Din  = (g == gr_ampere::aux) ? dom.em0 : dom.em;
      float*       Dout = (g == gr_ampere::init) ? dom.em : dom.em0;
      return eb200_ampere_gr(dom.ctx, Din, Dout, dom.aux, dT, dom.prm->fbc, dom.stream);
    }

    // fieldsolvers.h:224-262
    int AmpereCurrents(Domain& dom, gr_ampere g) {
      const float coeff = -dom.prm->dt * dom.prm->q0 / dom.prm->B0;
      const float* J    = (g == gr_ampere::aux) ? dom.cur : dom.cur0;
      return eb200_currents_ampere_gr(dom.ctx, dom.em0, J, coeff, dom.prm->fbc, dom.stream);
    }

    // utils.h:41-61
    int TimeAverageDB(Domain& dom) {
      return eb200_time_average(dom.ctx, dom.em0, dom.em, 6, dom.stream);
    }

    int TimeAverageJ(Domain& dom) {
      return eb200_time_average(dom.ctx, dom.cur, dom.cur0, 3, dom.stream);
    }

    int CopyFields(Domain& dom) {
      size_t n = 6;
      for (int a = 0; a < dom.grid.dim; ++a) n *= (size_t)(dom.grid.n[a] + 2 * dom.grid.ng);
      return cudaMemcpyAsync(dom.em0, dom.em, n * sizeof(float), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)dom.stream) == cudaSuccess
               ? EB200_OK
               : EB200_ERR_CUDA;
    }

    void SwapFields(Domain& dom) {
      std::swap(dom.em, dom.em0);
      std::swap(dom.cur, dom.cur0);
    }

    // grpic::FieldBoundaries (fields_bcs.h:232-266): directions in the order of
    // dir::Directions<Dim::_2D>::orth = (-x1, -x2, +x2, +x1) (directions.h:190-195): the axis
    // rows are mirrored before the MATCH layer blends them. A single domain, so the local and
    // the global boundary of a face are the same thing
    int FieldBoundaries(Domain& dom, int tags, gr_bc g) {
      const eb200_grpic_params_t& p = *dom.prm;
      PHASE(dom, EB200_PHASE_FIELDSOLVER);
      static const int orth[4] = { 0, 2, 3, 1 }; // face = 2 * dim + (sign > 0)
      if (g == gr_bc::main) {
        for (int q = 0; q < 4; ++q) {
          const int face = orth[q];
          const int sign = (face & 1) ? +1 : -1;
          const int o    = face >> 1;
          const int bc   = p.fbc[face];
          if (bc == EB200_FBC_MATCH) {
            if (o != 0) return EB200_ERR_ARG; // "Invalid dimension"
            if (dom.match_target == nullptr) continue; // no init_flds: nothing to match to
            float* flds[2] = { dom.em, dom.em0 };
            for (float* f : flds) {
              TRY(eb200_match_fields_curv(dom.ctx, f, dom.match_target, 0, p.match_xg_edge,
                                          p.match_ds, tags, p.match_mask, p.match_range_min,
                                          p.match_range_max, p.fbc, dom.stream));
            }
          } else if (bc == EB200_FBC_AXIS) {
            TRY(eb200_axis_fields(dom.ctx, dom.em, sign, tags, dom.stream));
            TRY(eb200_axis_fields(dom.ctx, dom.em0, sign, tags, dom.stream));
          } else if (bc == EB200_FBC_HORIZON) {
            TRY(eb200_horizon_fields(dom.ctx, dom.em, tags, p.nfilter, dom.stream));
            TRY(eb200_horizon_fields(dom.ctx, dom.em0, tags, p.nfilter, dom.stream));
          }
        }
      } else if (g == gr_bc::aux) {
        // HorizonFieldsIn acts for gr_bc::main only (fields_bcs.h:150-167): nothing to do
      } else {
        for (int face = 0; face < 4; ++face) {
          if (p.pbc[face] != EB200_PBC_ABSORB || p.fbc[face] == EB200_FBC_HORIZON) continue;
          if (face != 1) return EB200_ERR_ARG; // "Absorption of currents only possible in +x1 (+r)"
          TRY(eb200_absorb_currents_gr(dom.ctx, dom.cur0, p.match_xg_edge, p.match_ds,
                                       p.match_range_min, p.match_range_max, dom.stream));
        }
      }
      return EB200_OK;
    }

    // particle_pusher.h:27-89
    int ParticlePush(Domain& dom) {
      const eb200_grpic_params_t& p = *dom.prm;
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.npart == 0 || sp.pusher_flags == EB200_PUSHER_NONE) continue;
        eb200_pusher_gr_t c;
        std::memset(&c, 0, sizeof(c));
        c.pusher_flags = sp.pusher_flags;
        c.mass = sp.mass, c.charge = sp.charge;
        c.dt = p.dt, c.omegaB0 = p.omegaB0;
        c.epsilon = p.pusher_eps, c.niter = p.pusher_niter;
        for (int k = 0; k < 6; ++k) c.pbc[k] = p.pbc[k];
        c.tag_outgoing = 0;
        TRY(eb200_push_gr(dom.ctx, &c, &sp.arrays, sp.npart, dom.em, dom.em0, dom.stream));
      }
      return EB200_OK;
    }

    // currents.h:29-73: into cur0
    int CurrentsDeposit(Domain& dom) {
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.npart == 0 || almost_zero(sp.charge)) continue;
        TRY(eb200_deposit(dom.ctx, &sp.arrays, sp.npart, sp.charge, dom.prm->dt, dom.cur0,
                          dom.prm->deposit_mode, dom.stream));
      }
      return EB200_OK;
    }

    int CurrentsFilter(Domain& dom) {
      return eb200_filter(dom.ctx, dom.cur0, dom.buff, dom.prm->nfilter, dom.prm->fbc, dom.stream);
    }

    int SortParticles(Domain& dom, uint32_t step) {
      const int  ci = dom.prm->clear_interval, si = dom.prm->sort_interval;
      const bool clear = (ci > 0) && (step % (uint32_t)ci == 0u) && (step > 0u);
      const bool sort  = (si > 0) && (step % (uint32_t)si == 0u);
      if (!clear && !sort) return EB200_OK;
      for (int s = 0; s < dom.nspecies; ++s) {
        eb200_species_t& sp = dom.species[s];
        if (sp.npart == 0) continue;
        uint32_t n = sp.npart;
        TRY(eb200_sort_particles(dom.ctx, &sp.arrays, &n, (clear ? 1 : 0) | eb200_ctx_sort_flags(dom.ctx),
                                 dom.stream));
        sp.npart = n;
      }
      return EB200_OK;
    }

#define FS(expr)                                                                               \
  do {                                                                                         \
    PHASE(dom, EB200_PHASE_FIELDSOLVER);                                                       \
    TRY(expr);                                                                                 \
  } while (0)

    // GRPICEngine::step_forward, grpic.hpp:66-634. A single domain has no neighbour: the
    // CommunicateFields / SynchronizeFields calls between the sub-steps have nothing to move
    // (no face of a 2D (r, theta) domain is periodic).
    int step_forward(Domain& dom, uint32_t step) {
      const eb200_grpic_params_t& p = *dom.prm;
      const int BC_DB = EB200_BC_E | EB200_BC_B;
      if (step == 0) {
        if (p.fieldsolver_enabled) {
          TRY(FieldBoundaries(dom, BC_DB, gr_bc::main));                 // :99-104
          TRY(CopyFields(dom));                                          // :112
          FS(ComputeAuxE(dom, gr_getE::D_B0));                           // :120-121
          FS(ComputeAuxH(dom, gr_getH::D_B0));
          TRY(FieldBoundaries(dom, BC_DB, gr_bc::aux));                  // :127-132
          FS(Faraday(dom, gr_faraday::aux, 0.5f));                       // :139-143
          TRY(FieldBoundaries(dom, EB200_BC_B, gr_bc::main));            // :149-154
          FS(Ampere(dom, gr_ampere::init, 0.5f));                        // :161-165
          TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::main));            // :171-176
          FS(ComputeAuxE(dom, gr_getE::D_B0));                           // :184-185
          FS(ComputeAuxH(dom, gr_getH::D_B0));
          TRY(FieldBoundaries(dom, BC_DB, gr_bc::aux));                  // :191-196
          FS(Faraday(dom, gr_faraday::main, 1.0f));                      // :205-209
          TRY(FieldBoundaries(dom, EB200_BC_B, gr_bc::main));            // :214-219
          FS(Ampere(dom, gr_ampere::aux, 1.0f));                         // :226
          TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::main));            // :231-236
          FS(ComputeAuxH(dom, gr_getH::D0_B0));                          // :243
          TRY(FieldBoundaries(dom, EB200_BC_B, gr_bc::aux));             // :248-253
          FS(Ampere(dom, gr_ampere::main, 1.0f));                        // :261
          TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::main));            // :266-271
          SwapFields(dom);                                               // :278
        } else {
          TRY(CopyFields(dom));                                          // :302
        }
      }
      if (p.fieldsolver_enabled) {
        FS(TimeAverageDB(dom));                                          // :329
        FS(ComputeAuxE(dom, gr_getE::D0_B));                             // :330
        TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::aux));
        FS(Faraday(dom, gr_faraday::aux, 1.0f));
        TRY(FieldBoundaries(dom, EB200_BC_B, gr_bc::main));
        FS(ComputeAuxH(dom, gr_getH::D_B0));
        TRY(FieldBoundaries(dom, EB200_BC_B, gr_bc::aux));
      }
      {
        {
          PHASE(dom, EB200_PHASE_PUSH_DEPOSIT);
          TRY(ParticlePush(dom));
          if (p.deposit_enabled) {
            size_t n = 3;
            for (int a = 0; a < dom.grid.dim; ++a) n *= (size_t)(dom.grid.n[a] + 2 * dom.grid.ng);
            if (cudaMemsetAsync(dom.cur0, 0, n * sizeof(float), (cudaStream_t)dom.stream) != cudaSuccess) {
              return EB200_ERR_CUDA;
            }
            TRY(CurrentsDeposit(dom));
          }
        }
        if (p.deposit_enabled) {
          TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::curr));
          PHASE(dom, EB200_PHASE_FILTER);
          TRY(CurrentsFilter(dom));
        }
      }
      if (p.fieldsolver_enabled) {
        if (p.deposit_enabled) {
          FS(TimeAverageJ(dom));
        }
        FS(ComputeAuxE(dom, gr_getE::D_B0));
        TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::aux));
        FS(Faraday(dom, gr_faraday::main, 1.0f));
        TRY(FieldBoundaries(dom, EB200_BC_B, gr_bc::main));
        FS(Ampere(dom, gr_ampere::aux, 1.0f));
        if (p.deposit_enabled) {
          FS(AmpereCurrents(dom, gr_ampere::aux));
        }
        TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::main));
        FS(ComputeAuxH(dom, gr_getH::D0_B0));
        TRY(FieldBoundaries(dom, EB200_BC_B, gr_bc::aux));
        FS(Ampere(dom, gr_ampere::main, 1.0f));
        if (p.deposit_enabled) {
          FS(AmpereCurrents(dom, gr_ampere::main));
        }
        SwapFields(dom);
        TRY(FieldBoundaries(dom, EB200_BC_E, gr_bc::main));
      }
      {
        PHASE(dom, EB200_PHASE_SORT);
        TRY(SortParticles(dom, step));
      }
      return EB200_OK;
    }
#undef FS

  } // namespace grpic
} // namespace eb200

extern "C" int eb200_grpic_step(eb200_ctx_t* ctx, const eb200_grpic_params_t* prm, float** em,
                                float** em0, float** cur, float** cur0, float* aux, float* buff,
                                const float* match_target, eb200_species_t* species, int nspecies,
                                uint32_t step, double time, eb200_stream_t stream) {
  (void)time;
  if (!ctx || !prm || !em || !em0 || !cur || !cur0 || !*em || !*em0 || !*cur || !*cur0 || !aux ||
      !buff || (nspecies > 0 && !species)) {
    return EB200_ERR_ARG;
  }
  eb200::grpic::Domain dom;
  dom.ctx = ctx;
  dom.prm = prm;
  float dx_unused, xmin_unused[3];
  int   rc = eb200_ctx_grid(ctx, &dom.grid, &dx_unused, xmin_unused);
  if (rc != EB200_OK) return rc;
  if (dom.grid.dim != 2) return EB200_ERR_ARG;
  dom.em = *em, dom.em0 = *em0, dom.cur = *cur, dom.cur0 = *cur0;
  dom.aux = aux, dom.buff = buff;
  dom.match_target = match_target;
  dom.species = species, dom.nspecies = nspecies;
  dom.stream = stream;
  dom.prof   = &eb200_ctx_engine_state(ctx)->prof;
  rc         = eb200::grpic::step_forward(dom, step);
  *em = dom.em, *em0 = dom.em0, *cur = dom.cur, *cur0 = dom.cur0;
  return rc;
}

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
  dom.prof     = &eb200_ctx_engine_state(ctx)->prof;
  dom.eng      = eb200_ctx_engine_state(ctx);
  const int metric = eb200_ctx_metric(ctx);
  dom.curv         = metric == EB200_METRIC_SPHERICAL || metric == EB200_METRIC_QSPHERICAL;
  if (metric != EB200_METRIC_MINKOWSKI && !dom.curv) return EB200_ERR_UNSUPPORTED; // GRPIC: eb200_grpic_step
  return eb200::srpic::step_forward(dom, step, time);
}

namespace eb200 {
  EngineState* engine_state_new() { return new EngineState(); }

  void engine_state_delete(EngineState* e) {
    if (!e) return;
    for (auto& r : e->prof.recs) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    for (auto ev : e->prof.pool) cudaEventDestroy(ev);
    for (auto p : e->mirror.ptrs)
      if (p) cudaFree(p);
    delete e;
  }
} // namespace eb200

extern "C" int eb200_srpic_set_match(eb200_ctx_t* ctx, const eb200_match_face_t* faces, int nfaces,
                                     const float* target, int components_mask) {
  if (!ctx || nfaces < 0 || (nfaces > 0 && (!faces || !target))) return EB200_ERR_ARG;
  eb200::EngineState* e = eb200_ctx_engine_state(ctx);
  e->match.assign(faces, faces + nfaces);
  e->match_target = nfaces > 0 ? target : nullptr;
  e->match_mask   = components_mask;
  return EB200_OK;
}

extern "C" int eb200_srpic_set_field_bcs(eb200_ctx_t* ctx, const eb200_field_bc_t* bcs, int n) {
  if (!ctx || n < 0 || (n > 0 && !bcs)) return EB200_ERR_ARG;
  for (int k = 0; k < n; ++k) {
    if ((bcs[k].kind != EB200_FBC_MATCH && bcs[k].kind != EB200_FBC_ATMOSPHERE) || !bcs[k].target ||
        (bcs[k].o != 0 && bcs[k].o != 1) || bcs[k].sign == 0) {
      return EB200_ERR_ARG;
    }
  }
  eb200_ctx_engine_state(ctx)->curv_bcs.assign(bcs, bcs + n);
  return EB200_OK;
}

extern "C" int eb200_srpic_set_ext_current(eb200_ctx_t* ctx, const eb200_ext_current_t* ext) {
  if (!ctx) return EB200_ERR_ARG;
  eb200::EngineState* e = eb200_ctx_engine_state(ctx);
  if (ext == nullptr) {
    e->has_ext = false;
    return EB200_OK;
  }
  if (ext->nmodes < 0 || ext->nmodes > EB200_MAX_MODES) return EB200_ERR_ARG;
  e->ext     = *ext;
  e->has_ext = true;
  return EB200_OK;
}

extern "C" int eb200_srpic_set_emission(eb200_ctx_t* ctx, int species, int photon_species,
                                        const eb200_emission_t* policy) {
  if (!ctx || species < 0 || species >= 64) return EB200_ERR_ARG;
  eb200::EngineState* e = eb200_ctx_engine_state(ctx);
  if ((int)e->emission.size() <= species) e->emission.resize(species + 1);
  if (policy == nullptr) {
    e->emission[species].on = false;
    return EB200_OK;
  }
  if ((policy->kind != EB200_EMISSION_SYNCHROTRON && policy->kind != EB200_EMISSION_COMPTON) ||
      photon_species < 0 || photon_species == species) {
    return EB200_ERR_ARG;
  }
  e->emission[species].on             = true;
  e->emission[species].photon_species = photon_species;
  e->emission[species].policy         = *policy;
  return EB200_OK;
}

extern "C" int eb200_srpic_set_atmosphere_injector(eb200_ctx_t* ctx, const eb200_atmosphere_t* atm) {
  if (!ctx) return EB200_ERR_ARG;
  eb200::EngineState* e = eb200_ctx_engine_state(ctx);
  if (atm == nullptr) {
    e->has_atm = false;
    return EB200_OK;
  }
  if (atm->dim < 0 || atm->dim > 2 || atm->sign == 0 || atm->density <= 0.0f || atm->height <= 0.0f) {
    return EB200_ERR_ARG;
  }
  e->atm     = *atm;
  e->has_atm = true;
  return EB200_OK;
}

extern "C" int eb200_profile_enable(eb200_ctx_t* ctx, int on) {
  if (!ctx) return EB200_ERR_ARG;
  eb200::Profiler& p = eb200_ctx_engine_state(ctx)->prof;
  p.on               = on != 0;
  for (auto& r : p.recs) {
    p.pool.push_back(r.a);
    p.pool.push_back(r.b);
  }
  p.recs.clear();
  return EB200_OK;
}

extern "C" int eb200_profile_read(eb200_ctx_t* ctx, float* ms, int* calls) {
  if (!ctx || !ms || !calls) return EB200_ERR_ARG;
  eb200::Profiler& p = eb200_ctx_engine_state(ctx)->prof;
  for (int k = 0; k < EB200_NPHASES; ++k) {
    ms[k]    = 0.0f;
    calls[k] = 0;
  }
  for (auto& r : p.recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return EB200_ERR_CUDA;
    float t = 0.0f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return EB200_ERR_CUDA;
    ms[r.phase]    += t;
    calls[r.phase] += 1;
    p.pool.push_back(r.a);
    p.pool.push_back(r.b);
  }
  p.recs.clear();
  return EB200_OK;
}

// ---- host-buffer step: device mirrors owned by the context ---------------------------------
namespace {
  struct Slot {
    void** host_field; // address of the pointer inside the struct
    size_t elem;
  };

  void* mirror_get(eb200::HostMirror& m, size_t slot, size_t bytes) {
    if (m.ptrs.size() <= slot) {
      m.ptrs.resize(slot + 1, nullptr);
      m.sizes.resize(slot + 1, 0);
    }
    if (m.sizes[slot] < bytes) {
      if (m.ptrs[slot]) cudaFree(m.ptrs[slot]);
      m.ptrs[slot]  = nullptr;
      m.sizes[slot] = 0;
      if (cudaMalloc(&m.ptrs[slot], bytes) != cudaSuccess) return nullptr;
      m.sizes[slot] = bytes;
    }
    return m.ptrs[slot];
  }
} // namespace

extern "C" int eb200_srpic_step_host(eb200_ctx_t* ctx, const eb200_srpic_params_t* prm,
                                     float* em_host, float* cur_host,
                                     eb200_species_t* species_host, int nspecies, uint32_t step,
                                     double time, uint64_t* bytes_h2d, uint64_t* bytes_d2h) {
  if (!ctx || !prm || !em_host || !cur_host || (nspecies > 0 && !species_host)) return EB200_ERR_ARG;
  eb200_grid_t g;
  float        dx, xmin[3];
  if (eb200_ctx_grid(ctx, &g, &dx, xmin) != EB200_OK) return EB200_ERR_ARG;
  eb200::HostMirror& m = eb200_ctx_engine_state(ctx)->mirror;
  size_t cells = 1;
  for (int a = 0; a < g.dim; ++a) cells *= (size_t)(g.n[a] + 2 * g.ng);
  const size_t b6 = cells * 6 * sizeof(float), b3 = cells * 3 * sizeof(float);
  cudaStream_t st = 0;
  uint64_t     up = 0, down = 0;
  float*       d_em   = (float*)mirror_get(m, 0, b6);
  float*       d_cur  = (float*)mirror_get(m, 1, b3);
  float*       d_buff = (float*)mirror_get(m, 2, b3);
  if (!d_em || !d_cur || !d_buff) return EB200_ERR_CUDA;
  // Streamed variant: the fused Minkowski step of a single domain, on steps that neither sort
  // nor compact (those permute every array, weight included: they take the plain path below)
  {
    const int  ci = prm->clear_interval, si = prm->sort_interval;
    const bool reorders = ((ci > 0) && (step % (uint32_t)ci == 0u) && (step > 0u)) ||
                          ((si > 0) && (step % (uint32_t)si == 0u));
    static const bool disabled = getenv("EB200_HOST_STEP_PLAIN") != nullptr;
    if (!disabled && !reorders && prm->fuse_push_deposit && prm->deposit_enabled &&
        !eb200_ctx_has_comm(ctx) && m.streams()) {
      m.next_event = 0;
      std::vector<eb200_species_t> dev(species_host, species_host + nspecies);
      for (int s = 0; s < nspecies; ++s) {
        void** hp = (void**)&species_host[s].arrays;
        void** dp = (void**)&dev[s].arrays;
        for (int k = 0; k < 20; ++k) dp[k] = nullptr;
        for (int k = 0; k < 17; ++k) {
          if (!hp[k]) continue;
          dp[k] = mirror_get(m, 3 + (size_t)s * 17 + k, (size_t)species_host[s].maxnpart * eb200::srpic::kSlotElem[k]);
          if (!dp[k]) return EB200_ERR_CUDA;
        }
      }
      eb200::HostStreamer H { &m, species_host, size_t(8) << 20 };
      if (const char* e = getenv("EB200_HOST_CHUNK")) H.chunk = std::max<size_t>(1024, (size_t)atol(e) / 1024 * 1024);
      cudaMemcpyAsync(d_em, em_host, b6, cudaMemcpyHostToDevice, m.main);
      up += b6; // J is zeroed by the deposit before anything reads it: not uploaded
      eb200::srpic::Domain dom;
      dom.ctx = ctx;
      dom.prm = prm;
      if (eb200_ctx_grid(ctx, &dom.grid, &dom.dx, dom.xmin) != EB200_OK) return EB200_ERR_ARG;
      dom.em       = d_em;
      dom.cur      = d_cur;
      dom.buff     = d_buff;
      dom.species  = dev.data();
      dom.nspecies = nspecies;
      dom.stream   = m.main;
      dom.prof     = &eb200_ctx_engine_state(ctx)->prof;
      dom.eng      = eb200_ctx_engine_state(ctx);
      dom.host     = &H;
      int rc = eb200::srpic::step_forward(dom, step, time);
      if (rc == EB200_OK) {
        cudaMemcpyAsync(em_host, d_em, b6, cudaMemcpyDeviceToHost, m.main);
        cudaMemcpyAsync(cur_host, d_cur, b3, cudaMemcpyDeviceToHost, m.main);
        down += b6 + b3;
      }
      cudaError_t e1 = cudaStreamSynchronize(m.up), e2 = cudaStreamSynchronize(m.main),
                  e3 = cudaStreamSynchronize(m.down);
      if (rc != EB200_OK) return rc;
      if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return EB200_ERR_CUDA;
      for (int s = 0; s < nspecies; ++s) species_host[s].npart = dev[s].npart;
      if (bytes_h2d) *bytes_h2d = up + H.up_bytes;
      if (bytes_d2h) *bytes_d2h = down + H.down_bytes;
      return EB200_OK;
    }
  }
  cudaMemcpyAsync(d_em, em_host, b6, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_cur, cur_host, b3, cudaMemcpyHostToDevice, st);
  up += b6 + b3;
  std::vector<eb200_species_t> dev(species_host, species_host + nspecies);
  // per species 17 candidate arrays; element sizes in ParticleArrays order
  static const size_t esz[17] = { 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2 };
  const bool          lean_prev = eb200_ctx_lean_prev(ctx) != 0;
  for (int s = 0; s < nspecies; ++s) {
    void** hp = (void**)&species_host[s].arrays;
    void** dp = (void**)&dev[s].arrays;
    for (int k = 0; k < 20; ++k) dp[k] = nullptr;
    const size_t n = species_host[s].npart, cap = species_host[s].maxnpart;
    for (int k = 0; k < 17; ++k) {
      if (!hp[k]) continue;
      void* d = mirror_get(m, 3 + (size_t)s * 17 + k, cap * esz[k]);
      if (!d) return EB200_ERR_CUDA;
      dp[k] = d;
      if (lean_prev && k >= 10 && k <= 15) continue; // dead values on entry and on exit
      cudaMemcpyAsync(d, hp[k], n * esz[k], cudaMemcpyHostToDevice, st);
      up += n * esz[k];
    }
  }
  int rc = eb200_srpic_step(ctx, prm, d_em, d_cur, d_buff, dev.data(), nspecies, step, time, st);
  if (rc != EB200_OK) return rc;
  cudaMemcpyAsync(em_host, d_em, b6, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(cur_host, d_cur, b3, cudaMemcpyDeviceToHost, st);
  down += b6 + b3;
  for (int s = 0; s < nspecies; ++s) {
    species_host[s].npart = dev[s].npart;
    void** hp = (void**)&species_host[s].arrays;
    void** dp = (void**)&dev[s].arrays;
    const size_t n = dev[s].npart;
    for (int k = 0; k < 17; ++k) {
      if (!hp[k]) continue;
      if (lean_prev && k >= 10 && k <= 15) continue;
      cudaMemcpyAsync(hp[k], dp[k], n * esz[k], cudaMemcpyDeviceToHost, st);
      down += n * esz[k];
    }
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) return EB200_ERR_CUDA;
  if (bytes_h2d) *bytes_h2d = up;
  if (bytes_d2h) *bytes_d2h = down;
  return EB200_OK;
}
