#!/usr/bin/env python
"""Inserts the calls of integration/eb200_shim.hpp into a SCRATCH COPY of the reference (argument:
its root). Touches three files of the copy, each patch anchored on the Kokkos launch it replaces
and guarded by EB200_SHIM and the Minkowski branch:

  src/engines/srpic/fieldsolvers.h    srpic::Faraday / Ampere / CurrentsAmpere -> eb200_faraday / eb200_ampere /
                                      eb200_currents_ampere
  src/engines/srpic/currents.h        srpic::CurrentsDeposit -> eb200_zero_currents + eb200_deposit per species;
                                      srpic::CurrentsFilter -> eb200_filter (single domain)
  src/engines/srpic/particle_pusher.h srpic::ParticlePush -> eb200_push_deposit_sr (fused, charged species) or
                                      eb200_push_sr per species (species without emission policy; pgens
                                      without per-particle functors)
  src/framework/domain/metadomain_sort.cpp  Particles::SortSpatially -> eb200_sort_particles

Never run against /root/reference (refused)."""
import os
import sys


def patch(path, anchor, insert, where="before", count=1):
    s = open(path).read()
    assert s.count(anchor) >= count, f"{path}: anchor not found: {anchor[:60]!r}"
    if insert.strip() in s:
        return
    if where == "before":
        s = s.replace(anchor, insert + anchor, count)
    else:
        s = s.replace(anchor, anchor + insert, count)
    open(path, "w").write(s)


def main(root):
    root = os.path.abspath(root)
    assert not root.startswith("/root/reference"), "refusing to patch the reference tree"
    fs = os.path.join(root, "src/engines/srpic/fieldsolvers.h")
    cu = os.path.join(root, "src/engines/srpic/currents.h")
    pp = os.path.join(root, "src/engines/srpic/particle_pusher.h")
    here = os.path.dirname(os.path.abspath(__file__))
    # the scratch tree exists only for this build: the switch is defined in the patched headers, so
    # that no compile flag changes (the Kokkos objects of the plain CUDA build are reused as they are)
    # start from the pristine files every time (the patches are not cumulative)
    ref = "/root/reference"
    for f in (fs, cu, pp, os.path.join(root, "src/framework/domain/metadomain_sort.cpp")):
        src = os.path.join(ref, os.path.relpath(f, root))
        if os.path.exists(src):
            with open(src) as a, open(f, "w") as b:
                b.write(a.read())
    inc = f'#ifndef EB200_SHIM\n#define EB200_SHIM 1\n#endif\n#include "{here}/eb200_shim.hpp"\n'
    for f in (fs, cu, pp):
        patch(f, "namespace ntt {", inc)

    patch(fs, '        Kokkos::parallel_for("Faraday",\n                             domain.mesh.rangeActiveCells(),\n'
              '                             kernel::mink::Faraday_kernel<M::Dim>(',
          '''#ifdef EB200_SHIM
        {
          const float st[9] = { (float)deltax, (float)deltay, (float)betaxy, (float)betayx, (float)deltaz,
                                (float)betaxz, (float)betazx, (float)betayz, (float)betazy };
          auto* c = eb200shim::ctx(domain);
          eb200shim::check(c, eb200_faraday(c, domain.fields.em.data(), coeff1, coeff2, st, eb200shim::stream()),
                           "eb200_faraday");
          return;
        }
#endif
''')
    patch(fs, '        Kokkos::parallel_for(\n          "Ampere",\n          range,\n'
              '          kernel::mink::Ampere_kernel<M::Dim>(',
          '''#ifdef EB200_SHIM
        {
          auto* c = eb200shim::ctx(domain);
          eb200shim::check(c, eb200_ampere(c, domain.fields.em.data(), coeff1, coeff2, eb200shim::stream()),
                           "eb200_ampere");
          return;
        }
#endif
''')
    patch(fs, '          Kokkos::parallel_for(\n            "Ampere",\n            domain.mesh.rangeActiveCells(),\n'
              '            kernel::mink::CurrentsAmpere_kernel<M::Dim>(domain.fields.em,',
          '''#ifdef EB200_SHIM
          {
            auto* c = eb200shim::ctx(domain);
            eb200shim::check(c,
                             eb200_currents_ampere(c, domain.fields.em.data(), domain.fields.cur.data(), coeff, ppc0,
                                                   eb200shim::stream()),
                             "eb200_currents_ampere");
            return;
          }
#endif
''')
    patch(cu, "      Kokkos::deep_copy(domain.fields.cur, ZERO);\n      auto scatter_cur",
          '''#ifdef EB200_SHIM
      if constexpr (M::CoordType == Coord::Cartesian) {
        if (eb200shim::state().fused) {
          eb200shim::state().fused = false; // deposited by the fused pass of srpic::ParticlePush
          return;
        }
        auto* c = eb200shim::ctx(domain);
        eb200shim::check(c, eb200_zero_currents(c, domain.fields.cur.data(), eb200shim::stream()), "zero J");
        for (auto& species : domain.species) {
          if ((species.pusher() == ParticlePusher::NONE) or (species.npart() == 0) or
              cmp::AlmostZero_host(species.charge())) {
            continue;
          }
          const eb200_prtls_t p = eb200shim::prtls(species);
          eb200shim::check(c,
                           eb200_deposit(c, &p, (uint32_t)species.npart(), (float)species.charge(), (float)dt,
                                         domain.fields.cur.data(), EB200_DEPOSIT_AGGREGATED, eb200shim::stream()),
                           "eb200_deposit");
        }
        return;
      }
#endif
''')
    patch(pp, "        auto pusher_boundaries = kernel::sr::PusherBoundaries<M::Dim> {",
          '''#ifdef EB200_SHIM
        if constexpr (M::CoordType == Coord::Cartesian) {
          if (species.emission_policy_flag() == EmissionType::NONE) {
            auto*          c = eb200shim::ctx(domain);
            eb200_pusher_t q {};
            q.pusher_flags = (int)species.pusher();
            q.drag_flags   = (int)species.radiative_drag_flags();
            q.mass = species.mass(), q.charge = species.charge();
            q.time = (double)time, q.dt = (float)dt;
            q.omegaB0 = (float)pusher_ctx.omegaB0;
            if (species.pusher() & ParticlePusher::GCA) {
              q.gca_larmor_max      = params.template get<real_t>("algorithms.gca.larmor_max");
              q.gca_e_ovr_b_sqr_max = SQR(params.template get<real_t>("algorithms.gca.e_ovr_b_max"));
            }
            if (species.radiative_drag_flags() & RadiativeDrag::SYNCHROTRON) {
              q.sync_coeff = pusher_ctx.synchrotron_drag.coeff;
            }
            if (species.radiative_drag_flags() & RadiativeDrag::COMPTON) {
              q.compton_coeff = pusher_ctx.compton_drag.coeff;
            }
            q.has_atmosphere = has_atmosphere ? 1 : 0;
            q.atm_gx1 = gx1, q.atm_gx2 = gx2, q.atm_gx3 = gx3, q.atm_x_surf = x_surf, q.atm_ds = ds;
            eb200shim::particle_bcs(domain.mesh.prtl_bc(), (int)M::Dim, q.pbc);
            q.tag_outgoing = 0;
            q.dx           = (float)math::sqrt(domain.mesh.metric.template h_<1, 1>({}));
            const auto ext = domain.mesh.extent();
            for (int a = 0; a < (int)M::Dim; ++a) q.xmin[a] = (float)ext[a].first;
            const eb200_prtls_t p = eb200shim::prtls(species);
            // ParticlePush and CurrentsDeposit are adjacent in SRPICEngine::step_forward (srpic.hpp:103-121):
            // one fused pass per charged species; srpic::CurrentsDeposit then finds its work done
            const bool deposit = params.template get<bool>("algorithms.deposit.enable") and
                                 not cmp::AlmostZero_host(species.charge());
            if (deposit) {
              if (not eb200shim::state().fused) {
                eb200shim::check(c, eb200_zero_currents(c, domain.fields.cur.data(), eb200shim::stream()), "zero J");
                eb200shim::state().fused = true;
              }
              eb200shim::check(c,
                               eb200_push_deposit_sr(c, &q, &p, (uint32_t)species.npart(), domain.fields.em.data(),
                                                     domain.fields.cur.data(), EB200_DEPOSIT_AGGREGATED,
                                                     eb200shim::stream()),
                               "eb200_push_deposit_sr");
            } else {
              eb200shim::check(c,
                               eb200_push_sr(c, &q, &p, (uint32_t)species.npart(), domain.fields.em.data(),
                                             eb200shim::stream()),
                               "eb200_push_sr");
            }
            continue;
          }
        }
#endif
''')
    patch(cu, "      // !TODO: this needs to be done more efficiently\n      for (auto i { 0u }; i < nfilter; ++i) {",
          '''#ifdef EB200_SHIM
      if constexpr (M::CoordType == Coord::Cartesian) {
        if (metadomain.ndomains() == 1u and eb200shim::filter_enabled()) {
          // all passes in one call: temporally blocked sweeps, the periodic ghost fill included
          // (what the loop below does with deep_copy + kernel + CommunicateFields(Comm::J) per pass)
          auto* c = eb200shim::ctx(domain);
          int   fbc[6];
          eb200shim::field_bcs(domain.mesh.flds_bc(), (int)M::Dim, fbc);
          eb200shim::check(c,
                           eb200_filter(c, domain.fields.cur.data(), domain.fields.buff.data(), (int)nfilter, fbc,
                                        eb200shim::stream()),
                           "eb200_filter");
          return;
        }
      }
#endif
''')
    so = os.path.join(root, "src/framework/domain/metadomain_sort.cpp")
    patch(so, "namespace ntt {", inc)
    patch(so, "        species.SortSpatially(domain.mesh);",
          '''#ifdef EB200_SHIM
        if constexpr (M::CoordType == Coord::Cartesian) {
          // cell order with i1 fastest (the field layout of a LayoutLeft build); stable, all arrays
          auto*         c = eb200shim::ctx(domain);
          eb200_prtls_t p = eb200shim::prtls(species);
          p.pld_r = species.pld_r.data(), p.npld_r = (int)species.npld_r();
          p.pld_i = (uint32_t*)species.pld_i.data(), p.npld_i = (int)species.npld_i();
          p.pld_stride = (uint32_t)species.maxnpart();
          uint32_t n = (uint32_t)species.npart();
          eb200shim::check(c, eb200_sort_particles(c, &p, &n, 0, eb200shim::stream()), "eb200_sort_particles");
          continue;
        }
#endif
''')
    print("patched", root)


if __name__ == "__main__":
    main(sys.argv[1])
