// TEST INFRASTRUCTURE (oracle): wrapper problem generator that runs one of the reference's own
// pgens unmodified and dumps the raw state (fields + particle SoA) after chosen steps.
//
// Built by oracle/build_entity_xc.sh into the reference's own entity.xc (cmake -D pgen=<this dir>);
// the hook is the engine's CustomPostStep call (/root/reference/src/engines/engine.hpp:272-279).
// Nothing here is product code; nothing under entity_b200/ uses it.
//
//   EB_DUMP_DIR    output directory (default: no dump)
//   EB_DUMP_STEPS  comma-separated step indices; the dump for index s is the state AFTER
//                  step_forward of step s (engine.hpp:268-279). Default "0".
//
// File format: <dir>/s<step>.bin = sequence of records
//   [u32 name_len][name][u32 dtype (0=f32,1=i32,2=i16,3=f64,4=u32)][u32 ndim][u64 shape[ndim]][data]
// Field arrays are written with i1 FASTEST and the component SLOWEST (component planes),
// independent of the Kokkos layout of the build.
#ifndef EB_DUMP_COMMON_HPP
#define EB_DUMP_COMMON_HPP

#include "enums.h"
#include "global.h"

#include "arch/kokkos_aliases.h"
#include "traits/pgen.h"
#include "utils/error.h"
#include "utils/numeric.h"

#include "archetypes/utils.h"
#include "framework/domain/domain.h"
#include "framework/domain/metadomain.h"
#include "framework/parameters/parameters.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

// the reference's own problem generator, with its class renamed
#define PGen RefPGen
#include EB_REF_PGEN
#undef PGen

namespace ebdump {

  inline void rec(std::FILE* f, const std::string& name, std::uint32_t dtype,
                  const std::vector<std::uint64_t>& shape, const void* data,
                  std::size_t bytes) {
    const std::uint32_t nl = (std::uint32_t)name.size(), nd = (std::uint32_t)shape.size();
    std::fwrite(&nl, 4, 1, f);
    std::fwrite(name.data(), 1, nl, f);
    std::fwrite(&dtype, 4, 1, f);
    std::fwrite(&nd, 4, 1, f);
    std::fwrite(shape.data(), 8, nd, f);
    std::fwrite(data, 1, bytes, f);
  }

  template <class T>
  constexpr std::uint32_t code() {
    if constexpr (std::is_same_v<T, float>) return 0;
    else if constexpr (std::is_same_v<T, int>) return 1;
    else if constexpr (std::is_same_v<T, short>) return 2;
    else if constexpr (std::is_same_v<T, double>) return 3;
    else return 4;
  }

  // N-component field view -> component planes, i1 fastest
  template <class View>
  void field(std::FILE* f, const std::string& name, const View& v) {
    auto h = Kokkos::create_mirror_view(v);
    Kokkos::deep_copy(h, v);
    using T = typename View::non_const_value_type;
    constexpr int R = View::rank;
    std::vector<std::uint64_t> shp;
    std::size_t n = 1;
    // numpy shape (C order): [comp][i3][i2][i1]
    for (int d = R - 1; d >= 0; --d) { shp.push_back(v.extent(d)); n *= v.extent(d); }
    std::vector<T> buf(n);
    std::size_t k = 0;
    if constexpr (R == 2) {
      for (std::size_t c = 0; c < v.extent(1); ++c)
        for (std::size_t i = 0; i < v.extent(0); ++i) buf[k++] = h(i, c);
    } else if constexpr (R == 3) {
      for (std::size_t c = 0; c < v.extent(2); ++c)
        for (std::size_t j = 0; j < v.extent(1); ++j)
          for (std::size_t i = 0; i < v.extent(0); ++i) buf[k++] = h(i, j, c);
    } else {
      for (std::size_t c = 0; c < v.extent(3); ++c)
        for (std::size_t l = 0; l < v.extent(2); ++l)
          for (std::size_t j = 0; j < v.extent(1); ++j)
            for (std::size_t i = 0; i < v.extent(0); ++i) buf[k++] = h(i, j, l, c);
    }
    rec(f, name, code<T>(), shp, buf.data(), n * sizeof(T));
  }

  template <class View>
  void arr(std::FILE* f, const std::string& name, const View& v, std::size_t n) {
    using T = typename View::non_const_value_type;
    if (v.extent(0) < n) n = v.extent(0);
    auto h = Kokkos::create_mirror_view(v);
    Kokkos::deep_copy(h, v);
    std::vector<T> buf(n);
    for (std::size_t p = 0; p < n; ++p) buf[p] = h(p);
    rec(f, name, code<T>(), { (std::uint64_t)n }, buf.data(), n * sizeof(T));
  }

  inline auto steps() -> const std::set<long>& {
    static std::set<long> s;
    static bool init = false;
    if (!init) {
      init = true;
      const char* e = std::getenv("EB_DUMP_STEPS");
      std::string str = e ? e : "0";
      std::size_t pos = 0;
      while (pos < str.size()) {
        std::size_t q = str.find(',', pos);
        if (q == std::string::npos) q = str.size();
        if (q > pos) s.insert(std::atol(str.substr(pos, q - pos).c_str()));
        pos = q + 1;
      }
    }
    return s;
  }

  // npart of every species before the reference pgen's own CustomPostStep ran (its injectors
  // append particles: [npart_pre, npart) of a dump are the ones injected after the step)
  inline auto npart_pre() -> std::vector<std::uint32_t>& {
    static std::vector<std::uint32_t> v;
    return v;
  }

  template <ntt::SimEngine::type S, class M>
  void dump(long step, double time, ntt::Domain<S, M>& dom) {
    const char* dir = std::getenv("EB_DUMP_DIR");
    if (!dir || !steps().count(step)) return;
    Kokkos::fence();
    std::string fn = std::string(dir) + "/s" + std::to_string(step) + "_d" +
                     std::to_string(dom.index()) + ".bin";
    std::FILE* f = std::fopen(fn.c_str(), "wb");
    if (!f) { std::fprintf(stderr, "ebdump: cannot open %s\n", fn.c_str()); return; }
    rec(f, "time", 3, { 1 }, &time, 8);
    field(f, "em", dom.fields.em);
    field(f, "cur", dom.fields.cur);
    if constexpr (S == ntt::SimEngine::GRPIC) {
      field(f, "em0", dom.fields.em0);
      field(f, "cur0", dom.fields.cur0);
      field(f, "aux", dom.fields.aux);
    }
    int s = 0;
    for (auto& sp : dom.species) {
      const std::string p = "sp" + std::to_string(s++) + "_";
      const std::size_t n = sp.npart();
      const std::uint32_t np = (std::uint32_t)n;
      rec(f, p + "npart", 4, { 1 }, &np, 4);
      const std::uint32_t npre = ((std::size_t)(s - 1) < npart_pre().size()) ? npart_pre()[s - 1] : np;
      rec(f, p + "npart_pre", 4, { 1 }, &npre, 4);
      const float mq[2] = { sp.mass(), sp.charge() };
      rec(f, p + "mass_charge", 0, { 2 }, mq, 8);
      arr(f, p + "i1", sp.i1, n); arr(f, p + "i2", sp.i2, n); arr(f, p + "i3", sp.i3, n);
      arr(f, p + "dx1", sp.dx1, n); arr(f, p + "dx2", sp.dx2, n); arr(f, p + "dx3", sp.dx3, n);
      arr(f, p + "ux1", sp.ux1, n); arr(f, p + "ux2", sp.ux2, n); arr(f, p + "ux3", sp.ux3, n);
      arr(f, p + "weight", sp.weight, n);
      arr(f, p + "i1_prev", sp.i1_prev, n); arr(f, p + "i2_prev", sp.i2_prev, n);
      arr(f, p + "i3_prev", sp.i3_prev, n);
      arr(f, p + "dx1_prev", sp.dx1_prev, n); arr(f, p + "dx2_prev", sp.dx2_prev, n);
      arr(f, p + "dx3_prev", sp.dx3_prev, n);
      arr(f, p + "tag", sp.tag, n);
      arr(f, p + "phi", sp.phi, n);
    }
    std::fclose(f);
  }

  // the turbulence antenna (pgens/turbulence/pgen.hpp:139-298): wave vectors and the complex
  // amplitudes as they stand AFTER the pgen's own CustomPostStep, i.e. what the next step's
  // CurrentsAmpere adds as ext_current
  template <class PG>
  void antenna(long step, PG& pg) {
    if constexpr (requires { pg.ext_current.k; pg.ext_current.a_real; pg.ext_current.a_imag_inv; }) {
      const char* dir = std::getenv("EB_DUMP_DIR");
      if (!dir || !steps().count(step)) return;
      std::string fn = std::string(dir) + "/s" + std::to_string(step) + "_ant.bin";
      std::FILE* f = std::fopen(fn.c_str(), "wb");
      if (!f) return;
      auto& ec = pg.ext_current;
      auto  kh = Kokkos::create_mirror_view(ec.k);
      Kokkos::deep_copy(kh, ec.k);
      const std::size_t nd = ec.k.extent(0), nm = ec.k.extent(1);
      std::vector<float> kb(nd * nm);
      for (std::size_t d = 0; d < nd; ++d)
        for (std::size_t m = 0; m < nm; ++m) kb[d * nm + m] = kh(d, m);
      rec(f, "k", 0, { (std::uint64_t)nd, (std::uint64_t)nm }, kb.data(), kb.size() * 4);
      arr(f, "a_real", ec.a_real, nm);
      arr(f, "a_imag", ec.a_imag, nm);
      arr(f, "a_real_inv", ec.a_real_inv, nm);
      arr(f, "a_imag_inv", ec.a_imag_inv, nm);
      std::fclose(f);
    }
  }

  // What a field-setter functor of the pgen (init_flds for GRPIC MATCH, MatchFields / AtmFields
  // for SRPIC) evaluates to on every component's own node of the ghost-inclusive 2D mesh, in
  // the basis the boundary kernels blend with (fields_bcs.hpp:176-340, 975-1060): SRPIC =
  // transform<c, T, U>(node, f(x_Ph)), GRPIC = f(x_Ph) as is. Components the functor does not
  // define stay 0; `mask` bit c says which are defined. This is the table a host hands to the
  // C ABI in place of the functor.
  template <ntt::SimEngine::type S, class M, class FS>
  void target2d(std::FILE* f, const std::string& name, const FS& fs, ntt::Domain<S, M>& dom) {
    if constexpr (M::Dim == Dim::_2D) {
      const auto&       metric = dom.mesh.metric;
      const std::size_t N1 = dom.fields.em.extent(0), N2 = dom.fields.em.extent(1);
      std::vector<float> buf(6 * N1 * N2, 0.0f);
      std::uint32_t      mask = 0;
      auto put = [&](int c, std::size_t i, std::size_t j, real_t v) { buf[(c * N2 + j) * N1 + i] = v; };
      for (std::size_t j = 0; j < N2; ++j) {
        for (std::size_t i = 0; i < N1; ++i) {
          const real_t i1_ = static_cast<real_t>(static_cast<int>(i) - static_cast<int>(N_GHOSTS));
          const real_t i2_ = static_cast<real_t>(static_cast<int>(j) - static_cast<int>(N_GHOSTS));
          coord_t<Dim::_2D> x00 { ZERO }, x0H { ZERO }, xH0 { ZERO }, xHH { ZERO };
          metric.template convert<Crd::Cd, Crd::Ph>({ i1_, i2_ }, x00);
          metric.template convert<Crd::Cd, Crd::Ph>({ i1_, i2_ + HALF }, x0H);
          metric.template convert<Crd::Cd, Crd::Ph>({ i1_ + HALF, i2_ }, xH0);
          metric.template convert<Crd::Cd, Crd::Ph>({ i1_ + HALF, i2_ + HALF }, xHH);
          if constexpr (S == ntt::SimEngine::SRPIC) {
            if constexpr (requires { fs.ex1(xH0); }) {
              mask |= 1u;
              put(0, i, j, metric.template transform<1, Idx::T, Idx::U>({ i1_ + HALF, i2_ }, fs.ex1(xH0)));
            }
            if constexpr (requires { fs.ex2(x0H); }) {
              mask |= 2u;
              put(1, i, j, metric.template transform<2, Idx::T, Idx::U>({ i1_, i2_ + HALF }, fs.ex2(x0H)));
            }
            if constexpr (requires { fs.ex3(x00); }) {
              mask |= 4u;
              put(2, i, j, metric.template transform<3, Idx::T, Idx::U>({ i1_, i2_ }, fs.ex3(x00)));
            }
            if constexpr (requires { fs.bx1(x0H); }) {
              mask |= 8u;
              put(3, i, j, metric.template transform<1, Idx::T, Idx::U>({ i1_, i2_ + HALF }, fs.bx1(x0H)));
            }
            if constexpr (requires { fs.bx2(xH0); }) {
              mask |= 16u;
              put(4, i, j, metric.template transform<2, Idx::T, Idx::U>({ i1_ + HALF, i2_ }, fs.bx2(xH0)));
            }
            if constexpr (requires { fs.bx3(xHH); }) {
              mask |= 32u;
              put(5, i, j, metric.template transform<3, Idx::T, Idx::U>({ i1_ + HALF, i2_ + HALF }, fs.bx3(xHH)));
            }
          } else {
            if constexpr (requires { fs.dx1(xH0); }) { mask |= 1u; put(0, i, j, fs.dx1(xH0)); }
            if constexpr (requires { fs.dx2(x0H); }) { mask |= 2u; put(1, i, j, fs.dx2(x0H)); }
            if constexpr (requires { fs.dx3(x00); }) { mask |= 4u; put(2, i, j, fs.dx3(x00)); }
            if constexpr (requires { fs.bx1(x0H); }) { mask |= 8u; put(3, i, j, fs.bx1(x0H)); }
            if constexpr (requires { fs.bx2(xH0); }) { mask |= 16u; put(4, i, j, fs.bx2(xH0)); }
            if constexpr (requires { fs.bx3(xHH); }) { mask |= 32u; put(5, i, j, fs.bx3(xHH)); }
          }
        }
      }
      rec(f, name, 0, { 6, (std::uint64_t)N2, (std::uint64_t)N1 }, buf.data(), buf.size() * 4);
      rec(f, name + "_mask", 4, { 1 }, &mask, 4);
    }
  }

  template <ntt::SimEngine::type S, class M, class PG>
  void targets(long step, double time, PG& pg, ntt::Domain<S, M>& dom) {
    const char* dir = std::getenv("EB_DUMP_DIR");
    if (!dir || !steps().count(step)) return;
    if constexpr (M::Dim == Dim::_2D && M::CoordType != Coord::Cartesian) {
      std::string fn = std::string(dir) + "/s" + std::to_string(step) + "_tgt.bin";
      std::FILE* f = std::fopen(fn.c_str(), "wb");
      if (!f) return;
      if constexpr (S == ntt::SimEngine::GRPIC) {
        if constexpr (requires { pg.init_flds; }) target2d<S, M>(f, "init_flds", pg.init_flds, dom);
      } else {
        if constexpr (requires { pg.MatchFields(time); }) target2d<S, M>(f, "match", pg.MatchFields(time), dom);
        if constexpr (requires { pg.AtmFields(time); }) target2d<S, M>(f, "atm", pg.AtmFields(time), dom);
      }
      std::fclose(f);
    }
  }

  // the derived scalars the engines read from SimulationParams (parameters.cpp:47-78,
  // algorithms.cpp:18-30): what a host passes by value across the C ABI
  inline void scales(long step, const ntt::SimulationParams& params) {
    const char* dir = std::getenv("EB_DUMP_DIR");
    if (!dir || !steps().count(step)) return;
    std::string fn = std::string(dir) + "/s" + std::to_string(step) + "_scl.bin";
    std::FILE* f = std::fopen(fn.c_str(), "wb");
    if (!f) return;
    const char* names[] = { "algorithms.timestep.dt", "algorithms.timestep.correction",
                            "scales.q0", "scales.B0", "scales.omegaB0", "scales.V0", "scales.n0",
                            "scales.sigma0", "scales.larmor0", "scales.skindepth0", "scales.dx0",
                            "particles.ppc0", "grid.boundaries.atmosphere.g",
                            "grid.boundaries.atmosphere.ds", "grid.boundaries.atmosphere.height",
                            "grid.boundaries.atmosphere.temperature",
                            "grid.boundaries.atmosphere.density", "algorithms.gca.larmor_max",
                            "algorithms.gca.e_ovr_b_max" };
    for (const char* nm : names) {
      if (!params.contains(nm)) continue;
      const float v = (float)params.template get<real_t>(nm);
      rec(f, nm, 0, { 1 }, &v, 4);
    }
    std::fclose(f);
  }

  // EB_COUNT_FILE: one line per step, "step npart_0 npart_1 ..." (host-side counters only;
  // the reference's own stats writer needs -D output=ON)
  template <ntt::SimEngine::type S, class M>
  void counts(long step, ntt::Domain<S, M>& dom) {
    const char* fn = std::getenv("EB_COUNT_FILE");
    if (!fn) return;
    std::FILE* f = std::fopen(fn, "a");
    if (!f) return;
    std::fprintf(f, "%ld", step);
    for (auto& sp : dom.species) std::fprintf(f, " %lu", (unsigned long)sp.npart());
    std::fprintf(f, "\n");
    std::fclose(f);
  }

} // namespace ebdump

namespace user {
  using namespace ntt;

  template <SimEngine::type S, class M>
  struct PGen : public RefPGen<S, M> {
    using base_t = RefPGen<S, M>;

    // the reference's pgens differ in the constness of their constructor arguments
    template <class P, class MD>
    PGen(P&& p, MD&& m) : base_t { std::forward<P>(p), std::forward<MD>(m) }, eb_params { &p } {}

    const SimulationParams* eb_params;

    void CustomPostStep(timestep_t step, simtime_t time, Domain<S, M>& dom) {
      ebdump::npart_pre().clear();
      for (auto& sp : dom.species) ebdump::npart_pre().push_back((std::uint32_t)sp.npart());
      if constexpr (::traits::pgen::HasCustomPostStep<base_t, Domain<S, M>>) {
        base_t::CustomPostStep(step, time, dom);
      }
      ebdump::scales((long)step, *eb_params);
      ebdump::antenna((long)step, *static_cast<base_t*>(this));
      ebdump::targets((long)step, (double)time, *static_cast<base_t*>(this), dom);
      ebdump::counts((long)step, dom);
      ebdump::dump((long)step, (double)time, dom);
    }
  };
} // namespace user

#endif
