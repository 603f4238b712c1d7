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
    PGen(P&& p, MD&& m) : base_t { std::forward<P>(p), std::forward<MD>(m) } {}

    void CustomPostStep(timestep_t step, simtime_t time, Domain<S, M>& dom) {
      ebdump::npart_pre().clear();
      for (auto& sp : dom.species) ebdump::npart_pre().push_back((std::uint32_t)sp.npart());
      if constexpr (::traits::pgen::HasCustomPostStep<base_t, Domain<S, M>>) {
        base_t::CustomPostStep(step, time, dom);
      }
      ebdump::antenna((long)step, *static_cast<base_t*>(this));
      ebdump::counts((long)step, dom);
      ebdump::dump((long)step, (double)time, dom);
    }
  };
} // namespace user

#endif
