// TEST INFRASTRUCTURE ONLY -- type stand-ins; no RNG is used on the hot path.
#pragma once
#include "Kokkos_Core.hpp"
namespace Kokkos {
  template <class Space>
  struct Random_XorShift64_Pool {
    struct generator_type {
      double drand() { return 0.0; }
      float  frand() { return 0.0f; }
      template <class... A> double drand(A...) { return 0.0; }
      template <class... A> float  frand(A...) { return 0.0f; }
      std::uint64_t urand64() { return 0; }
    };
    Random_XorShift64_Pool() = default;
    Random_XorShift64_Pool(std::uint64_t) {}
    generator_type get_state() const { return {}; }
    void           free_state(const generator_type&) const {}
  };
  template <class Space>
  using Random_XorShift1024_Pool = Random_XorShift64_Pool<Space>;
  template <class G, class T>
  struct rand {
    static T draw(G&) { return T(0); }
    static T draw(G&, T) { return T(0); }
    static T draw(G&, T, T) { return T(0); }
  };
}
