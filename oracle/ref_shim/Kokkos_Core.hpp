// TEST INFRASTRUCTURE ONLY.
// A serial, host-only stand-in for the small part of the Kokkos API that the
// reference's kernel headers (src/kernels/*.hpp, src/metrics/*.h) touch, so that
// those headers can be compiled IN PLACE from /root/reference with plain g++
// (no cmake, no libkokkos) into oracle/_ref/libref_*.so. Written from scratch
// for this repo; it is not Kokkos code. Views wrap caller-owned memory in
// LayoutLeft order (first index fastest), which is the layout the reference's
// CUDA build uses and the layout of this repo's C ABI.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
#include <initializer_list>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_CLASS_LAMBDA [ =, this ]
#define KOKKOS_FORCEINLINE_FUNCTION inline

namespace Kokkos {

  struct HostSpace {};
  struct Serial {};
  using DefaultExecutionSpace     = Serial;
  using DefaultHostExecutionSpace = Serial;
  struct LayoutLeft {};
  struct LayoutRight {};

  enum : unsigned { RandomAccess = 1u, Unmanaged = 2u, Atomic = 4u };
  template <unsigned T>
  struct MemoryTraits {};

  struct ALL_t {};
  inline constexpr ALL_t ALL {};

  template <class A, class B>
  using pair = std::pair<A, B>;
  using std::make_pair;

  // ---- data type decoding: T*...[N]... -> value type, rank, static extents
  namespace shim {
    template <class T>
    struct dt {
      using value = T;
      static constexpr int rank = 0;
      static void          statics(std::size_t*, int) {}
    };
    template <class T>
    struct dt<T*> {
      using value = typename dt<T>::value;
      static constexpr int rank = dt<T>::rank + 1;
      static void statics(std::size_t* e, int pos) { dt<T>::statics(e, pos + 1); }
    };
    template <class T, std::size_t N>
    struct dt<T[N]> {
      using value = typename dt<T>::value;
      static constexpr int rank = dt<T>::rank + 1;
      static void          statics(std::size_t* e, int pos) {
        e[pos] = N;
        dt<T>::statics(e, pos + 1);
      }
    };
    // number of leading dynamic ('*') dimensions
    template <class T>
    struct ndyn { static constexpr int value = 0; };
    template <class T>
    struct ndyn<T*> { static constexpr int value = ndyn<T>::value + 1; };
    template <class T, std::size_t N>
    struct ndyn<T[N]> { static constexpr int value = ndyn<T>::value; };
  } // namespace shim

  template <class DataType, class... Props>
  class View {
  public:
    using value_type           = typename shim::dt<DataType>::value;
    using non_const_value_type = std::remove_const_t<value_type>;
    static constexpr int rank  = shim::dt<DataType>::rank;
    static constexpr int Rank  = rank;
    using host_mirror_type     = View<DataType>;
    using HostMirror           = View<DataType>;

    value_type*                             ptr { nullptr };
    std::size_t                             ext[8] { 1, 1, 1, 1, 1, 1, 1, 1 };
    std::shared_ptr<non_const_value_type[]> owner;

    View() = default;

    // unmanaged wrap (LayoutLeft)
    template <class... N>
    explicit View(value_type* p, N... n) : ptr { p } {
      set_extents(n...);
    }

    // allocating constructor (zero-initialised, like Kokkos)
    template <class... N>
    explicit View(const std::string&, N... n) {
      set_extents(n...);
      std::size_t tot = 1;
      for (int r = 0; r < rank; ++r) tot *= ext[r];
      owner = std::shared_ptr<non_const_value_type[]>(new non_const_value_type[tot ? tot : 1]());
      ptr   = owner.get();
    }

    // const / memory-trait conversion
    template <class DT2, class... P2,
              class = std::enable_if_t<std::is_same_v<std::remove_const_t<typename shim::dt<DT2>::value>,
                                                      non_const_value_type>>>
    View(const View<DT2, P2...>& o) : ptr { o.ptr }, owner { o.owner } {
      for (int r = 0; r < 8; ++r) ext[r] = o.ext[r];
    }

    template <class... I>
    inline value_type& operator()(I... idx) const {
      std::size_t off = 0, stride = 1;
      int         r   = 0;
      ((off += static_cast<std::size_t>(idx) * stride, stride *= ext[r++]), ...);
      return ptr[off];
    }

    std::size_t extent(int r) const { return ext[r]; }
    int         extent_int(int r) const { return static_cast<int>(ext[r]); }
    std::size_t size() const {
      std::size_t tot = 1;
      for (int r = 0; r < rank; ++r) tot *= ext[r];
      return tot;
    }
    value_type* data() const { return ptr; }
    bool        is_allocated() const { return ptr != nullptr; }
    std::size_t stride(int r) const {
      std::size_t s = 1;
      for (int q = 0; q < r; ++q) s *= ext[q];
      return s;
    }

  private:
    template <class... N>
    void set_extents(N... n) {
      // Kokkos convention for `T**[N]`: run-time extents first, compile-time last
      std::size_t statics[8] { 0, 0, 0, 0, 0, 0, 0, 0 };
      shim::dt<DataType>::statics(statics, 0);
      const std::size_t dyn[] = { static_cast<std::size_t>(n)..., 0 };
      constexpr int     nd    = shim::ndyn<DataType>::value;
      int               k     = 0;
      for (int r = 0; r < nd; ++r) {
        ext[k++] = (r < static_cast<int>(sizeof...(N))) ? dyn[r] : 0;
      }
      for (int r = 0; r < 8 && k < rank; ++r) {
        if (statics[r] != 0) ext[k++] = statics[r];
      }
    }
  };

  template <class V>
  inline V create_mirror_view(const V& v) { return v; }
  template <class A, class B>
  inline void deep_copy(const A&, const B&) {}
  inline void fence() {}

  template <class... T>
  struct RangePolicy {
    RangePolicy() = default;
    template <class... A>
    RangePolicy(A...) {}
  };
  template <unsigned N>
  struct Rank {};
  template <class... T>
  struct MDRangePolicy {
    MDRangePolicy() = default;
    template <class A, class B>
    MDRangePolicy(const A&, const B&) {}
    template <class I>
    MDRangePolicy(std::initializer_list<I>, std::initializer_list<I>) {}
  };
  template <class T, std::size_t N>
  struct Array {
    T v[N];
    T&       operator[](std::size_t i) { return v[i]; }
    const T& operator[](std::size_t i) const { return v[i]; }
  };

  // ---- math (namespace math = Kokkos in the reference)
  using std::abs;
  using std::acos;
  using std::asin;
  using std::atan;
  using std::atan2;
  using std::cbrt;
  using std::ceil;
  using std::cos;
  using std::cosh;
  using std::exp;
  using std::fabs;
  using std::floor;
  using std::fmod;
  using std::log;
  using std::log10;
  using std::pow;
  using std::round;
  using std::sin;
  using std::sinh;
  using std::sqrt;
  using std::tan;
  using std::tanh;
  using std::isnan;
  using std::isinf;
  using std::isfinite;
  using std::min;
  using std::max;

  template <class... A>
  inline void printf(const char* fmt, A... a) {
    if constexpr (sizeof...(A) == 0) {
      std::fputs(fmt, stderr);
    } else {
      std::fprintf(stderr, fmt, a...);
    }
  }
  [[noreturn]] inline void abort(const char* msg) {
    std::fprintf(stderr, "Kokkos(shim)::abort: %s\n", msg);
    std::abort();
  }

  template <class T>
  inline T atomic_fetch_add(T* p, T v) {
    T old = *p;
    *p += v;
    return old;
  }
  template <class T, class U>
  inline T atomic_fetch_add(T* p, U v) {
    T old = *p;
    *p += static_cast<T>(v);
    return old;
  }
  template <class T, class U>
  inline void atomic_add(T* p, U v) { *p += static_cast<T>(v); }

  namespace Experimental {
    template <class T>
    struct epsilon { static constexpr T value = std::numeric_limits<T>::epsilon(); };
    template <class T>
    struct finite_max { static constexpr T value = std::numeric_limits<T>::max(); };
    template <class T>
    struct infinity { static constexpr T value = std::numeric_limits<T>::infinity(); };
  } // namespace Experimental

  template <class T>
  struct Sum {};
  template <class T>
  struct Max {};
  template <class T>
  struct Min {};

  inline void initialize(int&, char**) {}
  inline void finalize() {}

  // serial parallel_for over [0, n)
  template <class F>
  inline void parallel_for(const std::string&, std::size_t n, const F& f) {
    for (std::size_t i = 0; i < n; ++i) f(static_cast<std::uint32_t>(i));
  }

} // namespace Kokkos
