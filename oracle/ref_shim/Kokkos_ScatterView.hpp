// TEST INFRASTRUCTURE ONLY -- serial stand-in (see Kokkos_Core.hpp in this dir).
// On the Serial backend the real ScatterView is NonDuplicated/NonAtomic: access()
// hands back the target view and `+=` goes straight to memory in program order.
#pragma once
#include "Kokkos_Core.hpp"
namespace Kokkos { namespace Experimental {
  template <class DataType, class... P>
  class ScatterView {
  public:
    View<DataType> target;
    ScatterView() = default;
    ScatterView(const View<DataType>& v) : target { v } {}
    const View<DataType>& access() const { return target; }
  };
  template <class DataType, class... P>
  inline ScatterView<DataType> create_scatter_view(const View<DataType, P...>& v) {
    return ScatterView<DataType>(View<DataType>(v));
  }
  template <class A, class B>
  inline void contribute(const A&, const B&) {}
}} // namespace Kokkos::Experimental
