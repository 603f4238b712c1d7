// TEST INFRASTRUCTURE ONLY.
// extern "C" driver around the REFERENCE's own output-staging kernels
// (kernel::FieldsToPhys_kernel, src/kernels/fields_to_phys.hpp; kernel::PrtlToPhys_kernel,
// src/kernels/prtls_to_phys.hpp), compiled in place from $(REF)/src against the serial
// mini-Kokkos in ref_shim/ (no reference source is copied). 2D meshes: Minkowski and the five
// curvilinear metrics. Checker of eb200_fields_to_phys / eb200_prtls_to_phys
// (tests/golden/make_out_golden.py -> tests/golden/out_golden.npz).
#include "oracle.h"

#include "enums.h"
#include "global.h"

#include "arch/kokkos_aliases.h"
#include "utils/numeric.h"

#include "metrics/kerr_schild.h"
#include "metrics/kerr_schild_0.h"
#include "metrics/minkowski.h"
#include "metrics/qkerr_schild.h"
#include "metrics/qspherical.h"
#include "metrics/spherical.h"

#include "kernels/fields_to_phys.hpp"
#include "kernels/prtls_to_phys.hpp"

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

using namespace ntt;

extern "C" {
typedef struct {
  int   kind; // 0 minkowski, 1 spherical, 2 qspherical, 3 kerr_schild, 4 qkerr_schild, 5 kerr_schild_0
  int   n1, n2;
  float x1min, x1max, x2min, x2max;
  float r0, h, a;
} refo_metric_t;
}

namespace {
  constexpr auto D2 = Dim::_2D;

  template <class M>
  M make(const refo_metric_t* m) {
    std::vector<ncells_t> res { (ncells_t)m->n1, (ncells_t)m->n2 };
    boundaries_t<real_t>  ext { { m->x1min, m->x1max }, { m->x2min, m->x2max } };
    if constexpr (std::is_same_v<M, metric::Minkowski<D2>>) {
      return M(res, ext);
    } else {
      std::map<std::string, real_t> prm { { "r0", m->r0 }, { "h", m->h }, { "a", m->a } };
      return M(res, ext, prm);
    }
  }

  template <class M>
  void fields_run(const refo_metric_t* mm, const orc_grid_t* g, float* from, float* to,
                  const int* cf, const int* ct, int flags) {
    const std::size_t G2 = 2 * (std::size_t)g->ng;
    ndfield_t<D2, 6>  F(from, g->n[0] + G2, g->n[1] + G2);
    ndfield_t<D2, 6>  T(to, g->n[0] + G2, g->n[1] + G2);
    const auto        metric = make<M>(mm);
    list_t<uint8_t, 3> lcf { (uint8_t)cf[0], (uint8_t)cf[1], (uint8_t)cf[2] };
    list_t<uint8_t, 3> lct { (uint8_t)ct[0], (uint8_t)ct[1], (uint8_t)ct[2] };
    kernel::FieldsToPhys_kernel<M, 6, 6> k(F, T, lcf, lct, (PrepareOutputFlags)flags, metric);
    // Mesh::rangeActiveCells
    for (ncells_t i = g->ng; i < (ncells_t)(g->n[0] + g->ng); ++i)
      for (ncells_t j = g->ng; j < (ncells_t)(g->n[1] + g->ng); ++j) k(i, j);
  }

  template <SimEngine::type S, class M>
  void prtls_run(const refo_metric_t* mm, const orc_prtls_t* p, uint32_t npart, uint32_t stride,
                 uint32_t nout, float* x1, float* x2, float* x3, float* u1, float* u2, float* u3,
                 float* w) {
    const auto         metric = make<M>(mm);
    array_t<npart_t*>  idx;
    array_t<real_t*>   bx1(x1, nout), bx2(x2, nout), bx3(x3, nout), bu1(u1, nout), bu2(u2, nout),
      bu3(u3, nout), bw(w, nout);
    // no payload columns (extent(1) == 0): the kernel's payload loops do nothing
    array_t<real_t**>  bpr((real_t*)nullptr, nout, 0), pr((real_t*)nullptr, npart, 0);
    array_t<npart_t**> bpi((npart_t*)nullptr, nout, 0), pi((npart_t*)nullptr, npart, 0);
    array_t<int*>      i1(p->i1, npart), i2(p->i2, npart), i3(p->i3, npart);
    array_t<prtldx_t*> d1(p->dx1, npart), d2(p->dx2, npart), d3(p->dx3, npart);
    array_t<real_t*>   v1(p->ux1, npart), v2(p->ux2, npart), v3(p->ux3, npart), ph(p->phi, npart),
      wt(p->weight, npart);
    kernel::PrtlToPhys_kernel<S, M, false> k(stride, idx, bx1, bx2, bx3, bu1, bu2, bu3, bw, bpr, bpi,
                                             i1, i2, i3, d1, d2, d3, v1, v2, v3, ph, wt, pr, pi,
                                             metric);
    for (npart_t q = 0; q < nout; ++q) k(q);
  }
} // namespace

extern "C" {
int refo_flag(int which) {
  switch (which) {
    case 0: return (int)PrepareOutput::InterpToCellCenterFromEdges;
    case 1: return (int)PrepareOutput::InterpToCellCenterFromFaces;
    case 2: return (int)PrepareOutput::ConvertToHat;
    case 3: return (int)PrepareOutput::ConvertToPhysCntrv;
    default: return (int)PrepareOutput::ConvertToPhysCov;
  }
}

void refo_fields_to_phys(const refo_metric_t* m, const orc_grid_t* g, float* from, float* to,
                         const int* cf, const int* ct, int flags) {
  switch (m->kind) {
    case 0: fields_run<metric::Minkowski<D2>>(m, g, from, to, cf, ct, flags); break;
    case 1: fields_run<metric::Spherical<D2>>(m, g, from, to, cf, ct, flags); break;
    case 2: fields_run<metric::QSpherical<D2>>(m, g, from, to, cf, ct, flags); break;
    case 3: fields_run<metric::KerrSchild<D2>>(m, g, from, to, cf, ct, flags); break;
    case 4: fields_run<metric::QKerrSchild<D2>>(m, g, from, to, cf, ct, flags); break;
    default: fields_run<metric::KerrSchild0<D2>>(m, g, from, to, cf, ct, flags); break;
  }
}

void refo_prtls_to_phys(const refo_metric_t* m, const orc_prtls_t* p, uint32_t npart,
                        uint32_t stride, uint32_t nout, float* x1, float* x2, float* x3,
                        float* u1, float* u2, float* u3, float* w) {
  constexpr auto SR = SimEngine::SRPIC, GR = SimEngine::GRPIC;
  switch (m->kind) {
    case 0: prtls_run<SR, metric::Minkowski<D2>>(m, p, npart, stride, nout, x1, x2, x3, u1, u2, u3, w); break;
    case 1: prtls_run<SR, metric::Spherical<D2>>(m, p, npart, stride, nout, x1, x2, x3, u1, u2, u3, w); break;
    case 2: prtls_run<SR, metric::QSpherical<D2>>(m, p, npart, stride, nout, x1, x2, x3, u1, u2, u3, w); break;
    case 3: prtls_run<GR, metric::KerrSchild<D2>>(m, p, npart, stride, nout, x1, x2, x3, u1, u2, u3, w); break;
    case 4: prtls_run<GR, metric::QKerrSchild<D2>>(m, p, npart, stride, nout, x1, x2, x3, u1, u2, u3, w); break;
    default: prtls_run<GR, metric::KerrSchild0<D2>>(m, p, npart, stride, nout, x1, x2, x3, u1, u2, u3, w); break;
  }
}
}
