// entity_b200 -- matching ("MATCH") field boundaries of a Minkowski SRPIC domain (SURVEY.md
// section 8f-1): kernel::bc::MatchBoundaries_kernel (src/kernels/fields_bcs.hpp:42-560) as
// called by srpic::MatchFieldsIn (src/engines/srpic/fields_bcs.h:38-215).
//
// The reference evaluates the problem generator's MatchFields functor inside the kernel; a
// functor cannot cross a C ABI, so the host evaluates it once per call site (it is a function
// of position and time only) into `target`: six component planes in the layout of `em`, each
// value taken at that component's own (staggered) node, in the tetrad basis -- what
// fset.ex1(x_Ph) ... fset.bx3(x_Ph) return. Here every component in `mask` becomes
//   F = s * F + (1 - s) * transform<T -> U>(target),  s = tanh(|x_o - xg_edge| * 4 / ds)
// with x_o the physical coordinate of the component's node along the matching direction o.
// Compiled with --fmad=false: the same fp32 operations in the same order as the reference
// (tanhf differs from glibc's in the last ulp: the stated tolerance of the parity test).
#include "common.cuh"
#include "launch.h"

namespace eb200 {
  namespace {
    struct MatchArgs {
      int   lo[3], n[3]; // ghost-inclusive start index and extent of the range per dimension
      int   G, o, tags, mask;
      float dx, xmin_o, xg_edge, ds;
    };

    template <int D>
    __global__ void __launch_bounds__(256)
      match_fields_kernel(const __grid_constant__ MatchArgs A, FieldView<D> F, FieldView<D> T) {
      const long ncell = (long)A.n[0] * A.n[1] * A.n[2];
      const long idx   = (long)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= ncell) return;
      int ijk[3];
      ijk[0] = (int)(idx % A.n[0]) + A.lo[0];
      ijk[1] = (D > 1) ? (int)((idx / A.n[0]) % A.n[1]) + A.lo[1] : 0;
      ijk[2] = (D > 2) ? (int)(idx / ((long)A.n[0] * A.n[1])) + A.lo[2] : 0;
      const float io = static_cast<float>(ijk[A.o]) - static_cast<float>(A.G); // COORD(i)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (!((A.mask >> c) & 1)) continue;
        const bool is_b = c >= 3;
        if (!(A.tags & (is_b ? EB200_BC_B : EB200_BC_E))) continue;
        const int   a         = is_b ? c - 3 : c;
        const bool  staggered = (A.o < D) && (is_b ? (A.o != a) : (A.o == a));
        const float xi        = staggered ? io + 0.5f : io;
        const float xph       = xi * A.dx + A.xmin_o; // metric.convert<o, Cd, Ph>
        const float s         = tanhf(fabsf(xph - A.xg_edge) * 4.0f / A.ds);
        const float t         = T.ld(ijk[0], ijk[1], ijk[2], c);
        const float tu        = (a < D) ? t / A.dx : t; // transform<a, Idx::T, Idx::U>
        float&      f         = F.at(ijk[0], ijk[1], ijk[2], c);
        f                     = s * f + (1.0f - s) * tu;
      }
    }
  } // namespace

  cudaError_t match_fields(const eb200_grid_t& g, float* em, const float* target, int o, float dx,
                           float xmin_o, float xg_edge, float ds, int tags, int mask,
                           const int* rmin, const int* rmax, cudaStream_t st) {
    MatchArgs A;
    long      ncell = 1;
    for (int a = 0; a < 3; ++a) {
      A.lo[a] = (a < g.dim) ? rmin[a] : 0;
      A.n[a]  = (a < g.dim) ? rmax[a] - rmin[a] : 1;
      if (A.n[a] <= 0) return cudaSuccess; // empty intersection: nothing to match
      ncell *= A.n[a];
    }
    A.G = g.ng, A.o = o, A.tags = tags, A.mask = mask;
    A.dx = dx, A.xmin_o = xmin_o, A.xg_edge = xg_edge, A.ds = ds;
    const unsigned nb = (unsigned)((ncell + 255) / 256);
    float*         tp = const_cast<float*>(target);
    switch (g.dim) {
      case 1: match_fields_kernel<1><<<nb, 256, 0, st>>>(A, FieldView<1>(g, em), FieldView<1>(g, tp)); break;
      case 2: match_fields_kernel<2><<<nb, 256, 0, st>>>(A, FieldView<2>(g, em), FieldView<2>(g, tp)); break;
      case 3: match_fields_kernel<3><<<nb, 256, 0, st>>>(A, FieldView<3>(g, em), FieldView<3>(g, tp)); break;
      default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
  }
} // namespace eb200
