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
#include "metrics.cuh"

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

  /* ================= field boundaries of 2D curvilinear (SRPIC) and GRPIC domains ================= */
  // What srpic::FieldBoundaries / grpic::FieldBoundaries launch besides the Minkowski MATCH
  // above (src/engines/srpic/fields_bcs.h:39-672, src/engines/grpic/fields_bcs.h:42-270). All
  // fields are 2D (r, theta) arrays of 6 (em, em0, aux) or 3 (cur0) component planes.
  namespace {
    __device__ __forceinline__ float& at2(float* f, long plane, int N1, int i, int j, int c) {
      return f[c * plane + (long)j * N1 + i];
    }

    // kernel::bc::AxisBoundaries_kernel<Dim::_2D, P> (fields_bcs.hpp:824-868): one thread per i1
    // over the whole ghost-inclusive x1 extent
    __global__ void __launch_bounds__(256)
      axis_fields_kernel(float* F, long plane, int N1, int i_edge, bool pos, bool setE, bool setB) {
      const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
      if (i1 >= N1) return;
      auto f = [&](int j, int c) -> float& { return at2(F, plane, N1, i1, j, c); };
      if (!pos) {
        if (setE) {
          f(i_edge - 1, 1) = -f(i_edge, 1);
          f(i_edge, 2)     = ZERO;
          f(i_edge - 1, 2) = f(i_edge + 1, 2);
        }
        if (setB) {
          f(i_edge - 1, 3) = f(i_edge, 3);
          f(i_edge, 4)     = ZERO;
          f(i_edge - 1, 4) = -f(i_edge + 1, 4);
          f(i_edge - 1, 5) = f(i_edge, 5);
        }
      } else {
        if (setE) {
          f(i_edge, 1)     = -f(i_edge - 1, 1);
          f(i_edge, 2)     = ZERO;
          f(i_edge + 1, 2) = f(i_edge - 1, 2);
        }
        if (setB) {
          f(i_edge, 3)     = f(i_edge - 1, 3);
          f(i_edge, 4)     = ZERO;
          f(i_edge + 1, 4) = -f(i_edge - 1, 4);
          f(i_edge, 5)     = f(i_edge - 1, 5);
        }
      }
    }

    // kernel::bc::gr::HorizonBoundaries_kernel<Dim::_2D> (fields_bcs.hpp:1187-1240): one thread
    // per i2 in [i2_min, i2_max]; the cells i1_min - G .. i1_min - G + 2 + nfilter take the value
    // of cell i1_min + 1 + nfilter
    __global__ void __launch_bounds__(256)
      horizon_fields_kernel(float* F, long plane, int N1, int j_lo, int j_hi, int i1_min, int G,
                            int nfilter, bool setE, bool setB) {
      const int j = j_lo + blockIdx.x * blockDim.x + threadIdx.x;
      if (j >= j_hi) return;
      const int src = i1_min + 1 + nfilter;
      for (int i = 0; i <= 2 + nfilter; ++i) {
        const int dst = i1_min - G + i;
        if (setE) {
#pragma unroll
          for (int c = 0; c < 3; ++c) at2(F, plane, N1, dst, j, c) = at2(F, plane, N1, src, j, c);
        }
        if (setB) {
#pragma unroll
          for (int c = 3; c < 6; ++c) at2(F, plane, N1, dst, j, c) = at2(F, plane, N1, src, j, c);
        }
      }
    }

    struct RangeArgs {
      int lo[2], n[2];
      int G, N1, N2;
    };

    // kernel::bc::MatchBoundaries_kernel<S, M, FS, o>::operator()(i1, i2) for a non-Cartesian M
    // (fields_bcs.hpp:176-340): F = s F + (1 - s) T on every component in `mask` whose tag is
    // set, s = tanh(|convert<o, Cd, Ph>(x_o) - xg_edge| 4 / ds) at the component's own node.
    // T = the functor's value in the basis the kernel blends with (SRPIC: contravariant,
    // GRPIC: as returned), tabulated by the host. The third component of E / D gets no target
    // on the axis rows (i2 == G with an axis at i2min, i2 == N2 - G with one at i2max).
    template <class M>
    __global__ void __launch_bounds__(256)
      match_fields_curv_kernel(const __grid_constant__ RangeArgs R, const MetricParams mp,
                               float* F, const float* __restrict__ T, long plane, int o,
                               float xg_edge, float ds, bool tagE, bool tagB, int mask,
                               bool axis_min, bool axis_max) {
      const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= (long)R.n[0] * R.n[1]) return;
      const int   i1 = (int)(idx % R.n[0]) + R.lo[0], i2 = (int)(idx / R.n[0]) + R.lo[1];
      const float c1 = static_cast<float>(i1 - R.G), c2 = static_cast<float>(i2 - R.G);
      auto        shape = [&](bool stag) {
        const float xi = ((o == 0) ? c1 : c2) + (stag ? HALF : ZERO);
        const float ph = (o == 0) ? M::r(mp, xi) : M::theta(mp, xi);
        return tanhf(fabsf(ph - xg_edge) * FOUR / ds);
      };
      const long n = (long)i2 * R.N1 + i1;
      // node of each component: ex1 (H,0) ex2 (0,H) ex3 (0,0) bx1 (0,H) bx2 (H,0) bx3 (H,H)
      const bool stag_x1[6] = { true, false, false, false, true, true };
      const bool stag_x2[6] = { false, true, false, true, false, true };
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (!((mask >> c) & 1)) continue;
        if (!((c < 3) ? tagE : tagB)) continue;
        const float s = shape((o == 0) ? stag_x1[c] : stag_x2[c]);
        float&      f = F[c * plane + n];
        const float t = T[c * plane + n];
        if (c == 2) {
          f = s * f;
          if ((!axis_min || (i2 > R.G)) && (!axis_max || (i2 < R.N2 - R.G))) {
            f += (ONE - s) * t;
          }
        } else {
          f = s * f + (ONE - s) * t;
        }
      }
    }

    // kernel::bc::EnforcedBoundaries_kernel<M, FS, P, O>::operator()(i1, i2)
    // (fields_bcs.hpp:975-1060): inside the range every defined component is SET to the
    // functor's value (contravariant, tabulated by the host), the normal E and the tangential B
    // only on the far side of i_edge (ghost-inclusive; >= for P, < otherwise)
    __global__ void __launch_bounds__(256)
      enforce_fields_kernel(const __grid_constant__ RangeArgs R, float* F,
                            const float* __restrict__ T, long plane, int o, bool pos, int i_edge,
                            bool tagE, bool tagB, int mask) {
      const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= (long)R.n[0] * R.n[1]) return;
      const int  i1 = (int)(idx % R.n[0]) + R.lo[0], i2 = (int)(idx / R.n[0]) + R.lo[1];
      const int  io = (o == 0) ? i1 : i2;
      const bool beyond = pos ? (io >= i_edge) : (io < i_edge);
      const long n = (long)i2 * R.N1 + i1;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (!((mask >> c) & 1)) continue;
        if (!((c < 3) ? tagE : tagB)) continue;
        const int  a = (c < 3) ? c : c - 3;
        // normal E (a == o) and tangential B (a != o) are restricted to the far side
        const bool restricted = (c < 3) ? (a == o) : (a != o);
        if (restricted && !beyond) continue;
        F[c * plane + n] = T[c * plane + n];
      }
    }

    // kernel::bc::gr::AbsorbCurrents_kernel<M, 1> (fields_bcs.hpp:1242-1283): every component of
    // J times tanh(|r(i1) - xg_edge| / (ds / 4)), the same factor for all three
    template <class M>
    __global__ void __launch_bounds__(256)
      absorb_currents_kernel(const __grid_constant__ RangeArgs R, const MetricParams mp, float* J,
                             long plane, float xg_edge, float ds) {
      const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= (long)R.n[0] * R.n[1]) return;
      const int   i1 = (int)(idx % R.n[0]) + R.lo[0], i2 = (int)(idx / R.n[0]) + R.lo[1];
      const float dx = fabsf(M::r(mp, static_cast<float>(i1 - R.G)) - xg_edge);
      const float s  = tanhf(dx / (INV_4 * ds));
      const long  n  = (long)i2 * R.N1 + i1;
#pragma unroll
      for (int c = 0; c < 3; ++c) J[c * plane + n] *= s;
    }

    // kernel::bc::ConductorBoundaries_kernel<Dim::_2D, o, P> (fields_bcs.hpp:523-660): mirror
    // images across a perfectly conducting face of a Cartesian domain
    __global__ void __launch_bounds__(256)
      conductor_fields2d_kernel(float* F, long plane, int N1, int N2, int o, bool pos, int i_edge,
                                int depth, bool tagE, bool tagB) {
      const long idx   = (long)blockIdx.x * blockDim.x + threadIdx.x;
      const int  other = (o == 0) ? N2 : N1;
      if (idx >= (long)depth * other) return;
      const int  k = (int)(idx % depth), t = (int)(idx / depth); // k: distance index, t: transverse
      auto f = [&](int q, int c) -> float& {
        return (o == 0) ? at2(F, plane, N1, q, t, c) : at2(F, plane, N1, t, q, c);
      };
      // components: n = normal (o), the other in-plane one and the out-of-plane one are tangential
      const int en = o, et = 1 - o, bn = 3 + o, bt = 3 + (1 - o);
      if (tagE) {
        if (k == 0) {
          f(i_edge, et) = ZERO;
          f(i_edge, 2)  = ZERO;
        } else if (!pos) {
          f(i_edge - k, en) = f(i_edge + k - 1, en);
          f(i_edge - k, et) = -f(i_edge + k, et);
          f(i_edge - k, 2)  = -f(i_edge + k, 2);
        } else {
          f(i_edge + k - 1, en) = f(i_edge - k, en);
          f(i_edge + k, et)     = -f(i_edge - k, et);
          f(i_edge + k, 2)      = -f(i_edge - k, 2);
        }
      }
      if (tagB) {
        if (k == 0) {
          f(i_edge, bn) = ZERO;
        } else if (!pos) {
          f(i_edge - k, bn) = -f(i_edge + k, bn);
          f(i_edge - k, bt) = f(i_edge + k - 1, bt);
          f(i_edge - k, 5)  = f(i_edge + k - 1, 5);
        } else {
          f(i_edge + k, bn)     = -f(i_edge - k, bn);
          f(i_edge + k - 1, bt) = f(i_edge - k, bt);
          f(i_edge + k - 1, 5)  = f(i_edge - k, 5);
        }
      }
    }

    static bool make_range(const eb200_grid_t& g, const int* rmin, const int* rmax, RangeArgs& R) {
      R.G  = g.ng;
      R.N1 = g.n[0] + 2 * g.ng;
      R.N2 = g.n[1] + 2 * g.ng;
      for (int a = 0; a < 2; ++a) {
        R.lo[a] = rmin[a];
        R.n[a]  = rmax[a] - rmin[a];
        if (R.n[a] <= 0) return false;
      }
      return true;
    }
  } // namespace

  namespace curv {
    cudaError_t axis_fields(const eb200_grid_t& g, float* fld, bool pos, int tags, cudaStream_t st) {
      const int  N1 = g.n[0] + 2 * g.ng, N2 = g.n[1] + 2 * g.ng;
      const int  i_edge = pos ? (g.ng + g.n[1]) : g.ng; // i_max(x2) / i_min(x2)
      axis_fields_kernel<<<(N1 + 255) / 256, 256, 0, st>>>(fld, (long)N1 * N2, N1, i_edge, pos,
                                                           (tags & EB200_BC_E) != 0,
                                                           (tags & EB200_BC_B) != 0);
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t horizon_fields(const eb200_grid_t& g, float* fld, int tags, int nfilter,
                               cudaStream_t st) {
      const int N1 = g.n[0] + 2 * g.ng, N2 = g.n[1] + 2 * g.ng;
      const int j_lo = g.ng, j_hi = g.ng + g.n[1] + 1; // [i_min(x2), i_max(x2) + 1)
      horizon_fields_kernel<<<(j_hi - j_lo + 255) / 256, 256, 0, st>>>(
        fld, (long)N1 * N2, N1, j_lo, j_hi, g.ng, g.ng, nfilter, (tags & EB200_BC_E) != 0,
        (tags & EB200_BC_B) != 0);
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t match_fields_curv(const MetricParams& m, const eb200_grid_t& g, float* fld,
                                  const float* target, int o, float xg_edge, float ds, int tags,
                                  int mask, const int* rmin, const int* rmax, const int* fbc,
                                  cudaStream_t st) {
      RangeArgs R;
      if (!make_range(g, rmin, rmax, R)) return cudaSuccess;
      const long     plane = (long)R.N1 * R.N2;
      const unsigned nb    = (unsigned)(((long)R.n[0] * R.n[1] + 255) / 256);
      const bool     tE = (tags & EB200_BC_E) != 0, tB = (tags & EB200_BC_B) != 0;
      const bool     amin = fbc[2] == EB200_FBC_AXIS, amax = fbc[3] == EB200_FBC_AXIS;
#define CALL(MM)                                                                               \
  match_fields_curv_kernel<MM><<<nb, 256, 0, st>>>(R, m, fld, target, plane, o, xg_edge, ds, tE, \
                                                   tB, mask, amin, amax)
      switch (m.kind) {
        case EB200_METRIC_SPHERICAL: CALL(Spherical); break;
        case EB200_METRIC_QSPHERICAL: CALL(QSpherical); break;
        case EB200_METRIC_KERR_SCHILD: CALL(KerrSchild); break;
        case EB200_METRIC_QKERR_SCHILD: CALL(QKerrSchild); break;
        case EB200_METRIC_KERR_SCHILD_0: CALL(KerrSchild0); break;
        default: return cudaErrorInvalidValue;
      }
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t enforce_fields(const eb200_grid_t& g, float* em, const float* target, int o,
                               bool pos, int i_edge, int tags, int mask, const int* rmin,
                               const int* rmax, cudaStream_t st) {
      RangeArgs R;
      if (!make_range(g, rmin, rmax, R)) return cudaSuccess;
      const unsigned nb = (unsigned)(((long)R.n[0] * R.n[1] + 255) / 256);
      enforce_fields_kernel<<<nb, 256, 0, st>>>(R, em, target, (long)R.N1 * R.N2, o, pos, i_edge,
                                                (tags & EB200_BC_E) != 0, (tags & EB200_BC_B) != 0,
                                                mask);
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t absorb_currents(const MetricParams& m, const eb200_grid_t& g, float* cur,
                                float xg_edge, float ds, const int* rmin, const int* rmax,
                                cudaStream_t st) {
      RangeArgs R;
      if (!make_range(g, rmin, rmax, R)) return cudaSuccess;
      const long     plane = (long)R.N1 * R.N2;
      const unsigned nb    = (unsigned)(((long)R.n[0] * R.n[1] + 255) / 256);
#define CALL(MM) absorb_currents_kernel<MM><<<nb, 256, 0, st>>>(R, m, cur, plane, xg_edge, ds)
      switch (m.kind) {
        case EB200_METRIC_KERR_SCHILD: CALL(KerrSchild); break;
        case EB200_METRIC_QKERR_SCHILD: CALL(QKerrSchild); break;
        case EB200_METRIC_KERR_SCHILD_0: CALL(KerrSchild0); break;
        default: return cudaErrorInvalidValue;
      }
#undef CALL
      count_launch();
      return cudaGetLastError();
    }
  } // namespace curv

  cudaError_t conductor_fields2d(const eb200_grid_t& g, float* em, int o, bool pos, int tags,
                                 cudaStream_t st) {
    const int N1 = g.n[0] + 2 * g.ng, N2 = g.n[1] + 2 * g.ng;
    // srpic::PerfectConductorFieldsIn (fields_bcs.h:384-470): distance index 0 .. G (- side) or
    // 0 .. G - 1 (+ side), the whole transverse extent
    const int  depth  = pos ? g.ng : g.ng + 1;
    const int  i_edge = pos ? (g.ng + g.n[o]) : g.ng;
    const long n      = (long)depth * ((o == 0) ? N2 : N1);
    conductor_fields2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      em, (long)N1 * N2, N1, N2, o, pos, i_edge, depth, (tags & EB200_BC_E) != 0,
      (tags & EB200_BC_B) != 0);
    count_launch();
    return cudaGetLastError();
  }
} // namespace eb200
