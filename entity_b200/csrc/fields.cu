// entity_b200 -- cell kernels: Faraday / Ampere / CurrentsAmpere (Minkowski), the binomial
// current filter, and the single-domain ghost exchange.
//
// Reference behaviour: src/kernels/faraday_mink.hpp:71-166, src/kernels/ampere_mink.hpp:48-215,
// src/kernels/digital_filter.hpp:99-388, src/framework/domain/metadomain_comm.cpp:122-562,
// src/framework/domain/comm_nompi.hpp:29-119. Compiled twice (EB200_STRICT=0/1).
//
// All kernels are one thread per cell with i1 along threadIdx.x, so every load/store of a
// warp is a contiguous 128 B line of one component plane; the +-1 / +-2 neighbours of the
// stencils are served by L1/L2 (a row of the largest configured mesh is 16 KB).
#include "common.cuh"
#include "launch.h"

namespace eb200 {
  namespace EB200_VARIANT {

    struct Box {
      int n[3]; // active cells
      int G;
    };

    // flat thread index -> active cell (ghost-inclusive coordinates)
    template <int D>
    __device__ __forceinline__ bool cell_of(const Box& b, long t, int& i, int& j, int& k) {
      const long total = (long)b.n[0] * (D > 1 ? b.n[1] : 1) * (D > 2 ? b.n[2] : 1);
      if (t >= total) return false;
      i = (int)(t % b.n[0]) + b.G;
      j = 0;
      k = 0;
      if constexpr (D > 1) {
        const long r = t / b.n[0];
        j            = (int)(r % b.n[1]) + b.G;
        if constexpr (D > 2) {
          k = (int)(r / b.n[1]) + b.G;
        }
      }
      return true;
    }

    struct Stencil {
      float dx, dy, bxy, byx, dz, bxz, bzx, byz, bzy;
    };

    /* ------------------------------------------------------------------ Faraday */
    template <int D, bool EXT>
    __global__ void __launch_bounds__(256)
      faraday_kernel(Box box, FieldView<D> F, float coeff1, float coeff2, Stencil s) {
      int i1, i2, i3;
      if (!cell_of<D>(box, (long)blockIdx.x * blockDim.x + threadIdx.x, i1, i2, i3)) return;
      auto EB = [&](int i, int j, int k, int c) -> float { return F.at(i, j, k, c); };
      if constexpr (D == 1) {
        if constexpr (EXT) {
          const float ax = ONE - THREE * s.dx;
          F.at(i1, 0, 0, bx2) += coeff1 * (+ax * (EB(i1 + 1, 0, 0, ex3) - EB(i1, 0, 0, ex3)) +
                                           s.dx * (EB(i1 + 2, 0, 0, ex3) - EB(i1 - 1, 0, 0, ex3)));
          F.at(i1, 0, 0, bx3) += coeff1 * (-ax * (EB(i1 + 1, 0, 0, ex2) - EB(i1, 0, 0, ex2)) -
                                           s.dx * (EB(i1 + 2, 0, 0, ex2) - EB(i1 - 1, 0, 0, ex2)));
        } else {
          F.at(i1, 0, 0, bx2) += coeff1 * (EB(i1 + 1, 0, 0, ex3) - EB(i1, 0, 0, ex3));
          F.at(i1, 0, 0, bx3) += coeff1 * (EB(i1, 0, 0, ex2) - EB(i1 + 1, 0, 0, ex2));
        }
      } else if constexpr (D == 2) {
        if constexpr (EXT) {
          const float ax = ONE - TWO * s.bxy - THREE * s.dx;
          const float ay = ONE - TWO * s.byx - THREE * s.dy;
          // clang-format off
          F.at(i1, i2, 0, bx1) += coeff1 * (
              - ay    * (EB(i1    , i2 + 1, 0, ex3) - EB(i1    , i2    , 0, ex3))
              - s.dy  * (EB(i1    , i2 + 2, 0, ex3) - EB(i1    , i2 - 1, 0, ex3))
              - s.byx * (EB(i1 + 1, i2 + 1, 0, ex3) - EB(i1 + 1, i2    , 0, ex3))
              - s.byx * (EB(i1 - 1, i2 + 1, 0, ex3) - EB(i1 - 1, i2    , 0, ex3)));
          F.at(i1, i2, 0, bx2) += coeff1 * (
              + ax    * (EB(i1 + 1, i2    , 0, ex3) - EB(i1    , i2    , 0, ex3))
              + s.dx  * (EB(i1 + 2, i2    , 0, ex3) - EB(i1 - 1, i2    , 0, ex3))
              + s.bxy * (EB(i1 + 1, i2 + 1, 0, ex3) - EB(i1    , i2 + 1, 0, ex3))
              + s.bxy * (EB(i1 + 1, i2 - 1, 0, ex3) - EB(i1    , i2 - 1, 0, ex3)));
          F.at(i1, i2, 0, bx3) += coeff2 * (
              + ay    * (EB(i1    , i2 + 1, 0, ex1) - EB(i1    , i2    , 0, ex1))
              + s.dy  * (EB(i1    , i2 + 2, 0, ex1) - EB(i1    , i2 - 1, 0, ex1))
              + s.byx * (EB(i1 + 1, i2 + 1, 0, ex1) - EB(i1 + 1, i2    , 0, ex1))
              + s.byx * (EB(i1 - 1, i2 + 1, 0, ex1) - EB(i1 - 1, i2    , 0, ex1))
              - ax    * (EB(i1 + 1, i2    , 0, ex2) - EB(i1    , i2    , 0, ex2))
              - s.dx  * (EB(i1 + 2, i2    , 0, ex2) - EB(i1 - 1, i2    , 0, ex2))
              - s.bxy * (EB(i1 + 1, i2 + 1, 0, ex2) - EB(i1    , i2 + 1, 0, ex2))
              - s.bxy * (EB(i1 + 1, i2 - 1, 0, ex2) - EB(i1    , i2 - 1, 0, ex2)));
          // clang-format on
        } else {
          const float e3  = EB(i1, i2, 0, ex3);
          F.at(i1, i2, 0, bx1) += coeff1 * (e3 - EB(i1, i2 + 1, 0, ex3));
          F.at(i1, i2, 0, bx2) += coeff1 * (EB(i1 + 1, i2, 0, ex3) - e3);
          F.at(i1, i2, 0, bx3) += coeff2 * ((EB(i1, i2 + 1, 0, ex1) - EB(i1, i2, 0, ex1)) -
                                            (EB(i1 + 1, i2, 0, ex2) - EB(i1, i2, 0, ex2)));
        }
      } else {
        if constexpr (EXT) {
          const float ax = ONE - TWO * s.bxy - TWO * s.bxz - THREE * s.dx;
          const float ay = ONE - TWO * s.byx - TWO * s.byz - THREE * s.dy;
          const float az = ONE - TWO * s.bzx - TWO * s.bzy - THREE * s.dz;
          // clang-format off
          F.at(i1, i2, i3, bx1) += coeff1 * (
              + az    * (EB(i1    , i2    , i3 + 1, ex2) - EB(i1    , i2    , i3    , ex2))
              + s.dz  * (EB(i1    , i2    , i3 + 2, ex2) - EB(i1    , i2    , i3 - 1, ex2))
              + s.bzx * (EB(i1 + 1, i2    , i3 + 1, ex2) - EB(i1 + 1, i2    , i3    , ex2))
              + s.bzx * (EB(i1 - 1, i2    , i3 + 1, ex2) - EB(i1 - 1, i2    , i3    , ex2))
              + s.bzy * (EB(i1    , i2 + 1, i3 + 1, ex2) - EB(i1    , i2 + 1, i3    , ex2))
              + s.bzy * (EB(i1    , i2 - 1, i3 + 1, ex2) - EB(i1    , i2 - 1, i3    , ex2))
              - ay    * (EB(i1    , i2 + 1, i3    , ex3) - EB(i1    , i2    , i3    , ex3))
              - s.dy  * (EB(i1    , i2 + 2, i3    , ex3) - EB(i1    , i2 - 1, i3    , ex3))
              - s.byx * (EB(i1 + 1, i2 + 1, i3    , ex3) - EB(i1 + 1, i2    , i3    , ex3))
              - s.byx * (EB(i1 - 1, i2 + 1, i3    , ex3) - EB(i1 - 1, i2    , i3    , ex3))
              - s.byz * (EB(i1    , i2 + 1, i3 + 1, ex3) - EB(i1    , i2    , i3 + 1, ex3))
              - s.byz * (EB(i1    , i2 + 1, i3 - 1, ex3) - EB(i1    , i2    , i3 - 1, ex3)));
          F.at(i1, i2, i3, bx2) += coeff1 * (
              + ax    * (EB(i1 + 1, i2    , i3    , ex3) - EB(i1    , i2    , i3    , ex3))
              + s.dx  * (EB(i1 + 2, i2    , i3    , ex3) - EB(i1 - 1, i2    , i3    , ex3))
              + s.bxy * (EB(i1 + 1, i2 + 1, i3    , ex3) - EB(i1    , i2 + 1, i3    , ex3))
              + s.bxy * (EB(i1 + 1, i2 - 1, i3    , ex3) - EB(i1    , i2 - 1, i3    , ex3))
              + s.bxz * (EB(i1 + 1, i2    , i3 + 1, ex3) - EB(i1    , i2    , i3 + 1, ex3))
              + s.bxz * (EB(i1 + 1, i2    , i3 - 1, ex3) - EB(i1    , i2    , i3 - 1, ex3))
              - az    * (EB(i1    , i2    , i3 + 1, ex1) - EB(i1    , i2    , i3    , ex1))
              - s.dz  * (EB(i1    , i2    , i3 + 2, ex1) - EB(i1    , i2    , i3 - 1, ex1))
              - s.bzx * (EB(i1 + 1, i2    , i3 + 1, ex1) - EB(i1 + 1, i2    , i3    , ex1))
              - s.bzx * (EB(i1 - 1, i2    , i3 + 1, ex1) - EB(i1 - 1, i2    , i3    , ex1))
              - s.bzy * (EB(i1    , i2 + 1, i3 + 1, ex1) - EB(i1    , i2 + 1, i3    , ex1))
              - s.bzy * (EB(i1    , i2 - 1, i3 + 1, ex1) - EB(i1    , i2 - 1, i3    , ex1)));
          F.at(i1, i2, i3, bx3) += coeff1 * (
              + ay    * (EB(i1    , i2 + 1, i3    , ex1) - EB(i1    , i2    , i3    , ex1))
              + s.dy  * (EB(i1    , i2 + 2, i3    , ex1) - EB(i1    , i2 - 1, i3    , ex1))
              + s.byx * (EB(i1 + 1, i2 + 1, i3    , ex1) - EB(i1 + 1, i2    , i3    , ex1))
              + s.byx * (EB(i1 - 1, i2 + 1, i3    , ex1) - EB(i1 - 1, i2    , i3    , ex1))
              + s.byz * (EB(i1    , i2 + 1, i3 + 1, ex1) - EB(i1    , i2    , i3 + 1, ex1))
              + s.byz * (EB(i1    , i2 + 1, i3 - 1, ex1) - EB(i1    , i2    , i3 - 1, ex1))
              - ax    * (EB(i1 + 1, i2    , i3    , ex2) - EB(i1    , i2    , i3    , ex2))
              - s.dx  * (EB(i1 + 2, i2    , i3    , ex2) - EB(i1 - 1, i2    , i3    , ex2))
              - s.bxy * (EB(i1 + 1, i2 + 1, i3    , ex2) - EB(i1    , i2 + 1, i3    , ex2))
              - s.bxy * (EB(i1 + 1, i2 - 1, i3    , ex2) - EB(i1    , i2 - 1, i3    , ex2))
              - s.bxz * (EB(i1 + 1, i2    , i3 + 1, ex2) - EB(i1    , i2    , i3 + 1, ex2))
              - s.bxz * (EB(i1 + 1, i2    , i3 - 1, ex2) - EB(i1    , i2    , i3 - 1, ex2)));
          // clang-format on
        } else {
          const float e1 = EB(i1, i2, i3, ex1), e2 = EB(i1, i2, i3, ex2), e3 = EB(i1, i2, i3, ex3);
          F.at(i1, i2, i3, bx1) += coeff1 * ((EB(i1, i2, i3 + 1, ex2) - e2) -
                                             (EB(i1, i2 + 1, i3, ex3) - e3));
          F.at(i1, i2, i3, bx2) += coeff1 * ((EB(i1 + 1, i2, i3, ex3) - e3) -
                                             (EB(i1, i2, i3 + 1, ex1) - e1));
          F.at(i1, i2, i3, bx3) += coeff1 * ((EB(i1, i2 + 1, i3, ex1) - e1) -
                                             (EB(i1 + 1, i2, i3, ex2) - e2));
        }
      }
    }

    /* ------------------------------------------------------------------- Ampere */
    template <int D>
    __global__ void __launch_bounds__(256)
      ampere_kernel(Box box, FieldView<D> F, float coeff1, float coeff2) {
      int i1, i2, i3;
      if (!cell_of<D>(box, (long)blockIdx.x * blockDim.x + threadIdx.x, i1, i2, i3)) return;
      auto EB = [&](int i, int j, int k, int c) -> float { return F.at(i, j, k, c); };
      if constexpr (D == 1) {
        F.at(i1, 0, 0, ex2) += coeff1 * (EB(i1 - 1, 0, 0, bx3) - EB(i1, 0, 0, bx3));
        F.at(i1, 0, 0, ex3) += coeff1 * (EB(i1, 0, 0, bx2) - EB(i1 - 1, 0, 0, bx2));
      } else if constexpr (D == 2) {
        F.at(i1, i2, 0, ex1) += coeff1 * (EB(i1, i2, 0, bx3) - EB(i1, i2 - 1, 0, bx3));
        F.at(i1, i2, 0, ex2) += coeff1 * (EB(i1 - 1, i2, 0, bx3) - EB(i1, i2, 0, bx3));
        F.at(i1, i2, 0, ex3) += coeff2 * (EB(i1, i2 - 1, 0, bx1) - EB(i1, i2, 0, bx1) +
                                          EB(i1, i2, 0, bx2) - EB(i1 - 1, i2, 0, bx2));
      } else {
        F.at(i1, i2, i3, ex1) += coeff1 * (EB(i1, i2, i3 - 1, bx2) - EB(i1, i2, i3, bx2) +
                                           EB(i1, i2, i3, bx3) - EB(i1, i2 - 1, i3, bx3));
        F.at(i1, i2, i3, ex2) += coeff1 * (EB(i1 - 1, i2, i3, bx3) - EB(i1, i2, i3, bx3) +
                                           EB(i1, i2, i3, bx1) - EB(i1, i2, i3 - 1, bx1));
        F.at(i1, i2, i3, ex3) += coeff1 * (EB(i1, i2 - 1, i3, bx1) - EB(i1, i2, i3, bx1) +
                                           EB(i1, i2, i3, bx2) - EB(i1 - 1, i2, i3, bx2));
      }
    }

    template <int D>
    __global__ void __launch_bounds__(256)
      currents_ampere_kernel(Box box, FieldView<D> E, FieldView<D> J, float coeff, float ppc0) {
      int i1, i2, i3;
      if (!cell_of<D>(box, (long)blockIdx.x * blockDim.x + threadIdx.x, i1, i2, i3)) return;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float j = J.at(i1, i2, i3, c);
        E.at(i1, i2, i3, c) += j * coeff;
        J.at(i1, i2, i3, c) = j / ppc0;
      }
    }

    // CurrentsAmpere_kernel<D, ExtCurrent> (ampere_mink.hpp:134-215) with the external current
    // given as a table of Fourier modes (eb200_ext_current_t); the table arrives by value
    __device__ __forceinline__ float ext_current_at(const eb200_ext_current_t& X, int c, float x1,
                                                    float x2, float x3, int dim) {
      float j = ZERO;
      for (int m = 0; m < X.nmodes; ++m) {
        float kr = X.k[0][m] * x1 + X.k[1][m] * x2;
        if (dim == 3) kr = kr + X.k[2][m] * x3;
        const float cs = cosf(kr), sn = sinf(kr);
        j += X.pref[c][m] * (X.a_real[m] * cs - X.a_imag[m] * sn);
        if (X.pref2[c][m] != ZERO) {
          j += X.pref2[c][m] * (X.a_real2[m] * cs - X.a_imag2[m] * sn);
        }
      }
      return j;
    }

    template <int D>
    __global__ void __launch_bounds__(256)
      currents_ampere_ext_kernel(Box box, FieldView<D> E, FieldView<D> J, float coeff, float ppc0,
                                 const __grid_constant__ eb200_ext_current_t X, float dx,
                                 float x1min, float x2min, float x3min) {
      int i1, i2, i3;
      if (!cell_of<D>(box, (long)blockIdx.x * blockDim.x + threadIdx.x, i1, i2, i3)) return;
      const float f1 = static_cast<float>(i1 - box.G), f2 = static_cast<float>(i2 - box.G),
                  f3 = static_cast<float>(i3 - box.G);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // the component's own node: staggered by half a cell along its own direction
        const float x1 = ((c == 0) ? (f1 + HALF) : f1) * dx + x1min;
        const float x2 = (D > 1) ? ((c == 1) ? (f2 + HALF) : f2) * dx + x2min : ZERO;
        const float x3 = (D > 2) ? ((c == 2) ? (f3 + HALF) : f3) * dx + x3min : ZERO;
        float       j  = J.at(i1, i2, i3, c);
        if (c < D || D >= 1) {
          j += ppc0 * ext_current_at(X, c, x1, x2, x3, D);
        }
        E.at(i1, i2, i3, c) += j * coeff;
        J.at(i1, i2, i3, c) = j / ppc0;
      }
    }

    /* ------------------------------------------------------------ binomial filter */
    struct FilterBC {
      bool cmin[3], cmax[3]; // conductor faces
    };

    template <int D>
    __global__ void __launch_bounds__(256)
      filter_kernel(Box box, FieldView<D> A, FieldView<D> B, FilterBC bc) {
      int i1, i2, i3;
      if (!cell_of<D>(box, (long)blockIdx.x * blockDim.x + threadIdx.x, i1, i2, i3)) return;
      auto      buf = [&](int i, int j, int k, int c) -> float { return B.ld(i, j, k, c); };
      const int G      = box.G;
      const int i1_max = box.n[0] + G, i2_max = box.n[1] + G, i3_max = box.n[2] + G;
      if constexpr (D == 1) {
        if ((bc.cmin[0] && i1 == G) || (bc.cmax[0] && i1 == i1_max - 1)) {
          const int s          = bc.cmin[0] ? (i1 + 1) : (i1 - 1);
          A.at(i1, 0, 0, jx1) = (THREE * INV_4) * buf(i1, 0, 0, jx1) + (INV_4)*buf(s, 0, 0, jx1);
        } else if ((bc.cmin[0] && i1 == G + 1) || (bc.cmax[0] && i1 == i1_max - 2)) {
          const int s          = bc.cmin[0] ? (i1 + 1) : (i1 - 1);
          A.at(i1, 0, 0, jx1) = INV_2 * buf(i1, 0, 0, jx1) +
                                INV_4 * (buf(i1 - 1, 0, 0, jx1) + buf(i1 + 1, 0, 0, jx1));
          A.at(i1, 0, 0, jx2) = (INV_2)*buf(i1, 0, 0, jx2) + (INV_4)*buf(s, 0, 0, jx2);
          A.at(i1, 0, 0, jx3) = (INV_2)*buf(i1, 0, 0, jx3) + (INV_4)*buf(s, 0, 0, jx3);
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            A.at(i1, 0, 0, c) = INV_2 * buf(i1, 0, 0, c) +
                                INV_4 * (buf(i1 - 1, 0, 0, c) + buf(i1 + 1, 0, 0, c));
          }
        }
      } else if constexpr (D == 2) {
        auto F1 = [&](int c, int i, int j) {
          return INV_2 * buf(i, j, 0, c) + INV_4 * (buf(i - 1, j, 0, c) + buf(i + 1, j, 0, c));
        };
        auto F2 = [&](int c, int i, int j) {
          return INV_2 * buf(i, j, 0, c) + INV_4 * (buf(i, j - 1, 0, c) + buf(i, j + 1, 0, c));
        };
        if ((bc.cmin[0] && i1 == G) || (bc.cmax[0] && i1 == i1_max - 1)) {
          const int s           = bc.cmin[0] ? (i1 + 1) : (i1 - 1);
          A.at(i1, i2, 0, jx1) = (THREE * INV_4) * (F2(jx1, i1, i2)) + (INV_4) * (F2(jx1, s, i2));
        } else if ((bc.cmin[0] && i1 == G + 1) || (bc.cmax[0] && i1 == i1_max - 2)) {
          const int s           = bc.cmin[0] ? (i1 + 1) : (i1 - 1);
          A.at(i1, i2, 0, jx1) = INV_2 * (F2(jx1, i1, i2)) +
                                 INV_4 * ((F2(jx1, i1 - 1, i2)) + (F2(jx1, i1 + 1, i2)));
          A.at(i1, i2, 0, jx2) = INV_2 * (F2(jx2, i1, i2)) + INV_4 * (F2(jx2, s, i2));
          A.at(i1, i2, 0, jx3) = INV_2 * (F2(jx3, i1, i2)) + INV_4 * (F2(jx3, s, i2));
        } else if ((bc.cmin[1] && i2 == G) || (bc.cmax[1] && i2 == i2_max - 1)) {
          const int s           = bc.cmin[1] ? (i2 + 1) : (i2 - 1);
          A.at(i1, i2, 0, jx2) = (THREE * INV_4) * (F1(jx2, i1, i2)) + (INV_4) * (F1(jx2, i1, s));
        } else if ((bc.cmin[1] && i2 == G + 1) || (bc.cmax[1] && i2 == i2_max - 2)) {
          const int s           = bc.cmin[1] ? (i2 + 1) : (i2 - 1);
          A.at(i1, i2, 0, jx1) = INV_2 * (F1(jx1, i1, i2)) + INV_4 * (F1(jx1, i1, s));
          A.at(i1, i2, 0, jx2) = INV_2 * (F1(jx2, i1, i2)) +
                                 INV_4 * ((F1(jx2, i1, i2 - 1)) + (F1(jx2, i1, i2 + 1)));
          A.at(i1, i2, 0, jx3) = INV_2 * (F1(jx3, i1, i2)) + INV_4 * (F1(jx3, i1, s));
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            A.at(i1, i2, 0, c) =
              INV_4 * buf(i1, i2, 0, c) +
              INV_8 * (buf(i1 - 1, i2, 0, c) + buf(i1 + 1, i2, 0, c) + buf(i1, i2 - 1, 0, c) +
                       buf(i1, i2 + 1, 0, c)) +
              INV_16 * (buf(i1 - 1, i2 - 1, 0, c) + buf(i1 + 1, i2 + 1, 0, c) +
                        buf(i1 - 1, i2 + 1, 0, c) + buf(i1 + 1, i2 - 1, 0, c));
          }
        }
      } else {
        auto F12 = [&](int c, int i, int j, int k) {
          return INV_4 * buf(i, j, k, c) +
                 INV_8 * (buf(i - 1, j, k, c) + buf(i + 1, j, k, c) + buf(i, j - 1, k, c) +
                          buf(i, j + 1, k, c)) +
                 INV_16 * (buf(i - 1, j - 1, k, c) + buf(i + 1, j + 1, k, c) +
                           buf(i - 1, j + 1, k, c) + buf(i + 1, j - 1, k, c));
        };
        auto F23 = [&](int c, int i, int j, int k) {
          return INV_4 * buf(i, j, k, c) +
                 INV_8 * (buf(i, j - 1, k, c) + buf(i, j + 1, k, c) + buf(i, j, k - 1, c) +
                          buf(i, j, k + 1, c)) +
                 INV_16 * (buf(i, j - 1, k - 1, c) + buf(i, j + 1, k + 1, c) +
                           buf(i, j - 1, k + 1, c) + buf(i, j + 1, k - 1, c));
        };
        auto F13 = [&](int c, int i, int j, int k) {
          return INV_4 * buf(i, j, k, c) +
                 INV_8 * (buf(i - 1, j, k, c) + buf(i + 1, j, k, c) + buf(i, j, k - 1, c) +
                          buf(i, j, k + 1, c)) +
                 INV_16 * (buf(i - 1, j, k - 1, c) + buf(i + 1, j, k + 1, c) +
                           buf(i - 1, j, k + 1, c) + buf(i + 1, j, k - 1, c));
        };
        if ((bc.cmin[0] && i1 == G) || (bc.cmax[0] && i1 == i1_max - 1)) {
          const int s            = bc.cmin[0] ? (i1 + 1) : (i1 - 1);
          A.at(i1, i2, i3, jx1) = (THREE * INV_4) * (F23(jx1, i1, i2, i3)) +
                                  (INV_4) * (F23(jx1, s, i2, i3));
        } else if ((bc.cmin[0] && i1 == G + 1) || (bc.cmax[0] && i1 == i1_max - 2)) {
          const int s            = bc.cmin[0] ? (i1 + 1) : (i1 - 1);
          A.at(i1, i2, i3, jx1) = INV_2 * (F23(jx1, i1, i2, i3)) +
                                  INV_4 * ((F23(jx1, i1 - 1, i2, i3)) + (F23(jx1, i1 + 1, i2, i3)));
          A.at(i1, i2, i3, jx2) = INV_2 * (F23(jx2, i1, i2, i3)) + INV_4 * (F23(jx2, s, i2, i3));
          A.at(i1, i2, i3, jx3) = INV_2 * (F23(jx3, i1, i2, i3)) + INV_4 * (F23(jx3, s, i2, i3));
        } else if ((bc.cmin[1] && i2 == G) || (bc.cmax[1] && i2 == i2_max - 1)) {
          const int s            = bc.cmin[1] ? (i2 + 1) : (i2 - 1);
          A.at(i1, i2, i3, jx2) = (THREE * INV_4) * (F13(jx2, i1, i2, i3)) +
                                  (INV_4) * (F13(jx2, i1, s, i3));
        } else if ((bc.cmin[1] && i2 == G + 1) || (bc.cmax[1] && i2 == i2_max - 2)) {
          const int s            = bc.cmin[1] ? (i2 + 1) : (i2 - 1);
          A.at(i1, i2, i3, jx1) = INV_2 * (F13(jx1, i1, i2, i3)) + INV_4 * (F13(jx1, i1, s, i3));
          A.at(i1, i2, i3, jx2) = INV_2 * (F13(jx2, i1, i2, i3)) +
                                  INV_4 * ((F13(jx2, i1, i2 - 1, i3)) + (F13(jx2, i1, i2 + 1, i3)));
          A.at(i1, i2, i3, jx3) = INV_2 * (F13(jx3, i1, i2, i3)) + INV_4 * (F13(jx3, i1, s, i3));
        } else if ((bc.cmin[2] && i3 == G) || (bc.cmax[2] && i3 == i3_max - 1)) {
          const int s            = bc.cmin[2] ? (i3 + 1) : (i3 - 1);
          A.at(i1, i2, i3, jx3) = (THREE * INV_4) * (F12(jx3, i1, i2, i3)) +
                                  (INV_4) * (F12(jx3, i1, i2, s));
        } else if ((bc.cmin[2] && i3 == G + 1) || (bc.cmax[2] && i3 == i3_max - 2)) {
          const int s            = bc.cmin[2] ? (i3 + 1) : (i3 - 1);
          A.at(i1, i2, i3, jx1) = INV_2 * (F12(jx1, i1, i2, i3)) + INV_4 * (F12(jx1, i1, i2, s));
          A.at(i1, i2, i3, jx2) = INV_2 * (F12(jx2, i1, i2, i3)) + INV_4 * (F12(jx2, i1, i2, s));
          A.at(i1, i2, i3, jx3) = INV_2 * (F12(jx3, i1, i2, i3)) +
                                  INV_4 * ((F12(jx3, i1, i2, i3 - 1)) + (F12(jx3, i1, i2, i3 + 1)));
        } else {
          // the 1/32 group repeats (0,0,-+1) where the (0,-+1,+-1) pair would be expected;
          // that is what the reference computes (digital_filter.hpp:358-369)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            A.at(i1, i2, i3, c) =
              INV_8 * buf(i1, i2, i3, c) +
              INV_16 * (buf(i1 - 1, i2, i3, c) + buf(i1 + 1, i2, i3, c) + buf(i1, i2 - 1, i3, c) +
                        buf(i1, i2 + 1, i3, c) + buf(i1, i2, i3 - 1, c) + buf(i1, i2, i3 + 1, c)) +
              INV_32 * (buf(i1 - 1, i2 - 1, i3, c) + buf(i1 + 1, i2 + 1, i3, c) +
                        buf(i1 - 1, i2 + 1, i3, c) + buf(i1 + 1, i2 - 1, i3, c) +
                        buf(i1, i2 - 1, i3 - 1, c) + buf(i1, i2 + 1, i3 + 1, c) +
                        buf(i1, i2, i3 - 1, c) + buf(i1, i2, i3 + 1, c) +
                        buf(i1 - 1, i2, i3 - 1, c) + buf(i1 + 1, i2, i3 + 1, c) +
                        buf(i1 - 1, i2, i3 + 1, c) + buf(i1 + 1, i2, i3 - 1, c)) +
              INV_64 * (buf(i1 - 1, i2 - 1, i3 - 1, c) + buf(i1 + 1, i2 + 1, i3 + 1, c) +
                        buf(i1 - 1, i2 + 1, i3 + 1, c) + buf(i1 + 1, i2 - 1, i3 - 1, c) +
                        buf(i1 - 1, i2 - 1, i3 + 1, c) + buf(i1 + 1, i2 + 1, i3 - 1, c) +
                        buf(i1 - 1, i2 + 1, i3 - 1, c) + buf(i1 + 1, i2 - 1, i3 + 1, c));
          }
        }
      }
    }

    /* ------------------------------------------------- single-domain ghost exchange */
    struct Periodic {
      bool per[3];
    };

    // one thread per cell of the ghost-inclusive box; a ghost cell takes the value of its
    // periodic image (an active cell) when every face it lies beyond is periodic
    // Threads cover only the 2 D ghost slabs (thickness G, full extent in the other dimensions;
    // corner cells belong to several slabs and are written more than once with the same value).
    template <int D>
    __global__ void __launch_bounds__(256)
      ghost_fill_kernel(Box box, FieldView<D> F, int c0, int c1, Periodic per) {
      long      t    = (long)blockIdx.x * blockDim.x + threadIdx.x;
      const int G    = box.G;
      const int N[3] = { F.N1, F.N2, F.N3 };
      int       x[3] = { 0, 0, 0 };
      bool      found = false;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        // slab pair of dimension a: 2 G layers x the full extents of the other dimensions
        long other = 1;
#pragma unroll
        for (int b = 0; b < D; ++b) other *= (b == a) ? 1 : N[b];
        const long sz = 2L * G * other;
        if (!found && t < sz) {
          const int layer = (int)(t / other); // 0 .. 2G-1
          long      r     = t % other;
          x[a]            = (layer < G) ? layer : (box.n[a] + layer); // hi: n + G + (layer - G)
#pragma unroll
          for (int b = 0; b < D; ++b) {
            if (b != a) {
              x[b] = (int)(r % N[b]);
              r   /= N[b];
            }
          }
          found = true;
        }
        if (!found) t -= sz;
      }
      if (!found) return;
      int s[3] = { x[0], x[1], x[2] };
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const int n = box.n[a];
        if (x[a] < G) {
          if (!per.per[a]) return;
          s[a] = x[a] + n;
        } else if (x[a] >= n + G) {
          if (!per.per[a]) return;
          s[a] = x[a] - n;
        }
      }
      for (int c = c0; c < c1; ++c) {
        F.at(x[0], x[1], x[2], c) = F.at(s[0], s[1], s[2], c);
      }
    }

    // additive synchronisation of deposited currents: every active cell within G of a
    // periodic face receives what was deposited into the image ghost/edge cells. The
    // contributions are summed from zero in the reference's direction order (lexicographic
    // over {-1,0,1}^D) and then added to the cell, reproducing buff-then-add exactly.
    // All sources have at least one ghost coordinate, so the update is safe in place.
    template <int D>
    __global__ void __launch_bounds__(256)
      sync_currents_kernel(Box box, FieldView<D> J, Periodic per) {
      int x[3];
      if (!cell_of<D>(box, (long)blockIdx.x * blockDim.x + threadIdx.x, x[0], x[1], x[2])) return;
      const int G = box.G;
      bool      lo[3] = { false, false, false }, hi[3] = { false, false, false };
      bool      any   = false;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        lo[a] = per.per[a] && (x[a] < 2 * G);
        hi[a] = per.per[a] && (x[a] >= box.n[a]);
        any   = any || lo[a] || hi[a];
      }
      if (!any) return;
      float acc[3] = { ZERO, ZERO, ZERO };
      constexpr int ND = (D == 1) ? 3 : ((D == 2) ? 9 : 27);
      for (int lin = 0; lin < ND; ++lin) {
        int  d[3] = { 0, 0, 0 };
        int  r    = lin;
        bool ok   = true, zero = true;
#pragma unroll
        for (int a = D - 1; a >= 0; --a) {
          d[a] = (r % 3) - 1;
          r   /= 3;
        }
        int s[3] = { x[0], x[1], x[2] };
#pragma unroll
        for (int a = 0; a < D; ++a) {
          if (d[a] == 1) {
            ok   = ok && lo[a];
            s[a] = x[a] + box.n[a];
            zero = false;
          } else if (d[a] == -1) {
            ok   = ok && hi[a];
            s[a] = x[a] - box.n[a];
            zero = false;
          }
        }
        if (zero || !ok) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          acc[c] += J.at(s[0], s[1], s[2], c);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        J.at(x[0], x[1], x[2], c) += acc[c];
      }
    }

    /* ------------------------------------- fused binomial passes (2D, doubly periodic) */
    // `P` consecutive filter passes in one sweep over memory (temporal blocking): a CTA stages
    // its FT_X x FT_Y output tile plus a P-cell halo of one component in shared memory (halo
    // cells read the periodic image of the ACTIVE cells, which is what the per-pass ghost
    // exchange of srpic::CurrentsFilter provides, currents.h:108-118), runs the passes between
    // two shared buffers on a region that shrinks by one cell per pass, and writes the tile.
    // Every cell of every pass is computed with the reference's interior expression
    // (digital_filter.hpp:267-280) from the same nine values as the pass-by-pass path, so the
    // strict build gives identical bits. Traffic per P passes: ~1.7 reads + 1 write per value
    // instead of P x (copy + stencil + ghost fill).
    constexpr int FT_X = 64, FT_Y = 16, FT_PMAX = 4;

    // GHOSTS: the halo of a tile is read from the ghost cells as they are (they hold the
    // neighbour domains' exchanged values, P <= G) instead of the periodic image
    // st1 / st2 (GHOSTS = false): the dimension is not periodic and not exchanged -- its ghost
    // cells hold values that no pass changes (the reference never refreshes them between
    // passes either): the halo is read from them as they are and cells outside the active
    // range are carried, not filtered
    template <int P, bool GHOSTS = false, int ST = 0>
    __global__ void __launch_bounds__(256)
      filter_fused2d_kernel(FieldView<2> src, FieldView<2> dst, int n1, int n2, int G) {
      constexpr bool st1 = (ST & 2) != 0, st2 = (ST & 4) != 0; // static dimension 1 / 2
      constexpr int W = FT_X + 2 * P, H = FT_Y + 2 * P;
      __shared__ float buf[2][H * W];
      const int c  = blockIdx.z;
      const int x0 = blockIdx.x * FT_X - P, y0 = blockIdx.y * FT_Y - P; // active coordinates
      // periodic image of an active coordinate; one wrap suffices unless the mesh is smaller
      // than tile + halo, then the general modulo is taken
      auto wrap = [](int g, int n) {
        if (g < 0) g += n;
        if (g >= n) g -= n;
        if (g < 0 || g >= n) {
          g %= n;
          g += (g < 0) ? n : 0;
        }
        return g;
      };
      for (int e = threadIdx.x; e < W * H; e += 256) {
        const int ly = e / W, lx = e - ly * W;
        if constexpr (GHOSTS) {
          // tiles past the active edge (partial tiles) clamp: those cells are never stored
          const int gx = min(x0 + lx, n1 + G - 1), gy = min(y0 + ly, n2 + G - 1);
          buf[0][e] = src.ld(gx + G, gy + G, 0, c);
        } else {
          // only the first ghost layer of a static dimension is ever read by a carried cell's
          // neighbours; anything further out is clamped (and never used)
          // (the corner ghosts next to a static face are not periodic images either: the
          // reference exchanges corners only where both dimensions communicate)
          const int  x = x0 + lx, y = y0 + ly;
          const bool out1 = st1 && (x < 0 || x >= n1), out2 = st2 && (y < 0 || y >= n2);
          const int  gx = (st1 || out2) ? max(-1, min(x, n1)) : wrap(x, n1);
          const int  gy = (st2 || out1) ? max(-1, min(y, n2)) : wrap(y, n2);
          buf[0][e] = src.ld(gx + G, gy + G, 0, c);
        }
      }
      __syncthreads();
      int cur = 0;
#pragma unroll
      for (int p = 1; p <= P; ++p) {
        const float* b = buf[cur];
        float*       a = buf[cur ^ 1];
        const int    w = W - 2 * p, h = H - 2 * p; // compile-time after unrolling
        for (int e = threadIdx.x; e < w * h; e += 256) {
          const int    ly = e / w, lx = e - ly * w;
          const float* q  = b + (ly + p) * W + (lx + p);
          if constexpr (!GHOSTS) {
            if constexpr (st1 || st2) {
              const int gx = x0 + lx + p, gy = y0 + ly + p;
              if ((st1 && (gx < 0 || gx >= n1)) || (st2 && (gy < 0 || gy >= n2))) {
                a[(ly + p) * W + (lx + p)] = q[0];
                continue;
              }
              // next to a static face the neighbours in the static ghost layer are read from the
              // array itself at their true place (corner ghosts included -- they are not
              // periodic images), never from the tile; the cell this position stands for: a halo
              // position of a periodic dimension is the image of a cell at the other end
              if ((st1 && (gx == 0 || gx == n1 - 1)) || (st2 && (gy == 0 || gy == n2 - 1))) {
                const int X = st1 ? gx : wrap(gx, n1), Y = st2 ? gy : wrap(gy, n2);
                auto nb = [&](int dx, int dy) {
                  const int  nx = X + dx, ny = Y + dy;
                  const bool out = (st1 && (nx < 0 || nx >= n1)) || (st2 && (ny < 0 || ny >= n2));
                  return out ? src.ld(nx + G, ny + G, 0, c) : q[dy * W + dx];
                };
                a[(ly + p) * W + (lx + p)] =
                  INV_4 * q[0] + INV_8 * (nb(-1, 0) + nb(1, 0) + nb(0, -1) + nb(0, 1)) +
                  INV_16 * (nb(-1, -1) + nb(1, 1) + nb(-1, 1) + nb(1, -1));
                continue;
              }
            }
          }
          a[(ly + p) * W + (lx + p)] =
            INV_4 * q[0] + INV_8 * (q[-1] + q[1] + q[-W] + q[W]) +
            INV_16 * (q[-W - 1] + q[W + 1] + q[W - 1] + q[-W + 1]);
        }
        cur ^= 1;
        __syncthreads();
      }
      const float* r = buf[cur];
      for (int e = threadIdx.x; e < FT_X * FT_Y; e += 256) {
        const int ly = e / FT_X, lx = e - ly * FT_X;
        const int gx = blockIdx.x * FT_X + lx, gy = blockIdx.y * FT_Y + ly;
        if (gx < n1 && gy < n2) {
          dst.at(gx + G, gy + G, 0, c) = r[(ly + P) * W + (lx + P)];
        }
      }
    }

    /* ----------------------------------------------------------------- launchers */
    static Box make_box(const eb200_grid_t& g) {
      Box b;
      b.n[0] = g.n[0];
      b.n[1] = g.dim > 1 ? g.n[1] : 1;
      b.n[2] = g.dim > 2 ? g.n[2] : 1;
      b.G    = g.ng;
      return b;
    }

    static unsigned blocks_for(long n) { return (unsigned)((n + 255) / 256); }

    static long n_active(const eb200_grid_t& g) {
      return (long)g.n[0] * (g.dim > 1 ? g.n[1] : 1) * (g.dim > 2 ? g.n[2] : 1);
    }

    static long n_total(const eb200_grid_t& g) {
      long t = 1;
      for (int a = 0; a < g.dim; ++a) t *= g.n[a] + 2 * g.ng;
      return t;
    }

#define BY_DIM(g, CALL)                                                                        \
  switch ((g).dim) {                                                                           \
    case 1: CALL(1); break;                                                                    \
    case 2: CALL(2); break;                                                                    \
    case 3: CALL(3); break;                                                                    \
    default: return cudaErrorInvalidValue;                                                     \
  }

    cudaError_t faraday(const eb200_grid_t& g, float* em, float c1, float c2,
                        const float* st9, cudaStream_t st) {
      Stencil s { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
      bool    ext = false;
      if (st9) {
        s = Stencil { st9[0], st9[1], st9[2], st9[3], st9[4], st9[5], st9[6], st9[7], st9[8] };
        for (int q = 0; q < 9; ++q) ext = ext || (st9[q] != 0.0f);
      }
#if EB200_STRICT
      ext = true; // keep the reference's full expression (signed zeros included)
#endif
      const Box  box = make_box(g);
      const long n   = n_active(g);
#define CALL(D)                                                                                \
  if (ext)                                                                                     \
    faraday_kernel<D, true><<<blocks_for(n), 256, 0, st>>>(box, FieldView<D>(g, em), c1, c2, s); \
  else                                                                                         \
    faraday_kernel<D, false><<<blocks_for(n), 256, 0, st>>>(box, FieldView<D>(g, em), c1, c2, s);
      BY_DIM(g, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t ampere(const eb200_grid_t& g, float* em, float c1, float c2, cudaStream_t st) {
      const Box  box = make_box(g);
      const long n   = n_active(g);
#define CALL(D) ampere_kernel<D><<<blocks_for(n), 256, 0, st>>>(box, FieldView<D>(g, em), c1, c2);
      BY_DIM(g, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t currents_ampere(const eb200_grid_t& g, float* em, float* cur, float coeff,
                                float ppc0, cudaStream_t st) {
      const Box  box = make_box(g);
      const long n   = n_active(g);
#define CALL(D)                                                                                \
  currents_ampere_kernel<D>                                                                    \
    <<<blocks_for(n), 256, 0, st>>>(box, FieldView<D>(g, em), FieldView<D>(g, cur), coeff, ppc0);
      BY_DIM(g, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t currents_ampere_ext(const eb200_grid_t& g, float* em, float* cur, float coeff,
                                    float ppc0, const eb200_ext_current_t& ext, float dx,
                                    const float* xmin, cudaStream_t st) {
      const Box  box = make_box(g);
      const long n   = n_active(g);
#define CALL(D)                                                                                \
  currents_ampere_ext_kernel<D><<<blocks_for(n), 256, 0, st>>>(                                \
    box, FieldView<D>(g, em), FieldView<D>(g, cur), coeff, ppc0, ext, dx, xmin[0], xmin[1], xmin[2]);
      BY_DIM(g, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    // `extend` > 0: the pass also covers that many ghost layers (they must hold exchanged
    // data `extend` + 1 deep). Lets two passes share one halo exchange: the ghost-layer values
    // computed here are what the neighbour computes for its own edge cells.
    cudaError_t filter_pass(const eb200_grid_t& g, float* cur, const float* buff, const int* fbc,
                            int extend, cudaStream_t st) {
      Box box = make_box(g);
      if (extend > 0) {
        for (int a = 0; a < g.dim; ++a) box.n[a] += 2 * extend;
        box.G -= extend;
      }
      FilterBC  bc;
      for (int a = 0; a < 3; ++a) {
        bc.cmin[a] = (a < g.dim) && fbc[2 * a] == EB200_FBC_CONDUCTOR;
        bc.cmax[a] = (a < g.dim) && fbc[2 * a + 1] == EB200_FBC_CONDUCTOR;
      }
      const long n = (long)box.n[0] * box.n[1] * box.n[2];
#define CALL(D)                                                                                \
  filter_kernel<D><<<blocks_for(n), 256, 0, st>>>(box, FieldView<D>(g, cur),                   \
                                                  FieldView<D>(g, const_cast<float*>(buff)), bc);
      BY_DIM(g, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    // ghosts: 0 periodic images, 1 exchanged ghost cells (<= 2 passes), 2 | 4 = dimension 1 | 2
    // is static (non-periodic, not exchanged), the other one periodic unless also static
    cudaError_t filter_fused(const eb200_grid_t& g, const float* src, float* dst, int passes,
                             int ghosts, cudaStream_t st) {
      if (g.dim != 2 || passes < 1 || passes > FT_PMAX) return cudaErrorInvalidValue;
      const dim3 grid((g.n[0] + FT_X - 1) / FT_X, (g.n[1] + FT_Y - 1) / FT_Y, 3);
      const FieldView<2> S(g, const_cast<float*>(src)), Dd(g, dst);
      if (ghosts & 6) {
#define FUSED_ST(PP, SS) \
  filter_fused2d_kernel<PP, false, SS><<<grid, 256, 0, st>>>(S, Dd, g.n[0], g.n[1], g.ng)
#define FUSED_ST_P(SS)                                                                         \
  switch (passes) {                                                                            \
    case 1: FUSED_ST(1, SS); break;                                                            \
    case 2: FUSED_ST(2, SS); break;                                                            \
    case 3: FUSED_ST(3, SS); break;                                                            \
    default: FUSED_ST(4, SS); break;                                                           \
  }
        switch (ghosts & 6) {
          case 2: FUSED_ST_P(2) break;
          case 4: FUSED_ST_P(4) break;
          default: FUSED_ST_P(6) break;
        }
#undef FUSED_ST_P
#undef FUSED_ST
        count_launch();
        return cudaGetLastError();
      }
      if (ghosts) {
        if (passes > 2 || passes > g.ng) return cudaErrorInvalidValue;
        if (passes == 1) {
          filter_fused2d_kernel<1, true><<<grid, 256, 0, st>>>(S, Dd, g.n[0], g.n[1], g.ng);
        } else {
          filter_fused2d_kernel<2, true><<<grid, 256, 0, st>>>(S, Dd, g.n[0], g.n[1], g.ng);
        }
        count_launch();
        return cudaGetLastError();
      }
      switch (passes) {
        case 1: filter_fused2d_kernel<1><<<grid, 256, 0, st>>>(S, Dd, g.n[0], g.n[1], g.ng); break;
        case 2: filter_fused2d_kernel<2><<<grid, 256, 0, st>>>(S, Dd, g.n[0], g.n[1], g.ng); break;
        case 3: filter_fused2d_kernel<3><<<grid, 256, 0, st>>>(S, Dd, g.n[0], g.n[1], g.ng); break;
        default: filter_fused2d_kernel<4><<<grid, 256, 0, st>>>(S, Dd, g.n[0], g.n[1], g.ng); break;
      }
      count_launch();
      return cudaGetLastError();
    }

    static Periodic periodic_of(const eb200_grid_t& g, const int* fbc) {
      Periodic p;
      for (int a = 0; a < 3; ++a) {
        p.per[a] = (a < g.dim) && fbc[2 * a] == EB200_FBC_PERIODIC &&
                   fbc[2 * a + 1] == EB200_FBC_PERIODIC;
      }
      return p;
    }

    cudaError_t comm_fields_self(const eb200_grid_t& g, float* fld, int c0, int c1,
                                 const int* fbc, cudaStream_t st) {
      const Box      box = make_box(g);
      const Periodic per = periodic_of(g, fbc);
      if (!(per.per[0] || per.per[1] || per.per[2])) return cudaSuccess;
      long n = 0; // cells of the ghost slabs
      for (int a = 0; a < g.dim; ++a) {
        long other = 1;
        for (int b = 0; b < g.dim; ++b) other *= (b == a) ? 1 : (g.n[b] + 2 * g.ng);
        n += 2L * g.ng * other;
      }
#define CALL(D)                                                                                \
  ghost_fill_kernel<D><<<blocks_for(n), 256, 0, st>>>(box, FieldView<D>(g, fld), c0, c1, per);
      BY_DIM(g, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t sync_currents_self(const eb200_grid_t& g, float* cur, float* buff,
                                   const int* fbc, cudaStream_t st) {
      (void)buff; // the in-place kernel needs no staging buffer
      const Box      box = make_box(g);
      const Periodic per = periodic_of(g, fbc);
      if (!(per.per[0] || per.per[1] || per.per[2])) return cudaSuccess;
      const long n = n_active(g);
#define CALL(D) sync_currents_kernel<D><<<blocks_for(n), 256, 0, st>>>(box, FieldView<D>(g, cur), per);
      BY_DIM(g, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

  } // namespace EB200_VARIANT
} // namespace eb200
