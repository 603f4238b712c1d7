// TEST INFRASTRUCTURE ONLY -- CPU oracle for the PIC hot path (see oracle.h).
//
// Scalar fp32 restatement, in serial program order, of the reference kernels
// for the Minkowski (Cartesian) SRPIC path. Every function cites the reference
// lines it follows. Build with -ffp-contract=off so that no FMA is formed: the
// result then equals the reference's Kokkos Serial/OpenMP(1 thread) build on a
// baseline x86-64 target bit for bit (checked against oracle/_ref in
// tests/test_oracle_vs_ref.py when the reference is compiled here).
#include "oracle.h"
#include "shapes.hpp"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace orc {

  enum { ex1 = 0, ex2 = 1, ex3 = 2, bx1 = 3, bx2 = 4, bx3 = 5 };
  enum { jx1 = 0, jx2 = 1, jx3 = 2 };

  // LayoutLeft field accessor: i1 fastest, component planes slowest.
  template <int D>
  struct Fld {
    float* p;
    long   N1, N2, N3;

    Fld(const orc_grid_t* g, float* ptr) : p { ptr } {
      N1 = g->n[0] + 2 * g->ng;
      N2 = (D > 1) ? g->n[1] + 2 * g->ng : 1;
      N3 = (D > 2) ? g->n[2] + 2 * g->ng : 1;
    }

    inline float& operator()(long i, int c) const { return p[i + N1 * N2 * N3 * c]; }

    inline float& operator()(long i, long j, int c) const {
      return p[i + N1 * (j + N2 * N3 * c)];
    }

    inline float& operator()(long i, long j, long k, int c) const {
      return p[i + N1 * (j + N2 * (k + N3 * c))];
    }
  };

  /* ---------------------------------------------------------------------- */
  /* Faraday: src/kernels/faraday_mink.hpp:71-166                            */
  /* ---------------------------------------------------------------------- */
  template <int D>
  void faraday(const orc_grid_t* g, float* em, float coeff1, float coeff2, const float* st) {
    Fld<D>      EB(g, em);
    const float deltax = st ? st[0] : ZERO, deltay = st ? st[1] : ZERO;
    const float betaxy = st ? st[2] : ZERO, betayx = st ? st[3] : ZERO;
    const float deltaz = st ? st[4] : ZERO, betaxz = st ? st[5] : ZERO;
    const float betazx = st ? st[6] : ZERO, betayz = st ? st[7] : ZERO;
    const float betazy = st ? st[8] : ZERO;
    const int   G      = g->ng;
    if constexpr (D == 1) {
      const float alphax = ONE - THREE * deltax;
      for (long i1 = G; i1 < g->n[0] + G; ++i1) {
        EB(i1, bx2) += coeff1 * (+alphax * (EB(i1 + 1, ex3) - EB(i1, ex3)) +
                                 deltax * (EB(i1 + 2, ex3) - EB(i1 - 1, ex3)));
        EB(i1, bx3) += coeff1 * (-alphax * (EB(i1 + 1, ex2) - EB(i1, ex2)) -
                                 deltax * (EB(i1 + 2, ex2) - EB(i1 - 1, ex2)));
      }
    } else if constexpr (D == 2) {
      const float alphax = ONE - TWO * betaxy - THREE * deltax;
      const float alphay = ONE - TWO * betayx - THREE * deltay;
      for (long i2 = G; i2 < g->n[1] + G; ++i2) {
        for (long i1 = G; i1 < g->n[0] + G; ++i1) {
          // clang-format off
          EB(i1, i2, bx1) += coeff1 * (
                          - alphay * (EB(i1    , i2 + 1, ex3) - EB(i1    , i2    , ex3))
                          - deltay * (EB(i1    , i2 + 2, ex3) - EB(i1    , i2 - 1, ex3))
                          - betayx * (EB(i1 + 1, i2 + 1, ex3) - EB(i1 + 1, i2    , ex3))
                          - betayx * (EB(i1 - 1, i2 + 1, ex3) - EB(i1 - 1, i2    , ex3)));
          EB(i1, i2, bx2) += coeff1 * (
                          + alphax * (EB(i1 + 1, i2    , ex3) - EB(i1    , i2    , ex3))
                          + deltax * (EB(i1 + 2, i2    , ex3) - EB(i1 - 1, i2    , ex3))
                          + betaxy * (EB(i1 + 1, i2 + 1, ex3) - EB(i1    , i2 + 1, ex3))
                          + betaxy * (EB(i1 + 1, i2 - 1, ex3) - EB(i1    , i2 - 1, ex3)));
          EB(i1, i2, bx3) += coeff2 * (
                          + alphay * (EB(i1    , i2 + 1, ex1) - EB(i1    , i2    , ex1))
                          + deltay * (EB(i1    , i2 + 2, ex1) - EB(i1    , i2 - 1, ex1))
                          + betayx * (EB(i1 + 1, i2 + 1, ex1) - EB(i1 + 1, i2    , ex1))
                          + betayx * (EB(i1 - 1, i2 + 1, ex1) - EB(i1 - 1, i2    , ex1))
                          - alphax * (EB(i1 + 1, i2    , ex2) - EB(i1    , i2    , ex2))
                          - deltax * (EB(i1 + 2, i2    , ex2) - EB(i1 - 1, i2    , ex2))
                          - betaxy * (EB(i1 + 1, i2 + 1, ex2) - EB(i1    , i2 + 1, ex2))
                          - betaxy * (EB(i1 + 1, i2 - 1, ex2) - EB(i1    , i2 - 1, ex2)));
          // clang-format on
        }
      }
    } else {
      const float alphax = ONE - TWO * betaxy - TWO * betaxz - THREE * deltax;
      const float alphay = ONE - TWO * betayx - TWO * betayz - THREE * deltay;
      const float alphaz = ONE - TWO * betazx - TWO * betazy - THREE * deltaz;
      for (long i3 = G; i3 < g->n[2] + G; ++i3) {
        for (long i2 = G; i2 < g->n[1] + G; ++i2) {
          for (long i1 = G; i1 < g->n[0] + G; ++i1) {
            // clang-format off
            EB(i1, i2, i3, bx1) += coeff1 * (
                  + alphaz * (EB(i1    , i2    , i3 + 1, ex2) - EB(i1    , i2    , i3    , ex2))
                  + deltaz * (EB(i1    , i2    , i3 + 2, ex2) - EB(i1    , i2    , i3 - 1, ex2))
                  + betazx * (EB(i1 + 1, i2    , i3 + 1, ex2) - EB(i1 + 1, i2    , i3    , ex2))
                  + betazx * (EB(i1 - 1, i2    , i3 + 1, ex2) - EB(i1 - 1, i2    , i3    , ex2))
                  + betazy * (EB(i1    , i2 + 1, i3 + 1, ex2) - EB(i1    , i2 + 1, i3    , ex2))
                  + betazy * (EB(i1    , i2 - 1, i3 + 1, ex2) - EB(i1    , i2 - 1, i3    , ex2))
                  - alphay * (EB(i1    , i2 + 1, i3    , ex3) - EB(i1    , i2    , i3    , ex3))
                  - deltay * (EB(i1    , i2 + 2, i3    , ex3) - EB(i1    , i2 - 1, i3    , ex3))
                  - betayx * (EB(i1 + 1, i2 + 1, i3    , ex3) - EB(i1 + 1, i2    , i3    , ex3))
                  - betayx * (EB(i1 - 1, i2 + 1, i3    , ex3) - EB(i1 - 1, i2    , i3    , ex3))
                  - betayz * (EB(i1    , i2 + 1, i3 + 1, ex3) - EB(i1    , i2    , i3 + 1, ex3))
                  - betayz * (EB(i1    , i2 + 1, i3 - 1, ex3) - EB(i1    , i2    , i3 - 1, ex3)));
            EB(i1, i2, i3, bx2) += coeff1 * (
                  + alphax * (EB(i1 + 1, i2    , i3    , ex3) - EB(i1    , i2    , i3    , ex3))
                  + deltax * (EB(i1 + 2, i2    , i3    , ex3) - EB(i1 - 1, i2    , i3    , ex3))
                  + betaxy * (EB(i1 + 1, i2 + 1, i3    , ex3) - EB(i1    , i2 + 1, i3    , ex3))
                  + betaxy * (EB(i1 + 1, i2 - 1, i3    , ex3) - EB(i1    , i2 - 1, i3    , ex3))
                  + betaxz * (EB(i1 + 1, i2    , i3 + 1, ex3) - EB(i1    , i2    , i3 + 1, ex3))
                  + betaxz * (EB(i1 + 1, i2    , i3 - 1, ex3) - EB(i1    , i2    , i3 - 1, ex3))
                  - alphaz * (EB(i1    , i2    , i3 + 1, ex1) - EB(i1    , i2    , i3    , ex1))
                  - deltaz * (EB(i1    , i2    , i3 + 2, ex1) - EB(i1    , i2    , i3 - 1, ex1))
                  - betazx * (EB(i1 + 1, i2    , i3 + 1, ex1) - EB(i1 + 1, i2    , i3    , ex1))
                  - betazx * (EB(i1 - 1, i2    , i3 + 1, ex1) - EB(i1 - 1, i2    , i3    , ex1))
                  - betazy * (EB(i1    , i2 + 1, i3 + 1, ex1) - EB(i1    , i2 + 1, i3    , ex1))
                  - betazy * (EB(i1    , i2 - 1, i3 + 1, ex1) - EB(i1    , i2 - 1, i3    , ex1)));
            EB(i1, i2, i3, bx3) += coeff1 * (
                  + alphay * (EB(i1    , i2 + 1, i3    , ex1) - EB(i1    , i2    , i3    , ex1))
                  + deltay * (EB(i1    , i2 + 2, i3    , ex1) - EB(i1    , i2 - 1, i3    , ex1))
                  + betayx * (EB(i1 + 1, i2 + 1, i3    , ex1) - EB(i1 + 1, i2    , i3    , ex1))
                  + betayx * (EB(i1 - 1, i2 + 1, i3    , ex1) - EB(i1 - 1, i2    , i3    , ex1))
                  + betayz * (EB(i1    , i2 + 1, i3 + 1, ex1) - EB(i1    , i2    , i3 + 1, ex1))
                  + betayz * (EB(i1    , i2 + 1, i3 - 1, ex1) - EB(i1    , i2    , i3 - 1, ex1))
                  - alphax * (EB(i1 + 1, i2    , i3    , ex2) - EB(i1    , i2    , i3    , ex2))
                  - deltax * (EB(i1 + 2, i2    , i3    , ex2) - EB(i1 - 1, i2    , i3    , ex2))
                  - betaxy * (EB(i1 + 1, i2 + 1, i3    , ex2) - EB(i1    , i2 + 1, i3    , ex2))
                  - betaxy * (EB(i1 + 1, i2 - 1, i3    , ex2) - EB(i1    , i2 - 1, i3    , ex2))
                  - betaxz * (EB(i1 + 1, i2    , i3 + 1, ex2) - EB(i1    , i2    , i3 + 1, ex2))
                  - betaxz * (EB(i1 + 1, i2    , i3 - 1, ex2) - EB(i1    , i2    , i3 - 1, ex2)));
            // clang-format on
          }
        }
      }
    }
  }

  /* ---------------------------------------------------------------------- */
  /* Ampere: src/kernels/ampere_mink.hpp:48-89                               */
  /* ---------------------------------------------------------------------- */
  template <int D>
  void ampere(const orc_grid_t* g, float* em, float coeff1, float coeff2) {
    Fld<D>    EB(g, em);
    const int G = g->ng;
    if constexpr (D == 1) {
      for (long i1 = G; i1 < g->n[0] + G; ++i1) {
        EB(i1, ex2) += coeff1 * (EB(i1 - 1, bx3) - EB(i1, bx3));
        EB(i1, ex3) += coeff1 * (EB(i1, bx2) - EB(i1 - 1, bx2));
      }
    } else if constexpr (D == 2) {
      for (long i2 = G; i2 < g->n[1] + G; ++i2) {
        for (long i1 = G; i1 < g->n[0] + G; ++i1) {
          EB(i1, i2, ex1) += coeff1 * (EB(i1, i2, bx3) - EB(i1, i2 - 1, bx3));
          EB(i1, i2, ex2) += coeff1 * (EB(i1 - 1, i2, bx3) - EB(i1, i2, bx3));
          EB(i1, i2, ex3) += coeff2 * (EB(i1, i2 - 1, bx1) - EB(i1, i2, bx1) +
                                       EB(i1, i2, bx2) - EB(i1 - 1, i2, bx2));
        }
      }
    } else {
      for (long i3 = G; i3 < g->n[2] + G; ++i3) {
        for (long i2 = G; i2 < g->n[1] + G; ++i2) {
          for (long i1 = G; i1 < g->n[0] + G; ++i1) {
            EB(i1, i2, i3, ex1) += coeff1 * (EB(i1, i2, i3 - 1, bx2) - EB(i1, i2, i3, bx2) +
                                             EB(i1, i2, i3, bx3) - EB(i1, i2 - 1, i3, bx3));
            EB(i1, i2, i3, ex2) += coeff1 * (EB(i1 - 1, i2, i3, bx3) - EB(i1, i2, i3, bx3) +
                                             EB(i1, i2, i3, bx1) - EB(i1, i2, i3 - 1, bx1));
            EB(i1, i2, i3, ex3) += coeff1 * (EB(i1, i2 - 1, i3, bx1) - EB(i1, i2, i3, bx1) +
                                             EB(i1, i2, i3, bx2) - EB(i1 - 1, i2, i3, bx2));
          }
        }
      }
    }
  }

  /* ---------------------------------------------------------------------- */
  /* CurrentsAmpere (no external current): src/kernels/ampere_mink.hpp:134-215 */
  /* ---------------------------------------------------------------------- */
  template <int D>
  void currents_ampere(const orc_grid_t* g, float* em, float* cur, float coeff, float ppc0) {
    Fld<D>    E(g, em), J(g, cur);
    const int G  = g->ng;
    const long n2 = (D > 1) ? g->n[1] : 1, n3 = (D > 2) ? g->n[2] : 1;
    for (long k = 0; k < n3; ++k) {
      for (long j = 0; j < n2; ++j) {
        for (long i = 0; i < g->n[0]; ++i) {
          for (int c = 0; c < 3; ++c) {
            float *e, *jc;
            if constexpr (D == 1) {
              e  = &E(i + G, c);
              jc = &J(i + G, c);
            } else if constexpr (D == 2) {
              e  = &E(i + G, j + G, c);
              jc = &J(i + G, j + G, c);
            } else {
              e  = &E(i + G, j + G, k + G, c);
              jc = &J(i + G, j + G, k + G, c);
            }
            *e  += *jc * coeff;
            *jc /= ppc0;
          }
        }
      }
    }
  }

  /* ---------------------------------------------------------------------- */
  /* Digital filter, Cartesian: src/kernels/digital_filter.hpp:99-388        */
  /* ---------------------------------------------------------------------- */
  template <int D>
  void filter_pass(const orc_grid_t* g, float* cur, const float* buff_, const int* fbc) {
    Fld<D>     array(g, cur);
    Fld<D>     buffer(g, const_cast<float*>(buff_));
    const int  G      = g->ng;
    const long i1_min = G, i2_min = G, i3_min = G;
    const long i1_max = g->n[0] + G;
    const long i2_max = (D > 1) ? g->n[1] + G : 0;
    const long i3_max = (D > 2) ? g->n[2] + G : 0;
    const bool c1min = fbc[0] == ORC_FBC_CONDUCTOR, c1max = fbc[1] == ORC_FBC_CONDUCTOR;
    const bool c2min = (D > 1) && fbc[2] == ORC_FBC_CONDUCTOR;
    const bool c2max = (D > 1) && fbc[3] == ORC_FBC_CONDUCTOR;
    const bool c3min = (D > 2) && fbc[4] == ORC_FBC_CONDUCTOR;
    const bool c3max = (D > 2) && fbc[5] == ORC_FBC_CONDUCTOR;

    if constexpr (D == 1) {
      for (long i1 = i1_min; i1 < i1_max; ++i1) {
        if ((c1min && i1 == i1_min) || (c1max && i1 == i1_max - 1)) {
          const long i1side = c1min ? (i1 + 1) : (i1 - 1);
          array(i1, jx1)    = (THREE * INV_4) * buffer(i1, jx1) + (INV_4)*buffer(i1side, jx1);
        } else if ((c1min && i1 == i1_min + 1) || (c1max && i1 == i1_max - 2)) {
          const long i1side = c1min ? (i1 + 1) : (i1 - 1);
          array(i1, jx1)    = INV_2 * buffer(i1, jx1) +
                           INV_4 * (buffer(i1 - 1, jx1) + buffer(i1 + 1, jx1));
          array(i1, jx2) = (INV_2)*buffer(i1, jx2) + (INV_4)*buffer(i1side, jx2);
          array(i1, jx3) = (INV_2)*buffer(i1, jx3) + (INV_4)*buffer(i1side, jx3);
        } else {
          for (int comp = 0; comp < 3; ++comp) {
            array(i1, comp) = INV_2 * buffer(i1, comp) +
                              INV_4 * (buffer(i1 - 1, comp) + buffer(i1 + 1, comp));
          }
        }
      }
    } else if constexpr (D == 2) {
      // digital_filter.hpp:20-26
      auto F_I1 = [&](int c, long i, long j) {
        return INV_2 * buffer(i, j, c) + INV_4 * (buffer(i - 1, j, c) + buffer(i + 1, j, c));
      };
      auto F_I2 = [&](int c, long i, long j) {
        return INV_2 * buffer(i, j, c) + INV_4 * (buffer(i, j - 1, c) + buffer(i, j + 1, c));
      };
      for (long i2 = i2_min; i2 < i2_max; ++i2) {
        for (long i1 = i1_min; i1 < i1_max; ++i1) {
          if ((c1min && i1 == i1_min) || (c1max && i1 == i1_max - 1)) {
            const long i1side  = c1min ? (i1 + 1) : (i1 - 1);
            array(i1, i2, jx1) = (THREE * INV_4) * (F_I2(jx1, i1, i2)) +
                                 (INV_4) * (F_I2(jx1, i1side, i2));
          } else if ((c1min && i1 == i1_min + 1) || (c1max && i1 == i1_max - 2)) {
            const long i1side  = c1min ? (i1 + 1) : (i1 - 1);
            array(i1, i2, jx1) = INV_2 * (F_I2(jx1, i1, i2)) +
                                 INV_4 * ((F_I2(jx1, i1 - 1, i2)) + (F_I2(jx1, i1 + 1, i2)));
            array(i1, i2, jx2) = INV_2 * (F_I2(jx2, i1, i2)) + INV_4 * (F_I2(jx2, i1side, i2));
            array(i1, i2, jx3) = INV_2 * (F_I2(jx3, i1, i2)) + INV_4 * (F_I2(jx3, i1side, i2));
          } else if ((c2min && i2 == i2_min) || (c2max && i2 == i2_max - 1)) {
            const long i2side  = c2min ? (i2 + 1) : (i2 - 1);
            array(i1, i2, jx2) = (THREE * INV_4) * (F_I1(jx2, i1, i2)) +
                                 (INV_4) * (F_I1(jx2, i1, i2side));
          } else if ((c2min && i2 == i2_min + 1) || (c2max && i2 == i2_max - 2)) {
            const long i2side  = c2min ? (i2 + 1) : (i2 - 1);
            array(i1, i2, jx1) = INV_2 * (F_I1(jx1, i1, i2)) + INV_4 * (F_I1(jx1, i1, i2side));
            array(i1, i2, jx2) = INV_2 * (F_I1(jx2, i1, i2)) +
                                 INV_4 * ((F_I1(jx2, i1, i2 - 1)) + (F_I1(jx2, i1, i2 + 1)));
            array(i1, i2, jx3) = INV_2 * (F_I1(jx3, i1, i2)) + INV_4 * (F_I1(jx3, i1, i2side));
          } else {
            for (int comp = 0; comp < 3; ++comp) {
              array(i1, i2, comp) = INV_4 * buffer(i1, i2, comp) +
                                    INV_8 * (buffer(i1 - 1, i2, comp) + buffer(i1 + 1, i2, comp) +
                                             buffer(i1, i2 - 1, comp) + buffer(i1, i2 + 1, comp)) +
                                    INV_16 * (buffer(i1 - 1, i2 - 1, comp) +
                                              buffer(i1 + 1, i2 + 1, comp) +
                                              buffer(i1 - 1, i2 + 1, comp) +
                                              buffer(i1 + 1, i2 - 1, comp));
            }
          }
        }
      }
    } else {
      // digital_filter.hpp:28-56
      auto F_I1_I2 = [&](int c, long i, long j, long k) {
        return INV_4 * buffer(i, j, k, c) +
               INV_8 * (buffer(i - 1, j, k, c) + buffer(i + 1, j, k, c) + buffer(i, j - 1, k, c) +
                        buffer(i, j + 1, k, c)) +
               INV_16 * (buffer(i - 1, j - 1, k, c) + buffer(i + 1, j + 1, k, c) +
                         buffer(i - 1, j + 1, k, c) + buffer(i + 1, j - 1, k, c));
      };
      auto F_I2_I3 = [&](int c, long i, long j, long k) {
        return INV_4 * buffer(i, j, k, c) +
               INV_8 * (buffer(i, j - 1, k, c) + buffer(i, j + 1, k, c) + buffer(i, j, k - 1, c) +
                        buffer(i, j, k + 1, c)) +
               INV_16 * (buffer(i, j - 1, k - 1, c) + buffer(i, j + 1, k + 1, c) +
                         buffer(i, j - 1, k + 1, c) + buffer(i, j + 1, k - 1, c));
      };
      auto F_I1_I3 = [&](int c, long i, long j, long k) {
        return INV_4 * buffer(i, j, k, c) +
               INV_8 * (buffer(i - 1, j, k, c) + buffer(i + 1, j, k, c) + buffer(i, j, k - 1, c) +
                        buffer(i, j, k + 1, c)) +
               INV_16 * (buffer(i - 1, j, k - 1, c) + buffer(i + 1, j, k + 1, c) +
                         buffer(i - 1, j, k + 1, c) + buffer(i + 1, j, k - 1, c));
      };
      for (long i3 = i3_min; i3 < i3_max; ++i3) {
        for (long i2 = i2_min; i2 < i2_max; ++i2) {
          for (long i1 = i1_min; i1 < i1_max; ++i1) {
            if ((c1min && i1 == i1_min) || (c1max && i1 == i1_max - 1)) {
              const long i1side      = c1min ? (i1 + 1) : (i1 - 1);
              array(i1, i2, i3, jx1) = (THREE * INV_4) * (F_I2_I3(jx1, i1, i2, i3)) +
                                       (INV_4) * (F_I2_I3(jx1, i1side, i2, i3));
            } else if ((c1min && i1 == i1_min + 1) || (c1max && i1 == i1_max - 2)) {
              const long i1side      = c1min ? (i1 + 1) : (i1 - 1);
              array(i1, i2, i3, jx1) = INV_2 * (F_I2_I3(jx1, i1, i2, i3)) +
                                       INV_4 * ((F_I2_I3(jx1, i1 - 1, i2, i3)) +
                                                (F_I2_I3(jx1, i1 + 1, i2, i3)));
              array(i1, i2, i3, jx2) = INV_2 * (F_I2_I3(jx2, i1, i2, i3)) +
                                       INV_4 * (F_I2_I3(jx2, i1side, i2, i3));
              array(i1, i2, i3, jx3) = INV_2 * (F_I2_I3(jx3, i1, i2, i3)) +
                                       INV_4 * (F_I2_I3(jx3, i1side, i2, i3));
            } else if ((c2min && i2 == i2_min) || (c2max && i2 == i2_max - 1)) {
              const long i2side      = c2min ? (i2 + 1) : (i2 - 1);
              array(i1, i2, i3, jx2) = (THREE * INV_4) * (F_I1_I3(jx2, i1, i2, i3)) +
                                       (INV_4) * (F_I1_I3(jx2, i1, i2side, i3));
            } else if ((c2min && i2 == i2_min + 1) || (c2max && i2 == i2_max - 2)) {
              const long i2side      = c2min ? (i2 + 1) : (i2 - 1);
              array(i1, i2, i3, jx1) = INV_2 * (F_I1_I3(jx1, i1, i2, i3)) +
                                       INV_4 * (F_I1_I3(jx1, i1, i2side, i3));
              array(i1, i2, i3, jx2) = INV_2 * (F_I1_I3(jx2, i1, i2, i3)) +
                                       INV_4 * ((F_I1_I3(jx2, i1, i2 - 1, i3)) +
                                                (F_I1_I3(jx2, i1, i2 + 1, i3)));
              array(i1, i2, i3, jx3) = INV_2 * (F_I1_I3(jx3, i1, i2, i3)) +
                                       INV_4 * (F_I1_I3(jx3, i1, i2side, i3));
            } else if ((c3min && i3 == i3_min) || (c3max && i3 == i3_max - 1)) {
              const long i3side      = c3min ? (i3 + 1) : (i3 - 1);
              array(i1, i2, i3, jx3) = (THREE * INV_4) * (F_I1_I2(jx3, i1, i2, i3)) +
                                       (INV_4) * (F_I1_I2(jx3, i1, i2, i3side));
            } else if ((c3min && i3 == i3_min + 1) || (c3max && i3 == i3_max - 2)) {
              const long i3side      = c3min ? (i3 + 1) : (i3 - 1);
              array(i1, i2, i3, jx1) = INV_2 * (F_I1_I2(jx1, i1, i2, i3)) +
                                       INV_4 * (F_I1_I2(jx1, i1, i2, i3side));
              array(i1, i2, i3, jx2) = INV_2 * (F_I1_I2(jx2, i1, i2, i3)) +
                                       INV_4 * (F_I1_I2(jx2, i1, i2, i3side));
              array(i1, i2, i3, jx3) = INV_2 * (F_I1_I2(jx3, i1, i2, i3)) +
                                       INV_4 * ((F_I1_I2(jx3, i1, i2, i3 - 1)) +
                                                (F_I1_I2(jx3, i1, i2, i3 + 1)));
            } else {
              // NB the reference's 1/32 group lists (0,0,+-1) a second time where the
              // (0,-+1,+-1) diagonals would be expected (digital_filter.hpp:358-369);
              // that is reproduced here because it changes the result.
              for (int comp = 0; comp < 3; ++comp) {
                array(i1, i2, i3, comp) =
                  INV_8 * buffer(i1, i2, i3, comp) +
                  INV_16 * (buffer(i1 - 1, i2, i3, comp) + buffer(i1 + 1, i2, i3, comp) +
                            buffer(i1, i2 - 1, i3, comp) + buffer(i1, i2 + 1, i3, comp) +
                            buffer(i1, i2, i3 - 1, comp) + buffer(i1, i2, i3 + 1, comp)) +
                  INV_32 * (buffer(i1 - 1, i2 - 1, i3, comp) + buffer(i1 + 1, i2 + 1, i3, comp) +
                            buffer(i1 - 1, i2 + 1, i3, comp) + buffer(i1 + 1, i2 - 1, i3, comp) +
                            buffer(i1, i2 - 1, i3 - 1, comp) + buffer(i1, i2 + 1, i3 + 1, comp) +
                            buffer(i1, i2, i3 - 1, comp) + buffer(i1, i2, i3 + 1, comp) +
                            buffer(i1 - 1, i2, i3 - 1, comp) + buffer(i1 + 1, i2, i3 + 1, comp) +
                            buffer(i1 - 1, i2, i3 + 1, comp) + buffer(i1 + 1, i2, i3 - 1, comp)) +
                  INV_64 *
                    (buffer(i1 - 1, i2 - 1, i3 - 1, comp) + buffer(i1 + 1, i2 + 1, i3 + 1, comp) +
                     buffer(i1 - 1, i2 + 1, i3 + 1, comp) + buffer(i1 + 1, i2 - 1, i3 - 1, comp) +
                     buffer(i1 - 1, i2 - 1, i3 + 1, comp) + buffer(i1 + 1, i2 + 1, i3 - 1, comp) +
                     buffer(i1 - 1, i2 + 1, i3 - 1, comp) + buffer(i1 + 1, i2 - 1, i3 + 1, comp));
              }
            }
          }
        }
      }
    }
  }

  /* ---------------------------------------------------------------------- */
  /* SR pusher, Minkowski: src/kernels/pushers/sr.hpp:117-1370               */
  /* ---------------------------------------------------------------------- */
  inline float dot3(float a1, float a2, float a3, float b1, float b2, float b3) {
    return a1 * b1 + a2 * b2 + a3 * b3; // DOT: numeric.h:82-83
  }

  inline float nsq(float a1, float a2, float a3) { return dot3(a1, a2, a3, a1, a2, a3); }

  inline float crs1(float, float a2, float a3, float, float b2, float b3) {
    return a2 * b3 - a3 * b2;
  }

  inline float crs2(float a1, float, float a3, float b1, float, float b3) {
    return a3 * b1 - a1 * b3;
  }

  inline float crs3(float a1, float a2, float, float b1, float b2, float) {
    return a1 * b2 - a2 * b1;
  }

  // sr.hpp:851-1370 (getInterpolatedEMFields)
  template <int D, int O>
  void interpolate(const Fld<D>& EB, int ng, const orc_prtls_t* P, uint32_t p, float* e0,
                   float* b0) {
    if constexpr (O == 0) {
      if constexpr (D == 1) {
        const int   i    = P->i1[p] + ng;
        const float dx1_ = P->dx1[p];
        const int   indx = static_cast<int>(dx1_ + HALF);
        float       c0, c1;
        const float ponpmx = ONE - dx1_, ponppx = dx1_;
        const float pondmx = static_cast<float>(indx + 1) - (dx1_ + HALF);
        const float pondpx = ONE - pondmx;
        c0    = EB(i - 1 + indx, ex1);
        c1    = EB(i + indx, ex1);
        e0[0] = c0 * pondmx + c1 * pondpx;
        c0    = EB(i, ex2);
        c1    = EB(i + 1, ex2);
        e0[1] = c0 * ponpmx + c1 * ponppx;
        c0    = EB(i, ex3);
        c1    = EB(i + 1, ex3);
        e0[2] = c0 * ponpmx + c1 * ponppx;
        c0    = EB(i, bx1);
        c1    = EB(i + 1, bx1);
        b0[0] = c0 * ponpmx + c1 * ponppx;
        c0    = EB(i - 1 + indx, bx2);
        c1    = EB(i + indx, bx2);
        b0[1] = c0 * pondmx + c1 * pondpx;
        c0    = EB(i - 1 + indx, bx3);
        c1    = EB(i + indx, bx3);
        b0[2] = c0 * pondmx + c1 * pondpx;
      } else if constexpr (D == 2) {
        const int   i = P->i1[p] + ng, j = P->i2[p] + ng;
        const float dx1_ = P->dx1[p], dx2_ = P->dx2[p];
        const int   indx = static_cast<int>(dx1_ + HALF);
        const int   indy = static_cast<int>(dx2_ + HALF);
        float       c000, c100, c010, c110, c00, c10;
        const float ponpmx = ONE - dx1_, ponppx = dx1_;
        const float ponpmy = ONE - dx2_, ponppy = dx2_;
        const float pondmx = static_cast<float>(indx + 1) - (dx1_ + HALF);
        const float pondpx = ONE - pondmx;
        const float pondmy = static_cast<float>(indy + 1) - (dx2_ + HALF);
        const float pondpy = ONE - pondmy;
        // Ex1 (dual, primal)
        c000  = EB(i - 1 + indx, j, ex1);
        c100  = EB(i + indx, j, ex1);
        c010  = EB(i - 1 + indx, j + 1, ex1);
        c110  = EB(i + indx, j + 1, ex1);
        c00   = c000 * pondmx + c100 * pondpx;
        c10   = c010 * pondmx + c110 * pondpx;
        e0[0] = c00 * ponpmy + c10 * ponppy;
        // Ex2 (primal, dual)
        c000  = EB(i, j - 1 + indy, ex2);
        c100  = EB(i + 1, j - 1 + indy, ex2);
        c010  = EB(i, j + indy, ex2);
        c110  = EB(i + 1, j + indy, ex2);
        c00   = c000 * ponpmx + c100 * ponppx;
        c10   = c010 * ponpmx + c110 * ponppx;
        e0[1] = c00 * pondmy + c10 * pondpy;
        // Ex3 (primal, primal)
        c000  = EB(i, j, ex3);
        c100  = EB(i + 1, j, ex3);
        c010  = EB(i, j + 1, ex3);
        c110  = EB(i + 1, j + 1, ex3);
        c00   = c000 * ponpmx + c100 * ponppx;
        c10   = c010 * ponpmx + c110 * ponppx;
        e0[2] = c00 * ponpmy + c10 * ponppy;
        // Bx1 (primal, dual)
        c000  = EB(i, j - 1 + indy, bx1);
        c100  = EB(i + 1, j - 1 + indy, bx1);
        c010  = EB(i, j + indy, bx1);
        c110  = EB(i + 1, j + indy, bx1);
        c00   = c000 * ponpmx + c100 * ponppx;
        c10   = c010 * ponpmx + c110 * ponppx;
        b0[0] = c00 * pondmy + c10 * pondpy;
        // Bx2 (dual, primal)
        c000  = EB(i - 1 + indx, j, bx2);
        c100  = EB(i + indx, j, bx2);
        c010  = EB(i - 1 + indx, j + 1, bx2);
        c110  = EB(i + indx, j + 1, bx2);
        c00   = c000 * pondmx + c100 * pondpx;
        c10   = c010 * pondmx + c110 * pondpx;
        b0[1] = c00 * ponpmy + c10 * ponppy;
        // Bx3 (dual, dual)
        c000  = EB(i - 1 + indx, j - 1 + indy, bx3);
        c100  = EB(i + indx, j - 1 + indy, bx3);
        c010  = EB(i - 1 + indx, j + indy, bx3);
        c110  = EB(i + indx, j + indy, bx3);
        c00   = c000 * pondmx + c100 * pondpx;
        c10   = c010 * pondmx + c110 * pondpx;
        b0[2] = c00 * pondmy + c10 * pondpy;
      } else {
        const int   i = P->i1[p] + ng, j = P->i2[p] + ng, k = P->i3[p] + ng;
        const float d[3] = { P->dx1[p], P->dx2[p], P->dx3[p] };
        int         ind[3];
        float       wp[3][2], wd[3][2]; // primal / dual weights {minus, plus}
        for (int a = 0; a < 3; ++a) {
          ind[a]   = static_cast<int>(d[a] + HALF);
          wp[a][0] = ONE - d[a];
          wp[a][1] = d[a];
          wd[a][0] = static_cast<float>(ind[a] + 1) - (d[a] + HALF);
          wd[a][1] = ONE - wd[a][0];
        }
        // generic trilinear with per-axis staggering; the nesting (x, then y, then z)
        // and operand order follow sr.hpp:1014-1116 exactly.
        // `wy_dual` is separate from `sy` because the reference weights Bx3 with the
        // PRIMAL x2 weights although it reads the dual-staggered x2 nodes
        // (sr.hpp:1102-1116); reproduced, it changes the result.
        auto tri = [&](int comp, bool sx, bool sy, bool sz, bool wy_dual) {
          const int    i0 = sx ? (i - 1 + ind[0]) : i;
          const int    j0 = sy ? (j - 1 + ind[1]) : j;
          const int    k0 = sz ? (k - 1 + ind[2]) : k;
          const float* wx = sx ? wd[0] : wp[0];
          const float* wy = wy_dual ? wd[1] : wp[1];
          const float* wz = sz ? wd[2] : wp[2];
          const float  c000 = EB(i0, j0, k0, comp), c100 = EB(i0 + 1, j0, k0, comp);
          const float  c010 = EB(i0, j0 + 1, k0, comp), c110 = EB(i0 + 1, j0 + 1, k0, comp);
          const float  c001 = EB(i0, j0, k0 + 1, comp), c101 = EB(i0 + 1, j0, k0 + 1, comp);
          const float  c011 = EB(i0, j0 + 1, k0 + 1, comp);
          const float  c111 = EB(i0 + 1, j0 + 1, k0 + 1, comp);
          const float  c00  = c000 * wx[0] + c100 * wx[1];
          const float  c10  = c010 * wx[0] + c110 * wx[1];
          const float  c0   = c00 * wy[0] + c10 * wy[1];
          const float  c01  = c001 * wx[0] + c101 * wx[1];
          const float  c11  = c011 * wx[0] + c111 * wx[1];
          const float  c1   = c01 * wy[0] + c11 * wy[1];
          return c0 * wz[0] + c1 * wz[1];
        };
        e0[0] = tri(ex1, true, false, false, false);
        e0[1] = tri(ex2, false, true, false, true);
        e0[2] = tri(ex3, false, false, true, false);
        b0[0] = tri(bx1, false, true, true, true);
        b0[1] = tri(bx2, true, false, true, false);
        b0[2] = tri(bx3, true, true, false, false);
      }
    } else {
      // O >= 1: tensor-product splines, separate primal/dual weights (sr.hpp:1118-1369)
      int   pmin[3] = { 0, 0, 0 }, dmin[3] = { 0, 0, 0 };
      float Sp[3][O + 1], Sd[3][O + 1];
      const int   ii[3] = { P->i1[p] + ng, D > 1 ? P->i2[p] + ng : 0, D > 2 ? P->i3[p] + ng : 0 };
      const float dd[3] = { P->dx1[p], D > 1 ? P->dx2[p] : ZERO, D > 2 ? P->dx3[p] : ZERO };
      for (int a = 0; a < D; ++a) {
        shape_order<false, O>(ii[a], dd[a], pmin[a], Sp[a]);
        shape_order<true, O>(ii[a], dd[a], dmin[a], Sd[a]);
      }
      auto gather = [&](int comp, bool sx, bool sy, bool sz) {
        const float* S1 = sx ? Sd[0] : Sp[0];
        const int    m1 = sx ? dmin[0] : pmin[0];
        if constexpr (D == 1) {
          float r = ZERO;
          for (int a = 0; a < O + 1; a++) r += S1[a] * EB(m1 + a, comp);
          return r;
        } else if constexpr (D == 2) {
          const float* S2 = sy ? Sd[1] : Sp[1];
          const int    m2 = sy ? dmin[1] : pmin[1];
          float        r  = ZERO;
          for (int b = 0; b < O + 1; b++) {
            float c0 = ZERO;
            for (int a = 0; a < O + 1; a++) c0 += S1[a] * EB(m1 + a, m2 + b, comp);
            r += c0 * S2[b];
          }
          return r;
        } else {
          const float* S2 = sy ? Sd[1] : Sp[1];
          const int    m2 = sy ? dmin[1] : pmin[1];
          const float* S3_ = sz ? Sd[2] : Sp[2];
          const int    m3 = sz ? dmin[2] : pmin[2];
          float        r  = ZERO;
          for (int c = 0; c < O + 1; c++) {
            float c0 = ZERO;
            for (int b = 0; b < O + 1; b++) {
              float c00 = ZERO;
              for (int a = 0; a < O + 1; a++) c00 += S1[a] * EB(m1 + a, m2 + b, m3 + c, comp);
              c0 += c00 * S2[b];
            }
            r += c0 * S3_[c];
          }
          return r;
        }
      };
      e0[0] = gather(ex1, true, false, false);
      e0[1] = gather(ex2, false, true, false);
      e0[2] = gather(ex3, false, false, true);
      b0[0] = gather(bx1, false, true, true);
      b0[1] = gather(bx2, true, false, true);
      b0[2] = gather(bx3, true, true, false);
    }
  }

  // sr.hpp:337-368
  inline void boris(float ndh, float* u, float* e0, float* b0) {
    float COEFF = ndh;
    e0[0] *= COEFF;
    e0[1] *= COEFF;
    e0[2] *= COEFF;
    float u0[3] = { u[0] + e0[0], u[1] + e0[1], u[2] + e0[2] };
    COEFF *= ONE / std::sqrt(ONE + nsq(u0[0], u0[1], u0[2]));
    b0[0] *= COEFF;
    b0[1] *= COEFF;
    b0[2] *= COEFF;
    COEFF = TWO / (ONE + nsq(b0[0], b0[1], b0[2]));
    const float u1[3] = {
      (u0[0] + crs1(u0[0], u0[1], u0[2], b0[0], b0[1], b0[2])) * COEFF,
      (u0[1] + crs2(u0[0], u0[1], u0[2], b0[0], b0[1], b0[2])) * COEFF,
      (u0[2] + crs3(u0[0], u0[1], u0[2], b0[0], b0[1], b0[2])) * COEFF
    };
    u0[0] += crs1(u1[0], u1[1], u1[2], b0[0], b0[1], b0[2]) + e0[0];
    u0[1] += crs2(u1[0], u1[1], u1[2], b0[0], b0[1], b0[2]) + e0[1];
    u0[2] += crs3(u1[0], u1[1], u1[2], b0[0], b0[1], b0[2]) + e0[2];
    u[0] = u0[0];
    u[1] = u0[1];
    u[2] = u0[2];
  }

  // sr.hpp:370-437
  inline void vay(float ndh, float* u, float* e0, float* b0) {
    float COEFF = ndh;
    e0[0] *= COEFF;
    e0[1] *= COEFF;
    e0[2] *= COEFF;
    b0[0] *= COEFF;
    b0[1] *= COEFF;
    b0[2] *= COEFF;
    COEFF = ONE / std::sqrt(ONE + nsq(u[0], u[1], u[2]));
    const float u1[3] = {
      (u[0] + TWO * e0[0] + crs1(u[0], u[1], u[2], b0[0], b0[1], b0[2]) * COEFF),
      (u[1] + TWO * e0[1] + crs2(u[0], u[1], u[2], b0[0], b0[1], b0[2]) * COEFF),
      (u[2] + TWO * e0[2] + crs3(u[0], u[1], u[2], b0[0], b0[1], b0[2]) * COEFF)
    };
    COEFF        = dot3(u1[0], u1[1], u1[2], b0[0], b0[1], b0[2]);
    float COEFF2 = ONE + nsq(u1[0], u1[1], u1[2]) - nsq(b0[0], b0[1], b0[2]);
    COEFF        = ONE / std::sqrt(INV_2 * (COEFF2 + std::sqrt(SQR(COEFF2) +
                                                        FOUR * (SQR(b0[0]) + SQR(b0[1]) +
                                                                SQR(b0[2]) + SQR(COEFF)))));
    COEFF2 = ONE / (ONE + SQR(b0[0] * COEFF) + SQR(b0[1] * COEFF) + SQR(b0[2] * COEFF));
    const float udb = dot3(u1[0], u1[1], u1[2], b0[0], b0[1], b0[2]);
    u[0] = COEFF2 * (u1[0] + COEFF * udb * (b0[0] * COEFF) + u1[1] * b0[2] * COEFF -
                     u1[2] * b0[1] * COEFF);
    u[1] = COEFF2 * (u1[1] + COEFF * udb * (b0[1] * COEFF) + u1[2] * b0[0] * COEFF -
                     u1[0] * b0[2] * COEFF);
    u[2] = COEFF2 * (u1[2] + COEFF * udb * (b0[2] * COEFF) + u1[0] * b0[1] * COEFF -
                     u1[1] * b0[0] * COEFF);
  }

  // sr.hpp:439-521 (f0 == nullptr: no external force)
  inline void gca(float ndh, float dt, float* u, const float* f0, float* e0, float* b0) {
    const float eb_sqr = nsq(e0[0], e0[1], e0[2]) + nsq(b0[0], b0[1], b0[2]);
    const float wE[3]  = { crs1(e0[0], e0[1], e0[2], b0[0], b0[1], b0[2]) / eb_sqr,
                           crs2(e0[0], e0[1], e0[2], b0[0], b0[1], b0[2]) / eb_sqr,
                           crs3(e0[0], e0[1], e0[2], b0[0], b0[1], b0[2]) / eb_sqr };
    {
      const float b_norm_inv = ONE / std::sqrt(nsq(b0[0], b0[1], b0[2]));
      b0[0] *= b_norm_inv;
      b0[1] *= b_norm_inv;
      b0[2] *= b_norm_inv;
    }
    float upar = dot3(u[0], u[1], u[2], b0[0], b0[1], b0[2]) +
                 ndh * TWO * dot3(e0[0], e0[1], e0[2], b0[0], b0[1], b0[2]);
    if (f0 != nullptr) {
      upar = dot3(u[0], u[1], u[2], b0[0], b0[1], b0[2]) +
             ndh * TWO * dot3(e0[0], e0[1], e0[2], b0[0], b0[1], b0[2]) +
             dt * dot3(f0[0], f0[1], f0[2], b0[0], b0[1], b0[2]);
    }
    float factor;
    {
      const float wE_sqr = nsq(wE[0], wE[1], wE[2]);
      if (wE_sqr < 0.01f) {
        factor = ONE + wE_sqr + TWO * SQR(wE_sqr) + FIVE * SQR(wE_sqr) * wE_sqr;
      } else {
        factor = (ONE - std::sqrt(ONE - FOUR * wE_sqr)) / (TWO * wE_sqr);
      }
    }
    const float vE[3] = { wE[0] * factor, wE[1] * factor, wE[2] * factor };
    const float Gamma = std::sqrt(ONE + SQR(upar)) /
                        std::sqrt(ONE - nsq(vE[0], vE[1], vE[2]));
    u[0] = upar * b0[0] + vE[0] * Gamma;
    u[1] = upar * b0[1] + vE[1] * Gamma;
    u[2] = upar * b0[2] + vE[2] * Gamma;
  }

  // sr.hpp:1372-1424
  inline void synchrotron_drag(float coeff, float* u, float* u_prime, const float* e0,
                               const float* b0) {
    float gamma_prime_sqr = ONE / std::sqrt(ONE + nsq(u_prime[0], u_prime[1], u_prime[2]));
    u_prime[0] *= gamma_prime_sqr;
    u_prime[1] *= gamma_prime_sqr;
    u_prime[2] *= gamma_prime_sqr;
    gamma_prime_sqr        = SQR(ONE / gamma_prime_sqr);
    const float beta_dot_e = dot3(u_prime[0], u_prime[1], u_prime[2], e0[0], e0[1], e0[2]);
    const float epb[3]     = {
      e0[0] + crs1(u_prime[0], u_prime[1], u_prime[2], b0[0], b0[1], b0[2]),
      e0[1] + crs2(u_prime[0], u_prime[1], u_prime[2], b0[0], b0[1], b0[2]),
      e0[2] + crs3(u_prime[0], u_prime[1], u_prime[2], b0[0], b0[1], b0[2])
    };
    const float kappaR[3] = {
      crs1(epb[0], epb[1], epb[2], b0[0], b0[1], b0[2]) + beta_dot_e * e0[0],
      crs2(epb[0], epb[1], epb[2], b0[0], b0[1], b0[2]) + beta_dot_e * e0[1],
      crs3(epb[0], epb[1], epb[2], b0[0], b0[1], b0[2]) + beta_dot_e * e0[2],
    };
    const float chiR_sqr = nsq(epb[0], epb[1], epb[2]) - SQR(beta_dot_e);
    u[0] += coeff * (kappaR[0] - gamma_prime_sqr * u_prime[0] * chiR_sqr);
    u[1] += coeff * (kappaR[1] - gamma_prime_sqr * u_prime[1] * chiR_sqr);
    u[2] += coeff * (kappaR[2] - gamma_prime_sqr * u_prime[2] * chiR_sqr);
  }

  // sr.hpp:1487-1499
  inline void compton_drag(float coeff, float* u, float* u_prime) {
    float gamma_prime_sqr = ONE / std::sqrt(ONE + nsq(u_prime[0], u_prime[1], u_prime[2]));
    u_prime[0] *= gamma_prime_sqr;
    u_prime[1] *= gamma_prime_sqr;
    u_prime[2] *= gamma_prime_sqr;
    gamma_prime_sqr = SQR(ONE / gamma_prime_sqr);
    u[0] -= coeff * gamma_prime_sqr * u_prime[0];
    u[1] -= coeff * gamma_prime_sqr * u_prime[1];
    u[2] -= coeff * gamma_prime_sqr * u_prime[2];
  }

  // mpi::SendTag, src/global/arch/mpi_tags.h:175-233: tag = 2 + lexicographic index of
  // the direction in {-1,0,1}^D with the null direction skipped; tag*1 when staying.
  template <int D>
  inline short send_tag(short tag, const int* dir) {
    int lin = 0, centre = 0;
    for (int a = 0; a < D; ++a) {
      lin    = lin * 3 + (dir[a] + 1);
      centre = centre * 3 + 1;
    }
    if (lin == centre) {
      return tag;
    }
    const int t = 2 + lin - (lin > centre ? 1 : 0);
    return static_cast<short>(((t - 1) + 1) * tag);
  }

  // sr.hpp:659-814 for one axis (Cartesian): returns nothing, mutates particle
  inline void bc_axis(int* i, int* i_prev, float* dx, float* u, short* tag, int ni, int bcmin,
                      int bcmax) {
    bool invert_vel = false;
    if (*i < 0) {
      if (bcmin == ORC_PBC_PERIODIC) {
        *i      += ni;
        *i_prev += ni;
      } else if (bcmin == ORC_PBC_ABSORB) {
        *tag = 0;
      } else if (bcmin == ORC_PBC_REFLECT) {
        *i         = 0;
        *dx        = ONE - *dx;
        invert_vel = true;
      } else if (bcmin == ORC_PBC_AXIS) {
        *i  = 0;
        *dx = ONE - *dx;
      }
    } else if (*i >= ni) {
      if (bcmax == ORC_PBC_PERIODIC) {
        *i      -= ni;
        *i_prev -= ni;
      } else if (bcmax == ORC_PBC_ABSORB) {
        *tag = 0;
      } else if (bcmax == ORC_PBC_REFLECT) {
        *i         = ni - 1;
        *dx        = ONE - *dx;
        invert_vel = true;
      } else if (bcmax == ORC_PBC_AXIS) {
        *i  = ni - 1;
        *dx = ONE - *dx;
      }
    }
    if (invert_vel) {
      *u = -*u;
    }
  }

  // sr.hpp:117-332 + 526-657 for M = Minkowski<D>, P = NoPolicy (+ optional atmosphere)
  template <int D, int O>
  void push_sr(const orc_grid_t* g, const orc_pusher_t* ctx, const orc_prtls_t* P, uint32_t npart,
               const float* em) {
    Fld<D>      EB(g, const_cast<float*>(em));
    const float ndh = HALF * (ctx->charge / ctx->mass) * ctx->omegaB0 * ctx->dt; // sr.hpp:111
    const float dt  = ctx->dt;
    const float dxc = ctx->dx;
    int*        ip[3]  = { P->i1, P->i2, P->i3 };
    int*        ipp[3] = { P->i1_prev, P->i2_prev, P->i3_prev };
    float*      dp[3]  = { P->dx1, P->dx2, P->dx3 };
    float*      dpp[3] = { P->dx1_prev, P->dx2_prev, P->dx3_prev };
    float*      up[3]  = { P->ux1, P->ux2, P->ux3 };

    for (uint32_t p = 0; p < npart; ++p) {
      if (P->tag[p] != 1) {
        continue; // sr.hpp:118-123 (invalid tags abort in the reference)
      }
      float u[3]  = { P->ux1[p], P->ux2[p], P->ux3[p] };
      bool  massive = true;
      if (ctx->pusher_flags == ORC_PUSHER_PHOTON) {
        massive = false; // sr.hpp:126-161 without emission: position push only
      } else {
        float ei[3] = { ZERO, ZERO, ZERO }, bi[3] = { ZERO, ZERO, ZERO };
        float ec[3], bc[3];
        float fext[3] = { ZERO, ZERO, ZERO };
        float u_prime[3] = { ZERO, ZERO, ZERO }, e_rad[3] = { ZERO, ZERO, ZERO },
              b_rad[3] = { ZERO, ZERO, ZERO };
        bool  is_gca = false;
        interpolate<D, O>(EB, g->ng, P, p, ei, bi);
        // transform_xyz<U,XYZ>: minkowski.h:263-299 -> in-plane components times dx
        for (int a = 0; a < 3; ++a) {
          ec[a] = (a < D) ? ei[a] * dxc : ei[a];
          bc[a] = (a < D) ? bi[a] * dxc : bi[a];
        }
        if (ctx->drag_flags != ORC_DRAG_NONE) { // sr.hpp:234-244
          for (int a = 0; a < 3; ++a) {
            e_rad[a]   = ec[a];
            b_rad[a]   = bc[a];
            u_prime[a] = u[a];
          }
        }
        if (ctx->has_atmosphere) { // sr.hpp:1426-1485 (Cartesian branch)
          float      f[3] = { ZERO, ZERO, ZERO };
          const float gg[3] = { ctx->atm_gx1, ctx->atm_gx2, ctx->atm_gx3 };
          for (int a = 0; a < D; ++a) {
            const float xcd = static_cast<float>(ip[a][p]) + static_cast<float>(dp[a][p]);
            const float xph = xcd * dxc + ctx->xmin[a]; // convert<Cd,Ph>: minkowski.h:156-180
            if (!(std::fabs(gg[a]) <= std::numeric_limits<float>::epsilon()) &&
                ((ctx->atm_ds < ZERO || xph <= ctx->atm_x_surf + ctx->atm_ds) &&
                 (ctx->atm_ds > ZERO || xph >= ctx->atm_x_surf + ctx->atm_ds))) {
              f[a] += gg[a];
            }
          }
          for (int a = 0; a < 3; ++a) fext[a] = f[a]; // transform_xyz<T,XYZ> is identity
        }
        auto conventional = [&]() {
          if (ctx->has_atmosphere) {
            u[0] += HALF * dt * fext[0];
            u[1] += HALF * dt * fext[1];
            u[2] += HALF * dt * fext[2];
          }
          if (ctx->pusher_flags & ORC_PUSHER_BORIS) {
            boris(ndh, u, ec, bc);
          } else if (ctx->pusher_flags & ORC_PUSHER_VAY) {
            vay(ndh, u, ec, bc);
          }
          if (ctx->has_atmosphere) {
            u[0] += HALF * dt * fext[0];
            u[1] += HALF * dt * fext[1];
            u[2] += HALF * dt * fext[2];
          }
        };
        if (ctx->pusher_flags & ORC_PUSHER_GCA) { // sr.hpp:251-288
          const float E2 = nsq(ec[0], ec[1], ec[2]);
          const float B2 = nsq(bc[0], bc[1], bc[2]);
          const float rL = std::sqrt(ONE + nsq(u[0], u[1], u[2])) * dt /
                           (TWO * std::fabs(ndh) * std::sqrt(B2));
          if (B2 > ZERO && rL < ctx->gca_larmor_max && (E2 / B2) < ctx->gca_e_ovr_b_sqr_max) {
            is_gca = true;
            gca(ndh, dt, u, ctx->has_atmosphere ? fext : nullptr, ec, bc);
          } else {
            conventional();
          }
        } else {
          conventional();
        }
        if (!is_gca && ctx->drag_flags != ORC_DRAG_NONE) { // sr.hpp:311-322
          u_prime[0] = HALF * (u_prime[0] + u[0]);
          u_prime[1] = HALF * (u_prime[1] + u[1]);
          u_prime[2] = HALF * (u_prime[2] + u[2]);
          if (ctx->drag_flags & ORC_DRAG_SYNCHROTRON) {
            synchrotron_drag(ctx->sync_coeff, u, u_prime, e_rad, b_rad);
          }
          if (ctx->drag_flags & ORC_DRAG_COMPTON) {
            compton_drag(ctx->compton_coeff, u, u_prime);
          }
        }
      }
      // positionPush, Cartesian branch: sr.hpp:526-572
      const float dt_inv_energy = massive
                                    ? (dt / std::sqrt(ONE + SQR(u[0]) + SQR(u[1]) + SQR(u[2])))
                                    : (dt / std::sqrt(SQR(u[0]) + SQR(u[1]) + SQR(u[2])));
      short       tag           = P->tag[p];
      for (int a = 0; a < D; ++a) {
        int   i  = ip[a][p];
        float dx = dp[a][p];
        ipp[a][p] = i;
        dpp[a][p] = dx;
        dx += (u[a] / dxc) * dt_inv_energy; // transform<XYZ,U> = v / sqrt(h_ii)
        i  += static_cast<int>(dx >= ONE) - static_cast<int>(dx < ZERO);
        dx -= (dx >= ONE);
        dx += (dx < ZERO);
        ip[a][p] = i;
        dp[a][p] = dx;
      }
      // boundaryConditions: sr.hpp:659-814
      const int ni[3] = { g->n[0], g->n[1], g->n[2] };
      for (int a = 0; a < D; ++a) {
        bc_axis(&ip[a][p], &ipp[a][p], &dp[a][p], &u[a], &tag, ni[a], ctx->pbc[2 * a],
                ctx->pbc[2 * a + 1]);
      }
      if (ctx->tag_outgoing) {
        int dir[3] = { 0, 0, 0 };
        for (int a = 0; a < D; ++a) {
          dir[a] = (ip[a][p] < 0) ? -1 : ((ip[a][p] >= ni[a]) ? 1 : 0);
        }
        tag = send_tag<D>(tag, dir);
      }
      P->tag[p] = tag;
      for (int a = 0; a < 3; ++a) up[a][p] = u[a];
    }
  }

  /* ---------------------------------------------------------------------- */
  /* Current deposit: src/kernels/currents_deposit.hpp:108-761 (SRPIC, Minkowski) */
  /* ---------------------------------------------------------------------- */
  template <int D, int O>
  void deposit(const orc_grid_t* g, const orc_prtls_t* P, uint32_t npart, float charge, float dt,
               float dxc, float* cur) {
    Fld<D>      J(g, cur);
    const int   G      = g->ng;
    const float inv_dt = ONE / dt;
    for (uint32_t p = 0; p < npart; ++p) {
      if (P->tag[p] == 0) {
        continue;
      }
      float vp[3];
      {
        // transform_xyz<XYZ,U> (minkowski.h:263-299): in-plane components divided by dx
        const float ux = P->ux1[p], uy = P->ux2[p], uz = P->ux3[p];
        vp[0] = (0 < D) ? ux / dxc : ux;
        vp[1] = (1 < D) ? uy / dxc : uy;
        vp[2] = (2 < D) ? uz / dxc : uz;
        const float inv_energy = ONE / std::sqrt(ONE + nsq(ux, uy, uz));
        if (std::isnan(vp[2]) || std::isinf(vp[2])) {
          vp[2] = ZERO;
        }
        vp[0] *= inv_energy;
        vp[1] *= inv_energy;
        vp[2] *= inv_energy;
      }
      const float coeff = P->weight[p] * charge;

      if constexpr (O == 0) {
        // zig-zag: currents_deposit.hpp:171-405
        const int   i1 = P->i1[p], i1p = P->i1_prev[p];
        const float dx1 = P->dx1[p], dx1p = P->dx1_prev[p];
        const float dxp_r_1 = static_cast<float>(i1 == i1p) * (dx1 + dx1p) * INV_2;
        const float Wx1_1   = INV_2 * (dxp_r_1 + dx1p + static_cast<float>(i1 > i1p));
        const float Wx1_2 =
          INV_2 * (dx1 + dxp_r_1 + static_cast<float>(static_cast<int>(i1 > i1p) + i1p - i1));
        const float Fx1_1 = (static_cast<float>(i1 > i1p) + dxp_r_1 - dx1p) * coeff * inv_dt;
        const float Fx1_2 =
          (static_cast<float>(i1 - i1p - static_cast<int>(i1 > i1p)) + dx1 - dxp_r_1) * coeff *
          inv_dt;
        if constexpr (D == 1) {
          const float Fx2_1 = HALF * vp[1] * coeff, Fx2_2 = HALF * vp[1] * coeff;
          const float Fx3_1 = HALF * vp[2] * coeff, Fx3_2 = HALF * vp[2] * coeff;
          J(i1p + G, jx1)     += Fx1_1;
          J(i1 + G, jx1)      += Fx1_2;
          J(i1p + G, jx2)     += Fx2_1 * (ONE - Wx1_1);
          J(i1p + G + 1, jx2) += Fx2_1 * Wx1_1;
          J(i1 + G, jx2)      += Fx2_2 * (ONE - Wx1_2);
          J(i1 + G + 1, jx2)  += Fx2_2 * Wx1_2;
          J(i1p + G, jx3)     += Fx3_1 * (ONE - Wx1_1);
          J(i1p + G + 1, jx3) += Fx3_1 * Wx1_1;
          J(i1 + G, jx3)      += Fx3_2 * (ONE - Wx1_2);
          J(i1 + G + 1, jx3)  += Fx3_2 * Wx1_2;
        } else {
          const int   i2 = P->i2[p], i2p = P->i2_prev[p];
          const float dx2 = P->dx2[p], dx2p = P->dx2_prev[p];
          const float dxp_r_2 = static_cast<float>(i2 == i2p) * (dx2 + dx2p) * INV_2;
          const float Wx2_1   = INV_2 * (dxp_r_2 + dx2p + static_cast<float>(i2 > i2p));
          const float Wx2_2 =
            INV_2 * (dx2 + dxp_r_2 + static_cast<float>(static_cast<int>(i2 > i2p) + i2p - i2));
          const float Fx2_1 = (static_cast<float>(i2 > i2p) + dxp_r_2 - dx2p) * coeff * inv_dt;
          const float Fx2_2 =
            (static_cast<float>(i2 - i2p - static_cast<int>(i2 > i2p)) + dx2 - dxp_r_2) * coeff *
            inv_dt;
          if constexpr (D == 2) {
            const float Fx3_1 = HALF * vp[2] * coeff, Fx3_2 = HALF * vp[2] * coeff;
            J(i1p + G, i2p + G, jx1)     += Fx1_1 * (ONE - Wx2_1);
            J(i1p + G, i2p + G + 1, jx1) += Fx1_1 * Wx2_1;
            J(i1 + G, i2 + G, jx1)       += Fx1_2 * (ONE - Wx2_2);
            J(i1 + G, i2 + G + 1, jx1)   += Fx1_2 * Wx2_2;

            J(i1p + G, i2p + G, jx2)     += Fx2_1 * (ONE - Wx1_1);
            J(i1p + G + 1, i2p + G, jx2) += Fx2_1 * Wx1_1;
            J(i1 + G, i2 + G, jx2)       += Fx2_2 * (ONE - Wx1_2);
            J(i1 + G + 1, i2 + G, jx2)   += Fx2_2 * Wx1_2;

            J(i1p + G, i2p + G, jx3)         += Fx3_1 * (ONE - Wx1_1) * (ONE - Wx2_1);
            J(i1p + G + 1, i2p + G, jx3)     += Fx3_1 * Wx1_1 * (ONE - Wx2_1);
            J(i1p + G, i2p + G + 1, jx3)     += Fx3_1 * (ONE - Wx1_1) * Wx2_1;
            J(i1p + G + 1, i2p + G + 1, jx3) += Fx3_1 * Wx1_1 * Wx2_1;

            J(i1 + G, i2 + G, jx3)         += Fx3_2 * (ONE - Wx1_2) * (ONE - Wx2_2);
            J(i1 + G + 1, i2 + G, jx3)     += Fx3_2 * Wx1_2 * (ONE - Wx2_2);
            J(i1 + G, i2 + G + 1, jx3)     += Fx3_2 * (ONE - Wx1_2) * Wx2_2;
            J(i1 + G + 1, i2 + G + 1, jx3) += Fx3_2 * Wx1_2 * Wx2_2;
          } else {
            const int   i3 = P->i3[p], i3p = P->i3_prev[p];
            const float dx3 = P->dx3[p], dx3p = P->dx3_prev[p];
            const float dxp_r_3 = static_cast<float>(i3 == i3p) * (dx3 + dx3p) * INV_2;
            const float Wx3_1   = INV_2 * (dxp_r_3 + dx3p + static_cast<float>(i3 > i3p));
            const float Wx3_2 =
              INV_2 * (dx3 + dxp_r_3 + static_cast<float>(static_cast<int>(i3 > i3p) + i3p - i3));
            const float Fx3_1 = (static_cast<float>(i3 > i3p) + dxp_r_3 - dx3p) * coeff * inv_dt;
            const float Fx3_2 =
              (static_cast<float>(i3 - i3p - static_cast<int>(i3 > i3p)) + dx3 - dxp_r_3) *
              coeff * inv_dt;
            const int a = i1p + G, b = i2p + G, c = i3p + G;
            const int A = i1 + G, B = i2 + G, C = i3 + G;
            J(a, b, c, jx1)         += Fx1_1 * (ONE - Wx2_1) * (ONE - Wx3_1);
            J(a, b + 1, c, jx1)     += Fx1_1 * Wx2_1 * (ONE - Wx3_1);
            J(a, b, c + 1, jx1)     += Fx1_1 * (ONE - Wx2_1) * Wx3_1;
            J(a, b + 1, c + 1, jx1) += Fx1_1 * Wx2_1 * Wx3_1;

            J(A, B, C, jx1)         += Fx1_2 * (ONE - Wx2_2) * (ONE - Wx3_2);
            J(A, B + 1, C, jx1)     += Fx1_2 * Wx2_2 * (ONE - Wx3_2);
            J(A, B, C + 1, jx1)     += Fx1_2 * (ONE - Wx2_2) * Wx3_2;
            J(A, B + 1, C + 1, jx1) += Fx1_2 * Wx2_2 * Wx3_2;

            J(a, b, c, jx2)         += Fx2_1 * (ONE - Wx1_1) * (ONE - Wx3_1);
            J(a + 1, b, c, jx2)     += Fx2_1 * Wx1_1 * (ONE - Wx3_1);
            J(a, b, c + 1, jx2)     += Fx2_1 * (ONE - Wx1_1) * Wx3_1;
            J(a + 1, b, c + 1, jx2) += Fx2_1 * Wx1_1 * Wx3_1;

            J(A, B, C, jx2)         += Fx2_2 * (ONE - Wx1_2) * (ONE - Wx3_2);
            J(A + 1, B, C, jx2)     += Fx2_2 * Wx1_2 * (ONE - Wx3_2);
            J(A, B, C + 1, jx2)     += Fx2_2 * (ONE - Wx1_2) * Wx3_2;
            J(A + 1, B, C + 1, jx2) += Fx2_2 * Wx1_2 * Wx3_2;

            J(a, b, c, jx3)         += Fx3_1 * (ONE - Wx1_1) * (ONE - Wx2_1);
            J(a + 1, b, c, jx3)     += Fx3_1 * Wx1_1 * (ONE - Wx2_1);
            J(a, b + 1, c, jx3)     += Fx3_1 * (ONE - Wx1_1) * Wx2_1;
            J(a + 1, b + 1, c, jx3) += Fx3_1 * Wx1_1 * Wx2_1;

            J(A, B, C, jx3)         += Fx3_2 * (ONE - Wx1_2) * (ONE - Wx2_2);
            J(A + 1, B, C, jx3)     += Fx3_2 * Wx1_2 * (ONE - Wx2_2);
            J(A, B + 1, C, jx3)     += Fx3_2 * (ONE - Wx1_2) * Wx2_2;
            J(A + 1, B + 1, C, jx3) += Fx3_2 * Wx1_2 * Wx2_2;
          }
        }
      } else {
        // Esirkepov: currents_deposit.hpp:406-754
        constexpr int N = O + 2;
        float iS_x1[N], fS_x1[N];
        int   i1_min, i1_max;
        for_deposit<O>(P->i1_prev[p], P->dx1_prev[p], P->i1[p], P->dx1[p], i1_min, i1_max, iS_x1,
                       fS_x1);
        if constexpr (D == 1) {
          float Wx1[N], Wx23[N];
          for (int i = 0; i < N; ++i) {
            Wx1[i]  = fS_x1[i] - iS_x1[i];
            Wx23[i] = HALF * (fS_x1[i] + iS_x1[i]);
          }
          float       jx1_[N];
          const float Qdx1dt = coeff * inv_dt;
          const float QVx2   = coeff * vp[1];
          const float QVx3   = coeff * vp[2];
          jx1_[0]            = -Qdx1dt * Wx1[0];
          for (int i = 1; i < N; ++i) {
            jx1_[i] = jx1_[i - 1] - Qdx1dt * Wx1[i];
          }
          i1_min += G;
          i1_max += G;
          const int di_x1 = i1_max - i1_min;
          for (int i = 0; i < di_x1; ++i) J(i1_min + i, jx1) += jx1_[i];
          for (int i = 0; i <= di_x1; ++i) J(i1_min + i, jx2) += QVx2 * Wx23[i];
          for (int i = 0; i <= di_x1; ++i) J(i1_min + i, jx3) += QVx3 * Wx23[i];
        } else if constexpr (D == 2) {
          float iS_x2[N], fS_x2[N];
          int   i2_min, i2_max;
          for_deposit<O>(P->i2_prev[p], P->dx2_prev[p], P->i2[p], P->dx2[p], i2_min, i2_max,
                         iS_x2, fS_x2);
          float Wx1[N][N], Wx2[N][N], Wx3[N][N];
          for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
              Wx1[i][j] = HALF * (fS_x1[i] - iS_x1[i]) * (fS_x2[j] + iS_x2[j]);
              Wx2[i][j] = HALF * (fS_x1[i] + iS_x1[i]) * (fS_x2[j] - iS_x2[j]);
              Wx3[i][j] = THIRD * (fS_x2[j] * (HALF * iS_x1[i] + fS_x1[i]) +
                                   iS_x2[j] * (HALF * fS_x1[i] + iS_x1[i]));
            }
          }
          float       jx1_[N][N], jx2_[N][N];
          const float Qdx1dt = coeff * inv_dt;
          const float Qdx2dt = coeff * inv_dt;
          const float QVx3   = coeff * vp[2];
          for (int j = 0; j < N; ++j) jx1_[0][j] = -Qdx1dt * Wx1[0][j];
          for (int i = 1; i < N; ++i) {
            for (int j = 0; j < N; ++j) jx1_[i][j] = jx1_[i - 1][j] - Qdx1dt * Wx1[i][j];
          }
          for (int i = 0; i < N; ++i) jx2_[i][0] = -Qdx2dt * Wx2[i][0];
          for (int j = 1; j < N; ++j) {
            for (int i = 0; i < N; ++i) jx2_[i][j] = jx2_[i][j - 1] - Qdx2dt * Wx2[i][j];
          }
          i1_min += G;
          i2_min += G;
          i1_max += G;
          i2_max += G;
          const int di_x1 = i1_max - i1_min, di_x2 = i2_max - i2_min;
          for (int i = 0; i < di_x1; ++i) {
            for (int j = 0; j <= di_x2; ++j) J(i1_min + i, i2_min + j, jx1) += jx1_[i][j];
          }
          for (int i = 0; i <= di_x1; ++i) {
            for (int j = 0; j < di_x2; ++j) J(i1_min + i, i2_min + j, jx2) += jx2_[i][j];
          }
          for (int i = 0; i <= di_x1; ++i) {
            for (int j = 0; j <= di_x2; ++j) J(i1_min + i, i2_min + j, jx3) += QVx3 * Wx3[i][j];
          }
        } else {
          float iS_x2[N], fS_x2[N], iS_x3[N], fS_x3[N];
          int   i2_min, i2_max, i3_min, i3_max;
          for_deposit<O>(P->i2_prev[p], P->dx2_prev[p], P->i2[p], P->dx2[p], i2_min, i2_max,
                         iS_x2, fS_x2);
          for_deposit<O>(P->i3_prev[p], P->dx3_prev[p], P->i3[p], P->dx3[p], i3_min, i3_max,
                         iS_x3, fS_x3);
          float Wx1[N][N][N], Wx2[N][N][N], Wx3[N][N][N];
          for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
              for (int k = 0; k < N; ++k) {
                Wx1[i][j][k] = THIRD * (fS_x1[i] - iS_x1[i]) *
                               ((iS_x2[j] * iS_x3[k] + fS_x2[j] * fS_x3[k]) +
                                HALF * (iS_x3[k] * fS_x2[j] + iS_x2[j] * fS_x3[k]));
                Wx2[i][j][k] = THIRD * (fS_x2[j] - iS_x2[j]) *
                               (iS_x1[i] * iS_x3[k] + fS_x1[i] * fS_x3[k] +
                                HALF * (iS_x3[k] * fS_x1[i] + iS_x1[i] * fS_x3[k]));
                Wx3[i][j][k] = THIRD * (fS_x3[k] - iS_x3[k]) *
                               (iS_x1[i] * iS_x2[j] + fS_x1[i] * fS_x2[j] +
                                HALF * (iS_x1[i] * fS_x2[j] + iS_x2[j] * fS_x1[i]));
              }
            }
          }
          float       jx1_[N][N][N], jx2_[N][N][N], jx3_[N][N][N];
          const float Qdxdt = coeff * inv_dt, Qdydt = coeff * inv_dt, Qdzdt = coeff * inv_dt;
          for (int j = 0; j < N; ++j) {
            for (int k = 0; k < N; ++k) jx1_[0][j][k] = -Qdxdt * Wx1[0][j][k];
          }
          for (int i = 1; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
              for (int k = 0; k < N; ++k) {
                jx1_[i][j][k] = jx1_[i - 1][j][k] - Qdxdt * Wx1[i][j][k];
              }
            }
          }
          for (int i = 0; i < N; ++i) {
            for (int k = 0; k < N; ++k) jx2_[i][0][k] = -Qdydt * Wx2[i][0][k];
          }
          for (int i = 0; i < N; ++i) {
            for (int j = 1; j < N; ++j) {
              for (int k = 0; k < N; ++k) {
                jx2_[i][j][k] = jx2_[i][j - 1][k] - Qdydt * Wx2[i][j][k];
              }
            }
          }
          for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) jx3_[i][j][0] = -Qdydt * Wx3[i][j][0]; // sic :697
          }
          for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
              for (int k = 1; k < N; ++k) {
                jx3_[i][j][k] = jx3_[i][j][k - 1] - Qdzdt * Wx3[i][j][k];
              }
            }
          }
          i1_min += G;
          i2_min += G;
          i3_min += G;
          i1_max += G;
          i2_max += G;
          i3_max += G;
          const int di_x1 = i1_max - i1_min, di_x2 = i2_max - i2_min, di_x3 = i3_max - i3_min;
          for (int i = 0; i < di_x1; ++i) {
            for (int j = 0; j <= di_x2; ++j) {
              for (int k = 0; k <= di_x3; ++k) {
                J(i1_min + i, i2_min + j, i3_min + k, jx1) += jx1_[i][j][k];
              }
            }
          }
          for (int i = 0; i <= di_x1; ++i) {
            for (int j = 0; j < di_x2; ++j) {
              for (int k = 0; k <= di_x3; ++k) {
                J(i1_min + i, i2_min + j, i3_min + k, jx2) += jx2_[i][j][k];
              }
            }
          }
          for (int i = 0; i <= di_x1; ++i) {
            for (int j = 0; j <= di_x2; ++j) {
              for (int k = 0; k < di_x3; ++k) {
                J(i1_min + i, i2_min + j, i3_min + k, jx3) += jx3_[i][j][k];
              }
            }
          }
        }
      }
    }
  }

  /* ---------------------------------------------------------------------- */
  /* Single-domain ghost exchange: metadomain_comm.cpp:122-195, 276-367;     */
  /* comm_nompi.hpp:29-119. Directions in dir::Directions<D>::all order       */
  /* (src/global/arch/directions.h:171-233) = lexicographic over {-1,0,1}^D.  */
  /* ---------------------------------------------------------------------- */
  struct Slice {
    long lo[3], hi[3];
  };

  template <int D>
  bool slices_for(const orc_grid_t* g, const int* dir, const int* fbc, bool sync, Slice& snd,
                  Slice& rcv) {
    // periodic self-communication only happens when the face the direction points
    // to is PERIODIC (GetSendRecvRanks, metadomain_comm.cpp:36-115). For a diagonal
    // direction the reference looks the BC up through mesh.flds_bc_in(direction),
    // which is PERIODIC only if every involved face is periodic.
    for (int a = 0; a < D; ++a) {
      if (dir[a] == 1 && fbc[2 * a + 1] != ORC_FBC_PERIODIC) return false;
      if (dir[a] == -1 && fbc[2 * a] != ORC_FBC_PERIODIC) return false;
    }
    const long G = g->ng;
    for (int a = 0; a < 3; ++a) {
      snd.lo[a] = rcv.lo[a] = 0;
      snd.hi[a] = rcv.hi[a] = 1;
    }
    for (int a = 0; a < D; ++a) {
      const long imin = G, imax = g->n[a] + G;
      const int  d = dir[a];
      if (!sync) {
        if (d == 0) {
          snd.lo[a] = imin, snd.hi[a] = imax;
          rcv.lo[a] = imin, rcv.hi[a] = imax;
        } else if (d == 1) {
          snd.lo[a] = imax - G, snd.hi[a] = imax;
          rcv.lo[a] = imin - G, rcv.hi[a] = imin; // -dir == -1
        } else {
          snd.lo[a] = imin, snd.hi[a] = imin + G;
          rcv.lo[a] = imax, rcv.hi[a] = imax + G; // -dir == +1
        }
      } else {
        if (d == 0) {
          snd.lo[a] = imin - G, snd.hi[a] = imax + G;
          rcv.lo[a] = imin - G, rcv.hi[a] = imax + G;
        } else if (d == 1) {
          snd.lo[a] = imax - G, snd.hi[a] = imax + G;
          rcv.lo[a] = imin - G, rcv.hi[a] = imin + G;
        } else {
          snd.lo[a] = imin - G, snd.hi[a] = imin + G;
          rcv.lo[a] = imax - G, rcv.hi[a] = imax + G;
        }
      }
    }
    return true;
  }

  template <int D>
  void comm_self(const orc_grid_t* g, float* src_, float* dst_, int c0, int c1, const int* fbc,
                 bool sync) {
    Fld<D> src(g, src_), dst(g, dst_);
    int    dir[3] = { 0, 0, 0 };
    int    ndirs  = 1;
    for (int a = 0; a < D; ++a) ndirs *= 3;
    for (int lin = 0; lin < ndirs; ++lin) {
      int  r = lin;
      bool zero = true;
      for (int a = D - 1; a >= 0; --a) {
        dir[a] = (r % 3) - 1;
        r     /= 3;
        zero   = zero && (dir[a] == 0);
      }
      if (zero) continue;
      Slice s, t;
      if (!slices_for<D>(g, dir, fbc, sync, s, t)) continue;
      for (int c = c0; c < c1; ++c) {
        for (long k = t.lo[2]; k < t.hi[2]; ++k) {
          for (long j = t.lo[1]; j < t.hi[1]; ++j) {
            for (long i = t.lo[0]; i < t.hi[0]; ++i) {
              const long si = i - (t.lo[0] - s.lo[0]);
              const long sj = j - (t.lo[1] - s.lo[1]);
              const long sk = k - (t.lo[2] - s.lo[2]);
              float      v;
              float*     d;
              if constexpr (D == 1) {
                v = src(si, c);
                d = &dst(i, c);
              } else if constexpr (D == 2) {
                v = src(si, sj, c);
                d = &dst(i, j, c);
              } else {
                v = src(si, sj, sk, c);
                d = &dst(i, j, k, c);
              }
              if (sync) {
                *d += v;
              } else {
                *d = v;
              }
            }
          }
        }
      }
    }
  }

  template <int D>
  void sync_currents_self(const orc_grid_t* g, float* cur, float* buff, const int* fbc) {
    Fld<D> J(g, cur), B(g, buff);
    const size_t ntot = static_cast<size_t>(J.N1) * J.N2 * J.N3 * 3;
    std::memset(buff, 0, ntot * sizeof(float)); // metadomain_comm.cpp:441
    comm_self<D>(g, cur, buff, 0, 3, fbc, true);
    // AddBufferedFields over active cells: metadomain_comm.cpp:369-406, 536-546
    const long G = g->ng;
    const long n2 = (D > 1) ? g->n[1] : 1, n3 = (D > 2) ? g->n[2] : 1;
    for (int c = 0; c < 3; ++c) {
      for (long k = 0; k < n3; ++k) {
        for (long j = 0; j < n2; ++j) {
          for (long i = 0; i < g->n[0]; ++i) {
            if constexpr (D == 1) {
              J(i + G, c) += B(i + G, c);
            } else if constexpr (D == 2) {
              J(i + G, j + G, c) += B(i + G, j + G, c);
            } else {
              J(i + G, j + G, k + G, c) += B(i + G, j + G, k + G, c);
            }
          }
        }
      }
    }
  }

} // namespace orc

#define DISPATCH_D(g, fn, ...)                                                                 \
  do {                                                                                         \
    if ((g)->dim == 1) {                                                                       \
      orc::fn<1>(__VA_ARGS__);                                                                 \
    } else if ((g)->dim == 2) {                                                                \
      orc::fn<2>(__VA_ARGS__);                                                                 \
    } else if ((g)->dim == 3) {                                                                \
      orc::fn<3>(__VA_ARGS__);                                                                 \
    } else {                                                                                   \
      std::abort();                                                                            \
    }                                                                                          \
  } while (0)

#define DISPATCH_DO(g, order, fn, ...)                                                         \
  do {                                                                                         \
    const int key_ = (g)->dim * 10 + (order);                                                  \
    switch (key_) {                                                                            \
      case 10: orc::fn<1, 0>(__VA_ARGS__); break;                                              \
      case 11: orc::fn<1, 1>(__VA_ARGS__); break;                                              \
      case 12: orc::fn<1, 2>(__VA_ARGS__); break;                                              \
      case 13: orc::fn<1, 3>(__VA_ARGS__); break;                                              \
      case 20: orc::fn<2, 0>(__VA_ARGS__); break;                                              \
      case 21: orc::fn<2, 1>(__VA_ARGS__); break;                                              \
      case 22: orc::fn<2, 2>(__VA_ARGS__); break;                                              \
      case 23: orc::fn<2, 3>(__VA_ARGS__); break;                                              \
      case 30: orc::fn<3, 0>(__VA_ARGS__); break;                                              \
      case 31: orc::fn<3, 1>(__VA_ARGS__); break;                                              \
      case 32: orc::fn<3, 2>(__VA_ARGS__); break;                                              \
      case 33: orc::fn<3, 3>(__VA_ARGS__); break;                                              \
      default: std::abort();                                                                   \
    }                                                                                          \
  } while (0)

extern "C" {

void orc_faraday_mink(const orc_grid_t* g, float* em, float coeff1, float coeff2,
                      const float* stencil9) {
  DISPATCH_D(g, faraday, g, em, coeff1, coeff2, stencil9);
}

void orc_ampere_mink(const orc_grid_t* g, float* em, float coeff1, float coeff2) {
  DISPATCH_D(g, ampere, g, em, coeff1, coeff2);
}

void orc_currents_ampere_mink(const orc_grid_t* g, float* em, float* cur, float coeff,
                              float ppc0) {
  DISPATCH_D(g, currents_ampere, g, em, cur, coeff, ppc0);
}

void orc_filter_pass(const orc_grid_t* g, float* cur, const float* buff, const int* fbc) {
  DISPATCH_D(g, filter_pass, g, cur, buff, fbc);
}

void orc_push_sr_mink(const orc_grid_t* g, int order, const orc_pusher_t* ctx,
                      const orc_prtls_t* p, uint32_t npart, const float* em) {
  DISPATCH_DO(g, order, push_sr, g, ctx, p, npart, em);
}

void orc_deposit_mink(const orc_grid_t* g, int order, const orc_prtls_t* p, uint32_t npart,
                      float charge, float dt, float dx, float* cur) {
  DISPATCH_DO(g, order, deposit, g, p, npart, charge, dt, dx, cur);
}

void orc_comm_fields_self(const orc_grid_t* g, float* fld, int ncomp, int c0, int c1,
                          const int* fbc) {
  (void)ncomp;
  DISPATCH_D(g, comm_self, g, fld, fld, c0, c1, fbc, false);
}

void orc_sync_currents_self(const orc_grid_t* g, float* cur, float* buff, const int* fbc) {
  DISPATCH_D(g, sync_currents_self, g, cur, buff, fbc);
}

} // extern "C"
