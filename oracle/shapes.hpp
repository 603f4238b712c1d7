// TEST INFRASTRUCTURE ONLY -- CPU oracle for the PIC hot path. Not shipped, not
// linked by the product library. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// Particle shape functions: restatement of the reference's
//   src/kernels/particle_shapes.hpp:25-62   (S1, S2, S3)
//   src/kernels/particle_shapes.hpp:544-613 (order<STAGGERED,O>, O = 1..3)
//   src/kernels/particle_shapes.hpp:934-1024 (for_deposit<O>)
// in plain scalar C++ (fp32, evaluation order kept so that results are
// bit-identical with the reference compiled without FMA contraction).
#pragma once
#include <cmath>

namespace orc {

  constexpr float ONE = 1.0f, TWO = 2.0f, THREE = 3.0f, FOUR = 4.0f, FIVE = 5.0f;
  constexpr float ZERO = 0.0f, HALF = 0.5f;
  constexpr float THIRD = 0.333333f; // sic: src/global/utils/numeric.h:41
  constexpr float THREE_FOURTHS = 0.75f, THREE_HALFS = 1.5f;
  constexpr float INV_2 = 0.5f, INV_4 = 0.25f, INV_8 = 0.125f, INV_16 = 0.0625f;
  constexpr float INV_32 = 0.03125f, INV_64 = 0.015625f;

  inline float SQR(float x) { return x * x; }
  inline float CUBE(float x) { return x * x * x; }

  // particle_shapes.hpp:55-64
  inline float S3(float x) {
    if (x < ONE) {
      return static_cast<float>(2.0 / 3.0) - SQR(x) + HALF * CUBE(x);
    } else if (x < TWO) {
      return static_cast<float>(4.0 / 3.0) - TWO * x + SQR(x) -
             static_cast<float>(1.0 / 6.0) * CUBE(x);
    } else {
      return ZERO;
    }
  }

  // particle_shapes.hpp:544-613
  template <bool STAGGERED, int O>
  inline void shape_order(int i, float di, int& i_min, float* S) {
    static_assert(O >= 1 && O <= 3, "oracle restates shape orders 1..3");
    if constexpr (O == 1) {
      if constexpr (!STAGGERED) {
        i_min = i;
        S[0]  = ONE - di;
        S[1]  = di;
      } else {
        if (di < HALF) {
          i_min = i - 1;
          S[0]  = HALF - di;
          S[1]  = ONE - S[0];
        } else {
          i_min = i;
          S[0]  = THREE_HALFS - di;
          S[1]  = ONE - S[0];
        }
      }
    } else if constexpr (O == 2) {
      if constexpr (!STAGGERED) {
        if (di < HALF) {
          i_min = i - 1;
          S[0]  = HALF * SQR(HALF - di);
          S[1]  = THREE_FOURTHS - SQR(di);
          S[2]  = ONE - S[0] - S[1];
        } else {
          i_min = i;
          S[0]  = HALF * SQR(THREE_HALFS - di);
          S[1]  = THREE_FOURTHS - SQR(ONE - di);
          S[2]  = ONE - S[0] - S[1];
        }
      } else {
        i_min = i - 1;
        S[0]  = HALF * SQR(ONE - di);
        S[2]  = HALF * SQR(di);
        S[1]  = ONE - S[0] - S[2];
      }
    } else {
      if constexpr (!STAGGERED) {
        i_min = i - 1;
        for (int n = 0; n < 4; n++) {
          S[n] = S3(std::fabs(ONE + di - static_cast<float>(n)));
        }
      } else {
        if (di < HALF) {
          i_min = i - 2;
          for (int n = 0; n < 4; n++) {
            S[n] = S3(std::fabs(1.5f + di - static_cast<float>(n)));
          }
        } else {
          i_min = i - 1;
          for (int n = 0; n < 4; n++) {
            S[n] = S3(std::fabs(HALF + di - static_cast<float>(n)));
          }
        }
      }
    }
  }

  // particle_shapes.hpp:934-1024
  template <int O>
  inline void for_deposit(int i_init, float di_init, int i_fin, float di_fin,
                          int& i_min, int& i_max, float* iS, float* fS) {
    int   i_init_min, i_fin_min;
    float iS_[O + 1], fS_[O + 1];
    shape_order<false, O>(i_init, di_init, i_init_min, iS_);
    shape_order<false, O>(i_fin, di_fin, i_fin_min, fS_);
    if (i_init_min < i_fin_min) {
      i_min = i_init_min;
      i_max = i_min + O + 1;
      for (int j = 0; j < O + 1; j++) iS[j] = iS_[j];
      iS[O + 1] = ZERO;
      fS[0]     = ZERO;
      for (int j = 0; j < O + 1; j++) fS[j + 1] = fS_[j];
    } else if (i_init_min > i_fin_min) {
      i_min = i_fin_min;
      i_max = i_min + O + 1;
      iS[0] = ZERO;
      for (int j = 0; j < O + 1; j++) iS[j + 1] = iS_[j];
      for (int j = 0; j < O + 1; j++) fS[j] = fS_[j];
      fS[O + 1] = ZERO;
    } else {
      i_min = i_init_min;
      i_max = i_min + O;
      for (int j = 0; j < O + 1; j++) iS[j] = iS_[j];
      iS[O + 1] = ZERO;
      for (int j = 0; j < O + 1; j++) fS[j] = fS_[j];
      fS[O + 1] = ZERO;
    }
  }

} // namespace orc
