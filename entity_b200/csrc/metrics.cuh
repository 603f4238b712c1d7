// entity_b200 -- curvilinear and GR metrics, evaluated in registers.
//
// One POD parameter block (filled on the host exactly like the reference's constructors do,
// in fp32) and one stateless policy class per metric. Everything is 2D axisymmetric: the
// reference's curvilinear metrics are (static_assert) 2D only.
//   Spherical    src/metrics/spherical.h:38-403
//   QSpherical   src/metrics/qspherical.h:32-497
//   KerrSchild   src/metrics/kerr_schild.h:32-760
//   QKerrSchild  src/metrics/qkerr_schild.h:32-916
//   KerrSchild0  src/metrics/kerr_schild_0.h:32-591
// Coordinates: x1, x2 are code units (cell index + offset); "Ph"/"Sph" are r, theta.
// The reference mixes double constants (constant::PI ...) into fp32 expressions; here they are
// fp32 except inside theta2eta (cubic inversion, double like the reference). Parity against the
// reference is therefore to fp32 rounding, not bit-exact (see tests/test_gpu_curvilinear.py).
#pragma once
#include <math.h>

#include "common.cuh"

namespace eb200 {

  constexpr float PI_F         = 3.14159265358979323846f;
  constexpr float HALF_PI_F    = 1.57079632679489661923f;
  constexpr float TWO_PI_F     = 6.28318530717958647692f;
  constexpr float INV_PI_F     = 0.31830988618379067154f;
  constexpr float INV_PI_SQR_F = 0.10132118364233777144f;
  constexpr float PI_SQR_F     = 9.86960440108935861882f;
  constexpr float SMALL_ANGLE_F    = 1e-3f;
  constexpr float SMALL_ANGLE_GR_F = 1e-5f;
  constexpr float EPS_F = 1.1920929e-07f; // Kokkos::Experimental::epsilon<float>

#define EB200_HD __host__ __device__ __forceinline__

  struct MetricParams {
    int   kind;        // EB200_METRIC_*
    float nx1, nx2;    // active cells of the metric's mesh (the local domain)
    float x1min, x1max, x2min, x2max;
    float r0, h, a;    // qspherical r0 / h, Kerr spin
    // derived
    float d1, d2, d1_inv, d2_inv; // dr|dchi, dtheta|deta and their inverses
    float chi_min, eta_min;
    int   small_angle;
  };

  /* -------------------------------------------------- angular stretching (qspherical.h:440-493) */
  EB200_HD float q_dtheta_deta(float h, float eta) {
    if (fabsf(h) <= EPS_F) {
      return ONE;
    }
    return ONE + TWO * h + 12.0f * h * (eta * INV_PI_F) * ((eta * INV_PI_F) - ONE);
  }

  EB200_HD float q_eta2theta(float h, float eta) {
    if (fabsf(h) <= EPS_F) {
      return eta;
    }
    return eta + TWO * h * eta * (PI_F - TWO * eta) * (PI_F - eta) * INV_PI_SQR_F;
  }

  // The reference's expression mixes fp32 sub-expressions (SQR(h), CUBE(h), SQR(theta), h - ONE)
  // into a double formula (qspherical.h:462-493); the same promotions are kept here.
  EB200_HD float q_theta2eta(float hf, float thetaf) {
    if (fabsf(hf) <= EPS_F) {
      return thetaf;
    }
    const double PI = 3.14159265358979323846, TWO_PI = 6.28318530717958647692;
    const double SQRT3 = 1.73205080756887729352;
    const double h = hf, theta = thetaf;
    const double h2 = static_cast<double>(hf * hf), h3 = static_cast<double>(hf * hf * hf);
    const double th2 = static_cast<double>(thetaf * thetaf);
    const double hm1 = static_cast<double>(hf - ONE);
    const double R = pow(-9.0 * h2 * (PI - 2.0 * theta) +
                           SQRT3 * sqrt(h3 * ((4.0 - h) * ((PI + h * TWO_PI) * (PI + h * TWO_PI)) -
                                              108.0 * h * PI * theta + 108.0 * h * th2)),
                         1.0 / 3.0);
    const double PI_TO_TWO_THIRD = 2.14502939711102560008, PI_TO_ONE_THIRD = 1.46459188756152326302;
    const double TWO_TO_TWO_THIRD = 1.58740105196819947475,
                 THREE_TO_ONE_THIRD = 1.442249570307408382321;
    const double TWO_TO_ONE_THIRD = 1.2599210498948731647672,
                 THREE_PI_TO_TWO_THIRD = 4.46184094890142313715794;
    return static_cast<float>(PI_TO_TWO_THIRD *
                              (6.0 * PI_TO_ONE_THIRD +
                               2.0 * TWO_TO_ONE_THIRD * hm1 * THREE_PI_TO_TWO_THIRD / R +
                               TWO_TO_TWO_THIRD * THREE_TO_ONE_THIRD * R / h) /
                              12.0);
  }

  // qspherical.h:441-460 keeps constant::PI etc. as doubles (qkerr_schild.h:843-866 casts them
  // to real_t): two spellings of the same maps, each followed to the letter
  EB200_HD float qs_dtheta_deta(float h, float eta) {
    if (fabsf(h) <= EPS_F) {
      return ONE;
    }
    const double INV_PI = 0.31830988618379067154;
    return static_cast<float>((ONE + TWO * h) + static_cast<double>(12.0f * h) * (eta * INV_PI) *
                                                  ((eta * INV_PI) - 1.0));
  }

  EB200_HD float qs_eta2theta(float h, float eta) {
    if (fabsf(h) <= EPS_F) {
      return eta;
    }
    const double PI = 3.14159265358979323846, INV_PI_SQR = 0.10132118364233777144;
    return static_cast<float>(eta + static_cast<double>(TWO * h * eta) * (PI - TWO * eta) *
                                      (PI - eta) * INV_PI_SQR);
  }

  // dx_dt of qkerr_schild.h:871-881 (d theta / d x2)
  EB200_HD float q_dx_dt(float h0, float deta, float eta) {
    if (fabsf(h0) <= EPS_F) {
      return deta;
    }
    return deta * (ONE + TWO * h0 * INV_PI_SQR_F *
                           (TWO * THREE * SQR(eta) - TWO * THREE * PI_F * eta + PI_SQR_F));
  }

  /* ------------------------------------------------------------- host-side construction */
  // mirrors the member initialisers of the reference constructors (fp32 arithmetic)
  inline MetricParams make_metric(int kind, int n1, int n2, float x1min, float x1max, float x2min,
                                  float x2max, float r0, float h, float a) {
    MetricParams m {};
    m.kind  = kind;
    m.nx1   = (float)n1;
    m.nx2   = (float)n2;
    m.x1min = x1min;
    m.x1max = x1max;
    m.x2min = x2min;
    m.x2max = x2max;
    m.r0    = r0;
    m.h     = h;
    m.a     = a;
    const bool quasi = (kind == EB200_METRIC_QSPHERICAL || kind == EB200_METRIC_QKERR_SCHILD);
    if (quasi) {
      m.chi_min = logf(x1min - r0);
      m.eta_min = q_theta2eta(h, x2min);
      m.d1      = (logf(x1max - r0) - m.chi_min) / m.nx1;
      m.d2      = (q_theta2eta(h, x2max) - m.eta_min) / m.nx2;
      m.small_angle = (kind == EB200_METRIC_QSPHERICAL ? qs_eta2theta(h, HALF * m.d2)
                                                       : q_eta2theta(h, HALF * m.d2)) < SMALL_ANGLE_F;
    } else {
      m.d1          = (x1max - x1min) / m.nx1;
      m.d2          = (x2max - x2min) / m.nx2;
      m.small_angle = HALF * m.d2 < SMALL_ANGLE_F;
    }
    m.d1_inv = ONE / m.d1;
    m.d2_inv = ONE / m.d2;
    return m;
  }

  /* ========================================================================== SR metrics */
  // Diagonal spatial metrics. Interface used by the SR kernels:
  //   r(x1), theta(x2), x1_of_r, x2_of_theta, h11/h22/h33, sqrt_h11/22/33, sqrt_det_h,
  //   polar_area
  struct Spherical {
    static constexpr int kind = EB200_METRIC_SPHERICAL;
    EB200_HD static float r(const MetricParams& m, float x1) { return x1 * m.d1 + m.x1min; }
    EB200_HD static float theta(const MetricParams& m, float x2) { return x2 * m.d2 + m.x2min; }
    EB200_HD static float x1_of_r(const MetricParams& m, float r) { return (r - m.x1min) * m.d1_inv; }
    EB200_HD static float x2_of_theta(const MetricParams& m, float t) {
      return (t - m.x2min) * m.d2_inv;
    }
    EB200_HD static float h11(const MetricParams& m, float, float) { return SQR(m.d1); }
    EB200_HD static float h22(const MetricParams& m, float x1, float) {
      return SQR(m.d2) * SQR(x1 * m.d1 + m.x1min);
    }
    EB200_HD static float h33(const MetricParams& m, float x1, float x2) {
      return SQR(x1 * m.d1 + m.x1min) * SQR(sinf(x2 * m.d2 + m.x2min));
    }
    EB200_HD static float sqrt_h11(const MetricParams& m, float, float) { return m.d1; }
    EB200_HD static float sqrt_h22(const MetricParams& m, float x1, float) {
      return m.d2 * (x1 * m.d1 + m.x1min);
    }
    EB200_HD static float sqrt_h33(const MetricParams& m, float x1, float x2) {
      return (x1 * m.d1 + m.x1min) * sinf(x2 * m.d2 + m.x2min);
    }
    EB200_HD static float sqrt_det_h(const MetricParams& m, float x1, float x2) {
      return m.d1 * m.d2 * SQR(x1 * m.d1 + m.x1min) * sinf(x2 * m.d2 + m.x2min);
    }
    EB200_HD static float polar_area(const MetricParams& m, float x1) {
      if (m.small_angle) {
        return m.d1 * SQR(x1 * m.d1 + m.x1min) * (48.0f - SQR(m.d2)) * SQR(m.d2) / 384.0f;
      }
      return m.d1 * SQR(x1 * m.d1 + m.x1min) * (ONE - cosf(HALF * m.d2));
    }
  };

  struct QSpherical {
    static constexpr int kind = EB200_METRIC_QSPHERICAL;
    EB200_HD static float r(const MetricParams& m, float x1) {
      return m.r0 + expf(x1 * m.d1 + m.chi_min);
    }
    EB200_HD static float theta(const MetricParams& m, float x2) {
      return qs_eta2theta(m.h, x2 * m.d2 + m.eta_min);
    }
    EB200_HD static float x1_of_r(const MetricParams& m, float r) {
      return (logf(r - m.r0) - m.chi_min) * m.d1_inv;
    }
    EB200_HD static float x2_of_theta(const MetricParams& m, float t) {
      return (q_theta2eta(m.h, t) - m.eta_min) * m.d2_inv;
    }
    EB200_HD static float h11(const MetricParams& m, float x1, float) {
      return SQR(m.d1) * expf(TWO * (x1 * m.d1 + m.chi_min));
    }
    EB200_HD static float h22(const MetricParams& m, float x1, float x2) {
      return SQR(m.d2) * SQR(qs_dtheta_deta(m.h, x2 * m.d2 + m.eta_min)) *
             SQR(m.r0 + expf(x1 * m.d1 + m.chi_min));
    }
    EB200_HD static float h33(const MetricParams& m, float x1, float x2) {
      return SQR((m.r0 + expf(x1 * m.d1 + m.chi_min)) *
                 sinf(qs_eta2theta(m.h, x2 * m.d2 + m.eta_min)));
    }
    EB200_HD static float sqrt_h11(const MetricParams& m, float x1, float) {
      return m.d1 * expf(x1 * m.d1 + m.chi_min);
    }
    EB200_HD static float sqrt_h22(const MetricParams& m, float x1, float x2) {
      return m.d2 * qs_dtheta_deta(m.h, x2 * m.d2 + m.eta_min) *
             (m.r0 + expf(x1 * m.d1 + m.chi_min));
    }
    EB200_HD static float sqrt_h33(const MetricParams& m, float x1, float x2) {
      return (m.r0 + expf(x1 * m.d1 + m.chi_min)) *
             sinf(qs_eta2theta(m.h, x2 * m.d2 + m.eta_min));
    }
    EB200_HD static float sqrt_det_h(const MetricParams& m, float x1, float x2) {
      const float exp_chi = expf(x1 * m.d1 + m.chi_min);
      return m.d1 * m.d2 * exp_chi * qs_dtheta_deta(m.h, x2 * m.d2 + m.eta_min) *
             SQR(m.r0 + exp_chi) * sinf(qs_eta2theta(m.h, x2 * m.d2 + m.eta_min));
    }
    EB200_HD static float polar_area(const MetricParams& m, float x1) {
      const float exp_chi = expf(x1 * m.d1 + m.chi_min);
      if (m.small_angle) {
        const float dtheta = qs_eta2theta(m.h, HALF * m.d2);
        return m.d1 * exp_chi * SQR(m.r0 + exp_chi) * (48.0f - SQR(dtheta)) * SQR(dtheta) /
               384.0f;
      }
      return m.d1 * exp_chi * SQR(m.r0 + exp_chi) * (ONE - cosf(qs_eta2theta(m.h, HALF * m.d2)));
    }
  };

  // code -> Cartesian (convert_xyz<Cd, XYZ>, e.g. qspherical.h:301-312); x[2] = phi
  template <class M>
  EB200_HD void cd_to_xyz(const MetricParams& m, const float* x, float* out) {
    const float r = M::r(m, x[0]), th = M::theta(m, x[1]), ph = x[2];
    out[0] = r * sinf(th) * cosf(ph);
    out[1] = r * sinf(th) * sinf(ph);
    out[2] = r * cosf(th);
  }

  // Cartesian -> code (convert_xyz<XYZ, Cd>, qspherical.h:313-324)
  template <class M>
  EB200_HD void xyz_to_cd(const MetricParams& m, const float* x, float* out) {
    const float r  = sqrtf(SQR(x[0]) + SQR(x[1]) + SQR(x[2]));
    const float th = HALF_PI_F - atan2f(x[2], sqrtf(SQR(x[0]) + SQR(x[1])));
    const float ph = PI_F - atan2f(x[1], -x[0]);
    out[0]         = M::x1_of_r(m, r);
    out[1]         = M::x2_of_theta(m, th);
    out[2]         = ph;
  }

  // sin/cos of a particle's (theta, phi), evaluated once and shared by every vector transform
  // at that position (the reference re-evaluates them inside each transform_xyz call)
  struct Trig {
    float st, ct, sp, cp;
  };

  template <class M>
  EB200_HD Trig trig_at(const MetricParams& m, const float* x) {
    Trig        t;
    const float th = M::theta(m, x[1]);
    t.st           = sinf(th);
    t.ct           = cosf(th);
    t.sp           = sinf(x[2]);
    t.cp           = cosf(x[2]);
    return t;
  }

  // tetrad -> Cartesian (transform_xyz<T, XYZ>, qspherical.h:399-411)
  EB200_HD void tetrad_to_xyz(const Trig& t, const float* v, float* out) {
    out[0] = v[0] * t.st * t.cp + v[1] * t.ct * t.cp - v[2] * t.sp;
    out[1] = v[0] * t.st * t.sp + v[1] * t.ct * t.sp + v[2] * t.cp;
    out[2] = v[0] * t.ct - v[1] * t.st;
  }

  // Cartesian -> tetrad (transform_xyz<XYZ, T>, qspherical.h:412-424)
  EB200_HD void xyz_to_tetrad(const Trig& t, const float* v, float* out) {
    out[0] = v[0] * t.st * t.cp + v[1] * t.st * t.sp + v[2] * t.ct;
    out[1] = v[0] * t.ct * t.cp + v[1] * t.ct * t.sp - v[2] * t.st;
    out[2] = -v[0] * t.sp + v[1] * t.cp;
  }

  // contravariant -> Cartesian (transform_xyz<U, XYZ>): U -> tetrad is * sqrt(h_ii)
  template <class M>
  EB200_HD void cntrv_to_xyz(const MetricParams& m, const float* x, const Trig& t, const float* v,
                             float* out) {
    const float vt[3] = { v[0] * M::sqrt_h11(m, x[0], x[1]), v[1] * M::sqrt_h22(m, x[0], x[1]),
                          v[2] * M::sqrt_h33(m, x[0], x[1]) };
    tetrad_to_xyz(t, vt, out);
  }

  // Cartesian -> contravariant (transform_xyz<XYZ, U>): tetrad -> U is / sqrt(h_ii)
  template <class M>
  EB200_HD void xyz_to_cntrv(const MetricParams& m, const float* x, const Trig& t, const float* v,
                             float* out) {
    float vt[3];
    xyz_to_tetrad(t, v, vt);
    out[0] = vt[0] / M::sqrt_h11(m, x[0], x[1]);
    out[1] = vt[1] / M::sqrt_h22(m, x[0], x[1]);
    out[2] = vt[2] / M::sqrt_h33(m, x[0], x[1]);
  }

  /* ========================================================================== GR metrics */
  // 3+1 split of the Kerr metric in (quasi-)spherical Kerr-Schild coordinates. Interface used
  // by the GR kernels (arguments are code coordinates x1, x2):
  //   theta, x2_of_theta, h_11/h_22/h_33/h_13 (covariant), h11/h22/h33/h13 (contravariant),
  //   alpha, beta1, sqrt_det_h, sqrt_det_h_tilde, polar_area,
  //   dr_* / dt_* derivatives of alpha, beta1, h11, h22, h33, h13 with respect to x1 / x2
  namespace ks {
    EB200_HD float Delta(float a, float r) { return SQR(r) - TWO * r + SQR(a); }
    EB200_HD float Sigma(float a, float r, float th) { return SQR(r) + SQR(a) * SQR(cosf(th)); }
    EB200_HD float A(float a, float r, float th) {
      return SQR(SQR(r) + SQR(a)) - SQR(a) * Delta(a, r) * SQR(sinf(th));
    }
    EB200_HD float z(float a, float r, float th) { return TWO * r / Sigma(a, r, th); }
  } // namespace ks

  struct KerrSchild { // src/metrics/kerr_schild.h
    static constexpr int kind = EB200_METRIC_KERR_SCHILD;
    EB200_HD static float r(const MetricParams& m, float x1) { return x1 * m.d1 + m.x1min; }
    EB200_HD static float theta(const MetricParams& m, float x2) { return x2 * m.d2 + m.x2min; }
    EB200_HD static float x2_of_theta(const MetricParams& m, float t) {
      return (t - m.x2min) * m.d2_inv;
    }
    EB200_HD static float h_11(const MetricParams& m, float x1, float x2) {
      return SQR(m.d1) * (ONE + ks::z(m.a, r(m, x1), theta(m, x2)));
    }
    EB200_HD static float h_22(const MetricParams& m, float x1, float x2) {
      return SQR(m.d2) * ks::Sigma(m.a, r(m, x1), theta(m, x2));
    }
    EB200_HD static float h_33(const MetricParams& m, float x1, float x2) {
      return ks::A(m.a, r(m, x1), theta(m, x2)) * SQR(sinf(theta(m, x2))) /
             ks::Sigma(m.a, r(m, x1), theta(m, x2));
    }
    EB200_HD static float h_13(const MetricParams& m, float x1, float x2) {
      return -m.d1 * m.a * (ONE + ks::z(m.a, r(m, x1), theta(m, x2))) * SQR(sinf(theta(m, x2)));
    }
    EB200_HD static float h11(const MetricParams& m, float x1, float x2) {
      const float Sigma_ = ks::Sigma(m.a, r(m, x1), theta(m, x2));
      return SQR(m.d1_inv) * ks::A(m.a, r(m, x1), theta(m, x2)) /
             (Sigma_ * (Sigma_ + TWO * r(m, x1)));
    }
    EB200_HD static float h22(const MetricParams& m, float x1, float x2) {
      return SQR(m.d2_inv) / ks::Sigma(m.a, r(m, x1), theta(m, x2));
    }
    EB200_HD static float h33(const MetricParams& m, float x1, float x2) {
      return ONE / (ks::Sigma(m.a, r(m, x1), theta(m, x2)) * SQR(sinf(theta(m, x2))));
    }
    EB200_HD static float h13(const MetricParams& m, float x1, float x2) {
      return m.d1_inv * m.a / ks::Sigma(m.a, r(m, x1), theta(m, x2));
    }
    EB200_HD static float alpha(const MetricParams& m, float x1, float x2) {
      return ONE / sqrtf(ONE + ks::z(m.a, r(m, x1), theta(m, x2)));
    }
    EB200_HD static float dt_Sigma(const MetricParams& m, float th) {
      const float v = -TWO * SQR(m.a) * sinf(th) * cosf(th) * m.d2;
      return (fabsf(v) <= EPS_F) ? ZERO : v;
    }
    EB200_HD static float dt_A(const MetricParams& m, float r_, float th) {
      const float v = -TWO * SQR(m.a) * sinf(th) * cosf(th) * ks::Delta(m.a, r_) * m.d2;
      return (fabsf(v) <= EPS_F) ? ZERO : v;
    }
    EB200_HD static float dr_alpha(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * m.d1;
      return -(m.d1 * ks::Sigma(m.a, r_, th) - r_ * dr_Sigma) * CUBE(alpha(m, x1, x2)) /
             SQR(ks::Sigma(m.a, r_, th));
    }
    EB200_HD static float dt_alpha(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      return CUBE(alpha(m, x1, x2)) * r_ * dt_Sigma(m, th) / SQR(ks::Sigma(m.a, r_, th));
    }
    EB200_HD static float beta1(const MetricParams& m, float x1, float x2) {
      const float z_ = ks::z(m.a, r(m, x1), theta(m, x2));
      return m.d1_inv * z_ / (ONE + z_);
    }
    EB200_HD static float dr_beta1(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * m.d1;
      return m.d1_inv * TWO * (m.d1 * ks::Sigma(m.a, r_, th) - r_ * dr_Sigma) /
             SQR(ks::Sigma(m.a, r_, th) + TWO * r_);
    }
    EB200_HD static float dt_beta1(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      return -m.d1_inv * TWO * r_ * dt_Sigma(m, th) /
             SQR(ks::Sigma(m.a, r_, th) * (ONE + ks::z(m.a, r_, th)));
    }
    EB200_HD static float dr_h11(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * m.d1;
      const float dr_Delta = TWO * m.d1 * (r_ - ONE);
      const float dr_A = FOUR * r_ * m.d1 * (SQR(r_) + SQR(m.a)) - SQR(m.a) * SQR(sinf(th)) * dr_Delta;
      const float S = ks::Sigma(m.a, r_, th);
      return (S * (S + TWO * r_) * dr_A -
              TWO * ks::A(m.a, r_, th) * (r_ * dr_Sigma + S * (dr_Sigma + m.d1))) /
             (SQR(S * (S + TWO * r_))) * SQR(m.d1_inv);
    }
    EB200_HD static float dr_h22(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * m.d1;
      return -dr_Sigma / SQR(ks::Sigma(m.a, r_, th)) * SQR(m.d2_inv);
    }
    EB200_HD static float dr_h33(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * m.d1;
      return -dr_Sigma / SQR(ks::Sigma(m.a, r_, th)) / SQR(sinf(th));
    }
    EB200_HD static float dr_h13(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * m.d1;
      return -m.a * dr_Sigma / SQR(ks::Sigma(m.a, r_, th)) * m.d1_inv;
    }
    EB200_HD static float dt_h11(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float S = ks::Sigma(m.a, r_, th);
      return (S * (S + TWO * r_) * dt_A(m, r_, th) -
              TWO * ks::A(m.a, r_, th) * dt_Sigma(m, th) * (r_ + S)) /
             (SQR(S * (S + TWO * r_))) * SQR(m.d1_inv);
    }
    EB200_HD static float dt_h22(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      return -dt_Sigma(m, th) / SQR(ks::Sigma(m.a, r_, th)) * SQR(m.d2_inv);
    }
    EB200_HD static float dt_h33(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      return -TWO * m.d2 * cosf(th) * (ks::Sigma(m.a, r_, th) - SQR(m.a) * SQR(sinf(th))) /
             CUBE(sinf(th)) / SQR(ks::Sigma(m.a, r_, th));
    }
    EB200_HD static float dt_h13(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      return -m.a * dt_Sigma(m, th) / SQR(ks::Sigma(m.a, r_, th)) * m.d1_inv;
    }
    EB200_HD static float sqrt_det_h(const MetricParams& m, float x1, float x2) {
      return m.d1 * m.d2 * ks::Sigma(m.a, r(m, x1), theta(m, x2)) * sinf(theta(m, x2)) *
             sqrtf(ONE + ks::z(m.a, r(m, x1), theta(m, x2)));
    }
    EB200_HD static float sqrt_det_h_tilde(const MetricParams& m, float x1, float x2) {
      return m.d1 * m.d2 * ks::Sigma(m.a, r(m, x1), theta(m, x2)) *
             sqrtf(ONE + ks::z(m.a, r(m, x1), theta(m, x2)));
    }
    EB200_HD static float polar_area(const MetricParams& m, float x1) {
      const float r_ = x1 * m.d1 + m.x1min;
      const float f  = m.d1 * (SQR(r_) + SQR(m.a)) * sqrtf(ONE + TWO * r_ / (SQR(r_) + SQR(m.a)));
      if (m.small_angle) {
        return f * (48.0f - SQR(m.d2)) * SQR(m.d2) / 384.0f;
      }
      return f * (ONE - cosf(HALF * m.d2));
    }
  };

  struct QKerrSchild { // src/metrics/qkerr_schild.h
    static constexpr int kind = EB200_METRIC_QKERR_SCHILD;
    EB200_HD static float chi(const MetricParams& m, float x1) { return x1 * m.d1 + m.chi_min; }
    EB200_HD static float eta(const MetricParams& m, float x2) { return x2 * m.d2 + m.eta_min; }
    EB200_HD static float r(const MetricParams& m, float x1) { return m.r0 + expf(chi(m, x1)); }
    EB200_HD static float theta(const MetricParams& m, float x2) {
      return q_eta2theta(m.h, eta(m, x2));
    }
    EB200_HD static float x2_of_theta(const MetricParams& m, float t) {
      return (q_theta2eta(m.h, t) - m.eta_min) * m.d2_inv;
    }
    EB200_HD static float dx_dt(const MetricParams& m, float eta_) {
      return q_dx_dt(m.h, m.d2, eta_);
    }
    EB200_HD static float h_11(const MetricParams& m, float x1, float x2) {
      return SQR(m.d1) * expf(TWO * chi(m, x1)) * (ONE + ks::z(m.a, r(m, x1), theta(m, x2)));
    }
    EB200_HD static float h_22(const MetricParams& m, float x1, float x2) {
      return SQR(m.d2) * SQR(q_dtheta_deta(m.h, eta(m, x2))) *
             ks::Sigma(m.a, r(m, x1), theta(m, x2));
    }
    EB200_HD static float h_33(const MetricParams& m, float x1, float x2) {
      const float th = theta(m, x2);
      return ks::A(m.a, r(m, x1), th) * SQR(sinf(th)) / ks::Sigma(m.a, r(m, x1), th);
    }
    EB200_HD static float h_13(const MetricParams& m, float x1, float x2) {
      const float th = theta(m, x2);
      return -m.d1 * expf(chi(m, x1)) * m.a * (ONE + ks::z(m.a, r(m, x1), th)) * SQR(sinf(th));
    }
    EB200_HD static float h11(const MetricParams& m, float x1, float x2) {
      const float th     = theta(m, x2);
      const float Sigma_ = ks::Sigma(m.a, r(m, x1), th);
      return (expf(-TWO * chi(m, x1)) / SQR(m.d1)) * ks::A(m.a, r(m, x1), th) /
             (Sigma_ * (Sigma_ + TWO * r(m, x1)));
    }
    EB200_HD static float h22(const MetricParams& m, float x1, float x2) {
      const float th = theta(m, x2);
      return ONE / (ks::Sigma(m.a, r(m, x1), th) * SQR(q_dtheta_deta(m.h, eta(m, x2))) * SQR(m.d2));
    }
    EB200_HD static float h33(const MetricParams& m, float x1, float x2) {
      const float th = theta(m, x2);
      return ONE / (ks::Sigma(m.a, r(m, x1), th) * SQR(sinf(th)));
    }
    EB200_HD static float h13(const MetricParams& m, float x1, float x2) {
      const float th = theta(m, x2);
      return (expf(-chi(m, x1)) * m.d1_inv) * m.a / ks::Sigma(m.a, r(m, x1), th);
    }
    EB200_HD static float alpha(const MetricParams& m, float x1, float x2) {
      return ONE / sqrtf(ONE + ks::z(m.a, r(m, x1), theta(m, x2)));
    }
    EB200_HD static float dt_Sigma(const MetricParams& m, float eta_) {
      const float th = q_eta2theta(m.h, eta_);
      const float v  = -TWO * SQR(m.a) * sinf(th) * cosf(th) * dx_dt(m, eta_);
      return (fabsf(v) <= EPS_F) ? ZERO : v;
    }
    EB200_HD static float dt_A(const MetricParams& m, float r_, float eta_) {
      const float th = q_eta2theta(m.h, eta_);
      const float v  = -TWO * SQR(m.a) * sinf(th) * cosf(th) * ks::Delta(m.a, r_) * dx_dt(m, eta_);
      return (fabsf(v) <= EPS_F) ? ZERO : v;
    }
    EB200_HD static float dr_alpha(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), th = theta(m, x2);
      const float dx_r = m.d1 * expf(chi(m, x1));
      const float dr_Sigma = TWO * r_ * dx_r;
      return -(dx_r * ks::Sigma(m.a, r_, th) - r_ * dr_Sigma) * CUBE(alpha(m, x1, x2)) /
             SQR(ks::Sigma(m.a, r_, th));
    }
    EB200_HD static float dt_alpha(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), eta_ = eta(m, x2), th = q_eta2theta(m.h, eta_);
      // the reference inlines dx_dt here without the h == 0 shortcut (qkerr_schild.h:353-365)
      const float dxdt = m.d2 * (ONE + TWO * m.h * INV_PI_SQR_F *
                                         (TWO * THREE * SQR(eta_) - TWO * THREE * PI_F * eta_ +
                                          PI_SQR_F));
      const float dtS = -TWO * SQR(m.a) * sinf(th) * cosf(th) * dxdt;
      return r_ * dtS * CUBE(alpha(m, x1, x2)) / SQR(ks::Sigma(m.a, r_, th));
    }
    EB200_HD static float beta1(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1);
      const float z_ = ks::z(m.a, m.r0 + expf(c), theta(m, x2));
      return expf(-c) * m.d1_inv * z_ / (ONE + z_);
    }
    EB200_HD static float dr_beta1(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), th = theta(m, x2);
      const float z_ = ks::z(m.a, r_, th);
      const float dx_r = m.d1 * expf(c);
      const float dr_Sigma = TWO * r_ * dx_r;
      return expf(-c) * m.d1_inv * TWO * (dx_r * ks::Sigma(m.a, r_, th) - r_ * dr_Sigma) /
               SQR(ks::Sigma(m.a, r_, th) + TWO * r_) -
             m.d1 * expf(-c) * m.d1_inv * z_ / (ONE + z_);
    }
    EB200_HD static float dt_beta1(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), eta_ = eta(m, x2);
      const float th = q_eta2theta(m.h, eta_);
      return -expf(-c) * m.d1_inv * TWO * r_ * dt_Sigma(m, eta_) /
             SQR(ks::Sigma(m.a, r_, th) * (ONE + ks::z(m.a, r_, th)));
    }
    EB200_HD static float dr_h11(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), th = theta(m, x2);
      const float dx_r = m.d1 * expf(c);
      const float dr_Sigma = TWO * r_ * dx_r;
      const float dr_Delta = TWO * dx_r * (r_ - ONE);
      const float dr_A = FOUR * r_ * dx_r * (SQR(r_) + SQR(m.a)) - SQR(m.a) * SQR(sinf(th)) * dr_Delta;
      const float S = ks::Sigma(m.a, r_, th);
      return (expf(-TWO * c) / SQR(m.d1) *
              (S * (S + TWO * r_) * dr_A -
               TWO * ks::A(m.a, r_, th) * (r_ * dr_Sigma + S * (dr_Sigma + dx_r))) /
              (SQR(S * (S + TWO * r_)))) -
             TWO * m.d1 * expf(-TWO * c) / SQR(m.d1) * ks::A(m.a, r_, th) / (S * (S + TWO * r_));
    }
    EB200_HD static float dr_h22(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * (m.d1 * expf(c));
      return -dr_Sigma / SQR(ks::Sigma(m.a, r_, th)) / SQR(m.d2);
    }
    EB200_HD static float dr_h33(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * (m.d1 * expf(c));
      return -dr_Sigma / SQR(ks::Sigma(m.a, r_, th)) / SQR(sinf(th));
    }
    EB200_HD static float dr_h13(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), th = theta(m, x2);
      const float dr_Sigma = TWO * r_ * (m.d1 * expf(c));
      return -m.a * dr_Sigma / SQR(ks::Sigma(m.a, r_, th)) * (expf(-c) * m.d1_inv) -
             m.d1 * (expf(-c) * m.d1_inv) * m.a / ks::Sigma(m.a, r_, th);
    }
    EB200_HD static float dt_h11(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), eta_ = eta(m, x2);
      const float th = q_eta2theta(m.h, eta_);
      const float S  = ks::Sigma(m.a, r_, th);
      return expf(-TWO * c) / SQR(m.d1) *
             (S * (S + TWO * r_) * dt_A(m, r_, eta_) -
              TWO * ks::A(m.a, r_, th) * dt_Sigma(m, eta_) * (r_ + S)) /
             (SQR(S * (S + TWO * r_)));
    }
    EB200_HD static float dt_h22(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), eta_ = eta(m, x2), th = q_eta2theta(m.h, eta_);
      return -dt_Sigma(m, eta_) / SQR(ks::Sigma(m.a, r_, th)) / SQR(m.d2);
    }
    EB200_HD static float dt_h33(const MetricParams& m, float x1, float x2) {
      const float r_ = r(m, x1), eta_ = eta(m, x2), th = q_eta2theta(m.h, eta_);
      return -(dt_Sigma(m, eta_) +
               TWO * cosf(th) / sinf(th) * ks::Sigma(m.a, r_, th) * dx_dt(m, eta_)) /
             SQR(ks::Sigma(m.a, r_, th) * sinf(th));
    }
    EB200_HD static float dt_h13(const MetricParams& m, float x1, float x2) {
      const float c = chi(m, x1), r_ = m.r0 + expf(c), eta_ = eta(m, x2);
      const float th = q_eta2theta(m.h, eta_);
      return -m.a * dt_Sigma(m, eta_) / SQR(ks::Sigma(m.a, r_, th)) * (expf(-c) * m.d1_inv);
    }
    EB200_HD static float sqrt_det_h(const MetricParams& m, float x1, float x2) {
      const float expchi = expf(chi(m, x1)), th = theta(m, x2);
      return m.d1 * expchi * q_dtheta_deta(m.h, eta(m, x2)) * m.d2 *
             ks::Sigma(m.a, m.r0 + expchi, th) * sinf(th) *
             sqrtf(ONE + ks::z(m.a, m.r0 + expchi, th));
    }
    EB200_HD static float sqrt_det_h_tilde(const MetricParams& m, float x1, float x2) {
      const float expchi = expf(chi(m, x1)), th = theta(m, x2);
      return m.d1 * expchi * q_dtheta_deta(m.h, eta(m, x2)) * m.d2 *
             ks::Sigma(m.a, m.r0 + expchi, th) * sqrtf(ONE + ks::z(m.a, m.r0 + expchi, th));
    }
    EB200_HD static float polar_area(const MetricParams& m, float x1) {
      const float e  = expf(x1 * m.d1 + m.chi_min);
      const float r_ = m.r0 + e;
      const float f  = m.d1 * e * (SQR(r_) + SQR(m.a)) * sqrtf(ONE + TWO * r_ / (SQR(r_) + SQR(m.a)));
      if (m.small_angle) {
        const float dtheta = q_eta2theta(m.h, HALF * m.d2);
        return f * (48.0f - SQR(dtheta)) * SQR(dtheta) / 384.0f;
      }
      return f * (ONE - cosf(q_eta2theta(m.h, HALF * m.d2)));
    }
  };

  struct KerrSchild0 { // src/metrics/kerr_schild_0.h: flat space in Kerr-Schild form
    static constexpr int kind = EB200_METRIC_KERR_SCHILD_0;
    EB200_HD static float r(const MetricParams& m, float x1) { return x1 * m.d1 + m.x1min; }
    EB200_HD static float theta(const MetricParams& m, float x2) { return x2 * m.d2 + m.x2min; }
    EB200_HD static float x2_of_theta(const MetricParams& m, float t) {
      return (t - m.x2min) * m.d2_inv;
    }
    EB200_HD static float h_11(const MetricParams& m, float, float) { return SQR(m.d1); }
    EB200_HD static float h_22(const MetricParams& m, float x1, float) {
      return SQR(m.d2) * SQR(r(m, x1));
    }
    EB200_HD static float h_33(const MetricParams& m, float x1, float x2) {
      return SQR(r(m, x1) * sinf(theta(m, x2)));
    }
    EB200_HD static float h_13(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float h11(const MetricParams& m, float, float) { return SQR(m.d1_inv); }
    EB200_HD static float h22(const MetricParams& m, float x1, float) {
      return SQR(m.d2_inv / r(m, x1));
    }
    EB200_HD static float h33(const MetricParams& m, float x1, float x2) {
      return ONE / SQR(r(m, x1) * sinf(theta(m, x2)));
    }
    EB200_HD static float h13(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float alpha(const MetricParams&, float, float) { return ONE; }
    EB200_HD static float dr_alpha(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dt_alpha(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float beta1(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dr_beta1(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dt_beta1(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dr_h11(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dr_h22(const MetricParams& m, float x1, float) {
      return -TWO / CUBE(r(m, x1)) * SQR(m.d2_inv) * m.d1;
    }
    EB200_HD static float dr_h33(const MetricParams& m, float x1, float x2) {
      return -TWO / CUBE(r(m, x1)) / SQR(sinf(theta(m, x2))) * m.d1;
    }
    EB200_HD static float dr_h13(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dt_h11(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dt_h22(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float dt_h33(const MetricParams& m, float x1, float x2) {
      const float th = theta(m, x2);
      return -TWO * cosf(th) / SQR(r(m, x1)) / CUBE(sinf(th)) * m.d2;
    }
    EB200_HD static float dt_h13(const MetricParams&, float, float) { return ZERO; }
    EB200_HD static float sqrt_det_h(const MetricParams& m, float x1, float x2) {
      return m.d1 * m.d2 * SQR(r(m, x1)) * sinf(theta(m, x2));
    }
    EB200_HD static float sqrt_det_h_tilde(const MetricParams& m, float x1, float) {
      return m.d1 * m.d2 * SQR(r(m, x1));
    }
    EB200_HD static float polar_area(const MetricParams& m, float x1) {
      return m.d1 * SQR(r(m, x1)) * (ONE - cosf(HALF * m.d2));
    }
  };

  /* ----------------------------------------------- vector transforms shared by the GR metrics */
  // kerr_schild.h:560-640 (identical in qkerr_schild.h / kerr_schild_0.h up to h_13 = 0)
  template <class M>
  EB200_HD void gr_cov_to_cntrv(const MetricParams& m, float x1, float x2, const float* v,
                                float* out) {
    const float H11 = M::h11(m, x1, x2), H13 = M::h13(m, x1, x2);
    out[0] = v[0] * H11 + v[2] * H13;
    out[1] = v[1] * M::h22(m, x1, x2);
    out[2] = v[0] * H13 + v[2] * M::h33(m, x1, x2);
  }

  template <class M>
  EB200_HD void gr_cov_to_tetrad(const MetricParams& m, float x1, float x2, const float* v,
                                 float* out) {
    const float A0 = sqrtf(M::h11(m, x1, x2));
    out[0] = v[0] * A0 - v[2] * A0 * M::h_13(m, x1, x2) / M::h_33(m, x1, x2);
    out[1] = v[1] / sqrtf(M::h_22(m, x1, x2));
    out[2] = v[2] / sqrtf(M::h_33(m, x1, x2));
  }

  template <class M>
  EB200_HD void gr_tetrad_to_cov(const MetricParams& m, float x1, float x2, const float* v,
                                 float* out) {
    out[0] = v[0] / sqrtf(M::h11(m, x1, x2)) +
             v[2] * M::h_13(m, x1, x2) / sqrtf(M::h_33(m, x1, x2));
    out[1] = v[1] * sqrtf(M::h_22(m, x1, x2));
    out[2] = v[2] * sqrtf(M::h_33(m, x1, x2));
  }

  template <class M>
  EB200_HD void gr_cntrv_to_tetrad(const MetricParams& m, float x1, float x2, const float* v,
                                   float* out) {
    out[0] = v[0] / sqrtf(M::h11(m, x1, x2));
    out[1] = v[1] * sqrtf(M::h_22(m, x1, x2));
    out[2] = v[2] * sqrtf(M::h_33(m, x1, x2)) +
             v[0] * M::h_13(m, x1, x2) / sqrtf(M::h_33(m, x1, x2));
  }

} // namespace eb200
