// entity_b200 -- kernels of the curvilinear SR path (2D axisymmetric spherical / qspherical).
//
//   push_sr_curv_kernel        sr::Pusher_kernel for M::CoordType != Cartesian
//                              (src/kernels/pushers/sr.hpp:117-332, 526-657, 659-814)
//   deposit_curv_kernel        DepositCurrents_kernel with the metric's velocity transform
//                              (src/kernels/currents_deposit.hpp:108-163 + the Cartesian body)
//   faraday_sr / ampere_sr / currents_ampere_sr
//                              src/kernels/faraday_sr.hpp:53-89, ampere_sr.hpp:58-109, 163-209
//   filter_sph_kernel          DigitalFilter_kernel<Dim::_2D, Coord::Spherical>
//                              (src/kernels/digital_filter.hpp:196-281)
// Metric components are evaluated in registers from a by-value parameter block (metrics.cuh).
// Compiled once, with IEEE division / square root and without FMA contraction (the reference's
// operation order); transcendental functions differ from glibc's in the last ulp, so parity
// with the reference is to fp32 rounding.
#define EB200_STRICT 1
#include "metrics.cuh"
#include "particle.cuh"
#include "launch.h"

namespace eb200 {
  namespace curv {

    /* ------------------------------------------------------------------ SR pusher */
    __device__ __forceinline__ void xi_to_i_di(float xi, int& i, float& di) {
      i  = static_cast<int>(xi + ONE) - 1; // from_Xi_to_i (sr.hpp:41-50)
      di = xi - static_cast<float>(i);
    }

    // flips the contravariant component `axis` of the Cartesian velocity at xp (sr.hpp:689-702)
    template <class M>
    __device__ __forceinline__ void reflect_velocity(const MetricParams& m, const float* xp,
                                                     int axis, float* u) {
      const Trig t = trig_at<M>(m, xp);
      float      v[3], w[3];
      xyz_to_cntrv<M>(m, xp, t, u, v);
      v[axis] = -v[axis];
      cntrv_to_xyz<M>(m, xp, t, v, w);
      u[0] = w[0];
      u[1] = w[1];
      u[2] = w[2];
    }

    template <class M, int O>
    __global__ void __launch_bounds__(128)
      push_sr_curv_kernel(PushArgs A, MetricParams m, eb200_prtls_t S, uint32_t npart,
                          FieldView<2> EB) {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= npart) {
        return;
      }
      const short tag = S.tag[p];
      if (tag != 1) {
        return;
      }
      const eb200_pusher_t& c = A.c;
      Prtl<2>               P;
      P.i[0] = S.i1[p];
      P.i[1] = S.i2[p];
      P.i[2] = 0;
      P.d[0] = S.dx1[p];
      P.d[1] = S.dx2[p];
      P.d[2] = ZERO;
      P.u[0] = S.ux1[p];
      P.u[1] = S.ux2[p];
      P.u[2] = S.ux3[p];
      P.tag  = tag;
      float xp[3] = { static_cast<float>(P.i[0]) + P.d[0], static_cast<float>(P.i[1]) + P.d[1],
                      S.phi[p] };
      const bool massive = (c.pusher_flags != EB200_PUSHER_PHOTON);
      if (massive) {
        float ei[3], bi[3], ec[3], bc[3];
        gather_fields<2, O>([&](int i, int j, int k, int comp) { return EB.ld(i, j, k, comp); },
                            A.ng, P, ei, bi);
        const Trig t = trig_at<M>(m, xp);
        cntrv_to_xyz<M>(m, xp, t, ei, ec);
        cntrv_to_xyz<M>(m, xp, t, bi, bc);
        float fext[3] = { ZERO, ZERO, ZERO };
        if (c.has_atmosphere) {
          // radial gravity ~ 1/r^2 (sr.hpp:1442-1455); gx2/gx3 are rejected on the host
          const float r  = M::r(m, xp[0]);
          float       f1 = ZERO;
          if (!(fabsf(c.atm_gx1) <= EPS_F) &&
              ((c.atm_ds < ZERO || r <= c.atm_x_surf + c.atm_ds) &&
               (c.atm_ds > ZERO || r >= c.atm_x_surf + c.atm_ds))) {
            f1 += c.atm_gx1 * SQR(c.atm_x_surf / r);
          }
          const float ft[3] = { f1, ZERO, ZERO };
          tetrad_to_xyz(t, ft, fext);
        }
        velocity_update(c, A.ndh, P.u, ec, bc, fext);
      }
      // full Cartesian coordinate push in the curvilinear basis (sr.hpp:573-627)
      const float inv_energy = massive
                                 ? ONE / sqrtf(ONE + SQR(P.u[0]) + SQR(P.u[1]) + SQR(P.u[2]))
                                 : ONE / sqrtf(SQR(P.u[0]) + SQR(P.u[1]) + SQR(P.u[2]));
      float       xc[3];
      cd_to_xyz<M>(m, xp, xc);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        xc[a] += (P.u[a] * inv_energy) * c.dt;
      }
      xyz_to_cd<M>(m, xc, xp);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        P.ip[a] = P.i[a];
        P.dp[a] = P.d[a];
        xi_to_i_di(xp[a], P.i[a], P.d[a]);
      }
      const float phi = xp[2];
      // position aligned with the velocity: the reference averages with the old position
      // twice (sr.hpp:615-626 and 635-649), phi is left as is
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          xp[a] = HALF * (xp[a] + (static_cast<float>(P.ip[a]) + P.dp[a]));
        }
      }
      // boundaries (sr.hpp:659-751)
      int lin = 0, centre = 0;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int ni  = A.ni[a];
        bool      inv = false;
        if (P.i[a] < 0) {
          const int b = c.pbc[2 * a];
          if (b == EB200_PBC_PERIODIC) {
            P.i[a]  += ni;
            P.ip[a] += ni;
          } else if (b == EB200_PBC_ABSORB) {
            P.tag = 0;
          } else if (b == EB200_PBC_REFLECT || (a == 1 && b == EB200_PBC_AXIS)) {
            P.i[a] = 0;
            P.d[a] = ONE - P.d[a];
            inv    = (b == EB200_PBC_REFLECT);
          }
        } else if (P.i[a] >= ni) {
          const int b = c.pbc[2 * a + 1];
          if (b == EB200_PBC_PERIODIC) {
            P.i[a]  -= ni;
            P.ip[a] -= ni;
          } else if (b == EB200_PBC_ABSORB) {
            P.tag = 0;
          } else if (b == EB200_PBC_REFLECT || (a == 1 && b == EB200_PBC_AXIS)) {
            P.i[a] = ni - 1;
            P.d[a] = ONE - P.d[a];
            inv    = (b == EB200_PBC_REFLECT);
          }
        }
        if (inv) {
          reflect_velocity<M>(m, xp, a, P.u);
        }
        const int dir = (P.i[a] < 0) ? 0 : ((P.i[a] >= ni) ? 2 : 1);
        lin           = lin * 3 + dir;
        centre        = centre * 3 + 1;
      }
      if (c.tag_outgoing && lin != centre) {
        P.tag = static_cast<short>((2 + lin - (lin > centre ? 1 : 0)) * P.tag);
      }
      S.i1[p]       = P.i[0];
      S.i2[p]       = P.i[1];
      S.dx1[p]      = P.d[0];
      S.dx2[p]      = P.d[1];
      S.i1_prev[p]  = P.ip[0];
      S.i2_prev[p]  = P.ip[1];
      S.dx1_prev[p] = P.dp[0];
      S.dx2_prev[p] = P.dp[1];
      S.ux1[p]      = P.u[0];
      S.ux2[p]      = P.u[1];
      S.ux3[p]      = P.u[2];
      S.phi[p]      = phi;
      if (P.tag != tag) {
        S.tag[p] = P.tag;
      }
    }

    /* -------------------------------------------------------------------- deposit */
    // coordinate velocity for the out-of-plane current (currents_deposit.hpp:133-139, 155-162)
    template <class M>
    struct VpSR {
      MetricParams m;
      const float* phi;

      __device__ __forceinline__ void operator()(const Prtl<2>& P, uint32_t p, float* vp) const {
        const float xp[3] = { static_cast<float>(P.i[0]) + P.d[0],
                              static_cast<float>(P.i[1]) + P.d[1], phi[p] };
        const Trig  t     = trig_at<M>(m, xp);
        xyz_to_cntrv<M>(m, xp, t, P.u, vp);
        const float inv_energy = ONE / sqrtf(ONE + (SQR(P.u[0]) + SQR(P.u[1]) + SQR(P.u[2])));
        if (isnan(vp[2]) || isinf(vp[2])) {
          vp[2] = ZERO;
        }
        vp[0] *= inv_energy;
        vp[1] *= inv_energy;
        vp[2] *= inv_energy;
      }
    };

    template <class VP, int O, bool AGG>
    __global__ void __launch_bounds__(128)
      deposit_curv_kernel(VP vpfn, eb200_prtls_t S, uint32_t npart, float charge, float inv_dt,
                          int G, FieldView<2> J) {
      const uint32_t p      = blockIdx.x * blockDim.x + threadIdx.x;
      const bool     active = (p < npart) && (S.tag[p] != 0);
      if constexpr (!AGG) {
        if (!active) {
          return;
        }
      }
      Prtl<2> P;
      float   vp[3] = { ZERO, ZERO, ZERO };
      if (active) {
        P.i[0]  = S.i1[p];
        P.i[1]  = S.i2[p];
        P.ip[0] = S.i1_prev[p];
        P.ip[1] = S.i2_prev[p];
        P.d[0]  = S.dx1[p];
        P.d[1]  = S.dx2[p];
        P.dp[0] = S.dx1_prev[p];
        P.dp[1] = S.dx2_prev[p];
        P.i[2] = P.ip[2] = 0;
        P.d[2] = P.dp[2] = ZERO;
        P.u[0] = S.ux1[p];
        P.u[1] = S.ux2[p];
        P.u[2] = S.ux3[p];
        P.w    = S.weight[p];
        vpfn(P, p, vp);
      }
      if constexpr (AGG) {
        deposit_particle_aggregated<2, O>(P, active, charge, inv_dt, ONE, G, J, vp);
      } else {
        deposit_particle<2, O>(
          P, charge, inv_dt, ONE, G,
          [&](int i, int j, int k, int c, float v, bool guard = true) {
            if (guard) atomicAdd(&J.at(i, j, k, c), v);
          },
          vp);
      }
    }

    /* -------------------------------------------------------------- field solvers */
    template <class M>
    __global__ void __launch_bounds__(128)
      faraday_sr_kernel(MetricParams m, FieldView<2> EB, int n1, int n2, int G, float coeff,
                        bool axis_min) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 >= G + n2) {
        return;
      }
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      const float inv_sqrt_detH_0pH  = ONE / M::sqrt_det_h(m, x1, x2 + HALF);
      const float inv_sqrt_detH_pHpH = ONE / M::sqrt_det_h(m, x1 + HALF, x2 + HALF);
      const float h1_pHp1            = M::h11(m, x1 + HALF, x2 + ONE);
      const float h1_pH0             = M::h11(m, x1 + HALF, x2);
      const float h2_p1pH            = M::h22(m, x1 + ONE, x2 + HALF);
      const float h2_0pH             = M::h22(m, x1, x2 + HALF);
      const float h3_00              = M::h33(m, x1, x2);
      const float h3_0p1             = M::h33(m, x1, x2 + ONE);
      EB.at(i1, i2, 0, bx1) += coeff * inv_sqrt_detH_0pH *
                               (h3_00 * EB.at(i1, i2, 0, ex3) - h3_0p1 * EB.at(i1, i2 + 1, 0, ex3));
      if ((i2 != G) || !axis_min) {
        const float inv_sqrt_detH_pH0 = ONE / M::sqrt_det_h(m, x1 + HALF, x2);
        const float h3_p10            = M::h33(m, x1 + ONE, x2);
        EB.at(i1, i2, 0, bx2) += coeff * inv_sqrt_detH_pH0 *
                                 (h3_p10 * EB.at(i1 + 1, i2, 0, ex3) -
                                  h3_00 * EB.at(i1, i2, 0, ex3));
      }
      EB.at(i1, i2, 0, bx3) += coeff * inv_sqrt_detH_pHpH *
                               (h1_pHp1 * EB.at(i1, i2 + 1, 0, ex1) - h1_pH0 * EB.at(i1, i2, 0, ex1) +
                                h2_0pH * EB.at(i1, i2, 0, ex2) - h2_p1pH * EB.at(i1 + 1, i2, 0, ex2));
    }

    // rows: active + one more when the upper x2 face is an axis (srpic::RangeWithAxisBCs,
    // src/engines/srpic/utils.h:104-129)
    template <class M>
    __global__ void __launch_bounds__(128)
      ampere_sr_kernel(MetricParams m, FieldView<2> EB, int n1, int nrows, int n2, int G,
                       float coeff, bool axis_min, bool axis_max) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 >= G + nrows) {
        return;
      }
      const int   i2min = G, i2max = n2 + G;
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      const float inv_sqrt_detH_0pH = ONE / M::sqrt_det_h(m, x1, x2 + HALF);
      const float h3_mHpH           = M::h33(m, x1 - HALF, x2 + HALF);
      const float h3_pHpH           = M::h33(m, x1 + HALF, x2 + HALF);
      if ((i2 == i2min) && axis_min) {
        const float inv_polar_area_pH0 = ONE / M::polar_area(m, x1 + HALF);
        EB.at(i1, i2, 0, ex1) += inv_polar_area_pH0 * coeff * (h3_pHpH * EB.at(i1, i2, 0, bx3));
        EB.at(i1, i2, 0, ex2) += coeff * inv_sqrt_detH_0pH *
                                 (h3_mHpH * EB.at(i1 - 1, i2, 0, bx3) -
                                  h3_pHpH * EB.at(i1, i2, 0, bx3));
      } else if ((i2 == i2max) && axis_max) {
        const float inv_polar_area_pH0 = ONE / M::polar_area(m, x1 + HALF);
        const float h3_pHmH            = M::h33(m, x1 + HALF, x2 - HALF);
        // reads bx3 of the row AT the axis, as the reference does (ampere_sr.hpp:82-83)
        EB.at(i1, i2, 0, ex1) -= inv_polar_area_pH0 * coeff * (h3_pHmH * EB.at(i1, i2, 0, bx3));
      } else {
        const float inv_sqrt_detH_00  = ONE / M::sqrt_det_h(m, x1, x2);
        const float inv_sqrt_detH_pH0 = ONE / M::sqrt_det_h(m, x1 + HALF, x2);
        const float h1_0mH            = M::h11(m, x1, x2 - HALF);
        const float h1_0pH            = M::h11(m, x1, x2 + HALF);
        const float h2_pH0            = M::h22(m, x1 + HALF, x2);
        const float h2_mH0            = M::h22(m, x1 - HALF, x2);
        const float h3_pHmH           = M::h33(m, x1 + HALF, x2 - HALF);
        EB.at(i1, i2, 0, ex1) += coeff * inv_sqrt_detH_pH0 *
                                 (h3_pHpH * EB.at(i1, i2, 0, bx3) -
                                  h3_pHmH * EB.at(i1, i2 - 1, 0, bx3));
        EB.at(i1, i2, 0, ex2) += coeff * inv_sqrt_detH_0pH *
                                 (h3_mHpH * EB.at(i1 - 1, i2, 0, bx3) -
                                  h3_pHpH * EB.at(i1, i2, 0, bx3));
        EB.at(i1, i2, 0, ex3) += coeff * inv_sqrt_detH_00 *
                                 (h1_0mH * EB.at(i1, i2 - 1, 0, bx1) - h1_0pH * EB.at(i1, i2, 0, bx1) +
                                  h2_pH0 * EB.at(i1, i2, 0, bx2) - h2_mH0 * EB.at(i1 - 1, i2, 0, bx2));
      }
    }

    template <class M>
    __global__ void __launch_bounds__(128)
      currents_ampere_sr_kernel(MetricParams m, FieldView<2> E, FieldView<2> J, int n1, int nrows,
                                int n2, int G, float coeff, float inv_n0, bool axis_min,
                                bool axis_max) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 >= G + nrows) {
        return;
      }
      const int   i2min = G, i2max = n2 + G;
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      if ((i2 == i2min) && axis_min) {
        J.at(i1, i2, 0, jx1) *= inv_n0 * HALF / M::polar_area(m, x1 + HALF);
        J.at(i1, i2, 0, jx2) *= inv_n0 / M::sqrt_det_h(m, x1, x2 + HALF);
        E.at(i1, i2, 0, ex2) += J.at(i1, i2, 0, jx2) * coeff;
        J.at(i1, i2, 0, jx3) = ZERO;
      } else if ((i2 == i2max) && axis_max) {
        J.at(i1, i2, 0, jx1) *= inv_n0 * HALF / M::polar_area(m, x1 + HALF);
        J.at(i1, i2, 0, jx3) = ZERO;
      } else {
        J.at(i1, i2, 0, jx1) *= inv_n0 / M::sqrt_det_h(m, x1 + HALF, x2);
        J.at(i1, i2, 0, jx2) *= inv_n0 / M::sqrt_det_h(m, x1, x2 + HALF);
        E.at(i1, i2, 0, ex2) += J.at(i1, i2, 0, jx2) * coeff;
        J.at(i1, i2, 0, jx3) *= inv_n0 / M::sqrt_det_h(m, x1, x2);
        E.at(i1, i2, 0, ex3) += J.at(i1, i2, 0, jx3) * coeff;
      }
      E.at(i1, i2, 0, ex1) += J.at(i1, i2, 0, jx1) * coeff;
    }

    /* ------------------------------------------------------ binomial filter, spherical */
    __device__ __forceinline__ float filt_i1(const FieldView<2>& B, int c, int i1, int i2) {
      // FILTER2D_IN_I1 (digital_filter.hpp:27-30)
      return INV_2 * B.at(i1, i2, 0, c) + INV_4 * (B.at(i1 - 1, i2, 0, c) + B.at(i1 + 1, i2, 0, c));
    }

    __global__ void __launch_bounds__(128)
      filter_sph_kernel(FieldView<2> A, FieldView<2> B, int n1, int nrows, int n2, int G,
                        bool axis_min, bool axis_max) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 >= G + nrows) {
        return;
      }
      const int i2_min = G, i2_max = n2 + G;
      float     cur_00, cur_0p1, cur_0m1;
      if (axis_min && (i2 == i2_min)) {
        cur_00                = filt_i1(B, jx1, i1, i2);
        cur_0p1               = filt_i1(B, jx1, i1, i2 + 1);
        A.at(i1, i2, 0, jx1)  = INV_2 * cur_00 + INV_2 * cur_0p1;
        A.at(i1, i2, 0, jx3)  = ZERO;
        cur_00                = filt_i1(B, jx2, i1, i2);
        cur_0p1               = filt_i1(B, jx2, i1, i2 + 1);
        A.at(i1, i2, 0, jx2)  = INV_4 * (cur_00 + cur_0p1);
      } else if (axis_min && (i2 == i2_min + 1)) {
        cur_00                = filt_i1(B, jx1, i1, i2);
        cur_0p1               = filt_i1(B, jx1, i1, i2 + 1);
        cur_0m1               = filt_i1(B, jx1, i1, i2 - 1);
        A.at(i1, i2, 0, jx1)  = INV_2 * cur_00 + INV_4 * (cur_0p1 + cur_0m1);
        cur_00                = filt_i1(B, jx3, i1, i2);
        cur_0p1               = filt_i1(B, jx3, i1, i2 + 1);
        A.at(i1, i2, 0, jx3)  = INV_2 * cur_00 + INV_4 * cur_0p1;
        cur_00                = filt_i1(B, jx2, i1, i2);
        cur_0p1               = filt_i1(B, jx2, i1, i2 + 1);
        cur_0m1               = filt_i1(B, jx2, i1, i2 - 1);
        A.at(i1, i2, 0, jx2)  = INV_2 * cur_00 + INV_4 * (cur_0m1 + cur_0p1);
      } else if (axis_max && (i2 == i2_max - 1)) {
        cur_00                = filt_i1(B, jx1, i1, i2);
        cur_0p1               = filt_i1(B, jx1, i1, i2 + 1);
        cur_0m1               = filt_i1(B, jx1, i1, i2 - 1);
        A.at(i1, i2, 0, jx1)  = INV_2 * cur_00 + INV_4 * (cur_0m1 + cur_0p1);
        cur_00                = filt_i1(B, jx3, i1, i2);
        cur_0m1               = filt_i1(B, jx3, i1, i2 - 1);
        A.at(i1, i2, 0, jx3)  = INV_2 * cur_00 + INV_4 * cur_0m1;
        cur_00                = filt_i1(B, jx2, i1, i2);
        cur_0m1               = filt_i1(B, jx2, i1, i2 - 1);
        A.at(i1, i2, 0, jx2)  = INV_4 * (cur_00 + cur_0m1);
      } else if (axis_max && (i2 == i2_max)) {
        cur_00                = filt_i1(B, jx1, i1, i2);
        cur_0m1               = filt_i1(B, jx1, i1, i2 - 1);
        A.at(i1, i2, 0, jx1)  = INV_2 * cur_00 + INV_2 * cur_0m1;
        A.at(i1, i2, 0, jx3)  = ZERO;
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          A.at(i1, i2, 0, c) =
            INV_4 * B.at(i1, i2, 0, c) +
            INV_8 * (B.at(i1 - 1, i2, 0, c) + B.at(i1 + 1, i2, 0, c) + B.at(i1, i2 - 1, 0, c) +
                     B.at(i1, i2 + 1, 0, c)) +
            INV_16 * (B.at(i1 - 1, i2 - 1, 0, c) + B.at(i1 + 1, i2 + 1, 0, c) +
                      B.at(i1 - 1, i2 + 1, 0, c) + B.at(i1 + 1, i2 - 1, 0, c));
        }
      }
    }

    /* ------------------------------------------------------------------ launchers */
    static dim3 cell_grid(int n1, int nrows) { return dim3((n1 + 127) / 128, nrows, 1); }

#define EB200_SR_METRIC_DO(kind, CALL)                                                         \
  switch (kind) {                                                                              \
    case EB200_METRIC_SPHERICAL: CALL(Spherical); break;                                       \
    case EB200_METRIC_QSPHERICAL: CALL(QSpherical); break;                                     \
    default: return cudaErrorInvalidValue;                                                     \
  }

#define EB200_ORDER_DO(order, CALL)                                                            \
  switch (order) {                                                                             \
    case 0: CALL(0); break;                                                                    \
    case 1: CALL(1); break;                                                                    \
    case 2: CALL(2); break;                                                                    \
    case 3: CALL(3); break;                                                                    \
    default: return cudaErrorInvalidValue;                                                     \
  }

    template <class M>
    static cudaError_t push_sr_m(const MetricParams& m, const eb200_grid_t& g, int order,
                                 const PushArgs& A, const eb200_prtls_t& S, uint32_t npart,
                                 const float* em, cudaStream_t st) {
      FieldView<2> EB(g, const_cast<float*>(em));
      const unsigned nb = (npart + 127) / 128;
#define CALL(O) push_sr_curv_kernel<M, O><<<nb, 128, 0, st>>>(A, m, S, npart, EB)
      EB200_ORDER_DO(order, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t push_sr(const MetricParams& m, const eb200_grid_t& g, int order,
                        const eb200_pusher_t& c, const eb200_prtls_t& S, uint32_t npart,
                        const float* em, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      PushArgs A;
      A.c      = c;
      A.ndh    = HALF * (c.charge / c.mass) * c.omegaB0 * c.dt;
      A.ng     = g.ng;
      A.inv_dx = ONE;
      for (int a = 0; a < 3; ++a) A.ni[a] = g.n[a];
#define CALL(M) return push_sr_m<M>(m, g, order, A, S, npart, em, st)
      EB200_SR_METRIC_DO(m.kind, CALL)
#undef CALL
      return cudaErrorInvalidValue;
    }

    template <class VP>
    static cudaError_t deposit_vp(const VP& vp, const eb200_grid_t& g, int order,
                                  const eb200_prtls_t& S, uint32_t npart, float charge, float dt,
                                  float* cur, int mode, cudaStream_t st) {
      FieldView<2>   J(g, cur);
      const float    inv_dt = ONE / dt;
      const unsigned nb     = (npart + 127) / 128;
      if (mode == EB200_DEPOSIT_AGGREGATED) {
#define CALL(O) deposit_curv_kernel<VP, O, true><<<nb, 128, 0, st>>>(vp, S, npart, charge, inv_dt, g.ng, J)
        EB200_ORDER_DO(order, CALL)
#undef CALL
      } else {
#define CALL(O) deposit_curv_kernel<VP, O, false><<<nb, 128, 0, st>>>(vp, S, npart, charge, inv_dt, g.ng, J)
        EB200_ORDER_DO(order, CALL)
#undef CALL
      }
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t deposit_sr(const MetricParams& m, const eb200_grid_t& g, int order,
                           const eb200_prtls_t& S, uint32_t npart, float charge, float dt,
                           float* cur, int mode, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
#define CALL(M) return deposit_vp(VpSR<M> { m, S.phi }, g, order, S, npart, charge, dt, cur, mode, st)
      EB200_SR_METRIC_DO(m.kind, CALL)
#undef CALL
      return cudaErrorInvalidValue;
    }

    cudaError_t faraday_sr(const MetricParams& m, const eb200_grid_t& g, float* em, float coeff,
                           const int* fbc, cudaStream_t st) {
      FieldView<2> EB(g, em);
      const bool   axis_min = fbc[2] == EB200_FBC_AXIS;
#define CALL(M) faraday_sr_kernel<M><<<cell_grid(g.n[0], g.n[1]), 128, 0, st>>>(m, EB, g.n[0], g.n[1], g.ng, coeff, axis_min)
      EB200_SR_METRIC_DO(m.kind, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t ampere_sr(const MetricParams& m, const eb200_grid_t& g, float* em, float coeff,
                          const int* fbc, cudaStream_t st) {
      FieldView<2> EB(g, em);
      const bool   axis_min = fbc[2] == EB200_FBC_AXIS, axis_max = fbc[3] == EB200_FBC_AXIS;
      const int    nrows = g.n[1] + (axis_max ? 1 : 0);
#define CALL(M) ampere_sr_kernel<M><<<cell_grid(g.n[0], nrows), 128, 0, st>>>(m, EB, g.n[0], nrows, g.n[1], g.ng, coeff, axis_min, axis_max)
      EB200_SR_METRIC_DO(m.kind, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t currents_ampere_sr(const MetricParams& m, const eb200_grid_t& g, float* em,
                                   float* cur, float coeff, float inv_n0, const int* fbc,
                                   cudaStream_t st) {
      FieldView<2> E(g, em), J(g, cur);
      const bool   axis_min = fbc[2] == EB200_FBC_AXIS, axis_max = fbc[3] == EB200_FBC_AXIS;
      const int    nrows = g.n[1] + (axis_max ? 1 : 0);
#define CALL(M) currents_ampere_sr_kernel<M><<<cell_grid(g.n[0], nrows), 128, 0, st>>>(m, E, J, g.n[0], nrows, g.n[1], g.ng, coeff, inv_n0, axis_min, axis_max)
      EB200_SR_METRIC_DO(m.kind, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t filter_sph_pass(const eb200_grid_t& g, float* cur, const float* buff,
                                const int* fbc, cudaStream_t st) {
      FieldView<2> A(g, cur), B(g, const_cast<float*>(buff));
      const bool   axis_min = fbc[2] == EB200_FBC_AXIS, axis_max = fbc[3] == EB200_FBC_AXIS;
      const int    nrows = g.n[1] + (axis_max ? 1 : 0);
      filter_sph_kernel<<<cell_grid(g.n[0], nrows), 128, 0, st>>>(A, B, g.n[0], nrows, g.n[1], g.ng,
                                                                  axis_min, axis_max);
      count_launch();
      return cudaGetLastError();
    }

    /* ======================================================================== GRPIC */
    // gr::Pusher_kernel (src/kernels/pushers/gr.hpp:728-814 massive, :668-723 massless)
    struct GRPushArgs {
      eb200_pusher_gr_t c;
      float             ndh; // 1/2 (q/m) omegaB0 dt
      int               ni[2];
      int               ng;
    };

    // theta clamped away from the axis (gr.hpp:749-761)
    template <class M>
    __device__ __forceinline__ float clamp_x2(const MetricParams& m, float x2) {
      const float theta_Ph = M::theta(m, x2);
      const float small_angle = SMALL_ANGLE_GR_F, large_angle = PI_F - SMALL_ANGLE_GR_F;
      if (theta_Ph < small_angle) {
        return M::x2_of_theta(m, small_angle);
      } else if (theta_Ph >= large_angle) {
        return M::x2_of_theta(m, large_angle);
      }
      return x2;
    }

    // gr.hpp:177-217
    template <class M>
    __device__ __forceinline__ void em_half_push(const MetricParams& m, float ndh, const float* xp,
                                                 const float* vp, const float* Dp_hat,
                                                 const float* Bp_hat, float* vp_upd) {
      float D0[3] = { Dp_hat[0], Dp_hat[1], Dp_hat[2] };
      float B0[3] = { Bp_hat[0], Bp_hat[1], Bp_hat[2] };
      float vh[3], vu[3];
      gr_cov_to_tetrad<M>(m, xp[0], xp[1], vp, vu);
      float COEFF = ndh * HALF * M::alpha(m, xp[0], xp[1]);
      D0[0] *= COEFF;
      D0[1] *= COEFF;
      D0[2] *= COEFF;
      vu[0] += D0[0];
      vu[1] += D0[1];
      vu[2] += D0[2];
      COEFF *= ONE / sqrtf(ONE + SQR(vu[0]) + SQR(vu[1]) + SQR(vu[2]));
      B0[0] *= COEFF;
      B0[1] *= COEFF;
      B0[2] *= COEFF;
      COEFF = TWO / (ONE + SQR(B0[0]) + SQR(B0[1]) + SQR(B0[2]));
      vh[0] = (vu[0] + vu[1] * B0[2] - vu[2] * B0[1]) * COEFF;
      vh[1] = (vu[1] + vu[2] * B0[0] - vu[0] * B0[2]) * COEFF;
      vh[2] = (vu[2] + vu[0] * B0[1] - vu[1] * B0[0]) * COEFF;
      vu[0] += vh[1] * B0[2] - vh[2] * B0[1] + D0[0];
      vu[1] += vh[2] * B0[0] - vh[0] * B0[2] + D0[1];
      vu[2] += vh[0] * B0[1] - vh[1] * B0[0] + D0[2];
      gr_tetrad_to_cov<M>(m, xp[0], xp[1], vu, vp_upd);
    }

    template <bool MASSIVE>
    __device__ __forceinline__ float gr_gamma(const float* u_cov, const float* u_cntrv) {
      if constexpr (MASSIVE) {
        return sqrtf(ONE + u_cov[0] * u_cntrv[0] + u_cov[1] * u_cntrv[1] + u_cov[2] * u_cntrv[2]);
      } else {
        return sqrtf(u_cov[0] * u_cntrv[0] + u_cov[1] * u_cntrv[1] + u_cov[2] * u_cntrv[2]);
      }
    }

    // gr.hpp:262-331: every metric quantity sits at the fixed point xp -> evaluated once,
    // the niter fixed-point iterations run on registers
    template <class M, bool MASSIVE>
    __device__ __forceinline__ void geodesic_momentum_push(const MetricParams& m, float dt,
                                                           int niter, const float* xp,
                                                           const float* vp, float* vp_upd) {
      const float x1 = xp[0], x2 = xp[1];
      const float al = M::alpha(m, x1, x2), dr_al = M::dr_alpha(m, x1, x2),
                  dt_al = M::dt_alpha(m, x1, x2);
      const float dr_b = M::dr_beta1(m, x1, x2), dt_b = M::dt_beta1(m, x1, x2);
      const float dr11 = M::dr_h11(m, x1, x2), dr22 = M::dr_h22(m, x1, x2),
                  dr33 = M::dr_h33(m, x1, x2), dr13 = M::dr_h13(m, x1, x2);
      const float dt11 = M::dt_h11(m, x1, x2), dt22 = M::dt_h22(m, x1, x2),
                  dt33 = M::dt_h33(m, x1, x2), dt13 = M::dt_h13(m, x1, x2);
      const float H11 = M::h11(m, x1, x2), H22 = M::h22(m, x1, x2), H33 = M::h33(m, x1, x2),
                  H13 = M::h13(m, x1, x2);
      float vm[3], vc[3];
      vp_upd[0] = vp[0];
      vp_upd[1] = vp[1];
      vp_upd[2] = vp[2];
      for (int it = 0; it < niter; ++it) {
        vm[0] = HALF * (vp[0] + vp_upd[0]);
        vm[1] = HALF * (vp[1] + vp_upd[1]);
        vm[2] = vp[2];
        vc[0] = vm[0] * H11 + vm[2] * H13;
        vc[1] = vm[1] * H22;
        vc[2] = vm[0] * H13 + vm[2] * H33;
        const float u0 = gr_gamma<MASSIVE>(vm, vc) / al;
        vp_upd[0] = vp[0] + dt * (-al * u0 * dr_al + vm[0] * dr_b -
                                  (HALF / u0) * (dr11 * SQR(vm[0]) + dr22 * SQR(vm[1]) +
                                                 dr33 * SQR(vm[2]) + TWO * dr13 * vm[0] * vm[2]));
        vp_upd[1] = vp[1] + dt * (-al * u0 * dt_al + vm[0] * dt_b -
                                  (HALF / u0) * (dt11 * SQR(vm[0]) + dt22 * SQR(vm[1]) +
                                                 dt33 * SQR(vm[2]) + TWO * dt13 * vm[0] * vm[2]));
      }
    }

    // gr.hpp:335-383
    template <class M, bool MASSIVE>
    __device__ __forceinline__ void geodesic_coordinate_push(const MetricParams& m, float dt,
                                                             int niter, const float* xp,
                                                             const float* vp, float* xp_upd) {
      float vc[3];
      xp_upd[0] = xp[0];
      xp_upd[1] = xp[1];
      for (int it = 0; it < niter; ++it) {
        const float xm0 = HALF * (xp[0] + xp_upd[0]);
        const float xm1 = clamp_x2<M>(m, HALF * (xp[1] + xp_upd[1]));
        gr_cov_to_cntrv<M>(m, xm0, xm1, vp, vc);
        const float u0 = gr_gamma<MASSIVE>(vp, vc) / M::alpha(m, xm0, xm1);
        xp_upd[0]      = xp[0] + dt * (vc[0] / u0 - M::beta1(m, xm0, xm1));
        xp_upd[1]      = xp[1] + dt * (vc[1] / u0);
      }
    }

    template <class M, int O, bool MASSIVE>
    __global__ void __launch_bounds__(128)
      push_gr_kernel(GRPushArgs A, MetricParams m, eb200_prtls_t S, uint32_t npart,
                     FieldView<2> DB, FieldView<2> DB0) {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= npart) {
        return;
      }
      const short tag = S.tag[p];
      if (tag != 1) {
        return;
      }
      const eb200_pusher_gr_t& c = A.c;
      Prtl<2>                  P;
      P.i[0] = P.ip[0] = S.i1[p];
      P.i[1] = P.ip[1] = S.i2[p];
      P.i[2] = P.ip[2] = 0;
      P.d[0] = P.dp[0] = S.dx1[p];
      P.d[1] = P.dp[1] = S.dx2[p];
      P.d[2] = P.dp[2] = ZERO;
      P.tag            = tag;
      const float xp[2] = { static_cast<float>(P.i[0]) + P.d[0],
                            static_cast<float>(P.i[1]) + P.d[1] };
      float       vp[3] = { S.ux1[p], S.ux2[p], S.ux3[p] };
      float       vp_upd[3], xp_upd[2];
      if constexpr (MASSIVE) {
        const float xp_[2] = { xp[0], clamp_x2<M>(m, xp[1]) };
        float       Dc[3], Bc[3], Dh[3], Bh[3];
        // D from em (components 0..2), B from em0 (components 3..5): gr.hpp:484-663
        gather_fields<2, O>(
          [&](int i, int j, int k, int comp) {
            return (comp < 3) ? DB.ld(i, j, k, comp) : DB0.ld(i, j, k, comp);
          },
          A.ng, P, Dc, Bc);
        gr_cntrv_to_tetrad<M>(m, xp[0], xp[1], Dc, Dh);
        gr_cntrv_to_tetrad<M>(m, xp[0], xp[1], Bc, Bh);
        em_half_push<M>(m, A.ndh, xp, vp, Dh, Bh, vp_upd);
        vp[0] = vp_upd[0];
        vp[1] = vp_upd[1];
        vp[2] = vp_upd[2];
        geodesic_momentum_push<M, true>(m, c.dt, c.niter, xp_, vp, vp_upd);
        vp[0] = vp_upd[0];
        vp[1] = vp_upd[1];
        vp[2] = vp_upd[2];
        em_half_push<M>(m, A.ndh, xp, vp, Dh, Bh, vp_upd);
        geodesic_coordinate_push<M, true>(m, c.dt, c.niter, xp, vp_upd, xp_upd);
      } else {
        geodesic_momentum_push<M, false>(m, c.dt, c.niter, xp, vp, vp_upd);
        geodesic_coordinate_push<M, false>(m, c.dt, c.niter, xp, vp_upd, xp_upd);
      }
      // UpdatePhi takes phi by value in the reference (gr.hpp:465-480): phi stays as it is
      xi_to_i_di(xp_upd[0], P.i[0], P.d[0]);
      xi_to_i_di(xp_upd[1], P.i[1], P.d[1]);
      // boundaries (gr.hpp:819-865)
      if ((P.i[0] < 0 && c.pbc[0] == EB200_PBC_ABSORB) ||
          (P.i[0] >= A.ni[0] && c.pbc[1] == EB200_PBC_ABSORB)) {
        P.tag = 0;
      }
      if (P.i[1] < 0) {
        if (c.pbc[2] == EB200_PBC_AXIS) {
          P.i[1]    = 0;
          P.d[1]    = ONE - P.d[1];
          vp_upd[1] = -vp_upd[1];
        }
      } else if (P.i[1] >= A.ni[1]) {
        if (c.pbc[3] == EB200_PBC_AXIS) {
          P.i[1]    = A.ni[1] - 1;
          P.d[1]    = ONE - P.d[1];
          vp_upd[1] = -vp_upd[1];
        }
      }
      if (c.tag_outgoing) {
        const int d0 = (P.i[0] < 0) ? 0 : ((P.i[0] >= A.ni[0]) ? 2 : 1);
        const int d1 = (P.i[1] < 0) ? 0 : ((P.i[1] >= A.ni[1]) ? 2 : 1);
        const int lin = d0 * 3 + d1;
        if (lin != 4) {
          P.tag = static_cast<short>((2 + lin - (lin > 4 ? 1 : 0)) * P.tag);
        }
      }
      S.i1[p]       = P.i[0];
      S.i2[p]       = P.i[1];
      S.dx1[p]      = P.d[0];
      S.dx2[p]      = P.d[1];
      S.i1_prev[p]  = P.ip[0];
      S.i2_prev[p]  = P.ip[1];
      S.dx1_prev[p] = P.dp[0];
      S.dx2_prev[p] = P.dp[1];
      S.ux1[p]      = vp_upd[0];
      S.ux2[p]      = vp_upd[1];
      S.ux3[p]      = vp_upd[2];
      if (P.tag != tag) {
        S.tag[p] = P.tag;
      }
    }

    // currents_deposit.hpp:140-154
    template <class M>
    struct VpGR {
      MetricParams m;

      __device__ __forceinline__ void operator()(const Prtl<2>& P, uint32_t, float* vp) const {
        const float x1 = static_cast<float>(P.i[0]) + P.d[0];
        const float x2 = clamp_x2<M>(m, static_cast<float>(P.i[1]) + P.d[1]);
        gr_cov_to_cntrv<M>(m, x1, x2, P.u, vp);
        const float inv_energy = M::alpha(m, x1, x2) /
                                 sqrtf(ONE + P.u[0] * vp[0] + P.u[1] * vp[1] + P.u[2] * vp[2]);
        if (isnan(vp[2]) || isinf(vp[2])) {
          vp[2] = ZERO;
        }
        vp[0] *= inv_energy;
        vp[1] *= inv_energy;
        vp[2] *= inv_energy;
      }
    };

    /* ------------------------------------------------------------ GR field kernels */
    // rows/columns: i1 from G-1, i2 from G; extents passed by the launcher
    // (grpic::range_with_axis_BCs, src/engines/grpic/fieldsolvers.h:55-76)
    template <class M>
    __global__ void __launch_bounds__(128)
      aux_e_kernel(MetricParams m, FieldView<2> Df, FieldView<2> Bf, FieldView<2> Ef, int ncols,
                   int nrows, int G) {
      const int i1 = G - 1 + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G - 1 + ncols || i2 >= G + nrows) {
        return;
      }
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      const float h_11_pH0 = M::h_11(m, x1 + HALF, x2), h_22_0pH = M::h_22(m, x1, x2 + HALF),
                  h_33_00 = M::h_33(m, x1, x2);
      const float alpha_00 = M::alpha(m, x1, x2), alpha_pH0 = M::alpha(m, x1 + HALF, x2),
                  alpha_0pH = M::alpha(m, x1, x2 + HALF);
      float w1, w2, ww, h13_1, h13_2, sd1, sd2, b1, b2;
      w1    = M::sqrt_det_h_tilde(m, x1 - HALF, x2);
      w2    = M::sqrt_det_h_tilde(m, x1 + HALF, x2);
      ww    = TWO * M::sqrt_det_h_tilde(m, x1, x2);
      h13_1 = M::h_13(m, x1 - HALF, x2);
      h13_2 = M::h_13(m, x1 + HALF, x2);
      const float D1_half = (w1 * h13_1 * Df.at(i1 - 1, i2, 0, ex1) + w2 * h13_2 * Df.at(i1, i2, 0, ex1)) / (ww);
      sd1 = M::sqrt_det_h(m, x1 - HALF, x2);
      sd2 = M::sqrt_det_h(m, x1 + HALF, x2);
      b1  = M::beta1(m, x1 - HALF, x2);
      b2  = M::beta1(m, x1 + HALF, x2);
      const float B2_half = (w1 * sd1 * b1 * Bf.at(i1 - 1, i2, 0, bx2) + w2 * sd2 * b2 * Bf.at(i1, i2, 0, bx2)) / (ww);
      w1  = M::sqrt_det_h_tilde(m, x1 - HALF, x2 + HALF);
      w2  = M::sqrt_det_h_tilde(m, x1 + HALF, x2 + HALF);
      ww  = TWO * M::sqrt_det_h_tilde(m, x1, x2 + HALF);
      sd1 = M::sqrt_det_h(m, x1 - HALF, x2 + HALF);
      sd2 = M::sqrt_det_h(m, x1 + HALF, x2 + HALF);
      b1  = M::beta1(m, x1 - HALF, x2 + HALF);
      b2  = M::beta1(m, x1 + HALF, x2 + HALF);
      const float B3_half = (w1 * sd1 * b1 * Bf.at(i1 - 1, i2, 0, bx3) + w2 * sd2 * b2 * Bf.at(i1, i2, 0, bx3)) / (ww);
      w1    = M::sqrt_det_h_tilde(m, x1, x2);
      w2    = M::sqrt_det_h_tilde(m, x1 + ONE, x2);
      ww    = TWO * M::sqrt_det_h_tilde(m, x1 + HALF, x2);
      h13_1 = M::h_13(m, x1, x2);
      h13_2 = M::h_13(m, x1 + ONE, x2);
      const float D3_half = (w1 * h13_1 * Df.at(i1, i2, 0, ex3) + w2 * h13_2 * Df.at(i1 + 1, i2, 0, ex3)) / (ww);
      const float D1_cov = h_11_pH0 * Df.at(i1, i2, 0, ex1) + D3_half;
      const float D2_cov = h_22_0pH * Df.at(i1, i2, 0, ex2);
      const float D3_cov = h_33_00 * Df.at(i1, i2, 0, ex3) + D1_half;
      Ef.at(i1, i2, 0, ex1) = alpha_pH0 * D1_cov;
      Ef.at(i1, i2, 0, ex2) = alpha_0pH * D2_cov - B3_half;
      Ef.at(i1, i2, 0, ex3) = alpha_00 * D3_cov + B2_half;
    }

    template <class M>
    __global__ void __launch_bounds__(128)
      aux_h_kernel(MetricParams m, FieldView<2> Df, FieldView<2> Bf, FieldView<2> Hf, int ncols,
                   int nrows, int G) {
      const int i1 = G - 1 + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G - 1 + ncols || i2 >= G + nrows) {
        return;
      }
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      const float h_11_0pH = M::h_11(m, x1, x2 + HALF), h_22_pH0 = M::h_22(m, x1 + HALF, x2),
                  h_33_pHpH = M::h_33(m, x1 + HALF, x2 + HALF);
      const float alpha_0pH = M::alpha(m, x1, x2 + HALF), alpha_pH0 = M::alpha(m, x1 + HALF, x2),
                  alpha_pHpH = M::alpha(m, x1 + HALF, x2 + HALF);
      float w1, w2, ww, h13_1, h13_2, sd1, sd2, b1, b2;
      w1    = M::sqrt_det_h_tilde(m, x1, x2 + HALF);
      w2    = M::sqrt_det_h_tilde(m, x1 + ONE, x2 + HALF);
      ww    = TWO * M::sqrt_det_h_tilde(m, x1 + HALF, x2 + HALF);
      h13_1 = M::h_13(m, x1, x2 + HALF);
      h13_2 = M::h_13(m, x1 + ONE, x2 + HALF);
      const float B1_half = (w1 * h13_1 * Bf.at(i1, i2, 0, bx1) + w2 * h13_2 * Bf.at(i1 + 1, i2, 0, bx1)) / (ww);
      sd1 = M::sqrt_det_h(m, x1, x2 + HALF);
      sd2 = M::sqrt_det_h(m, x1 + ONE, x2 + HALF);
      b1  = M::beta1(m, x1, x2 + HALF);
      b2  = M::beta1(m, x1 + ONE, x2 + HALF);
      const float D2_half = (w1 * sd1 * b1 * Df.at(i1, i2, 0, ex2) + w2 * sd2 * b2 * Df.at(i1 + 1, i2, 0, ex2)) / (ww);
      w1  = M::sqrt_det_h_tilde(m, x1, x2);
      w2  = M::sqrt_det_h_tilde(m, x1 + ONE, x2);
      ww  = TWO * M::sqrt_det_h_tilde(m, x1 + HALF, x2);
      sd1 = M::sqrt_det_h(m, x1, x2);
      sd2 = M::sqrt_det_h(m, x1 + ONE, x2);
      b1  = M::beta1(m, x1, x2);
      b2  = M::beta1(m, x1 + ONE, x2);
      const float D3_half = (w1 * sd1 * b1 * Df.at(i1, i2, 0, ex3) + w2 * sd2 * b2 * Df.at(i1 + 1, i2, 0, ex3)) / (ww);
      w1    = M::sqrt_det_h_tilde(m, x1 - HALF, x2 + HALF);
      w2    = M::sqrt_det_h_tilde(m, x1 + HALF, x2 + HALF);
      ww    = TWO * M::sqrt_det_h_tilde(m, x1, x2 + HALF);
      h13_1 = M::h_13(m, x1 - HALF, x2 + HALF);
      h13_2 = M::h_13(m, x1 + HALF, x2 + HALF);
      const float B3_half = (w1 * h13_1 * Bf.at(i1 - 1, i2, 0, bx3) + w2 * h13_2 * Bf.at(i1, i2, 0, bx3)) / (ww);
      const float B1_cov = h_11_0pH * Bf.at(i1, i2, 0, bx1) + B3_half;
      const float B2_cov = h_22_pH0 * Bf.at(i1, i2, 0, bx2);
      const float B3_cov = h_33_pHpH * Bf.at(i1, i2, 0, bx3) + B1_half;
      Hf.at(i1, i2, 0, bx1) = alpha_0pH * B1_cov;
      Hf.at(i1, i2, 0, bx2) = alpha_pH0 * B2_cov + D3_half;
      Hf.at(i1, i2, 0, bx3) = alpha_pHpH * B3_cov - D2_half;
    }

    // faraday_gr.hpp:63-93 (element-wise in Bin/Bout: in place or out of place alike)
    template <class M>
    __global__ void __launch_bounds__(128)
      faraday_gr_kernel(MetricParams m, FieldView<2> Bin, FieldView<2> Bout, FieldView<2> E,
                        int n1, int n2, int G, float coeff, bool axis_min) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 >= G + n2) {
        return;
      }
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      const float inv_sqrt_detH_0pH  = ONE / M::sqrt_det_h(m, x1, x2 + HALF);
      const float inv_sqrt_detH_pHpH = ONE / M::sqrt_det_h(m, x1 + HALF, x2 + HALF);
      Bout.at(i1, i2, 0, bx1) = Bin.at(i1, i2, 0, bx1) +
                                coeff * inv_sqrt_detH_0pH *
                                  (E.at(i1, i2, 0, ex3) - E.at(i1, i2 + 1, 0, ex3));
      if ((i2 == G) && axis_min) {
        Bout.at(i1, i2, 0, bx2) = ZERO;
      } else {
        const float inv_sqrt_detH_pH0 = ONE / M::sqrt_det_h(m, x1 + HALF, x2);
        Bout.at(i1, i2, 0, bx2) = Bin.at(i1, i2, 0, bx2) +
                                  coeff * inv_sqrt_detH_pH0 *
                                    (E.at(i1 + 1, i2, 0, ex3) - E.at(i1, i2, 0, ex3));
      }
      Bout.at(i1, i2, 0, bx3) = Bin.at(i1, i2, 0, bx3) +
                                coeff * inv_sqrt_detH_pHpH *
                                  (E.at(i1, i2 + 1, 0, ex1) - E.at(i1, i2, 0, ex1) +
                                   E.at(i1, i2, 0, ex2) - E.at(i1 + 1, i2, 0, ex2));
    }

    // ampere_gr.hpp:65-102 over i2 in [G, G + n2]. The two axis rows copy dx1 from the
    // neighbouring row. In place (Din == Dout) the reference is order dependent; the serial
    // row-ascending order is reproduced: the row at theta = 0 gets the neighbour's OLD value,
    // the row at theta = pi the neighbour's NEW value. Both copies are done by the thread that
    // owns the neighbouring row, so the kernel has no race.
    template <class M>
    __global__ void __launch_bounds__(128)
      ampere_gr_kernel(MetricParams m, FieldView<2> Din, FieldView<2> Dout, FieldView<2> H, int n1,
                       int n2, int G, float coeff, bool axis_min, bool axis_max) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 > G + n2) {
        return;
      }
      const int   i2min = G, i2max = n2 + G;
      const bool  in_place = (Din.p == Dout.p);
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      const float inv_sqrt_detH_0pH = ONE / M::sqrt_det_h(m, x1, x2 + HALF);
      if ((i2 == i2min) && axis_min) {
        const float inv0 = ONE / M::sqrt_det_h(m, x1, HALF);
        Dout.at(i1, i2, 0, ex2) = Din.at(i1, i2, 0, ex2) +
                                  coeff * inv0 * (H.at(i1 - 1, i2, 0, bx3) - H.at(i1, i2, 0, bx3));
      } else if ((i2 == i2max) && axis_max) {
        // dx1 of this row is written by the owner of row i2max - 1
      } else {
        const float inv_sqrt_detH_00  = ONE / M::sqrt_det_h(m, x1, x2);
        const float inv_sqrt_detH_pH0 = ONE / M::sqrt_det_h(m, x1 + HALF, x2);
        const float d1_old            = Din.at(i1, i2, 0, ex1);
        const float d1_new            = d1_old + coeff * inv_sqrt_detH_pH0 *
                                          (H.at(i1, i2, 0, bx3) - H.at(i1, i2 - 1, 0, bx3));
        if ((i2 == i2min + 1) && axis_min) {
          Dout.at(i1, i2min, 0, ex1) = d1_old;
        }
        if ((i2 == i2max - 1) && axis_max) {
          Dout.at(i1, i2max, 0, ex1) = in_place ? d1_new : d1_old;
        }
        Dout.at(i1, i2, 0, ex1) = d1_new;
        Dout.at(i1, i2, 0, ex2) = Din.at(i1, i2, 0, ex2) +
                                  coeff * inv_sqrt_detH_0pH *
                                    (H.at(i1 - 1, i2, 0, bx3) - H.at(i1, i2, 0, bx3));
        Dout.at(i1, i2, 0, ex3) = Din.at(i1, i2, 0, ex3) +
                                  coeff * inv_sqrt_detH_00 *
                                    ((H.at(i1, i2 - 1, 0, bx1) - H.at(i1, i2, 0, bx1)) +
                                     (H.at(i1, i2, 0, bx2) - H.at(i1 - 1, i2, 0, bx2)));
      }
    }

    // ampere_gr.hpp:143-176
    template <class M>
    __global__ void __launch_bounds__(128)
      currents_ampere_gr_kernel(MetricParams m, FieldView<2> Df, FieldView<2> J, int n1, int n2,
                                int G, float coeff, bool axis_min, bool axis_max) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 > G + n2) {
        return;
      }
      const int   i2min = G, i2max = n2 + G;
      const float x1 = static_cast<float>(i1 - G), x2 = static_cast<float>(i2 - G);
      const float inv_sqrt_detH_0pH = ONE / M::sqrt_det_h(m, x1, x2 + HALF);
      if ((i2 == i2min) && axis_min) {
        Df.at(i1, i2, 0, ex1) += J.at(i1, i2, 0, jx1) * HALF * coeff / M::polar_area(m, x1 + HALF);
        Df.at(i1, i2, 0, ex2) += J.at(i1, i2, 0, jx2) * coeff * inv_sqrt_detH_0pH;
      } else if ((i2 == i2max) && axis_max) {
        Df.at(i1, i2, 0, ex1) += J.at(i1, i2, 0, jx1) * HALF * coeff / M::polar_area(m, x1 + HALF);
      } else {
        const float inv_sqrt_detH_00  = ONE / M::sqrt_det_h(m, x1, x2);
        const float inv_sqrt_detH_pH0 = ONE / M::sqrt_det_h(m, x1 + HALF, x2);
        Df.at(i1, i2, 0, ex1) += J.at(i1, i2, 0, jx1) * coeff * inv_sqrt_detH_pH0;
        Df.at(i1, i2, 0, ex2) += J.at(i1, i2, 0, jx2) * coeff * inv_sqrt_detH_0pH;
        Df.at(i1, i2, 0, ex3) += J.at(i1, i2, 0, jx3) * coeff * inv_sqrt_detH_00;
      }
    }

    // TimeAverageDB_kernel / TimeAverageJ_kernel (aux_fields_gr.hpp:253-302): a = (a + b) / 2
    __global__ void __launch_bounds__(256)
      average_kernel(float* a, const float* b, int N1, int n1, int n2, int G, long plane, int ncomp) {
      const int i1 = G + blockIdx.x * blockDim.x + threadIdx.x;
      const int i2 = G + blockIdx.y;
      if (i1 >= G + n1 || i2 >= G + n2) {
        return;
      }
      const long e = i1 + (long)N1 * i2;
      for (int c = 0; c < ncomp; ++c) {
        a[e + plane * c] = HALF * (a[e + plane * c] + b[e + plane * c]);
      }
    }

#define EB200_GR_METRIC_DO(kind, CALL)                                                         \
  switch (kind) {                                                                              \
    case EB200_METRIC_KERR_SCHILD: CALL(KerrSchild); break;                                    \
    case EB200_METRIC_QKERR_SCHILD: CALL(QKerrSchild); break;                                  \
    case EB200_METRIC_KERR_SCHILD_0: CALL(KerrSchild0); break;                                 \
    default: return cudaErrorInvalidValue;                                                     \
  }

    template <class M>
    static cudaError_t push_gr_m(const MetricParams& m, const eb200_grid_t& g, int order,
                                 const GRPushArgs& A, const eb200_prtls_t& S, uint32_t npart,
                                 const float* em, const float* em0, cudaStream_t st) {
      FieldView<2>   DB(g, const_cast<float*>(em)), DB0(g, const_cast<float*>(em0));
      const unsigned nb = (npart + 127) / 128;
      if (A.c.pusher_flags == EB200_PUSHER_PHOTON) {
#define CALL(O) push_gr_kernel<M, O, false><<<nb, 128, 0, st>>>(A, m, S, npart, DB, DB0)
        EB200_ORDER_DO(order, CALL)
#undef CALL
      } else {
#define CALL(O) push_gr_kernel<M, O, true><<<nb, 128, 0, st>>>(A, m, S, npart, DB, DB0)
        EB200_ORDER_DO(order, CALL)
#undef CALL
      }
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t push_gr(const MetricParams& m, const eb200_grid_t& g, int order,
                        const eb200_pusher_gr_t& c, const eb200_prtls_t& S, uint32_t npart,
                        const float* em, const float* em0, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      GRPushArgs A;
      A.c     = c;
      A.ndh   = HALF * (c.charge / c.mass) * c.omegaB0 * c.dt;
      A.ng    = g.ng;
      A.ni[0] = g.n[0];
      A.ni[1] = g.n[1];
#define CALL(M) return push_gr_m<M>(m, g, order, A, S, npart, em, em0, st)
      EB200_GR_METRIC_DO(m.kind, CALL)
#undef CALL
      return cudaErrorInvalidValue;
    }

    cudaError_t deposit_gr(const MetricParams& m, const eb200_grid_t& g, int order,
                           const eb200_prtls_t& S, uint32_t npart, float charge, float dt,
                           float* cur, int mode, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
#define CALL(M) return deposit_vp(VpGR<M> { m }, g, order, S, npart, charge, dt, cur, mode, st)
      EB200_GR_METRIC_DO(m.kind, CALL)
#undef CALL
      return cudaErrorInvalidValue;
    }

    cudaError_t aux_gr(const MetricParams& m, const eb200_grid_t& g, int which_h, const float* Df,
                       const float* Bf, float* out, const int* fbc, cudaStream_t st) {
      FieldView<2> D(g, const_cast<float*>(Df)), B(g, const_cast<float*>(Bf)), O(g, out);
      const int    ncols = g.n[0] + 1, nrows = g.n[1] + (fbc[3] == EB200_FBC_AXIS ? 1 : 0);
      const dim3   grid((ncols + 127) / 128, nrows, 1);
      if (which_h) {
#define CALL(M) aux_h_kernel<M><<<grid, 128, 0, st>>>(m, D, B, O, ncols, nrows, g.ng)
        EB200_GR_METRIC_DO(m.kind, CALL)
#undef CALL
      } else {
#define CALL(M) aux_e_kernel<M><<<grid, 128, 0, st>>>(m, D, B, O, ncols, nrows, g.ng)
        EB200_GR_METRIC_DO(m.kind, CALL)
#undef CALL
      }
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t faraday_gr(const MetricParams& m, const eb200_grid_t& g, const float* Bin,
                           float* Bout, const float* E, float coeff, const int* fbc,
                           cudaStream_t st) {
      FieldView<2> I(g, const_cast<float*>(Bin)), O(g, Bout), EE(g, const_cast<float*>(E));
      const bool   axis_min = fbc[2] == EB200_FBC_AXIS;
#define CALL(M) faraday_gr_kernel<M><<<cell_grid(g.n[0], g.n[1]), 128, 0, st>>>(m, I, O, EE, g.n[0], g.n[1], g.ng, coeff, axis_min)
      EB200_GR_METRIC_DO(m.kind, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t ampere_gr(const MetricParams& m, const eb200_grid_t& g, const float* Din,
                          float* Dout, const float* H, float coeff, const int* fbc,
                          cudaStream_t st) {
      FieldView<2> I(g, const_cast<float*>(Din)), O(g, Dout), HH(g, const_cast<float*>(H));
      const bool   axis_min = fbc[2] == EB200_FBC_AXIS, axis_max = fbc[3] == EB200_FBC_AXIS;
#define CALL(M) ampere_gr_kernel<M><<<cell_grid(g.n[0], g.n[1] + 1), 128, 0, st>>>(m, I, O, HH, g.n[0], g.n[1], g.ng, coeff, axis_min, axis_max)
      EB200_GR_METRIC_DO(m.kind, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t currents_ampere_gr(const MetricParams& m, const eb200_grid_t& g, float* Df,
                                   const float* cur, float coeff, const int* fbc,
                                   cudaStream_t st) {
      FieldView<2> D(g, Df), J(g, const_cast<float*>(cur));
      const bool   axis_min = fbc[2] == EB200_FBC_AXIS, axis_max = fbc[3] == EB200_FBC_AXIS;
#define CALL(M) currents_ampere_gr_kernel<M><<<cell_grid(g.n[0], g.n[1] + 1), 128, 0, st>>>(m, D, J, g.n[0], g.n[1], g.ng, coeff, axis_min, axis_max)
      EB200_GR_METRIC_DO(m.kind, CALL)
#undef CALL
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t time_average(const eb200_grid_t& g, float* a, const float* b, int ncomp,
                             cudaStream_t st) {
      const int  N1 = g.n[0] + 2 * g.ng, N2 = g.n[1] + 2 * g.ng;
      const dim3 grid((g.n[0] + 255) / 256, g.n[1], 1);
      average_kernel<<<grid, 256, 0, st>>>(a, b, N1, g.n[0], g.n[1], g.ng, (long)N1 * N2, ncomp);
      count_launch();
      return cudaGetLastError();
    }

  } // namespace curv
} // namespace eb200
