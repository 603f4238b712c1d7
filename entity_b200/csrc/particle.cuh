// entity_b200 -- per-particle device arithmetic of the SR pusher and the current deposit.
//
// Register-resident formulation: a particle is loaded once into a `Prtl` struct, pushed,
// deposited and stored once. The arithmetic (operand order included) follows
//   src/kernels/pushers/sr.hpp          (gather :851-1370, Boris :337-368, Vay :370-437,
//                                        GCA :439-521, drag :1372-1424/:1487-1499,
//                                        position push :526-572, boundaries :659-814)
//   src/kernels/currents_deposit.hpp    (zig-zag :171-405, Esirkepov :406-754)
//   src/kernels/particle_shapes.hpp     (order<> :544-613, for_deposit<> :934-1024)
// so that the EB200_STRICT build reproduces the reference bit for bit.
#pragma once
#include "common.cuh"
#include "launch.h"

#include <type_traits>

namespace eb200 {

  template <int D>
  struct Prtl {
    int   i[3], ip[3];
    float d[3], dp[3];
    float u[3];
    float w;
    short tag;
  };

  /* ------------------------------------------------------------------ shapes */
  __device__ __forceinline__ float S3(float x) {
    if (x < ONE) {
      return static_cast<float>(2.0 / 3.0) - SQR(x) + HALF * CUBE(x);
    } else if (x < TWO) {
      return static_cast<float>(4.0 / 3.0) - TWO * x + SQR(x) -
             static_cast<float>(1.0 / 6.0) * CUBE(x);
    }
    return ZERO;
  }

  // S4 .. S11 (particle_shapes.hpp:60-542): centred cardinal B-splines of degree O. The tables
  // (scripts/gen_bspline.py, exact rational arithmetic from the closed form) hold every piece
  // as a polynomial in t = |x| - (left end of the piece), 0 <= t < 1, coefficients of magnitude
  // <= 1: Horner in fp32 is then accurate to an ulp or two. The reference sums the monomials in
  // |x| itself, whose terms reach ~9 for a result of ~1e-5 at O = 11: its own weights carry a
  // cancellation error of up to ~1e-5 there (tests/test_gpu_hiorder.py measures both against
  // the fp64 closed form).
  template <int O>
  struct BSplineCoef;
#include "bspline_coef.inc"

  template <int O>
  __device__ __forceinline__ float bspline(float x) {
    constexpr float first = (O % 2 == 0) ? HALF : ONE;
    const int       p     = (x < first) ? 0 : static_cast<int>(x - first) + 1;
    if (p >= BSplineCoef<O>::NP) return ZERO;
    const float t = (p == 0) ? x : x - (first + static_cast<float>(p - 1));
    float       r = BSplineCoef<O>::c(p, O);
#pragma unroll
    for (int k = O - 1; k >= 0; --k) {
      r = r * t + BSplineCoef<O>::c(p, k);
    }
    return r;
  }

  template <bool STAG, int O>
  __device__ __forceinline__ void shape_w(int i, float di, int& i_min, float (&S)[O + 1]) {
    static_assert(O >= 1 && O <= 11, "shape orders 1..11");
    if constexpr (O >= 4) {
      // order<STAGGERED, O> for O >= 4 (particle_shapes.hpp:608-932): the window of O + 1 nodes
      // starts O/2 (rounded by the parity of O, the staggering and di < 1/2) nodes below i
      float base;
      if constexpr (O % 2 == 1) {
        if constexpr (!STAG) {
          i_min = i - (O - 1) / 2;
          base  = static_cast<float>((O - 1) / 2) + di;
        } else {
          const bool lo = di < HALF;
          i_min         = lo ? i - (O + 1) / 2 : i - (O - 1) / 2;
          base          = (lo ? static_cast<float>(O) * HALF : static_cast<float>(O) * HALF - ONE) + di;
        }
      } else {
        if constexpr (!STAG) {
          const bool lo = di < HALF;
          i_min         = lo ? i - O / 2 : i - O / 2 + 1;
          base          = static_cast<float>(lo ? O / 2 : O / 2 - 1) + di;
        } else {
          i_min = i - O / 2;
          base  = static_cast<float>(O - 1) * HALF + di;
        }
      }
#pragma unroll 1
      for (int n = 0; n <= O; n++) {
        S[n] = bspline<O>(fabsf(base - static_cast<float>(n)));
      }
    } else if constexpr (O == 1) {
      if constexpr (!STAG) {
        i_min = i;
        S[0]  = ONE - di;
        S[1]  = di;
      } else {
        const bool lo = di < HALF;
        i_min         = lo ? i - 1 : i;
        S[0]          = (lo ? HALF : THREE_HALFS) - di;
        S[1]          = ONE - S[0];
      }
    } else if constexpr (O == 2) {
      if constexpr (!STAG) {
        if (di < HALF) {
          i_min = i - 1;
          S[0]  = HALF * SQR(HALF - di);
          S[1]  = THREE_FOURTHS - SQR(di);
        } else {
          i_min = i;
          S[0]  = HALF * SQR(THREE_HALFS - di);
          S[1]  = THREE_FOURTHS - SQR(ONE - di);
        }
        S[2] = ONE - S[0] - S[1];
      } else {
        i_min = i - 1;
        S[0]  = HALF * SQR(ONE - di);
        S[2]  = HALF * SQR(di);
        S[1]  = ONE - S[0] - S[2];
      }
    } else {
      float base;
      if constexpr (!STAG) {
        i_min = i - 1;
        base  = ONE + di;
      } else {
        const bool lo = di < HALF;
        i_min         = lo ? i - 2 : i - 1;
        base          = (lo ? 1.5f : HALF) + di;
      }
#pragma unroll
      for (int n = 0; n < 4; n++) {
        S[n] = S3(fabsf(base - static_cast<float>(n)));
      }
    }
  }

  // initial/final shape arrays aligned on a common (O+2)-wide window
  template <int O>
  __device__ __forceinline__ void deposit_shapes(int i_init, float di_init, int i_fin,
                                                 float di_fin, int& i_min, int& i_max,
                                                 float (&iS)[O + 2], float (&fS)[O + 2]) {
    int   a_min, b_min;
    float a[O + 1], b[O + 1];
    shape_w<false, O>(i_init, di_init, a_min, a);
    shape_w<false, O>(i_fin, di_fin, b_min, b);
    const int sa = (a_min > b_min) ? 1 : 0; // shift of the initial shape inside the window
    const int sb = (a_min < b_min) ? 1 : 0; // shift of the final shape
    i_min        = (a_min < b_min) ? a_min : b_min;
    i_max        = i_min + O + ((a_min != b_min) ? 1 : 0);
#pragma unroll
    for (int j = 0; j < O + 2; ++j) {
      iS[j] = ZERO;
      fS[j] = ZERO;
    }
#pragma unroll
    for (int j = 0; j < O + 1; ++j) {
      // static indexing in both branches keeps the arrays in registers
      if (sa) {
        iS[j + 1] = a[j];
      } else {
        iS[j] = a[j];
      }
      if (sb) {
        fS[j + 1] = b[j];
      } else {
        fS[j] = b[j];
      }
    }
  }

  /* ------------------------------------------------- reciprocal / rsqrt / divide */
  // Strict build: IEEE division and square root in the reference's operation order.
  // Fast build: MUFU.RCP / MUFU.RSQ based forms (<= 2 ulp), an order of magnitude fewer
  // instructions than the IEEE sequences; covered by the fast-build tolerances in tests/.
  __device__ __forceinline__ float rcp_sqrt(float x) {
#if EB200_STRICT
    return ONE / sqrtf(x);
#else
    return rsqrtf(x);
#endif
  }

  __device__ __forceinline__ float div_by_sqrt(float a, float x) {
#if EB200_STRICT
    return a / sqrtf(x);
#else
    return a * rsqrtf(x);
#endif
  }

  __device__ __forceinline__ float fdiv(float a, float b) {
#if EB200_STRICT
    return a / b;
#else
    return __fdividef(a, b);
#endif
  }

  // a / dx with the cell size's reciprocal precomputed on the host (fast build only)
  __device__ __forceinline__ float div_dx(float a, float dx, float inv_dx) {
#if EB200_STRICT
    (void)inv_dx;
    return a / dx;
#else
    (void)dx;
    return a * inv_dx;
#endif
  }

  /* ---------------------------------------------------------------- vector ops */
  __device__ __forceinline__ float dot3(const float* a, const float* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  }

  __device__ __forceinline__ float nsq(const float* a) { return dot3(a, a); }

  __device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
  }

  /* ------------------------------------------------------------------- gather */
  // EM: callable (i, j, k, comp) -> float in ghost-inclusive indices.
  template <int D, int O, class EM>
  __device__ __forceinline__ void gather_fields(const EM& F, int ng, const Prtl<D>& P, float* e0,
                                                float* b0) {
    if constexpr (O == 0) {
      // staggered multilinear interpolation; per axis a primal pair {1-d, d} at nodes
      // (i, i+1) and a dual pair at nodes (i-1+ind, i+ind), ind = int(d + 1/2)
      int   base[3] = { 0, 0, 0 }, dbase[3] = { 0, 0, 0 };
      float wp[3][2], wd[3][2];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const int ind = static_cast<int>(P.d[a] + HALF);
        base[a]       = P.i[a] + ng;
        dbase[a]      = base[a] - 1 + ind;
        wp[a][0]      = ONE - P.d[a];
        wp[a][1]      = P.d[a];
        wd[a][0]      = static_cast<float>(ind + 1) - (P.d[a] + HALF);
        wd[a][1]      = ONE - wd[a][0];
      }
      // sx/sy/sz: which axes read the dual nodes; wy_dual: which x2 weights are applied.
      // The reference weights Bx3 in 3D with the primal x2 pair (sr.hpp:1102-1116).
      auto lerp = [&](int c, bool sx, bool sy, bool sz, bool wy_dual) -> float {
        const int    i0 = sx ? dbase[0] : base[0];
        const float* wx = sx ? wd[0] : wp[0];
        if constexpr (D == 1) {
          return F(i0, 0, 0, c) * wx[0] + F(i0 + 1, 0, 0, c) * wx[1];
        } else if constexpr (D == 2) {
          const int    j0  = sy ? dbase[1] : base[1];
          const float* wy  = wy_dual ? wd[1] : wp[1];
          const float  c00 = F(i0, j0, 0, c) * wx[0] + F(i0 + 1, j0, 0, c) * wx[1];
          const float  c10 = F(i0, j0 + 1, 0, c) * wx[0] + F(i0 + 1, j0 + 1, 0, c) * wx[1];
          return c00 * wy[0] + c10 * wy[1];
        } else {
          const int    j0  = sy ? dbase[1] : base[1];
          const int    k0  = sz ? dbase[2] : base[2];
          const float* wy  = wy_dual ? wd[1] : wp[1];
          const float* wz  = sz ? wd[2] : wp[2];
          const float  c00 = F(i0, j0, k0, c) * wx[0] + F(i0 + 1, j0, k0, c) * wx[1];
          const float  c10 = F(i0, j0 + 1, k0, c) * wx[0] + F(i0 + 1, j0 + 1, k0, c) * wx[1];
          const float  c0  = c00 * wy[0] + c10 * wy[1];
          const float  c01 = F(i0, j0, k0 + 1, c) * wx[0] + F(i0 + 1, j0, k0 + 1, c) * wx[1];
          const float  c11 = F(i0, j0 + 1, k0 + 1, c) * wx[0] +
                            F(i0 + 1, j0 + 1, k0 + 1, c) * wx[1];
          const float c1 = c01 * wy[0] + c11 * wy[1];
          return c0 * wz[0] + c1 * wz[1];
        }
      };
      e0[0] = lerp(ex1, true, false, false, false);
      e0[1] = lerp(ex2, false, true, false, true);
      e0[2] = lerp(ex3, false, false, true, false);
      b0[0] = lerp(bx1, false, true, true, true);
      b0[1] = lerp(bx2, true, false, true, false);
      b0[2] = lerp(bx3, true, true, false, D != 3);
    } else if constexpr (O > 3) {
      // orders 4..11: the same sums with rolled loops (up to 12^3 nodes per component)
      int   pmin[3] = { 0, 0, 0 }, dmin[3] = { 0, 0, 0 };
      float Sp[3][O + 1], Sd[3][O + 1];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        shape_w<false, O>(P.i[a] + ng, P.d[a], pmin[a], Sp[a]);
        shape_w<true, O>(P.i[a] + ng, P.d[a], dmin[a], Sd[a]);
      }
      auto spline = [&](int c, bool sx, bool sy, bool sz) -> float {
        const float* S1  = sx ? Sd[0] : Sp[0];
        const int    m1  = sx ? dmin[0] : pmin[0];
        const float* S2  = (D > 1) ? (sy ? Sd[1] : Sp[1]) : S1;
        const int    m2  = (D > 1) ? (sy ? dmin[1] : pmin[1]) : 0;
        const float* S3w = (D > 2) ? (sz ? Sd[2] : Sp[2]) : S1;
        const int    m3  = (D > 2) ? (sz ? dmin[2] : pmin[2]) : 0;
        float        r   = ZERO;
#pragma unroll 1
        for (int q = 0; q < ((D > 2) ? O + 1 : 1); q++) {
          float c0 = ZERO;
#pragma unroll 1
          for (int b = 0; b < ((D > 1) ? O + 1 : 1); b++) {
            float c00 = ZERO;
#pragma unroll 1
            for (int a = 0; a < O + 1; a++) c00 += S1[a] * F(m1 + a, m2 + b, m3 + q, c);
            if constexpr (D > 1) {
              c0 += c00 * S2[b];
            } else {
              c0 = c00;
            }
          }
          if constexpr (D > 2) {
            r += c0 * S3w[q];
          } else {
            r = c0;
          }
        }
        return r;
      };
      e0[0] = spline(ex1, true, false, false);
      e0[1] = spline(ex2, false, true, false);
      e0[2] = spline(ex3, false, false, true);
      b0[0] = spline(bx1, false, true, true);
      b0[1] = spline(bx2, true, false, true);
      b0[2] = spline(bx3, true, true, false);
    } else {
      int   pmin[3] = { 0, 0, 0 }, dmin[3] = { 0, 0, 0 };
      float Sp[3][O + 1], Sd[3][O + 1];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        shape_w<false, O>(P.i[a] + ng, P.d[a], pmin[a], Sp[a]);
        shape_w<true, O>(P.i[a] + ng, P.d[a], dmin[a], Sd[a]);
      }
      auto spline = [&](int c, bool sx, bool sy, bool sz) -> float {
        const float* S1 = sx ? Sd[0] : Sp[0];
        const int    m1 = sx ? dmin[0] : pmin[0];
        if constexpr (D == 1) {
          float r = ZERO;
#pragma unroll
          for (int a = 0; a < O + 1; a++) r += S1[a] * F(m1 + a, 0, 0, c);
          return r;
        } else if constexpr (D == 2) {
          const float* S2 = sy ? Sd[1] : Sp[1];
          const int    m2 = sy ? dmin[1] : pmin[1];
          float        r  = ZERO;
#pragma unroll
          for (int b = 0; b < O + 1; b++) {
            float c0 = ZERO;
#pragma unroll
            for (int a = 0; a < O + 1; a++) c0 += S1[a] * F(m1 + a, m2 + b, 0, c);
            r += c0 * S2[b];
          }
          return r;
        } else {
          const float* S2 = sy ? Sd[1] : Sp[1];
          const int    m2 = sy ? dmin[1] : pmin[1];
          const float* S3w = sz ? Sd[2] : Sp[2];
          const int    m3 = sz ? dmin[2] : pmin[2];
          float        r  = ZERO;
#pragma unroll
          for (int q = 0; q < O + 1; q++) {
            float c0 = ZERO;
#pragma unroll
            for (int b = 0; b < O + 1; b++) {
              float c00 = ZERO;
#pragma unroll
              for (int a = 0; a < O + 1; a++) c00 += S1[a] * F(m1 + a, m2 + b, m3 + q, c);
              c0 += c00 * S2[b];
            }
            r += c0 * S3w[q];
          }
          return r;
        }
      };
      e0[0] = spline(ex1, true, false, false);
      e0[1] = spline(ex2, false, true, false);
      e0[2] = spline(ex3, false, false, true);
      b0[0] = spline(bx1, false, true, true);
      b0[1] = spline(bx2, true, false, true);
      b0[2] = spline(bx3, true, true, false);
    }
  }

  // field-node load that asks L1 to keep the line (the particle stream does not allocate there)
  __device__ __forceinline__ float ld_keep(const float* p) {
#ifdef EB200_X_LDG
    return __ldg(p);
#else
    float v;
    asm("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#endif
  }

  // Zig-zag (O = 0) gather straight from a FieldView: one 32-bit element index per particle,
  // the primal/dual choice is an element offset, the 2^D nodes of a component are immediate or
  // row-stride offsets from one pointer. Same weights and summation order as gather_fields().
  template <int D>
  __device__ __forceinline__ void gather_fields_direct(const FieldView<D>& EB, int ng,
                                                       const Prtl<D>& P, float* e0, float* b0) {
    float     wp[3][2], wd[3][2];
    int       od[3]  = { 0, 0, 0 };
    const int st[3]  = { 1, EB.N1, EB.N1 * EB.N2 };
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const int ind = static_cast<int>(P.d[a] + HALF);
      od[a]         = (ind - 1) * st[a];
      wp[a][0]      = ONE - P.d[a];
      wp[a][1]      = P.d[a];
      wd[a][0]      = static_cast<float>(ind + 1) - (P.d[a] + HALF);
      wd[a][1]      = ONE - wd[a][0];
    }
    const int e = static_cast<int>(
      EB.idx(P.i[0] + ng, (D > 1) ? P.i[1] + ng : 0, (D > 2) ? P.i[2] + ng : 0));
    auto lerp = [&](int c, bool sx, bool sy, bool sz, bool wy_dual) -> float {
#ifdef EB200_X_NOGATHER
      return 1e-3f * static_cast<float>(c + 1) + 1e-6f * static_cast<float>(e) + wp[0][0];
#endif
      const float* bc = EB.p + EB.plane * c;
      const int    o  = e + (sx ? od[0] : 0) + ((D > 1 && sy) ? od[1] : 0) +
                    ((D > 2 && sz) ? od[2] : 0);
      const float* wx = sx ? wd[0] : wp[0];
      const float* q  = bc + o;
      if constexpr (D == 1) {
        return ld_keep(q) * wx[0] + ld_keep(q + 1) * wx[1];
      } else if constexpr (D == 2) {
        const float* wy  = wy_dual ? wd[1] : wp[1];
        const float* r   = bc + (o + st[1]);
        const float  c00 = ld_keep(q) * wx[0] + ld_keep(q + 1) * wx[1];
        const float  c10 = ld_keep(r) * wx[0] + ld_keep(r + 1) * wx[1];
        return c00 * wy[0] + c10 * wy[1];
      } else {
        const float* wy  = wy_dual ? wd[1] : wp[1];
        const float* wz  = sz ? wd[2] : wp[2];
        const float* r   = bc + (o + st[1]);
        const float* q2  = bc + (o + st[2]);
        const float* r2  = bc + (o + st[1] + st[2]);
        const float  c00 = ld_keep(q) * wx[0] + ld_keep(q + 1) * wx[1];
        const float  c10 = ld_keep(r) * wx[0] + ld_keep(r + 1) * wx[1];
        const float  c0  = c00 * wy[0] + c10 * wy[1];
        const float  c01 = ld_keep(q2) * wx[0] + ld_keep(q2 + 1) * wx[1];
        const float  c11 = ld_keep(r2) * wx[0] + ld_keep(r2 + 1) * wx[1];
        const float  c1  = c01 * wy[0] + c11 * wy[1];
        return c0 * wz[0] + c1 * wz[1];
      }
    };
    e0[0] = lerp(ex1, true, false, false, false);
    e0[1] = lerp(ex2, false, true, false, true);
    e0[2] = lerp(ex3, false, false, true, false);
    b0[0] = lerp(bx1, false, true, true, true);
    b0[1] = lerp(bx2, true, false, true, false);
    b0[2] = lerp(bx3, true, true, false, D != 3);
  }

  // E/B of a 2D mesh repacked node by node (24 bytes: {Ex, By | Ey, Bx | Ez, Bz}, i.e. the
  // components grouped by staggering): the 2 x 2 patch of a staggering group is four loads off
  // two row pointers (64-bit loads for the two-component groups), and every address is the
  // uniform base + an unsigned 32-bit byte offset -- 16 loads and ~25 integer instructions per
  // particle instead of 24 loads behind ~70 instructions of 64-bit index arithmetic on the
  // component planes. The copy is made once per step by pack_em2d_kernel (0.4 GB of traffic).
  struct PackedEM2 {
    const char* p;
    unsigned    rowb; // bytes per mesh row: 24 * N1
  };

  __device__ __forceinline__ float2 ld_keep2(const char* p) {
    float2 v;
    asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
  }

  __device__ __forceinline__ float ld_keep1(const char* p) {
    float v;
    asm("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
  }

  // same weights, nodes and summation order as gather_fields_direct<2>()
  __device__ __forceinline__ void gather_packed(const PackedEM2& F, int ng, const Prtl<2>& P,
                                                float* e0, float* b0) {
    float          wp[2][2], wd[2][2];
    unsigned       back[2]; // how far the dual (staggered) patch starts before the primal one
    const unsigned st[2] = { 24u, F.rowb };
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ind = static_cast<int>(P.d[a] + HALF);
      back[a]       = ind ? 0u : st[a];
      wp[a][0]      = ONE - P.d[a];
      wp[a][1]      = P.d[a];
      wd[a][0]      = static_cast<float>(ind + 1) - (P.d[a] + HALF);
      wd[a][1]      = ONE - wd[a][0];
    }
    const unsigned b00 = static_cast<unsigned>(P.i[0] + ng) * 24u +
                         static_cast<unsigned>(P.i[1] + ng) * F.rowb;
    const unsigned b10 = b00 - back[0], b01 = b00 - back[1], b11 = b10 - back[1];
    {
      // staggered in x1: Ex (.x), By (.y); wx dual, wy primal
      const char*  q  = F.p + b10;
      const char*  r  = F.p + (b10 + F.rowb);
      const float2 q0 = ld_keep2(q), q1 = ld_keep2(q + 24), r0 = ld_keep2(r), r1 = ld_keep2(r + 24);
      const float* wx = wd[0];
      const float* wy = wp[1];
      e0[0] = (q0.x * wx[0] + q1.x * wx[1]) * wy[0] + (r0.x * wx[0] + r1.x * wx[1]) * wy[1];
      b0[1] = (q0.y * wx[0] + q1.y * wx[1]) * wy[0] + (r0.y * wx[0] + r1.y * wx[1]) * wy[1];
    }
    {
      // staggered in x2: Ey (.x), Bx (.y); wx primal, wy dual
      const char*  q  = F.p + (b01 + 8u);
      const char*  r  = F.p + (b01 + 8u + F.rowb);
      const float2 q0 = ld_keep2(q), q1 = ld_keep2(q + 24), r0 = ld_keep2(r), r1 = ld_keep2(r + 24);
      const float* wx = wp[0];
      const float* wy = wd[1];
      e0[1] = (q0.x * wx[0] + q1.x * wx[1]) * wy[0] + (r0.x * wx[0] + r1.x * wx[1]) * wy[1];
      b0[0] = (q0.y * wx[0] + q1.y * wx[1]) * wy[0] + (r0.y * wx[0] + r1.y * wx[1]) * wy[1];
    }
    {
      // Ez on the nodes
      const char*  q  = F.p + (b00 + 16u);
      const char*  r  = F.p + (b00 + 16u + F.rowb);
      const float* wx = wp[0];
      const float* wy = wp[1];
      e0[2] = (ld_keep1(q) * wx[0] + ld_keep1(q + 24) * wx[1]) * wy[0] +
              (ld_keep1(r) * wx[0] + ld_keep1(r + 24) * wx[1]) * wy[1];
    }
    {
      // Bz staggered in both
      const char*  q  = F.p + (b11 + 20u);
      const char*  r  = F.p + (b11 + 20u + F.rowb);
      const float* wx = wd[0];
      const float* wy = wd[1];
      b0[2] = (ld_keep1(q) * wx[0] + ld_keep1(q + 24) * wx[1]) * wy[0] +
              (ld_keep1(r) * wx[0] + ld_keep1(r + 24) * wx[1]) * wy[1];
    }
  }

  // A CTA's shared-memory copy of the E/B nodes around its (cell-sorted) particles: TR rows of
  // TC columns per component, origin (c0, r0) in ghost-inclusive node indices. Particles whose
  // 3 x 3 gather neighbourhood lies inside read shared memory (immediate offsets from one
  // local index); the others -- strays of a not quite sorted array -- read global memory.
  template <int TC, int TR>
  struct TileEM {
    const float* t;
    int          c0, r0;
    FieldView<2> G;
  };

  template <class T>
  struct is_tile_em : std::false_type {};
  template <int TC, int TR>
  struct is_tile_em<TileEM<TC, TR>> : std::true_type {};

  template <int TC, int TR>
  __device__ __forceinline__ void gather_tile(const TileEM<TC, TR>& T, int ng, const Prtl<2>& P,
                                              float* e0, float* b0) {
    const int lc = P.i[0] + ng - T.c0, lr = P.i[1] + ng - T.r0;
    if (static_cast<unsigned>(lc - 1) <= static_cast<unsigned>(TC - 3) &&
        static_cast<unsigned>(lr - 1) <= static_cast<unsigned>(TR - 3)) {
      constexpr int TN = TC * TR;
      float         wp[2][2], wd[2][2];
      int           od[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int ind = static_cast<int>(P.d[a] + HALF);
        od[a]         = (ind - 1) * (a == 0 ? 1 : TC);
        wp[a][0]      = ONE - P.d[a];
        wp[a][1]      = P.d[a];
        wd[a][0]      = static_cast<float>(ind + 1) - (P.d[a] + HALF);
        wd[a][1]      = ONE - wd[a][0];
      }
      const float* pp = T.t + (lr * TC + lc);
      const float* dp = pp + od[0];
      const float* pd = pp + od[1];
      const float* dd = dp + od[1];
      auto lerp = [&](const float* q, int c, const float* wx, const float* wy) -> float {
        q += c * TN;
        const float c00 = q[0] * wx[0] + q[1] * wx[1];
        const float c10 = q[TC] * wx[0] + q[TC + 1] * wx[1];
        return c00 * wy[0] + c10 * wy[1];
      };
      e0[0] = lerp(dp, ex1, wd[0], wp[1]);
      e0[1] = lerp(pd, ex2, wp[0], wd[1]);
      e0[2] = lerp(pp, ex3, wp[0], wp[1]);
      b0[0] = lerp(pd, bx1, wp[0], wd[1]);
      b0[1] = lerp(dp, bx2, wd[0], wp[1]);
      b0[2] = lerp(dd, bx3, wd[0], wd[1]);
    } else {
      gather_fields_direct<2>(T.G, ng, P, e0, b0);
    }
  }

  // what push_particle() calls: a FieldView goes through the direct gather where one exists,
  // any other callable through the generic one
  template <int D, int O, class EM>
  __device__ __forceinline__ void gather_any(const EM& F, int ng, const Prtl<D>& P, float* e0,
                                             float* b0) {
    if constexpr (is_tile_em<EM>::value) {
      static_assert(D == 2 && O == 0, "field tiles: 2D zig-zag only");
      gather_tile(F, ng, P, e0, b0);
    } else if constexpr (std::is_same<EM, PackedEM2>::value) {
      static_assert(D == 2 && O == 0, "packed nodes: 2D zig-zag only");
      gather_packed(F, ng, P, e0, b0);
    } else if constexpr (std::is_same<EM, FieldView<D>>::value) {
      if constexpr (O == 0) {
        gather_fields_direct<D>(F, ng, P, e0, b0);
      } else {
        gather_fields<D, O>([&](int i, int j, int k, int c) { return F.ld(i, j, k, c); }, ng, P,
                            e0, b0);
      }
    } else {
      gather_fields<D, O>(F, ng, P, e0, b0);
    }
  }

  /* ---------------------------------------------------------- velocity updates */
  __device__ __forceinline__ void boris(float ndh, float* u, float* e0, float* b0) {
    float c = ndh;
    e0[0] *= c;
    e0[1] *= c;
    e0[2] *= c;
    float u0[3] = { u[0] + e0[0], u[1] + e0[1], u[2] + e0[2] };
    c *= rcp_sqrt(ONE + nsq(u0));
    b0[0] *= c;
    b0[1] *= c;
    b0[2] *= c;
    c = fdiv(TWO, ONE + nsq(b0));
    float x[3], u1[3];
    cross3(u0, b0, x);
    u1[0] = (u0[0] + x[0]) * c;
    u1[1] = (u0[1] + x[1]) * c;
    u1[2] = (u0[2] + x[2]) * c;
    cross3(u1, b0, x);
    u[0] = u0[0] + (x[0] + e0[0]);
    u[1] = u0[1] + (x[1] + e0[1]);
    u[2] = u0[2] + (x[2] + e0[2]);
  }

  __device__ __forceinline__ void vay(float ndh, float* u, float* e0, float* b0) {
    float c = ndh;
    e0[0] *= c;
    e0[1] *= c;
    e0[2] *= c;
    b0[0] *= c;
    b0[1] *= c;
    b0[2] *= c;
    c = ONE / sqrtf(ONE + nsq(u));
    float x[3];
    cross3(u, b0, x);
    const float u1[3] = { (u[0] + TWO * e0[0] + x[0] * c), (u[1] + TWO * e0[1] + x[1] * c),
                          (u[2] + TWO * e0[2] + x[2] * c) };
    c        = dot3(u1, b0);
    float c2 = ONE + nsq(u1) - nsq(b0);
    c        = ONE / sqrtf(INV_2 * (c2 + sqrtf(SQR(c2) + FOUR * (SQR(b0[0]) + SQR(b0[1]) +
                                                              SQR(b0[2]) + SQR(c)))));
    c2 = ONE / (ONE + SQR(b0[0] * c) + SQR(b0[1] * c) + SQR(b0[2] * c));
    const float udb = dot3(u1, b0);
    u[0] = c2 * (u1[0] + c * udb * (b0[0] * c) + u1[1] * b0[2] * c - u1[2] * b0[1] * c);
    u[1] = c2 * (u1[1] + c * udb * (b0[1] * c) + u1[2] * b0[0] * c - u1[0] * b0[2] * c);
    u[2] = c2 * (u1[2] + c * udb * (b0[2] * c) + u1[0] * b0[1] * c - u1[1] * b0[0] * c);
  }

  // guiding-centre update; `f0` may be null (no external force term)
  __device__ __forceinline__ void gca(float ndh, float dt, float* u, const float* f0, float* e0,
                                      float* b0) {
    const float eb_sqr = nsq(e0) + nsq(b0);
    float       wE[3];
    cross3(e0, b0, wE);
    wE[0] /= eb_sqr;
    wE[1] /= eb_sqr;
    wE[2] /= eb_sqr;
    const float b_norm_inv = ONE / sqrtf(nsq(b0));
    b0[0] *= b_norm_inv;
    b0[1] *= b_norm_inv;
    b0[2] *= b_norm_inv;
    float upar = dot3(u, b0) + ndh * TWO * dot3(e0, b0);
    if (f0 != nullptr) {
      upar = upar + dt * dot3(f0, b0);
    }
    const float w2 = nsq(wE);
    float       factor;
    if (w2 < 0.01f) {
      factor = ONE + w2 + TWO * SQR(w2) + FIVE * SQR(w2) * w2;
    } else {
      factor = (ONE - sqrtf(ONE - FOUR * w2)) / (TWO * w2);
    }
    const float vE[3] = { wE[0] * factor, wE[1] * factor, wE[2] * factor };
    const float Gamma = sqrtf(ONE + SQR(upar)) / sqrtf(ONE - nsq(vE));
    u[0]              = upar * b0[0] + vE[0] * Gamma;
    u[1]              = upar * b0[1] + vE[1] * Gamma;
    u[2]              = upar * b0[2] + vE[2] * Gamma;
  }

  __device__ __forceinline__ void synchrotron_drag(float coeff, float* u, float* up,
                                                   const float* e0, const float* b0) {
    float g = ONE / sqrtf(ONE + nsq(up));
    up[0] *= g;
    up[1] *= g;
    up[2] *= g;
    g                = SQR(ONE / g);
    const float bde  = dot3(up, e0);
    float       x[3], kap[3];
    cross3(up, b0, x);
    const float epb[3] = { e0[0] + x[0], e0[1] + x[1], e0[2] + x[2] };
    cross3(epb, b0, kap);
    kap[0] += bde * e0[0];
    kap[1] += bde * e0[1];
    kap[2] += bde * e0[2];
    const float chi = nsq(epb) - SQR(bde);
    u[0] += coeff * (kap[0] - g * up[0] * chi);
    u[1] += coeff * (kap[1] - g * up[1] * chi);
    u[2] += coeff * (kap[2] - g * up[2] * chi);
  }

  __device__ __forceinline__ void compton_drag(float coeff, float* u, float* up) {
    float g = ONE / sqrtf(ONE + nsq(up));
    up[0] *= g;
    up[1] *= g;
    up[2] *= g;
    g = SQR(ONE / g);
    u[0] -= coeff * g * up[0];
    u[1] -= coeff * g * up[1];
    u[2] -= coeff * g * up[2];
  }

  /* ---------------------------------------------- velocity update of a massive particle */
  // Everything between the Cartesian fields at the particle and the position update
  // (sr.hpp:206-309): optional half-kicks by an external force, hybrid GCA / Boris / Vay,
  // radiative drag. `ec`, `bc`, `fext` are Cartesian; `fext` is only read when
  // c.has_atmosphere is set.
  __device__ __forceinline__ void velocity_update(const eb200_pusher_t& c, float ndh, float* u,
                                                  float* ec, float* bc, const float* fext) {
    const float dt        = c.dt;
    float       up[3] = { ZERO, ZERO, ZERO }, er[3] = { ZERO, ZERO, ZERO },
          br[3]       = { ZERO, ZERO, ZERO };
    const bool drag   = c.drag_flags != EB200_DRAG_NONE;
    if (drag) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        er[a] = ec[a];
        br[a] = bc[a];
        up[a] = u[a];
      }
    }
    bool is_gca = false;
    if (c.pusher_flags & EB200_PUSHER_GCA) {
      const float E2 = nsq(ec), B2 = nsq(bc);
      const float rL = sqrtf(ONE + nsq(u)) * dt / (TWO * fabsf(ndh) * sqrtf(B2));
      is_gca = (B2 > ZERO) && (rL < c.gca_larmor_max) && ((E2 / B2) < c.gca_e_ovr_b_sqr_max);
    }
    if (is_gca) {
      gca(ndh, dt, u, c.has_atmosphere ? fext : nullptr, ec, bc);
    } else {
      if (c.has_atmosphere) {
        u[0] += HALF * dt * fext[0];
        u[1] += HALF * dt * fext[1];
        u[2] += HALF * dt * fext[2];
      }
      if (c.pusher_flags & EB200_PUSHER_BORIS) {
        boris(ndh, u, ec, bc);
      } else if (c.pusher_flags & EB200_PUSHER_VAY) {
        vay(ndh, u, ec, bc);
      }
      if (c.has_atmosphere) {
        u[0] += HALF * dt * fext[0];
        u[1] += HALF * dt * fext[1];
        u[2] += HALF * dt * fext[2];
      }
      if (drag) {
        up[0] = HALF * (up[0] + u[0]);
        up[1] = HALF * (up[1] + u[1]);
        up[2] = HALF * (up[2] + u[2]);
        if (c.drag_flags & EB200_DRAG_SYNCHROTRON) {
          synchrotron_drag(c.sync_coeff, u, up, er, br);
        }
        if (c.drag_flags & EB200_DRAG_COMPTON) {
          compton_drag(c.compton_coeff, u, up);
        }
      }
    }
  }

  /* ------------------------------------------------------------- emission policies */
  // arch::emission::{Synchrotron, Compton}::shouldEmit (synchrotron.h:139-224, compton.h:133-164):
  // up = (u_before + u_after) / 2, e / b the Cartesian fields at the particle. Returns the
  // probability; delta_u = the recoil of the emitter, energy = the photon's.
  __device__ __forceinline__ float emission_response(const EmitParams& E, const float* up,
                                                     const float* e, const float* b,
                                                     float* delta_u, float& energy,
                                                     float& gamma) {
    const float u_sqr     = nsq(up);
    const float gamma_sqr = ONE + u_sqr;
    energy                = gamma_sqr * E.nominal_photon_energy;
    if (E.kind == EB200_EMISSION_COMPTON) {
      const float du = -E.photon_weight * energy / (sqrtf(u_sqr) * E.species_mass);
      delta_u[0] = du * up[0], delta_u[1] = du * up[1], delta_u[2] = du * up[2];
      gamma = sqrtf(gamma_sqr);
      return E.nominal_probability * sqrtf(u_sqr / gamma_sqr);
    }
    const float u_mag = sqrtf(u_sqr);
    gamma             = sqrtf(gamma_sqr);
    const float beta  = u_mag / gamma;
    float       x[3];
    cross3(up, b, x);
    const float epb[3] = { e[0] + x[0] / gamma, e[1] + x[1] / gamma, e[2] + x[2] / gamma };
    const float bde    = dot3(up, e) / gamma;
    float       kap[3];
    cross3(epb, b, kap);
    kap[0] += bde * e[0], kap[1] += bde * e[1], kap[2] += bde * e[2];
    const float chi = nsq(epb) - SQR(bde);
    const float p   = E.nominal_probability * (-dot3(kap, up) / (gamma_sqr * u_mag) + beta * chi);
    const float dir[3] = { -kap[0] + gamma * up[0] * chi, -kap[1] + gamma * up[1] * chi,
                           -kap[2] + gamma * up[2] * chi };
    const float du = -E.photon_weight * energy / (sqrtf(nsq(dir)) * E.species_mass);
    delta_u[0] = du * dir[0], delta_u[1] = du * dir[1], delta_u[2] = du * dir[2];
    return p;
  }

  struct NoEmission {};

  /* ------------------------------------------------------------- one full push */
  struct PushArgs {
    eb200_pusher_t c;
    float          ndh; // 1/2 (q/m) omegaB0 dt  (sr.hpp:111)
    int            ni[3];
    int            ng;
    float          inv_dx; // 1 / c.dx
    // exception list of this launch (launch.h ExcList), null when nobody asked for it
    uint32_t*      exc_count = nullptr;
    uint32_t*      exc_idx   = nullptr;
    uint32_t       exc_cap   = 0;
  };

  // particle p is not alive after this launch: dead on entry, absorbed, or tagged to migrate
  __device__ __forceinline__ void exc_append(const PushArgs& A, uint32_t p) {
    if (A.exc_count != nullptr) {
      const uint32_t s = atomicAdd(A.exc_count, 1u);
      if (s < A.exc_cap) A.exc_idx[s] = p;
    }
  }

  // LEAN: the context is known (host-side check, see lean_pusher()) to be a plain Boris push
  // without drag, atmosphere or GCA; the optional branches are compiled out.
  __host__ __device__ __forceinline__ bool lean_pusher(const eb200_pusher_t& c) {
    return c.pusher_flags == EB200_PUSHER_BORIS && c.drag_flags == EB200_DRAG_NONE &&
           !c.has_atmosphere;
  }

  // sr.hpp:659-751 (Cartesian): what happens to a particle that left [0, ni) along some axis --
  // periodic wrap (shifts i_prev too), absorption, reflection -- and the migration send tag
  template <int D>
#ifdef EB200_BC_NOINLINE
  __device__ __noinline__ void particle_boundaries(const PushArgs& A, Prtl<D>& P) {
#else
  __device__ __forceinline__ void particle_boundaries(const PushArgs& A, Prtl<D>& P) {
#endif
    const eb200_pusher_t& c = A.c;
    int lin = 0, centre = 0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const int ni  = A.ni[a];
      bool      inv = false;
      if (P.i[a] < 0) {
        const int b = c.pbc[2 * a];
        if (b == EB200_PBC_PERIODIC) {
          P.i[a]  += ni;
          P.ip[a] += ni;
        } else if (b == EB200_PBC_ABSORB) {
          P.tag = 0;
        } else if (b == EB200_PBC_REFLECT || b == EB200_PBC_AXIS) {
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
        } else if (b == EB200_PBC_REFLECT || b == EB200_PBC_AXIS) {
          P.i[a] = ni - 1;
          P.d[a] = ONE - P.d[a];
          inv    = (b == EB200_PBC_REFLECT);
        }
      }
      if (inv) {
        P.u[a] = -P.u[a];
      }
      const int dir = (P.i[a] < 0) ? 0 : ((P.i[a] >= ni) ? 2 : 1);
      lin           = lin * 3 + dir;
      centre        = centre * 3 + 1;
    }
    if (c.tag_outgoing && lin != centre) {
      // mpi::SendTag: 2 + lexicographic index of the direction, null direction skipped
      P.tag = static_cast<short>((2 + lin - (lin > centre ? 1 : 0)) * P.tag);
    }
  }

  // HOOK (emission): callable (P, u_mid, e_cart, b_cart) run between the velocity update and the
  // position update of a massive particle (sr.hpp:323-328); with a hook the continuous radiative
  // drag is not applied (sr.hpp:311-322)
  template <int D, int O, class EM, bool LEAN = false, class HOOK = NoEmission>
  // returns true when the particle left [0, ni) along some axis, i.e. when the boundary block
  // ran (only then can tag, i_prev or u have been touched by a boundary condition)
  __device__ __forceinline__ bool push_particle(const PushArgs& A, const EM& F, Prtl<D>& P,
                                                HOOK&& hook = HOOK {}) {
    const eb200_pusher_t& c  = A.c;
    const float           dt = c.dt;
    bool                  massive = true;
    if constexpr (LEAN) {
      float ec[3], bc[3];
      gather_any<D, O>(F, A.ng, P, ec, bc);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        ec[a] = ec[a] * c.dx;
        bc[a] = bc[a] * c.dx;
      }
      boris(A.ndh, P.u, ec, bc);
    } else if (c.pusher_flags == EB200_PUSHER_PHOTON) {
      massive = false;
    } else {
      float ec[3], bc[3];
      gather_any<D, O>(F, A.ng, P, ec, bc);
      // contravariant -> Cartesian: in-plane components scale with the cell size
#pragma unroll
      for (int a = 0; a < D; ++a) {
        ec[a] = ec[a] * c.dx;
        bc[a] = bc[a] * c.dx;
      }
      float fext[3] = { ZERO, ZERO, ZERO };
      if (c.has_atmosphere) {
        const float gg[3] = { c.atm_gx1, c.atm_gx2, c.atm_gx3 };
#pragma unroll
        for (int a = 0; a < D; ++a) {
          const float xph = (static_cast<float>(P.i[a]) + P.d[a]) * c.dx + c.xmin[a];
          const bool  on  = !(fabsf(gg[a]) <= 1.1920929e-07f) &&
                          ((c.atm_ds < ZERO || xph <= c.atm_x_surf + c.atm_ds) &&
                           (c.atm_ds > ZERO || xph >= c.atm_x_surf + c.atm_ds));
          if (on) {
            fext[a] += gg[a];
          }
        }
      }
      if constexpr (std::is_same<std::decay_t<HOOK>, NoEmission>::value) {
        velocity_update(c, A.ndh, P.u, ec, bc, fext);
      } else {
        float       um[3] = { P.u[0], P.u[1], P.u[2] };
        const float er[3] = { ec[0], ec[1], ec[2] }, br[3] = { bc[0], bc[1], bc[2] };
        eb200_pusher_t nodrag = c;
        nodrag.drag_flags     = EB200_DRAG_NONE;
        velocity_update(nodrag, A.ndh, P.u, ec, bc, fext);
        um[0] = HALF * (um[0] + P.u[0]);
        um[1] = HALF * (um[1] + P.u[1]);
        um[2] = HALF * (um[2] + P.u[2]);
        hook(P, um, er, br);
      }
    }
    // Cartesian i+dx position update
    const float g2 = massive ? (ONE + SQR(P.u[0]) + SQR(P.u[1]) + SQR(P.u[2]))
                             : (SQR(P.u[0]) + SQR(P.u[1]) + SQR(P.u[2]));
    const float dt_inv_energy = div_by_sqrt(dt, g2);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      P.ip[a]  = P.i[a];
      P.dp[a]  = P.d[a];
      float dx = P.d[a] + div_dx(P.u[a], c.dx, A.inv_dx) * dt_inv_energy;
      P.i[a]  += static_cast<int>(dx >= ONE) - static_cast<int>(dx < ZERO);
      dx      -= static_cast<float>(dx >= ONE);
      dx      += static_cast<float>(dx < ZERO);
      P.d[a]   = dx;
    }
    // particle boundaries, then the migration tag. Nearly every particle stays inside the
    // domain: one unsigned compare per axis skips the whole block for those.
    bool inside = true;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      inside = inside && (static_cast<unsigned>(P.i[a]) < static_cast<unsigned>(A.ni[a]));
    }
    if (inside) {
      return false;
    }
    particle_boundaries<D>(A, P);
    return true;
  }

  /* ------------------------------------------------------------------ deposit */
  // SINK: callable (i, j, k, comp, value) in ghost-inclusive indices. Calls are issued in the
  // reference's program order so that an order-preserving sink reproduces its sums exactly.
  // `vp_ext` (optional): the particle's coordinate velocity u^i / gamma as a curvilinear or GR
  // metric gives it (currents_deposit.hpp:113-163); null = Cartesian, computed here.
  template <int D, int O, class SINK>
  __device__ __forceinline__ void deposit_particle(const Prtl<D>& P, float charge, float inv_dt,
                                                   float dxc, int G, SINK&& J,
                                                   const float* vp_ext = nullptr) {
    float vp[3];
    if (vp_ext != nullptr) {
      vp[0] = vp_ext[0];
      vp[1] = vp_ext[1];
      vp[2] = vp_ext[2];
    } else {
      vp[0] = (0 < D) ? fdiv(P.u[0], dxc) : P.u[0];
      vp[1] = (1 < D) ? fdiv(P.u[1], dxc) : P.u[1];
      vp[2] = (2 < D) ? fdiv(P.u[2], dxc) : P.u[2];
      const float inv_energy = rcp_sqrt(ONE + nsq(P.u));
      if (isnan(vp[2]) || isinf(vp[2])) {
        vp[2] = ZERO;
      }
      vp[0] *= inv_energy;
      vp[1] *= inv_energy;
      vp[2] *= inv_energy;
    }
    const float coeff = P.w * charge;

    if constexpr (O == 0) {
      // zig-zag: the move is split at a relay point into one segment in the old cell
      // (index 0) and one in the new cell (index 1)
      float W[3][2], Fl[3][2];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const int   up = static_cast<int>(P.i[a] > P.ip[a]);
        const float r  = static_cast<float>(P.i[a] == P.ip[a]) * (P.d[a] + P.dp[a]) * INV_2;
        W[a][0]        = INV_2 * (r + P.dp[a] + static_cast<float>(up));
        W[a][1]        = INV_2 * (P.d[a] + r + static_cast<float>(up + P.ip[a] - P.i[a]));
        Fl[a][0]       = (static_cast<float>(up) + r - P.dp[a]) * coeff * inv_dt;
        Fl[a][1] = (static_cast<float>(P.i[a] - P.ip[a] - up) + P.d[a] - r) * coeff * inv_dt;
      }
      if constexpr (D == 1) {
        const int   a0 = P.ip[0] + G, a1 = P.i[0] + G;
        const float F2 = HALF * vp[1] * coeff, F3 = HALF * vp[2] * coeff;
        J(a0, 0, 0, jx1, Fl[0][0]);
        J(a1, 0, 0, jx1, Fl[0][1]);
        J(a0, 0, 0, jx2, F2 * (ONE - W[0][0]));
        J(a0 + 1, 0, 0, jx2, F2 * W[0][0]);
        J(a1, 0, 0, jx2, F2 * (ONE - W[0][1]));
        J(a1 + 1, 0, 0, jx2, F2 * W[0][1]);
        J(a0, 0, 0, jx3, F3 * (ONE - W[0][0]));
        J(a0 + 1, 0, 0, jx3, F3 * W[0][0]);
        J(a1, 0, 0, jx3, F3 * (ONE - W[0][1]));
        J(a1 + 1, 0, 0, jx3, F3 * W[0][1]);
      } else if constexpr (D == 2) {
        const float F3 = HALF * vp[2] * coeff;
        const int   ci[2] = { P.ip[0] + G, P.i[0] + G }, cj[2] = { P.ip[1] + G, P.i[1] + G };
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          J(ci[s], cj[s], 0, jx1, Fl[0][s] * (ONE - W[1][s]));
          J(ci[s], cj[s] + 1, 0, jx1, Fl[0][s] * W[1][s]);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          J(ci[s], cj[s], 0, jx2, Fl[1][s] * (ONE - W[0][s]));
          J(ci[s] + 1, cj[s], 0, jx2, Fl[1][s] * W[0][s]);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          J(ci[s], cj[s], 0, jx3, F3 * (ONE - W[0][s]) * (ONE - W[1][s]));
          J(ci[s] + 1, cj[s], 0, jx3, F3 * W[0][s] * (ONE - W[1][s]));
          J(ci[s], cj[s] + 1, 0, jx3, F3 * (ONE - W[0][s]) * W[1][s]);
          J(ci[s] + 1, cj[s] + 1, 0, jx3, F3 * W[0][s] * W[1][s]);
        }
      } else {
        const int ci[2] = { P.ip[0] + G, P.i[0] + G }, cj[2] = { P.ip[1] + G, P.i[1] + G },
                  ck[2] = { P.ip[2] + G, P.i[2] + G };
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          J(ci[s], cj[s], ck[s], jx1, Fl[0][s] * (ONE - W[1][s]) * (ONE - W[2][s]));
          J(ci[s], cj[s] + 1, ck[s], jx1, Fl[0][s] * W[1][s] * (ONE - W[2][s]));
          J(ci[s], cj[s], ck[s] + 1, jx1, Fl[0][s] * (ONE - W[1][s]) * W[2][s]);
          J(ci[s], cj[s] + 1, ck[s] + 1, jx1, Fl[0][s] * W[1][s] * W[2][s]);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          J(ci[s], cj[s], ck[s], jx2, Fl[1][s] * (ONE - W[0][s]) * (ONE - W[2][s]));
          J(ci[s] + 1, cj[s], ck[s], jx2, Fl[1][s] * W[0][s] * (ONE - W[2][s]));
          J(ci[s], cj[s], ck[s] + 1, jx2, Fl[1][s] * (ONE - W[0][s]) * W[2][s]);
          J(ci[s] + 1, cj[s], ck[s] + 1, jx2, Fl[1][s] * W[0][s] * W[2][s]);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          J(ci[s], cj[s], ck[s], jx3, Fl[2][s] * (ONE - W[0][s]) * (ONE - W[1][s]));
          J(ci[s] + 1, cj[s], ck[s], jx3, Fl[2][s] * W[0][s] * (ONE - W[1][s]));
          J(ci[s], cj[s] + 1, ck[s], jx3, Fl[2][s] * (ONE - W[0][s]) * W[1][s]);
          J(ci[s] + 1, cj[s] + 1, ck[s], jx3, Fl[2][s] * W[0][s] * W[1][s]);
        }
      }
    } else if constexpr (O > 3) {
      // orders 4..11: the decomposition below with rolled loops (windows of up to 13^3 nodes)
      constexpr int N = O + 2;
      float         iS[3][N], fS[3][N];
      int           mn[3] = { 0, 0, 0 }, dd[3] = { 0, 0, 0 };
#pragma unroll
      for (int a = 0; a < D; ++a) {
        int mx;
        deposit_shapes<O>(P.ip[a], P.dp[a], P.i[a], P.d[a], mn[a], mx, iS[a], fS[a]);
        mn[a] += G;
        dd[a] = mx + G - mn[a];
      }
      const float Q = coeff * inv_dt;
      if constexpr (D == 1) {
        const float QV2 = coeff * vp[1], QV3 = coeff * vp[2];
        float       acc = ZERO;
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
          acc = (i == 0) ? (-Q * (fS[0][0] - iS[0][0])) : (acc - Q * (fS[0][i] - iS[0][i]));
          J(mn[0] + i, 0, 0, jx1, acc, i < dd[0]);
        }
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
          J(mn[0] + i, 0, 0, jx2, QV2 * (HALF * (fS[0][i] + iS[0][i])), i <= dd[0]);
        }
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
          J(mn[0] + i, 0, 0, jx3, QV3 * (HALF * (fS[0][i] + iS[0][i])), i <= dd[0]);
        }
      } else if constexpr (D == 2) {
        const float QV3 = coeff * vp[2];
        {
          float acc[N];
#pragma unroll 1
          for (int i = 0; i < N; ++i) {
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
              const float w = HALF * (fS[0][i] - iS[0][i]) * (fS[1][j] + iS[1][j]);
              acc[j]        = (i == 0) ? (-Q * w) : (acc[j] - Q * w);
              J(mn[0] + i, mn[1] + j, 0, jx1, acc[j], i < dd[0] && j <= dd[1]);
            }
          }
        }
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
          float acc = ZERO;
#pragma unroll 1
          for (int j = 0; j < N; ++j) {
            const float w = HALF * (fS[0][i] + iS[0][i]) * (fS[1][j] - iS[1][j]);
            acc           = (j == 0) ? (-Q * w) : (acc - Q * w);
            J(mn[0] + i, mn[1] + j, 0, jx2, acc, i <= dd[0] && j < dd[1]);
          }
        }
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
#pragma unroll 1
          for (int j = 0; j < N; ++j) {
            const float w = THIRD * (fS[1][j] * (HALF * iS[0][i] + fS[0][i]) +
                                     iS[1][j] * (HALF * fS[0][i] + iS[0][i]));
            J(mn[0] + i, mn[1] + j, 0, jx3, QV3 * w, i <= dd[0] && j <= dd[1]);
          }
        }
      } else {
        {
          float acc[N][N];
#pragma unroll 1
          for (int i = 0; i < N; ++i) {
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
#pragma unroll 1
              for (int k = 0; k < N; ++k) {
                const float w = THIRD * (fS[0][i] - iS[0][i]) *
                                ((iS[1][j] * iS[2][k] + fS[1][j] * fS[2][k]) +
                                 HALF * (iS[2][k] * fS[1][j] + iS[1][j] * fS[2][k]));
                acc[j][k] = (i == 0) ? (-Q * w) : (acc[j][k] - Q * w);
                J(mn[0] + i, mn[1] + j, mn[2] + k, jx1, acc[j][k],
                  i < dd[0] && j <= dd[1] && k <= dd[2]);
              }
            }
          }
        }
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
          float acc[N];
#pragma unroll 1
          for (int j = 0; j < N; ++j) {
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
              const float w = THIRD * (fS[1][j] - iS[1][j]) *
                              (iS[0][i] * iS[2][k] + fS[0][i] * fS[2][k] +
                               HALF * (iS[2][k] * fS[0][i] + iS[0][i] * fS[2][k]));
              acc[k] = (j == 0) ? (-Q * w) : (acc[k] - Q * w);
              J(mn[0] + i, mn[1] + j, mn[2] + k, jx2, acc[k], i <= dd[0] && j < dd[1] && k <= dd[2]);
            }
          }
        }
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
#pragma unroll 1
          for (int j = 0; j < N; ++j) {
            float acc = ZERO;
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
              const float w = THIRD * (fS[2][k] - iS[2][k]) *
                              (iS[0][i] * iS[1][j] + fS[0][i] * fS[1][j] +
                               HALF * (iS[0][i] * fS[1][j] + iS[1][j] * fS[0][i]));
              acc = (k == 0) ? (-Q * w) : (acc - Q * w);
              J(mn[0] + i, mn[1] + j, mn[2] + k, jx3, acc, i <= dd[0] && j <= dd[1] && k < dd[2]);
            }
          }
        }
      }
    } else {
      // Esirkepov density decomposition on an (O+2)^D window
      constexpr int N = O + 2;
      float         iS1[N], fS1[N];
      int           min1, max1;
      deposit_shapes<O>(P.ip[0], P.dp[0], P.i[0], P.d[0], min1, max1, iS1, fS1);
      const float Q = coeff * inv_dt;
      if constexpr (D == 1) {
        const float QV2 = coeff * vp[1], QV3 = coeff * vp[2];
        min1 += G;
        max1 += G;
        const int d1 = max1 - min1;
        float     acc = ZERO;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          acc = (i == 0) ? (-Q * (fS1[0] - iS1[0])) : (acc - Q * (fS1[i] - iS1[i]));
          J(min1 + i, 0, 0, jx1, acc, i < d1);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          J(min1 + i, 0, 0, jx2, QV2 * (HALF * (fS1[i] + iS1[i])), i <= d1);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          J(min1 + i, 0, 0, jx3, QV3 * (HALF * (fS1[i] + iS1[i])), i <= d1);
        }
      } else if constexpr (D == 2) {
        float iS2[N], fS2[N];
        int   min2, max2;
        deposit_shapes<O>(P.ip[1], P.dp[1], P.i[1], P.d[1], min2, max2, iS2, fS2);
        const float QV3 = coeff * vp[2];
        min1 += G;
        min2 += G;
        max1 += G;
        max2 += G;
        const int d1 = max1 - min1, d2 = max2 - min2;
        // jx1: prefix sums along x1 for every x2 row of the window
        {
          float acc[N];
#pragma unroll
          for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
              const float w = HALF * (fS1[i] - iS1[i]) * (fS2[j] + iS2[j]);
              acc[j]        = (i == 0) ? (-Q * w) : (acc[j] - Q * w);
              J(min1 + i, min2 + j, 0, jx1, acc[j], i < d1 && j <= d2);
            }
          }
        }
        // jx2: prefix sums along x2
        {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            float acc = ZERO;
#pragma unroll
            for (int j = 0; j < N; ++j) {
              const float w = HALF * (fS1[i] + iS1[i]) * (fS2[j] - iS2[j]);
              acc           = (j == 0) ? (-Q * w) : (acc - Q * w);
              J(min1 + i, min2 + j, 0, jx2, acc, i <= d1 && j < d2);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const float w = THIRD * (fS2[j] * (HALF * iS1[i] + fS1[i]) +
                                     iS2[j] * (HALF * fS1[i] + iS1[i]));
            J(min1 + i, min2 + j, 0, jx3, QV3 * w, i <= d1 && j <= d2);
          }
        }
      } else {
        float iS2[N], fS2[N], iS3[N], fS3[N];
        int   min2, max2, min3, max3;
        deposit_shapes<O>(P.ip[1], P.dp[1], P.i[1], P.d[1], min2, max2, iS2, fS2);
        deposit_shapes<O>(P.ip[2], P.dp[2], P.i[2], P.d[2], min3, max3, iS3, fS3);
        min1 += G;
        min2 += G;
        min3 += G;
        max1 += G;
        max2 += G;
        max3 += G;
        const int d1 = max1 - min1, d2 = max2 - min2, d3 = max3 - min3;
        // jx1: running sum over i for each (j,k)
        {
          float acc[N][N];
#pragma unroll
          for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
#pragma unroll
              for (int k = 0; k < N; ++k) {
                const float w = THIRD * (fS1[i] - iS1[i]) *
                                ((iS2[j] * iS3[k] + fS2[j] * fS3[k]) +
                                 HALF * (iS3[k] * fS2[j] + iS2[j] * fS3[k]));
                acc[j][k] = (i == 0) ? (-Q * w) : (acc[j][k] - Q * w);
                J(min1 + i, min2 + j, min3 + k, jx1, acc[j][k], i < d1 && j <= d2 && k <= d3);
              }
            }
          }
        }
        // jx2: running sum over j for each (i,k)
        {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            float acc[N];
#pragma unroll
            for (int j = 0; j < N; ++j) {
#pragma unroll
              for (int k = 0; k < N; ++k) {
                const float w = THIRD * (fS2[j] - iS2[j]) *
                                (iS1[i] * iS3[k] + fS1[i] * fS3[k] +
                                 HALF * (iS3[k] * fS1[i] + iS1[i] * fS3[k]));
                acc[k] = (j == 0) ? (-Q * w) : (acc[k] - Q * w);
                J(min1 + i, min2 + j, min3 + k, jx2, acc[k], i <= d1 && j < d2 && k <= d3);
              }
            }
          }
        }
        // jx3: running sum over k for each (i,j)
        {
#pragma unroll
          for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
              float acc = ZERO;
#pragma unroll
              for (int k = 0; k < N; ++k) {
                const float w = THIRD * (fS3[k] - iS3[k]) *
                                (iS1[i] * iS2[j] + fS1[i] * fS2[j] +
                                 HALF * (iS1[i] * fS2[j] + iS2[j] * fS1[i]));
                acc = (k == 0) ? (-Q * w) : (acc - Q * w);
                J(min1 + i, min2 + j, min3 + k, jx3, acc, i <= d1 && j <= d2 && k < d3);
              }
            }
          }
        }
      }
    }
  }

  /* ----------------------- 3D third-order Esirkepov deposit, rows with paired reductions */
  // The per-lane atomic deposit of a 5 x 5 x 5 window is bound by the number of reduction
  // requests that reach L2, not by arithmetic. The five nodes of a window row along x1 are
  // contiguous in memory: they go out as two 64-bit reductions (red.global.add.v2.f32, which
  // needs an 8-byte aligned address: the pairing shifts by one element when the row starts on
  // an odd element) and one scalar, branch free. Same window, guards and per-node weights as
  // deposit_particle<3, 3>(), evaluated row by row (the running sums of jx2 / jx3 are kept per
  // (i, k) / (i, j) while x2 / x3 advance), so the additions into a J element differ from the
  // node-by-node form only in their order. Fast build only: the strict build keeps the
  // reference's program order.
  __device__ __forceinline__ void red_pair(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
  }

  // v[0..4] -> J elements e .. e + 4 (zeros are not sent)
  __device__ __forceinline__ void red_row5(float* e, const float (&v)[5]) {
    const bool odd = (reinterpret_cast<uintptr_t>(e) >> 2) & 1u;
    float*     a   = e + (odd ? 1 : 0);
    const float p0 = odd ? v[1] : v[0], p1 = odd ? v[2] : v[1];
    const float p2 = odd ? v[3] : v[2], p3 = odd ? v[4] : v[3];
    const float s  = odd ? v[0] : v[4];
    if (p0 != ZERO || p1 != ZERO) red_pair(a, p0, p1);
    if (p2 != ZERO || p3 != ZERO) red_pair(a + 2, p2, p3);
    if (s != ZERO) atomicAdd(odd ? e : e + 4, s);
  }

  __device__ __forceinline__ void deposit_esirkepov3_rows(const Prtl<3>& P, float charge,
                                                          float inv_dt, float dxc, int G,
                                                          const FieldView<3>& J) {
    constexpr int N = 5;
    (void)dxc; // 3D: every component comes from the density decomposition, no velocity term
    const float coeff = P.w * charge;
    const float Q     = coeff * inv_dt;
    float       iS1[N], fS1[N], iS2[N], fS2[N], iS3[N], fS3[N];
    int         min1, max1, min2, max2, min3, max3;
    deposit_shapes<3>(P.ip[0], P.dp[0], P.i[0], P.d[0], min1, max1, iS1, fS1);
    deposit_shapes<3>(P.ip[1], P.dp[1], P.i[1], P.d[1], min2, max2, iS2, fS2);
    deposit_shapes<3>(P.ip[2], P.dp[2], P.i[2], P.d[2], min3, max3, iS3, fS3);
    min1 += G, min2 += G, min3 += G;
    max1 += G, max2 += G, max3 += G;
    const int  d1 = max1 - min1, d2 = max2 - min2, d3 = max3 - min3;
    const long N12 = (long)J.N1 * J.N2;
    float*     base = J.p + J.idx(min1, min2, min3);
    // jx1(i, j, k) = -Q sum_{i' <= i} THIRD DS1(i') F23(j, k): one row per (j, k)
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (j > d2 || k > d3) continue;
        const float f = (iS2[j] * iS3[k] + fS2[j] * fS3[k]) +
                        HALF * (iS3[k] * fS2[j] + iS2[j] * fS3[k]);
        float v[N];
        float acc = ZERO;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          acc -= Q * (THIRD * (fS1[i] - iS1[i]) * f);
          v[i] = (i < d1) ? acc : ZERO;
        }
        red_row5(base + (long)j * J.N1 + (long)k * N12, v);
      }
    }
    // jx2(i, j, k) = -Q sum_{j' <= j} THIRD DS2(j') F13(i, k): running sums per (i, k) while
    // x2 advances, one row per (j, k)
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (k > d3) continue;
      float g[N], acc[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        g[i]   = iS1[i] * iS3[k] + fS1[i] * fS3[k] + HALF * (iS3[k] * fS1[i] + iS1[i] * fS3[k]);
        acc[i] = ZERO;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float ds = THIRD * (fS2[j] - iS2[j]);
        float       v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          acc[i] -= Q * (ds * g[i]);
          v[i] = (i <= d1) ? acc[i] : ZERO;
        }
        if (j < d2) red_row5(base + J.plane + (long)j * J.N1 + (long)k * N12, v);
      }
    }
    // jx3(i, j, k) = -Q sum_{k' <= k} THIRD DS3(k') F12(i, j): running sums per (i, j) while
    // x3 advances
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j > d2) continue;
      float g[N], acc[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        g[i]   = iS1[i] * iS2[j] + fS1[i] * fS2[j] + HALF * (iS1[i] * fS2[j] + iS2[j] * fS1[i]);
        acc[i] = ZERO;
      }
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const float ds = THIRD * (fS3[k] - iS3[k]);
        float       v[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          acc[i] -= Q * (ds * g[i]);
          v[i] = (i <= d1) ? acc[i] : ZERO;
        }
        if (k < d3) red_row5(base + 2 * J.plane + (long)j * J.N1 + (long)k * N12, v);
      }
    }
  }

  /* ------------------------------------------------- warp-aggregated deposit */
  // Particles arrive (nearly) cell-sorted, so consecutive lanes of a warp mostly deposit
  // onto the same nodes. Instead of one atomic per lane and node, runs of lanes with the
  // same key (cell / window origin) are summed with a segmented shuffle reduction and only
  // the head lane of each run issues the atomic. Correct for any order of the particles:
  // an unsorted warp just degrades to one run per lane.
  struct WarpRun {
    bool  head;
    float m[5]; // m[l] = 1 if lane + 2^l is still inside this lane's run, else 0
  };

  __device__ __forceinline__ WarpRun warp_runs(int key) {
    const unsigned lane  = threadIdx.x & 31u;
    const int      prev  = __shfl_up_sync(0xffffffffu, key, 1);
    const bool     head  = (lane == 0u) || (prev != key);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned above = (lane == 31u) ? 0u : (heads & ~((2u << lane) - 1u));
    WarpRun        r;
    r.head        = head;
    const int end = above ? (__ffs(above) - 2) : 31; // last lane of this lane's run
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      r.m[l] = (static_cast<int>(lane) + (1 << l) <= end) ? ONE : ZERO;
    }
    return r;
  }

  // sum of v over the run; valid on the head lane
  __device__ __forceinline__ float run_sum(float v, const WarpRun& r) {
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      // one FFMA per level; o * 1 and o * 0 are exact for finite o
      v = fmaf(__shfl_down_sync(0xffffffffu, v, 1 << l), r.m[l], v);
    }
    return v;
  }

  template <int D>
  struct ZigZag {
    static constexpr int NV = (D == 1) ? 5 : ((D == 2) ? 8 : 12);
  };

  // Contributions of the two zig-zag segments (0: in the old cell, 1: in the new cell) to
  // the NV nodes around the respective cell, in the node order of zigzag_offsets().
  template <int D>
  __device__ __forceinline__ void zigzag_values(const Prtl<D>& P, float charge, float inv_dt,
                                                float dxc, float (&v)[2][ZigZag<D>::NV],
                                                const float* vp_ext = nullptr) {
    float vp[3];
    if (vp_ext != nullptr) {
      vp[0] = vp_ext[0];
      vp[1] = vp_ext[1];
      vp[2] = vp_ext[2];
    } else {
      vp[0] = (0 < D) ? fdiv(P.u[0], dxc) : P.u[0];
      vp[1] = (1 < D) ? fdiv(P.u[1], dxc) : P.u[1];
      vp[2] = (2 < D) ? fdiv(P.u[2], dxc) : P.u[2];
      const float inv_energy = rcp_sqrt(ONE + nsq(P.u));
      if (isnan(vp[2]) || isinf(vp[2])) {
        vp[2] = ZERO;
      }
      vp[0] *= inv_energy;
      vp[1] *= inv_energy;
      vp[2] *= inv_energy;
    }
    const float coeff = P.w * charge;
    float       W[3][2], Fl[3][2];
#pragma unroll
    for (int a = 0; a < D; ++a) {
#if EB200_STRICT
      const int   up = static_cast<int>(P.i[a] > P.ip[a]);
      const float r  = static_cast<float>(P.i[a] == P.ip[a]) * (P.d[a] + P.dp[a]) * INV_2;
      W[a][0]        = INV_2 * (r + P.dp[a] + static_cast<float>(up));
      W[a][1]        = INV_2 * (P.d[a] + r + static_cast<float>(up + P.ip[a] - P.i[a]));
      Fl[a][0]       = (static_cast<float>(up) + r - P.dp[a]) * coeff * inv_dt;
      Fl[a][1] = (static_cast<float>(P.i[a] - P.ip[a] - up) + P.d[a] - r) * coeff * inv_dt;
#else
      // same relay point, written with selects: r0 / r1 = relay coordinate seen from the old /
      // new cell (midpoint when the particle stays, the shared face when it crosses)
      const int   di = P.i[a] - P.ip[a];
      const float r0 = (di == 0) ? INV_2 * (P.d[a] + P.dp[a]) : ((di > 0) ? ONE : ZERO);
      const float r1 = (di == 0) ? r0 : ((di > 0) ? ZERO : ONE);
      const float Q  = coeff * inv_dt;
      W[a][0]        = INV_2 * (r0 + P.dp[a]);
      W[a][1]        = INV_2 * (P.d[a] + r1);
      Fl[a][0]       = (r0 - P.dp[a]) * Q;
      Fl[a][1]       = (P.d[a] - r1) * Q;
#endif
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if constexpr (D == 1) {
        const float F2 = HALF * vp[1] * coeff, F3 = HALF * vp[2] * coeff;
        v[s][0] = Fl[0][s];
        v[s][1] = F2 * (ONE - W[0][s]);
        v[s][2] = F2 * W[0][s];
        v[s][3] = F3 * (ONE - W[0][s]);
        v[s][4] = F3 * W[0][s];
      } else if constexpr (D == 2) {
        const float F3 = HALF * vp[2] * coeff;
        v[s][0] = Fl[0][s] * (ONE - W[1][s]);
        v[s][1] = Fl[0][s] * W[1][s];
        v[s][2] = Fl[1][s] * (ONE - W[0][s]);
        v[s][3] = Fl[1][s] * W[0][s];
        v[s][4] = F3 * (ONE - W[0][s]) * (ONE - W[1][s]);
        v[s][5] = F3 * W[0][s] * (ONE - W[1][s]);
        v[s][6] = F3 * (ONE - W[0][s]) * W[1][s];
        v[s][7] = F3 * W[0][s] * W[1][s];
      } else {
        v[s][0]  = Fl[0][s] * (ONE - W[1][s]) * (ONE - W[2][s]);
        v[s][1]  = Fl[0][s] * W[1][s] * (ONE - W[2][s]);
        v[s][2]  = Fl[0][s] * (ONE - W[1][s]) * W[2][s];
        v[s][3]  = Fl[0][s] * W[1][s] * W[2][s];
        v[s][4]  = Fl[1][s] * (ONE - W[0][s]) * (ONE - W[2][s]);
        v[s][5]  = Fl[1][s] * W[0][s] * (ONE - W[2][s]);
        v[s][6]  = Fl[1][s] * (ONE - W[0][s]) * W[2][s];
        v[s][7]  = Fl[1][s] * W[0][s] * W[2][s];
        v[s][8]  = Fl[2][s] * (ONE - W[0][s]) * (ONE - W[1][s]);
        v[s][9]  = Fl[2][s] * W[0][s] * (ONE - W[1][s]);
        v[s][10] = Fl[2][s] * (ONE - W[0][s]) * W[1][s];
        v[s][11] = Fl[2][s] * W[0][s] * W[1][s];
      }
    }
  }

  // Zig-zag contributions of a particle that stays inside its cell, both segments summed.
  // Algebraically identical to zigzag_values() (the relay point is the midpoint, so both
  // half-segments carry half of the displacement and the node weights are evaluated at the
  // quarter points); the closed form needs ~1/4 of the operations. Fast build only: the
  // rounding differs from the reference's operation order at the 1e-7 level.
  template <int D>
  __device__ __forceinline__ void zigzag_values_incell(const Prtl<D>& P, float charge,
                                                       float inv_dt, float dxc,
                                                       float (&v)[ZigZag<D>::NV]) {
    const float inv_energy = rcp_sqrt(ONE + nsq(P.u));
    const float coeff      = P.w * charge;
    const float Q          = coeff * inv_dt;
    float       m[3], dl[3];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      m[a]  = HALF * (P.d[a] + P.dp[a]);
      dl[a] = P.d[a] - P.dp[a];
    }
    if constexpr (D == 1) {
      float v2 = P.u[1] * inv_energy, v3 = P.u[2] * inv_energy;
      if (isnan(v3) || isinf(v3)) v3 = ZERO;
      const float F2 = v2 * coeff, F3 = v3 * coeff;
      v[0] = Q * dl[0];
      v[1] = F2 * (ONE - m[0]);
      v[2] = F2 * m[0];
      v[3] = F3 * (ONE - m[0]);
      v[4] = F3 * m[0];
    } else if constexpr (D == 2) {
      float v3 = P.u[2] * inv_energy;
      if (isnan(v3) || isinf(v3)) v3 = ZERO;
      const float A = Q * dl[0], B = Q * dl[1], F = v3 * coeff;
      v[0]          = A * (ONE - m[1]);
      v[1]          = A * m[1];
      v[2]          = B * (ONE - m[0]);
      v[3]          = B * m[0];
      const float c = INV_16 * dl[0] * dl[1]; // (dx/4)(dy/4): the two quarter-point products
      v[4]          = F * ((ONE - m[0]) * (ONE - m[1]) + c);
      v[5]          = F * (m[0] * (ONE - m[1]) - c);
      v[6]          = F * ((ONE - m[0]) * m[1] - c);
      v[7]          = F * (m[0] * m[1] + c);
    } else {
      const float A = Q * dl[0], B = Q * dl[1], C = Q * dl[2];
      const float c12 = INV_16 * dl[0] * dl[1], c13 = INV_16 * dl[0] * dl[2],
                  c23 = INV_16 * dl[1] * dl[2];
      v[0]  = A * ((ONE - m[1]) * (ONE - m[2]) + c23);
      v[1]  = A * (m[1] * (ONE - m[2]) - c23);
      v[2]  = A * ((ONE - m[1]) * m[2] - c23);
      v[3]  = A * (m[1] * m[2] + c23);
      v[4]  = B * ((ONE - m[0]) * (ONE - m[2]) + c13);
      v[5]  = B * (m[0] * (ONE - m[2]) - c13);
      v[6]  = B * ((ONE - m[0]) * m[2] - c13);
      v[7]  = B * (m[0] * m[2] + c13);
      v[8]  = C * ((ONE - m[0]) * (ONE - m[1]) + c12);
      v[9]  = C * (m[0] * (ONE - m[1]) - c12);
      v[10] = C * ((ONE - m[0]) * m[1] - c12);
      v[11] = C * (m[0] * m[1] + c12);
    }
    (void)dxc;
  }

  // element offset of node n of zigzag_values() relative to the cell's own node
  template <int D>
  __device__ __forceinline__ long zigzag_offset(int n, long N1, long N12, long plane) {
    if constexpr (D == 1) {
      const int  di[5] = { 0, 0, 1, 0, 1 };
      const int  cc[5] = { jx1, jx2, jx2, jx3, jx3 };
      return di[n] + plane * cc[n];
    } else if constexpr (D == 2) {
      const int di[8] = { 0, 0, 0, 1, 0, 1, 0, 1 };
      const int dj[8] = { 0, 1, 0, 0, 0, 0, 1, 1 };
      const int cc[8] = { jx1, jx1, jx2, jx2, jx3, jx3, jx3, jx3 };
      return di[n] + N1 * dj[n] + plane * cc[n];
    } else {
      const int di[12] = { 0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 0, 1 };
      const int dj[12] = { 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 1 };
      const int dk[12] = { 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 0, 0 };
      const int cc[12] = { jx1, jx1, jx1, jx1, jx2, jx2, jx2, jx2, jx3, jx3, jx3, jx3 };
      return di[n] + N1 * dj[n] + N12 * dk[n] + plane * cc[n];
    }
  }

  // Must be called by all 32 lanes of the warp; `active` = this lane has a particle to deposit.
  template <int D, int O>
  __device__ __forceinline__ void deposit_particle_aggregated(const Prtl<D>& P, bool active,
                                                              float charge, float inv_dt,
                                                              float dxc, int G,
                                                              const FieldView<D>& J,
                                                              const float* vp_ext = nullptr) {
    if constexpr (O == 0) {
      constexpr int NV = ZigZag<D>::NV;
      float         v[2][NV];
      int           key0 = -1, key1 = -1;
      bool          cross = false;
      if (active) {
        zigzag_values<D>(P, charge, inv_dt, dxc, v, vp_ext);
        key0 = (int)J.idx(P.ip[0] + G, (D > 1) ? P.ip[1] + G : 0, (D > 2) ? P.ip[2] + G : 0);
        key1 = (int)J.idx(P.i[0] + G, (D > 1) ? P.i[1] + G : 0, (D > 2) ? P.i[2] + G : 0);
        cross = key0 != key1;
        if (!cross) {
#pragma unroll
          for (int n = 0; n < NV; ++n) v[0][n] += v[1][n];
        }
      } else {
#pragma unroll
        for (int n = 0; n < NV; ++n) v[0][n] = ZERO;
      }
      const WarpRun run = warp_runs(key0);
      const long    N12 = (long)J.N1 * J.N2;
#pragma unroll
      for (int n = 0; n < NV; ++n) {
#ifdef EB200_X_NOREDUCE
        const float s = v[0][n];
#else
        const float s = run_sum(v[0][n], run);
#endif
#ifdef EB200_X_NORED
        if (run.head && key0 >= 0 && s == 1.2345e-30f) {
#else
        if (run.head && key0 >= 0) {
#endif
          atomicAdd(J.p + key0 + zigzag_offset<D>(n, J.N1, N12, J.plane), s);
        }
      }
      if (cross) {
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          atomicAdd(J.p + key1 + zigzag_offset<D>(n, J.N1, N12, J.plane), v[1][n]);
        }
      }
    } else {
      // Esirkepov: key = origin of the (O+2)^D window; every lane walks the full window with
      // zero for the nodes it does not touch, so all lanes issue the same shuffles
      Prtl<D> Q = P;
      if (!active) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          Q.i[a] = Q.ip[a] = 0;
          Q.d[a] = Q.dp[a] = ZERO;
          Q.u[a]           = ZERO;
        }
        Q.w = ZERO;
      }
      bool    first = true;
      WarpRun run;
      deposit_particle<D, O>(Q, charge, inv_dt, dxc, G,
                             [&](int i, int j, int k, int c, float val, bool guard = true) {
                               if (first) {
                                 run   = warp_runs(active ? (int)J.idx(i, j, k) : -1);
                                 first = false;
                               }
                               const float s = run_sum(guard ? val : ZERO, run);
                               if (run.head && active && s != ZERO) {
                                 atomicAdd(&J.at(i, j, k, c), s);
                               }
                             },
                             vp_ext);
    }
  }

} // namespace eb200
