// entity_b200 -- output staging: the device side of the reference's
// writers, kernel::FieldsToPhys_kernel (src/kernels/fields_to_phys.hpp:33-239) and
// kernel::PrtlToPhys_kernel (src/kernels/prtls_to_phys.hpp:30-218). The on-disk format (ADIOS2)
// stays the host's; these kernels produce what it writes: fields interpolated to cell centres
// and converted to the orthonormal / physical basis, and strided particle samples in physical
// coordinates. Compiled with --fmad=false (the reference's operation order).
#include "common.cuh"
#include "launch.h"
#include "metrics.cuh"

namespace eb200 {
  namespace {
    // per-metric pieces of metric.transform<in, out>() used by the writers
    //   hat:  U -> T at (x1, x2);  pu: U -> PU;  pd: D -> PD   (minkowski.h:263-299,
    //   spherical.h / qspherical.h / kerr_schild*.h / qkerr_schild.h `transform`)
    struct MinkOut { // dim = simulated dimensions: components beyond it pass through
      float dx;
      int   dim;
      __device__ void hat(float, float, const float* v, float* o) const {
        for (int c = 0; c < 3; ++c) o[c] = (c < dim) ? v[c] * dx : v[c];
      }
      __device__ void pu(float, float, const float* v, float* o) const {
        for (int c = 0; c < 3; ++c) o[c] = (c < dim) ? v[c] * dx : v[c];
      }
      __device__ void pd(float, float, const float* v, float* o) const {
        const float dx_inv = ONE / dx;
        for (int c = 0; c < 3; ++c) o[c] = (c < dim) ? v[c] * dx_inv : v[c];
      }
    };

    template <class M>
    struct SphOut { // Spherical / QSpherical: diagonal
      MetricParams m;
      __device__ void hat(float x1, float x2, const float* v, float* o) const {
        o[0] = v[0] * M::sqrt_h11(m, x1, x2);
        o[1] = v[1] * M::sqrt_h22(m, x1, x2);
        o[2] = v[2] * M::sqrt_h33(m, x1, x2);
      }
      __device__ void pu(float x1, float x2, const float* v, float* o) const {
        if constexpr (M::kind == EB200_METRIC_SPHERICAL) {
          o[0] = v[0] * m.d1;
          o[1] = v[1] * m.d2;
        } else {
          o[0] = v[0] * expf(x1 * m.d1 + m.chi_min) * m.d1;
          o[1] = v[1] * (qs_dtheta_deta(m.h, x2 * m.d2 + m.eta_min) * m.d2);
        }
        o[2] = v[2];
      }
      __device__ void pd(float x1, float x2, const float* v, float* o) const {
        if constexpr (M::kind == EB200_METRIC_SPHERICAL) {
          o[0] = v[0] * m.d1_inv;
          o[1] = v[1] * m.d2_inv;
        } else {
          o[0] = v[0] * m.d1_inv / expf(x1 * m.d1 + m.chi_min);
          o[1] = v[1] * m.d2_inv / qs_dtheta_deta(m.h, x2 * m.d2 + m.eta_min);
        }
        o[2] = v[2];
      }
    };

    template <class M>
    struct GrOut { // Kerr-Schild family
      MetricParams m;
      __device__ void hat(float x1, float x2, const float* v, float* o) const {
        if constexpr (M::kind == EB200_METRIC_KERR_SCHILD_0) {
          // kerr_schild_0.h:457-463: no h_13 term (it would be 0 / 0 on the axis)
          o[0] = v[0] / sqrtf(M::h11(m, x1, x2));
          o[1] = v[1] * sqrtf(M::h_22(m, x1, x2));
          o[2] = v[2] * sqrtf(M::h_33(m, x1, x2));
        } else {
          gr_cntrv_to_tetrad<M>(m, x1, x2, v, o);
        }
      }
      __device__ void pu(float x1, float x2, const float* v, float* o) const {
        if constexpr (M::kind == EB200_METRIC_QKERR_SCHILD) {
          o[0] = v[0] * expf(x1 * m.d1 + m.chi_min) * m.d1;
          o[1] = v[1] * (q_dtheta_deta(m.h, x2 * m.d2 + m.eta_min) * m.d2);
        } else {
          o[0] = v[0] * m.d1;
          o[1] = v[1] * m.d2;
        }
        o[2] = v[2];
      }
      __device__ void pd(float x1, float x2, const float* v, float* o) const {
        if constexpr (M::kind == EB200_METRIC_QKERR_SCHILD) {
          o[0] = v[0] * m.d1_inv / expf(x1 * m.d1 + m.chi_min);
          o[1] = v[1] * m.d2_inv / q_dtheta_deta(m.h, x2 * m.d2 + m.eta_min);
        } else {
          o[0] = v[0] * m.d1_inv;
          o[1] = v[1] * m.d2_inv;
        }
        o[2] = v[2];
      }
    };

    struct F2PArgs {
      int  n[3], G, dim;
      long N1, N12, plane_from, plane_to;
      int  cf[3], ct[3];
      int  interp; // 0 none, 1 from edges, 2 from faces
      int  conv;   // 0 none, 1 hat, 2 phys cntrv, 3 phys cov
    };

    // FieldsToPhys_kernel::operator() over Mesh::rangeActiveCells, 1D / 2D / 3D
    template <class OUT>
    __global__ void __launch_bounds__(256)
      fields_to_phys_kernel(const __grid_constant__ F2PArgs A, const OUT mo,
                            const float* __restrict__ from, float* __restrict__ to) {
      const long t     = (long)blockIdx.x * blockDim.x + threadIdx.x;
      const long total = (long)A.n[0] * A.n[1] * A.n[2];
      if (t >= total) return;
      const int  i1 = (int)(t % A.n[0]) + A.G;
      const int  i2 = (A.dim > 1) ? (int)((t / A.n[0]) % A.n[1]) + A.G : 0;
      const int  i3 = (A.dim > 2) ? (int)(t / ((long)A.n[0] * A.n[1])) + A.G : 0;
      const long s1 = 1, s2 = (A.dim > 1) ? A.N1 : 0, s3 = (A.dim > 2) ? A.N12 : 0;
      const long n  = i1 + (long)i2 * A.N1 + (long)i3 * A.N12;
      auto F = [&](int c, long off) { return from[(long)A.cf[c] * A.plane_from + n + off]; };
      float f[3], o[3];
      if (A.interp == 1) { // edges: component c lives on the edge along c
        if (A.dim == 1) {
          f[0] = F(0, 0);
          f[1] = INV_2 * (F(1, 0) + F(1, s1));
          f[2] = INV_2 * (F(2, 0) + F(2, s1));
        } else if (A.dim == 2) {
          f[0] = INV_2 * (F(0, 0) + F(0, s2));
          f[1] = INV_2 * (F(1, 0) + F(1, s1));
          f[2] = INV_4 * (F(2, 0) + F(2, s1) + F(2, s2) + F(2, s1 + s2));
        } else {
          f[0] = INV_4 * (F(0, 0) + F(0, s2) + F(0, s3) + F(0, s2 + s3));
          f[1] = INV_4 * (F(1, 0) + F(1, s1) + F(1, s3) + F(1, s1 + s3));
          f[2] = INV_4 * (F(2, 0) + F(2, s1) + F(2, s2) + F(2, s1 + s2));
        }
      } else if (A.interp == 2) { // faces: component c lives on the face normal to c
        f[0] = INV_2 * (F(0, 0) + F(0, s1));
        f[1] = (A.dim > 1) ? INV_2 * (F(1, 0) + F(1, s2)) : F(1, 0);
        f[2] = (A.dim > 2) ? INV_2 * (F(2, 0) + F(2, s3)) : F(2, 0);
      } else {
        f[0] = F(0, 0), f[1] = F(1, 0), f[2] = F(2, 0);
      }
      const float h  = A.interp ? HALF : ZERO;
      const float x1 = static_cast<float>(i1 - A.G) + h, x2 = static_cast<float>(i2 - A.G) + h;
      if (A.conv == 1) {
        mo.hat(x1, x2, f, o);
      } else if (A.conv == 2) {
        mo.pu(x1, x2, f, o);
      } else if (A.conv == 3) {
        mo.pd(x1, x2, f, o);
      } else {
        o[0] = f[0], o[1] = f[1], o[2] = f[2];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) to[(long)A.ct[c] * A.plane_to + n] = o[c];
    }

    struct P2PArgs {
      uint32_t stride, nout;
      int      dim;
      float    dx, xmin[3];
    };

    // PrtlToPhys_kernel<S, M, false>::operator()(p): sample p * stride -> buffers[p]
    // KIND: 0 Minkowski, 1 curvilinear SR (u Cartesian -> tetrad), 2 GR (u_i -> physical cov.)
    template <int KIND, class M>
    __global__ void __launch_bounds__(256)
      prtls_to_phys_kernel(const __grid_constant__ P2PArgs A, const MetricParams mp,
                           eb200_prtls_t S, float* x1, float* x2, float* x3, float* u1, float* u2,
                           float* u3, float* w) {
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= A.nout) return;
      const size_t p  = (size_t)q * A.stride;
      const float  c1 = static_cast<float>(S.i1[p]) + S.dx1[p];
      const float  c2 = (A.dim > 1) ? static_cast<float>(S.i2[p]) + S.dx2[p] : ZERO;
      const float  v[3] = { S.ux1[p], S.ux2[p], S.ux3[p] };
      float        o[3];
      if constexpr (KIND == 0) {
        x1[q] = c1 * A.dx + A.xmin[0];
        if (A.dim > 1) x2[q] = c2 * A.dx + A.xmin[1];
        if (A.dim > 2) x3[q] = (static_cast<float>(S.i3[p]) + S.dx3[p]) * A.dx + A.xmin[2];
        o[0] = v[0], o[1] = v[1], o[2] = v[2]; // XYZ == tetrad
      } else {
        x1[q] = M::r(mp, c1);
        x2[q] = M::theta(mp, c2);
        x3[q] = S.phi[p];
        if constexpr (KIND == 1) {
          const float xx[3] = { c1, c2, S.phi[p] };
          const Trig  t     = trig_at<M>(mp, xx);
          xyz_to_tetrad(t, v, o);
        } else {
          GrOut<M> g { mp };
          g.pd(c1, c2, v, o);
        }
      }
      u1[q] = o[0], u2[q] = o[1], u3[q] = o[2];
      w[q] = S.weight[p];
    }
  } // namespace

  cudaError_t fields_to_phys(const MetricParams* mp, const eb200_grid_t& g, float dx,
                             const float* from, int ncomp_from, float* to, int ncomp_to,
                             const int* cf, const int* ct, int interp, int conv, cudaStream_t st) {
    (void)ncomp_from, (void)ncomp_to;
    F2PArgs A;
    long    plane = 1;
    for (int a = 0; a < 3; ++a) {
      A.n[a] = (a < g.dim) ? g.n[a] : 1;
      if (a < g.dim) plane *= (g.n[a] + 2 * g.ng);
    }
    A.G = g.ng, A.dim = g.dim;
    A.N1 = g.n[0] + 2 * g.ng;
    A.N12 = (g.dim > 1) ? A.N1 * (g.n[1] + 2 * g.ng) : 0;
    A.plane_from = A.plane_to = plane;
    for (int c = 0; c < 3; ++c) A.cf[c] = cf[c], A.ct[c] = ct[c];
    A.interp = interp, A.conv = conv;
    const long     total = (long)A.n[0] * A.n[1] * A.n[2];
    const unsigned nb    = (unsigned)((total + 255) / 256);
    const int      kind  = mp ? mp->kind : EB200_METRIC_MINKOWSKI;
    switch (kind) {
      case EB200_METRIC_MINKOWSKI:
        fields_to_phys_kernel<<<nb, 256, 0, st>>>(A, MinkOut { dx, g.dim }, from, to);
        break;
      case EB200_METRIC_SPHERICAL:
        fields_to_phys_kernel<<<nb, 256, 0, st>>>(A, SphOut<Spherical> { *mp }, from, to);
        break;
      case EB200_METRIC_QSPHERICAL:
        fields_to_phys_kernel<<<nb, 256, 0, st>>>(A, SphOut<QSpherical> { *mp }, from, to);
        break;
      case EB200_METRIC_KERR_SCHILD:
        fields_to_phys_kernel<<<nb, 256, 0, st>>>(A, GrOut<KerrSchild> { *mp }, from, to);
        break;
      case EB200_METRIC_QKERR_SCHILD:
        fields_to_phys_kernel<<<nb, 256, 0, st>>>(A, GrOut<QKerrSchild> { *mp }, from, to);
        break;
      case EB200_METRIC_KERR_SCHILD_0:
        fields_to_phys_kernel<<<nb, 256, 0, st>>>(A, GrOut<KerrSchild0> { *mp }, from, to);
        break;
      default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
  }

  cudaError_t prtls_to_phys(const MetricParams* mp, const eb200_grid_t& g, float dx,
                            const float* xmin, const eb200_prtls_t& S, uint32_t stride,
                            uint32_t nout, float* x1, float* x2, float* x3, float* u1, float* u2,
                            float* u3, float* w, cudaStream_t st) {
    if (nout == 0) return cudaSuccess;
    P2PArgs A;
    A.stride = stride, A.nout = nout, A.dim = g.dim, A.dx = dx;
    for (int a = 0; a < 3; ++a) A.xmin[a] = xmin ? xmin[a] : ZERO;
    const unsigned nb   = (nout + 255) / 256;
    const int      kind = mp ? mp->kind : EB200_METRIC_MINKOWSKI;
    MetricParams   m0 {};
    const MetricParams& m = mp ? *mp : m0;
#define RUN(K, MM) prtls_to_phys_kernel<K, MM><<<nb, 256, 0, st>>>(A, m, S, x1, x2, x3, u1, u2, u3, w)
    switch (kind) {
      case EB200_METRIC_MINKOWSKI: RUN(0, Spherical); break;
      case EB200_METRIC_SPHERICAL: RUN(1, Spherical); break;
      case EB200_METRIC_QSPHERICAL: RUN(1, QSpherical); break;
      case EB200_METRIC_KERR_SCHILD: RUN(2, KerrSchild); break;
      case EB200_METRIC_QKERR_SCHILD: RUN(2, QKerrSchild); break;
      case EB200_METRIC_KERR_SCHILD_0: RUN(2, KerrSchild0); break;
      default: return cudaErrorInvalidValue;
    }
#undef RUN
    count_launch();
    return cudaGetLastError();
  }
} // namespace eb200
