// entity_b200 -- reduced statistics of a Minkowski domain:
// kernel::ReducedFields_kernel and kernel::ReducedParticleMoments_kernel
// (src/kernels/reduced_stats.hpp:25-541) as grid-wide reductions. Each thread evaluates the
// reference's per-cell / per-particle term in fp32 (same expression), the terms are summed in
// fp64 (warp shuffles, one atomicAdd per block). The value returned is the LOCAL sum, i.e. what
// Kokkos::parallel_reduce hands back in ReduceFields / ComputeMoments
// (src/framework/domain/metadomain_stats.cpp:88-183) before the MPI reduction and the division
// by totVolume * ppc0.
#include "common.cuh"
#include "launch.h"
#include "metrics.cuh"

namespace eb200 {
  namespace {

    __device__ __forceinline__ void block_add(double v, double* out) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      __shared__ double part[8];
      const int         lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      if (lane == 0) part[warp] = v;
      __syncthreads();
      if (warp == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? part[lane] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v != 0.0) atomicAdd(out, v);
      }
    }

    // value of component c (0..2 of the E, B or J triple starting at plane c0) interpolated to
    // the cell centre: averaged over the active dimensions in which it is NOT staggered
    // (reduced_stats.hpp:81-128, 186-251, 296-386), lower dimension fastest, HALF / INV_4 of
    // the plain sum. E_a and J_a are staggered in dimension a only, B_a in all but a.
    template <int D>
    __device__ __forceinline__ float centred(const FieldView<D>& F, int c0, int c, bool is_b, int i,
                                             int j, int k) {
      int avg[3] = { 0, 0, 0 }, na = 0;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const bool staggered = is_b ? (a != c) : (a == c);
        if (!staggered) avg[na++] = a;
      }
      float     s = 0.0f;
      const int n = 1 << na;
      for (int m = 0; m < n; ++m) {
        int o[3] = { 0, 0, 0 };
        for (int q = 0; q < na; ++q) o[avg[q]] = (m >> q) & 1;
        s += F.ld(i + o[0], j + o[1], k + o[2], c0 + c);
      }
      if (na == 0) return s;
      if (na == 1) return 0.5f * s;
      if (na == 2) return 0.25f * s;
      return 0.125f * s; // 3D only for a component staggered nowhere: there is none
    }

    template <int D>
    __global__ void __launch_bounds__(256)
      stats_fields_kernel(FieldView<D> EM, FieldView<D> J, int n1, int n2, int n3, int G, float dx,
                          int what, int comp, double* out) {
      const long n   = (long)n1 * n2 * n3;
      const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
      double     v   = 0.0;
      if (idx < n) {
        const int   i   = (int)(idx % n1) + G;
        const int   j   = (D > 1) ? (int)((idx / n1) % n2) + G : 0;
        const int   k   = (D > 2) ? (int)(idx / ((long)n1 * n2)) + G : 0;
        const float sdh = (D == 1) ? dx : ((D == 2) ? dx * dx : dx * dx * dx);
        auto        fT  = [&](int a) { return (a < D) ? dx : 1.0f; };      // sqrt(h_aa)
        auto        fD  = [&](int a) { return (a < D) ? dx * dx : 1.0f; }; // h_aa
        const int   c   = comp - 1;
        float       t   = 0.0f;
        if (what == EB200_STATS_B2 || what == EB200_STATS_E2) {
          const float u = EM.ld(i, j, k, (what == EB200_STATS_B2 ? 3 : 0) + c);
          t             = u * (u * fD(c)) * sdh;
        } else if (what == EB200_STATS_EXB) {
          const int   a = (c + 1) % 3, b = (c + 2) % 3;
          const float ea = centred<D>(EM, 0, a, false, i, j, k) * fT(a);
          const float eb = centred<D>(EM, 0, b, false, i, j, k) * fT(b);
          const float ba = centred<D>(EM, 3, a, true, i, j, k) * fT(a);
          const float bb = centred<D>(EM, 3, b, true, i, j, k) * fT(b);
          t              = (ea * bb - eb * ba) * sdh;
        } else {
          float e[3], q[3];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            e[a] = centred<D>(EM, 0, a, false, i, j, k) * fT(a);
            q[a] = centred<D>(J, 0, a, false, i, j, k) * fT(a);
          }
          t = (e[0] * q[0] + e[1] * q[1] + e[2] * q[2]) * sdh;
        }
        v = (double)t;
      }
      block_add(v, out);
    }

    template <int D>
    __global__ void __launch_bounds__(256)
      stats_particles_kernel(eb200_prtls_t S, uint32_t npart, float mass, float charge,
                             int use_weights, float dx, int what, int c1, int c2, double* out) {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      double         v = 0.0;
      if (p < npart && S.tag[p] == 1) {
        const float dV = (D == 1) ? dx : ((D == 2) ? dx * dx : dx * dx * dx);
        if (what == EB200_STATS_NPART) {
          v = 1.0;
        } else if (what == EB200_STATS_N || what == EB200_STATS_RHO || what == EB200_STATS_CHARGE) {
          const float contrib = (what == EB200_STATS_RHO) ? mass
                                                          : ((what == EB200_STATS_CHARGE) ? charge : 1.0f);
          v = (double)(dV * (use_weights ? S.weight[p] : contrib));
        } else {
          // stress-energy component (c1, c2) in the tetrad basis; as in the reference neither the
          // weight nor the mass-density contribution enters (reduced_stats.hpp:524-536)
          const float u[3]   = { S.ux1[p], S.ux2[p], S.ux3[p] };
          const float usq    = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
          const float energy = (mass == 0.0f) ? sqrtf(usq) : mass * sqrtf(1.0f + usq);
          float       coeff  = 1.0f;
          coeff *= (c1 == 0) ? energy : u[c1 - 1];
          coeff *= (c2 == 0) ? energy : u[c2 - 1];
          v = (double)(dV * coeff / energy);
        }
      }
      block_add(v, out);
    }
    // ReducedFields_kernel<SRPIC, M, F, I> for a diagonal 2D metric (reduced_stats.hpp:145-251):
    // |B|^2 and |E|^2 as u_a (h_aa u_a) sqrt_det_h on the component's own node; E x B and J.E from
    // the components averaged to the cell centre and taken to the tetrad basis there
    template <class M>
    __global__ void __launch_bounds__(256)
      stats_fields_curv_kernel(FieldView<2> EM, FieldView<2> J, int n1, int n2, int G,
                               const MetricParams mp, int what, int comp, double* out) {
      const long n   = (long)n1 * n2;
      const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
      double     v   = 0.0;
      if (idx < n) {
        const int   i = (int)(idx % n1) + G, j = (int)(idx / n1) + G;
        const float x1 = static_cast<float>(i - G), x2 = static_cast<float>(j - G);
        const int   c = comp - 1;
        auto hD = [&](int a, float y1, float y2) {
          return a == 0 ? M::h11(mp, y1, y2) : (a == 1 ? M::h22(mp, y1, y2) : M::h33(mp, y1, y2));
        };
        auto hT = [&](int a, float y1, float y2) {
          return a == 0 ? M::sqrt_h11(mp, y1, y2) : (a == 1 ? M::sqrt_h22(mp, y1, y2) : M::sqrt_h33(mp, y1, y2));
        };
        float t = 0.0f;
        if (what == EB200_STATS_B2 || what == EB200_STATS_E2) {
          const bool  is_b = what == EB200_STATS_B2;
          // node of the component: E_a staggered in a only, B_a in all but a (2D: dims 0, 1)
          const float y1 = x1 + (((is_b ? (c != 0) : (c == 0))) ? 0.5f : 0.0f);
          const float y2 = x2 + (((is_b ? (c != 1) : (c == 1))) ? 0.5f : 0.0f);
          const float u  = EM.ld(i, j, 0, (is_b ? 3 : 0) + c);
          t              = u * (hD(c, y1, y2) * u) * M::sqrt_det_h(mp, y1, y2);
        } else {
          const float y1 = x1 + 0.5f, y2 = x2 + 0.5f;
          const float sd = M::sqrt_det_h(mp, y1, y2);
          if (what == EB200_STATS_EXB) {
            const int   a = (c + 1) % 3, b = (c + 2) % 3;
            const float ea = centred<2>(EM, 0, a, false, i, j, 0) * hT(a, y1, y2);
            const float eb = centred<2>(EM, 0, b, false, i, j, 0) * hT(b, y1, y2);
            const float ba = centred<2>(EM, 3, a, true, i, j, 0) * hT(a, y1, y2);
            const float bb = centred<2>(EM, 3, b, true, i, j, 0) * hT(b, y1, y2);
            t              = (ea * bb - eb * ba) * sd;
          } else {
            float e[3], q[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              e[a] = centred<2>(EM, 0, a, false, i, j, 0) * hT(a, y1, y2);
              q[a] = centred<2>(J, 0, a, false, i, j, 0) * hT(a, y1, y2);
            }
            t = (e[0] * q[0] + e[1] * q[1] + e[2] * q[2]) * sd;
          }
        }
        v = (double)t;
      }
      block_add(v, out);
    }

    // ReducedParticleMoments_kernel<SRPIC, M, P> (reduced_stats.hpp:400-536): dV = sqrt_det_h at
    // the particle, T^{ab} from the momentum taken to the tetrad basis at (x, phi)
    template <class M>
    __global__ void __launch_bounds__(256)
      stats_particles_curv_kernel(eb200_prtls_t S, uint32_t npart, float mass, float charge,
                                  int use_weights, const MetricParams mp, int what, int c1, int c2,
                                  double* out) {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      double         v = 0.0;
      if (p < npart && S.tag[p] == 1) {
        if (what == EB200_STATS_NPART) {
          v = 1.0;
        } else {
          const float x[3] = { static_cast<float>(S.i1[p]) + S.dx1[p], static_cast<float>(S.i2[p]) + S.dx2[p],
                               S.phi[p] };
          const float dV = M::sqrt_det_h(mp, x[0], x[1]);
          if (what == EB200_STATS_N || what == EB200_STATS_RHO || what == EB200_STATS_CHARGE) {
            const float contrib = (what == EB200_STATS_RHO) ? mass
                                                            : ((what == EB200_STATS_CHARGE) ? charge : 1.0f);
            v = (double)(dV * (use_weights ? S.weight[p] : contrib));
          } else {
            const float uc[3] = { S.ux1[p], S.ux2[p], S.ux3[p] };
            float       u[3];
            xyz_to_tetrad(trig_at<M>(mp, x), uc, u);
            const float usq    = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
            const float energy = (mass == 0.0f) ? sqrtf(usq) : mass * sqrtf(1.0f + usq);
            float       coeff  = 1.0f;
            coeff *= (c1 == 0) ? energy : u[c1 - 1];
            coeff *= (c2 == 0) ? energy : u[c2 - 1];
            v = (double)(dV * coeff / energy);
          }
        }
      }
      block_add(v, out);
    }
  } // namespace

  cudaError_t stats_fields_curv(const MetricParams& mp, const eb200_grid_t& g, const float* em,
                                const float* cur, int what, int comp, double* out_dev, cudaStream_t st) {
    const long  n = (long)g.n[0] * g.n[1];
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(double), st);
    if (e != cudaSuccess) return e;
    const unsigned nb = (unsigned)((n + 255) / 256);
    float*         E  = const_cast<float*>(em);
    float*         Jp = const_cast<float*>(cur ? cur : em);
    if (mp.kind == EB200_METRIC_SPHERICAL) {
      stats_fields_curv_kernel<Spherical><<<nb, 256, 0, st>>>(FieldView<2>(g, E), FieldView<2>(g, Jp), g.n[0],
                                                             g.n[1], g.ng, mp, what, comp, out_dev);
    } else if (mp.kind == EB200_METRIC_QSPHERICAL) {
      stats_fields_curv_kernel<QSpherical><<<nb, 256, 0, st>>>(FieldView<2>(g, E), FieldView<2>(g, Jp), g.n[0],
                                                              g.n[1], g.ng, mp, what, comp, out_dev);
    } else {
      return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
  }

  cudaError_t stats_particles_curv(const MetricParams& mp, const eb200_prtls_t& S, uint32_t npart,
                                   float mass, float charge, int use_weights, int what, int c1, int c2,
                                   double* out_dev, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(double), st);
    if (e != cudaSuccess) return e;
    if (npart == 0) return cudaSuccess;
    const unsigned nb = (npart + 255) / 256;
    if (mp.kind == EB200_METRIC_SPHERICAL) {
      stats_particles_curv_kernel<Spherical><<<nb, 256, 0, st>>>(S, npart, mass, charge, use_weights, mp, what,
                                                                c1, c2, out_dev);
    } else if (mp.kind == EB200_METRIC_QSPHERICAL) {
      stats_particles_curv_kernel<QSpherical><<<nb, 256, 0, st>>>(S, npart, mass, charge, use_weights, mp, what,
                                                                 c1, c2, out_dev);
    } else {
      return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
  }

  cudaError_t stats_fields(const eb200_grid_t& g, const float* em, const float* cur, float dx,
                           int what, int comp, double* out_dev, cudaStream_t st) {
    const int  n1 = g.n[0], n2 = g.dim > 1 ? g.n[1] : 1, n3 = g.dim > 2 ? g.n[2] : 1;
    const long n  = (long)n1 * n2 * n3;
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(double), st);
    if (e != cudaSuccess) return e;
    const unsigned nb = (unsigned)((n + 255) / 256);
    float*         E  = const_cast<float*>(em);
    float*         Jp = const_cast<float*>(cur ? cur : em);
    switch (g.dim) {
      case 1:
        stats_fields_kernel<1><<<nb, 256, 0, st>>>(FieldView<1>(g, E), FieldView<1>(g, Jp), n1, n2, n3,
                                                   g.ng, dx, what, comp, out_dev);
        break;
      case 2:
        stats_fields_kernel<2><<<nb, 256, 0, st>>>(FieldView<2>(g, E), FieldView<2>(g, Jp), n1, n2, n3,
                                                   g.ng, dx, what, comp, out_dev);
        break;
      case 3:
        stats_fields_kernel<3><<<nb, 256, 0, st>>>(FieldView<3>(g, E), FieldView<3>(g, Jp), n1, n2, n3,
                                                   g.ng, dx, what, comp, out_dev);
        break;
      default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
  }

  cudaError_t stats_particles(const eb200_grid_t& g, const eb200_prtls_t& S, uint32_t npart,
                              float mass, float charge, int use_weights, float dx, int what, int c1,
                              int c2, double* out_dev, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out_dev, 0, sizeof(double), st);
    if (e != cudaSuccess) return e;
    if (npart == 0) return cudaSuccess;
    const unsigned nb = (npart + 255) / 256;
    switch (g.dim) {
      case 1:
        stats_particles_kernel<1><<<nb, 256, 0, st>>>(S, npart, mass, charge, use_weights, dx, what, c1,
                                                      c2, out_dev);
        break;
      case 2:
        stats_particles_kernel<2><<<nb, 256, 0, st>>>(S, npart, mass, charge, use_weights, dx, what, c1,
                                                      c2, out_dev);
        break;
      case 3:
        stats_particles_kernel<3><<<nb, 256, 0, st>>>(S, npart, mass, charge, use_weights, dx, what, c1,
                                                      c2, out_dev);
        break;
      default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
  }
} // namespace eb200
