// entity_b200 -- particle injection and per-cell particle moments.
//
//  * eb200_inject_nonuniform = arch::InjectNonUniform (src/archetypes/particle_injector.h:
//    296-387) with kernel::NonUniformInjector_kernel (src/kernels/injectors.hpp:526-859): in
//    every cell of a range, ppc = ppc0 * spatial_dist(cell centre) pairs (stochastic rounding of
//    the fraction), each pair at one uniformly drawn position of the cell, velocities drawn
//    independently for the two species from their energy distributions, weight from the
//    spatial distribution (x sqrt_det_h / V0 on curvilinear meshes). It covers the archetypes'
//    uniform injector too (spatial_dist = 1: InjectUniform draws the same expected number).
//  * spatial distributions that cannot cross a C ABI as functors arrive as a per-cell table;
//    arch::spatial_dist::ReplenishUniform (spatial_dist.h:87-125) is built in: it reads the
//    density moment this file also computes.
//  * energy distributions: arch::energy_dist::Maxwellian (energy_dist.h:77-290): Box-Muller
//    below T = 1/2, Sobol's method above, drift by the flipping method.
//  * eb200_particle_moment = arch::ComputeMomentWithSpecies / kernel::ParticleMoments_kernel
//    (src/kernels/particle_moments.hpp:37-419) for N, Rho, Charge and Nppc on Minkowski
//    meshes, smoothing order 0 (the archetypes' default).
//
// RNG: the reference draws from a Kokkos random pool (one XorShift stream per thread, so the
// particles differ between backends and thread counts). Here every cell owns a counter-based
// Philox4x32-10 stream keyed by (seed, step, call id, cell): the injected particles are a pure
// function of the inputs -- same on any GPU, any launch shape -- and land in the arrays in
// cell order (the reference's atomic slot counter makes the order non-deterministic).
#include "common.cuh"
#include "launch.h"
#include "metrics.cuh"
#include "philox.cuh"

#include <cub/device/device_scan.cuh>

namespace eb200 {
  namespace {
    struct Maxwell {
      float temperature;
      float drift_4vel, drift_3vel, dir[3];
      int   drift_dir; // 0 none, +-1..3 principal axes, 4 arbitrary (energy_dist.h:199-222)
    };

    // JuttnerSinge (energy_dist.h:77-131)
    __device__ void juttner_synge(Philox& g, float temp, float* v) {
      if (temp < 0.5f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float x1 = g.uniform();
          while (fabsf(x1) <= 1.1920929e-07f) x1 = g.uniform();
          x1             = sqrtf(-TWO * logf(x1));
          const float x2 = 6.2831855f * g.uniform();
          v[c]           = x1 * cosf(x2) * sqrtf(temp);
        }
      } else {
        float randu = ONE, randeta = g.uniform(), x1, x2;
        while (SQR(randeta) <= SQR(randu) + ONE) {
          x1 = g.uniform() * g.uniform() * g.uniform();
          while (fabsf(x1) <= 1.1920929e-07f) x1 = g.uniform() * g.uniform() * g.uniform();
          randu = -temp * logf(x1);
          x2    = g.uniform();
          while (fabsf(x2) <= 1.1920929e-07f) x2 = g.uniform();
          randeta = -temp * logf(x1 * x2);
        }
        x1   = g.uniform();
        x2   = g.uniform();
        v[0] = randu * (TWO * x1 - ONE);
        v[2] = TWO * randu * sqrtf(x1 * (ONE - x1));
        v[1] = v[2] * cosf(6.2831855f * x2);
        v[2] = v[2] * sinf(6.2831855f * x2);
      }
    }

    // Maxwellian::operator() (energy_dist.h:223-268), Cartesian drift included
    __device__ void sample(Philox& g, const Maxwell& m, float* v) {
      if (fabsf(m.temperature) <= 1.1920929e-07f) {
        v[0] = v[1] = v[2] = ZERO;
      } else {
        juttner_synge(g, m.temperature, v);
      }
      if (m.drift_dir != 0) {
        const float gamma = sqrtf(ONE + SQR(v[0]) + SQR(v[1]) + SQR(v[2]));
        if (-m.drift_3vel * v[0] > gamma * g.uniform()) v[0] = -v[0];
        v[0] = sqrtf(ONE + SQR(m.drift_4vel)) * (v[0] + m.drift_3vel * gamma);
        const int d = m.drift_dir;
        if (d == -1) {
          v[0] = -v[0];
        } else if (d == 2 || d == -2) {
          const float t = v[1];
          v[1]          = d > 0 ? v[0] : -v[0];
          v[0]          = t;
        } else if (d == 3 || d == -3) {
          const float t = v[2];
          v[2]          = d > 0 ? v[0] : -v[0];
          v[0]          = t;
        } else if (d == 4) {
          const float o0 = v[0], o1 = v[1], o2 = v[2];
          const float d1 = m.dir[0], d2 = m.dir[1], d3 = m.dir[2];
          v[0] = o0 * d1 - o1 * d2 - o2 * d3;
          v[1] = (o0 * d2 * (d1 + ONE) + o1 * (SQR(d1) + d1 + SQR(d3)) - o2 * d2 * d3) / (d1 + ONE);
          v[2] = (o0 * d3 * (d1 + ONE) - o1 * d2 * d3 - o2 * (-d1 + SQR(d3) - ONE)) / (d1 + ONE);
        }
      }
    }

    struct InjArgs {
      int      lo[3], n[3]; // ghost-inclusive start and extent of the cell range
      int      dim, G;
      long     N1, N12;
      float    ppc0;
      int      sd_kind;     // EB200_SDIST_*
      const float* field;   // table / density moment, component plane `comp` selected by the host
      float    target;
      const float* target_field; // REPLENISH_TABLE
      float    target_max;
      int      atm_dim, atm_sign; // ATMOSPHERE
      float    atm_nmax, atm_height, atm_xsurf, atm_ds;
      float    dx, xmin[3];  // Minkowski: x_Ph = (cell + 1/2) dx + xmin
      float    inv_V0;
      uint64_t seed;
      uint32_t step, call;
    };

    // arch::AtmosphereDensityProfile<D, C, P, O>::operator() (particle_injector.h:151-189); xi =
    // the physical coordinate of the cell centre along the boundary's dimension
    template <bool CURV>
    __device__ float atmosphere_profile(const InjArgs& A, float xi) {
      if (A.atm_sign > 0) {
        if (xi < A.atm_xsurf - A.atm_ds || xi >= A.atm_xsurf) return ZERO;
        return A.atm_nmax * expf(-(A.atm_xsurf - xi) / A.atm_height); // Cartesian only (checked on the host)
      }
      if (xi < A.atm_xsurf || xi >= A.atm_xsurf + A.atm_ds) return ZERO;
      if constexpr (!CURV) {
        return A.atm_nmax * expf(-(xi - A.atm_xsurf) / A.atm_height);
      } else {
        return A.atm_nmax * expf(-(A.atm_xsurf / A.atm_height) * (ONE - (A.atm_xsurf / xi)));
      }
    }

    // spatial_dist(cell centre) of the built-in kinds; c = ghost-inclusive cell indices
    template <bool CURV, class M>
    __device__ float spatial_value(const InjArgs& A, const MetricParams& mp, const int* c, long node) {
      if (A.sd_kind == EB200_SDIST_UNIFORM) return ONE;
      const float f = A.field[node];
      if (A.sd_kind == EB200_SDIST_TABLE) return f;
      if (A.sd_kind == EB200_SDIST_REPLENISH) {
        // ReplenishUniform (spatial_dist.h:104-124)
        return (0.9f * A.target > f) ? (A.target - f) / A.target : ZERO;
      }
      // Replenish<M, N, T> (spatial_dist.h:56-80)
      float target;
      if (A.sd_kind == EB200_SDIST_REPLENISH_TABLE) {
        target = A.target_field[node];
      } else {
        const float xc = static_cast<float>(c[A.atm_dim] - A.G) + HALF;
        float       xi;
        if constexpr (!CURV) {
          xi = xc * A.dx + A.xmin[A.atm_dim];
        } else {
          xi = (A.atm_dim == 0) ? M::r(mp, xc) : M::theta(mp, xc);
        }
        target = atmosphere_profile<CURV>(A, xi);
      }
      return (0.9f * target > f) ? (target - f) / A.target_max : ZERO;
    }

    __device__ void cell_of_range(const InjArgs& A, long t, int* c, long& node) {
      c[0] = (int)(t % A.n[0]) + A.lo[0];
      c[1] = (A.dim > 1) ? (int)((t / A.n[0]) % A.n[1]) + A.lo[1] : 0;
      c[2] = (A.dim > 2) ? (int)(t / ((long)A.n[0] * A.n[1])) + A.lo[2] : 0;
      node = c[0] + (long)c[1] * A.N1 + (long)c[2] * A.N12;
    }

    // pass 1: NonUniformInjector_kernel::injected_ppc (injectors.hpp:616-633) per cell
    template <bool CURV, class M>
    __global__ void __launch_bounds__(256)
      inject_count_kernel(const __grid_constant__ InjArgs A, const MetricParams mp,
                          uint32_t* __restrict__ counts) {
      const long t     = (long)blockIdx.x * blockDim.x + threadIdx.x;
      const long total = (long)A.n[0] * A.n[1] * A.n[2];
      if (t >= total) return;
      int  c[3];
      long node;
      cell_of_range(A, t, c, node);
      const float ppc_real = A.ppc0 * spatial_value<CURV, M>(A, mp, c, node);
      uint32_t    ppc      = (uint32_t)ppc_real;
      Philox      g(A.seed, A.step, A.call, (uint32_t)t);
      if (g.uniform() < ppc_real - (float)ppc) ppc += 1;
      counts[t] = ppc;
    }

    // pass 2: the particles of every cell, at the cell's offset of the exclusive scan
    template <bool CURV, class M>
    __global__ void __launch_bounds__(256)
      inject_fill_kernel(const __grid_constant__ InjArgs A, const MetricParams mp,
                         const uint32_t* __restrict__ counts,
                         const uint32_t* __restrict__ offsets, eb200_prtls_t S1, eb200_prtls_t S2,
                         uint32_t off1, uint32_t off2, Maxwell m1, Maxwell m2) {
      const long t     = (long)blockIdx.x * blockDim.x + threadIdx.x;
      const long total = (long)A.n[0] * A.n[1] * A.n[2];
      if (t >= total) return;
      const uint32_t ppc = counts[t];
      if (ppc == 0) return;
      int  c[3];
      long node;
      cell_of_range(A, t, c, node);
      Philox g(A.seed, A.step, A.call, (uint32_t)t);
      (void)g.uniform(); // the draw pass 1 used for the rounding
      const uint32_t base = offsets[t];
      int*           i1[2] = { S1.i1, S2.i1 }, *i2[2] = { S1.i2, S2.i2 }, *i3[2] = { S1.i3, S2.i3 };
      float*         d1[2] = { S1.dx1, S2.dx1 }, *d2[2] = { S1.dx2, S2.dx2 }, *d3[2] = { S1.dx3, S2.dx3 };
      float*         u1[2] = { S1.ux1, S2.ux1 }, *u2[2] = { S1.ux2, S2.ux2 }, *u3[2] = { S1.ux3, S2.ux3 };
      int*           p1[2] = { S1.i1_prev, S2.i1_prev }, *p2[2] = { S1.i2_prev, S2.i2_prev },
          *p3[2] = { S1.i3_prev, S2.i3_prev };
      float*         q1[2] = { S1.dx1_prev, S2.dx1_prev }, *q2[2] = { S1.dx2_prev, S2.dx2_prev },
            *q3[2] = { S1.dx3_prev, S2.dx3_prev };
      float*         w[2]   = { S1.weight, S2.weight };
      float*         phi[2] = { S1.phi, S2.phi };
      short*         tag[2] = { S1.tag, S2.tag };
      const uint32_t off[2] = { off1, off2 };
      const Maxwell* mm[2]  = { &m1, &m2 };
      // curvilinear: weight *= sqrt_det_h(cell centre) / V0 (injectors.hpp:746-748); velocities are
      // drawn in the tetrad basis and stored Cartesian at (centre, phi = 0) (:758-764)
      float weight = ONE;
      Trig  trig {};
      if constexpr (CURV) {
        const float xc[3] = { static_cast<float>(c[0] - A.G) + HALF,
                              static_cast<float>(c[1] - A.G) + HALF, ZERO };
        weight *= M::sqrt_det_h(mp, xc[0], xc[1]) * A.inv_V0;
        trig = trig_at<M>(mp, xc);
      }
      for (uint32_t k = 0; k < ppc; ++k) {
        float dx[3] = { ZERO, ZERO, ZERO };
        for (int a = 0; a < A.dim; ++a) dx[a] = g.uniform();
        for (int s = 0; s < 2; ++s) {
          float v[3];
          sample(g, *mm[s], v);
          if constexpr (CURV) {
            float vx[3];
            tetrad_to_xyz(trig, v, vx);
            v[0] = vx[0], v[1] = vx[1], v[2] = vx[2];
          }
          const size_t p = (size_t)off[s] + base + k;
          i1[s][p] = c[0] - A.G, d1[s][p] = dx[0];
          p1[s][p] = c[0] - A.G, q1[s][p] = dx[0];
          if (A.dim > 1) {
            i2[s][p] = c[1] - A.G, d2[s][p] = dx[1];
            p2[s][p] = c[1] - A.G, q2[s][p] = dx[1];
          }
          if (A.dim > 2) {
            i3[s][p] = c[2] - A.G, d3[s][p] = dx[2];
            p3[s][p] = c[2] - A.G, q3[s][p] = dx[2];
          }
          u1[s][p] = v[0], u2[s][p] = v[1], u3[s][p] = v[2];
          w[s][p]   = weight;
          tag[s][p] = 1;
          if constexpr (CURV) {
            if (phi[s]) phi[s][p] = ZERO;
          }
        }
      }
    }

    // ParticleMoments_kernel<S, M, F, N>::operator() with window 0 (particle_moments.hpp:293-345):
    // contrib * inv_n0 / sqrt_det_h(cell centre) (* weight) into the particle's cell.
    // KIND 0: Minkowski (the constant volume element is folded into coeff by the host)
    template <int KIND, class M>
    __global__ void __launch_bounds__(256)
      moment_kernel(eb200_prtls_t S, uint32_t npart, int dim, int G, long N1, long N12, float coeff,
                    bool use_weights, bool volume, const MetricParams mp, float* __restrict__ plane) {
      // whole warps stay: consecutive (nearly cell-sorted) particles of the same cell are summed
      // with a segmented shuffle reduction and the run's head lane issues one atomic
      const uint32_t p     = blockIdx.x * blockDim.x + threadIdx.x;
      const bool     alive = (p < npart) && (S.tag[p] != 0);
      float          c     = ZERO;
      long           node  = -1 - (long)(threadIdx.x & 31u); // distinct keys for idle lanes
      if (alive) {
        c = coeff;
        if constexpr (KIND != 0) {
          if (volume) {
            c = c / M::sqrt_det_h(mp, static_cast<float>(S.i1[p]) + HALF,
                                  static_cast<float>(S.i2[p]) + HALF);
          }
        }
        if (use_weights) c *= S.weight[p];
        node = S.i1[p] + G;
        if (dim > 1) node += (long)(S.i2[p] + G) * N1;
        if (dim > 2) node += (long)(S.i3[p] + G) * N12;
      }
      const LaneRun r   = lane_run(node);
      const float   sum = lane_run_sum(c, r);
      if (r.head && alive) atomicAdd(plane + node, sum);
    }
  } // namespace

  static Maxwell make_maxwell(const eb200_maxwellian_t& e) {
    Maxwell m {};
    m.temperature = e.temperature;
    const float n = std::sqrt(e.drift_u[0] * e.drift_u[0] + e.drift_u[1] * e.drift_u[1] +
                              e.drift_u[2] * e.drift_u[2]);
    m.drift_4vel  = n;
    const float eps = 1.1920929e-07f;
    if (std::fabs(n) <= eps) {
      m.drift_dir = 0;
      return m;
    }
    m.drift_3vel = n / std::sqrt(1.0f + n * n);
    for (int a = 0; a < 3; ++a) m.dir[a] = e.drift_u[a] / n;
    m.drift_dir = 4;
    for (int d = 0; d < 3; ++d) {
      const int dp = (d + 2) % 3, dn = (d + 1) % 3;
      if (std::fabs(e.drift_u[dp]) <= eps && std::fabs(e.drift_u[dn]) <= eps) {
        m.drift_dir = (e.drift_u[d] > 0 ? 1 : -1) * (d + 1);
        break;
      }
    }
    return m;
  }

  cudaError_t inject_nonuniform(const eb200_grid_t& g, const MetricParams* mp, float dx,
                                const float* xmin, const eb200_prtls_t& S1, uint32_t npart1,
                                uint32_t cap1, const eb200_prtls_t& S2, uint32_t npart2,
                                uint32_t cap2, float ppc0, const eb200_spatial_dist_t& sd,
                                const float* field, const eb200_maxwellian_t& e1,
                                const eb200_maxwellian_t& e2, const int* rmin, const int* rmax,
                                uint64_t seed, uint32_t step, uint32_t call, uint32_t* n_inj_host,
                                int* overflow, Scratch& scratch, cudaStream_t st) {
    InjArgs A {};
    long    total = 1;
    for (int a = 0; a < 3; ++a) {
      A.lo[a] = (a < g.dim) ? rmin[a] : 0;
      A.n[a]  = (a < g.dim) ? rmax[a] - rmin[a] : 1;
      if (A.n[a] <= 0) {
        *n_inj_host = 0;
        return cudaSuccess;
      }
      total *= A.n[a];
      A.xmin[a] = xmin ? xmin[a] : ZERO;
    }
    A.dim = g.dim, A.G = g.ng;
    A.N1 = g.n[0] + 2 * g.ng;
    A.N12 = (g.dim > 1) ? A.N1 * (g.n[1] + 2 * g.ng) : 0;
    A.ppc0 = ppc0, A.sd_kind = sd.kind, A.field = field, A.target = sd.target_density;
    A.target_field = sd.target_field;
    A.target_max   = (sd.kind == EB200_SDIST_ATMOSPHERE) ? sd.atm_nmax : sd.target_max;
    A.atm_dim = sd.atm_dim, A.atm_sign = sd.atm_sign;
    A.atm_nmax = sd.atm_nmax, A.atm_height = sd.atm_height;
    A.atm_xsurf = sd.atm_xsurf, A.atm_ds = sd.atm_ds;
    A.dx = dx, A.inv_V0 = sd.inv_V0;
    A.seed = seed, A.step = step, A.call = call;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)total + 1, st);
    const size_t nb4 = ((size_t)(total + 1) * 4 + 255) / 256 * 256;
    cudaError_t  e   = scratch.reserve(2 * nb4 + tmp + 256);
    if (e != cudaSuccess) return e;
    uint32_t* counts  = (uint32_t*)scratch.ptr;
    uint32_t* offsets = (uint32_t*)((char*)scratch.ptr + nb4);
    void*     ws      = (char*)scratch.ptr + 2 * nb4;
    const unsigned nb = (unsigned)((total + 255) / 256);
    e = cudaMemsetAsync(counts + total, 0, 4, st);
    if (e != cudaSuccess) return e;
    const int          kind = mp ? mp->kind : EB200_METRIC_MINKOWSKI;
    MetricParams       m0 {};
    const MetricParams& m = mp ? *mp : m0;
    switch (kind) {
      case EB200_METRIC_MINKOWSKI: inject_count_kernel<false, Spherical><<<nb, 256, 0, st>>>(A, m, counts); break;
      case EB200_METRIC_SPHERICAL: inject_count_kernel<true, Spherical><<<nb, 256, 0, st>>>(A, m, counts); break;
      case EB200_METRIC_QSPHERICAL: inject_count_kernel<true, QSpherical><<<nb, 256, 0, st>>>(A, m, counts); break;
      default: return cudaErrorInvalidValue;
    }
    count_launch();
    e = cub::DeviceScan::ExclusiveSum(ws, tmp, counts, offsets, (int)total + 1, st);
    if (e != cudaSuccess) return e;
    count_launch();
    uint32_t n_inj = 0;
    e = cudaMemcpyAsync(&n_inj, offsets + total, 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st); // the new particle count is a host-visible result
    if (e != cudaSuccess) return e;
    *n_inj_host = n_inj;
    *overflow   = ((uint64_t)npart1 + n_inj > cap1) || ((uint64_t)npart2 + n_inj > cap2);
    if (*overflow || n_inj == 0) return cudaSuccess;
    const Maxwell mx1 = make_maxwell(e1), mx2 = make_maxwell(e2);
#define FILL(C, MM) inject_fill_kernel<C, MM><<<nb, 256, 0, st>>>(A, m, counts, offsets, S1, S2, npart1, npart2, mx1, mx2)
    switch (kind) {
      case EB200_METRIC_MINKOWSKI: FILL(false, Spherical); break;
      case EB200_METRIC_SPHERICAL: FILL(true, Spherical); break;
      case EB200_METRIC_QSPHERICAL: FILL(true, QSpherical); break;
      default: return cudaErrorInvalidValue;
    }
#undef FILL
    count_launch();
    return cudaGetLastError();
  }

  cudaError_t particle_moment(const eb200_grid_t& g, const MetricParams* mp, const eb200_prtls_t& S,
                              uint32_t npart, float coeff, bool use_weights, bool volume,
                              float* plane, cudaStream_t st) {
    if (npart == 0) return cudaSuccess;
    const long N1 = g.n[0] + 2 * g.ng, N12 = (g.dim > 1) ? N1 * (g.n[1] + 2 * g.ng) : 0;
    const unsigned nb   = (npart + 255) / 256;
    const int      kind = mp ? mp->kind : EB200_METRIC_MINKOWSKI;
    MetricParams   m0 {};
    const MetricParams& m = mp ? *mp : m0;
#define RUN(K, MM) moment_kernel<K, MM><<<nb, 256, 0, st>>>(S, npart, g.dim, g.ng, N1, N12, coeff, use_weights, volume, m, plane)
    switch (kind) {
      case EB200_METRIC_MINKOWSKI: RUN(0, Spherical); break;
      case EB200_METRIC_SPHERICAL: RUN(1, Spherical); break;
      case EB200_METRIC_QSPHERICAL: RUN(1, QSpherical); break;
      case EB200_METRIC_KERR_SCHILD: RUN(1, KerrSchild); break;
      case EB200_METRIC_QKERR_SCHILD: RUN(1, QKerrSchild); break;
      case EB200_METRIC_KERR_SCHILD_0: RUN(1, KerrSchild0); break;
      default: return cudaErrorInvalidValue;
    }
#undef RUN
    count_launch();
    return cudaGetLastError();
  }
} // namespace eb200
