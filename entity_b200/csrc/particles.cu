// entity_b200 -- particle kernels: SR push, current deposit, fused push+deposit.
//
// Compiled twice (EB200_STRICT=0/1, see common.cuh). Launch helpers at the bottom are
// what capi.cu calls.
#include "particle.cuh"
#include "launch.h"

#include <cub/device/device_radix_sort.cuh>

namespace eb200 {
  namespace EB200_VARIANT {

    /* ------------------------------------------------------------ load / store */
    template <int D>
    __device__ __forceinline__ void load_prtl(const eb200_prtls_t& S, uint32_t p, Prtl<D>& P,
                                              bool with_prev) {
      const int* __restrict__ ii[3]    = { S.i1, S.i2, S.i3 };
      const float* __restrict__ dd[3]  = { S.dx1, S.dx2, S.dx3 };
      const int* __restrict__ iip[3]   = { S.i1_prev, S.i2_prev, S.i3_prev };
      const float* __restrict__ ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        P.i[a] = P.ip[a] = 0;
        P.d[a] = P.dp[a] = ZERO;
      }
#pragma unroll
      for (int a = 0; a < D; ++a) {
        P.i[a] = ii[a][p];
        P.d[a] = dd[a][p];
        if (with_prev) {
          P.ip[a] = iip[a][p];
          P.dp[a] = ddp[a][p];
        }
      }
      P.u[0] = S.ux1[p];
      P.u[1] = S.ux2[p];
      P.u[2] = S.ux3[p];
      P.w    = S.weight[p];
    }

    template <int D>
    __device__ __forceinline__ void store_pushed(const eb200_prtls_t& S, uint32_t p,
                                                 const Prtl<D>& P, short tag_in) {
      int*   ii[3]  = { S.i1, S.i2, S.i3 };
      float* dd[3]  = { S.dx1, S.dx2, S.dx3 };
      int*   iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
      float* ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
      for (int a = 0; a < D; ++a) {
        ii[a][p]  = P.i[a];
        dd[a][p]  = P.d[a];
        iip[a][p] = P.ip[a];
        ddp[a][p] = P.dp[a];
      }
      S.ux1[p] = P.u[0];
      S.ux2[p] = P.u[1];
      S.ux3[p] = P.u[2];
      if (P.tag != tag_in) {
        S.tag[p] = P.tag;
      }
    }

    /* ------------------------------------------------------------------ kernels */
    template <int D, int O>
    __global__ void __launch_bounds__(256)
      push_kernel(PushArgs A, eb200_prtls_t S, uint32_t npart, FieldView<D> EB) {
      const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
      if (p >= npart) {
        return;
      }
      const short tag = S.tag[p];
      if (tag != 1) {
        return;
      }
      Prtl<D> P;
      load_prtl<D>(S, p, P, false);
      P.tag  = tag;
      auto F = [&](int i, int j, int k, int c) { return EB.ld(i, j, k, c); };
      push_particle<D, O>(A, F, P);
      store_pushed<D>(S, p, P, tag);
    }

    template <int D, int O, bool AGG>
    __global__ void __launch_bounds__(256)
      deposit_atomic_kernel(eb200_prtls_t S, uint32_t npart, float charge, float inv_dt,
                            float dxc, int G, FieldView<D> J) {
      const uint32_t p      = blockIdx.x * blockDim.x + threadIdx.x;
      const bool     active = (p < npart) && (S.tag[p] != 0);
      if constexpr (!AGG) {
        if (!active) {
          return;
        }
      }
      Prtl<D> P;
      if (active) {
        load_prtl<D>(S, p, P, true);
      }
      if constexpr (AGG) {
        deposit_particle_aggregated<D, O>(P, active, charge, inv_dt, dxc, G, J);
      } else {
        deposit_particle<D, O>(P, charge, inv_dt, dxc, G,
                               [&](int i, int j, int k, int c, float v, bool guard = true) {
                                 if (guard) atomicAdd(&J.at(i, j, k, c), v);
                               });
      }
    }

    template <int D, int O, bool AGG>
    __global__ void __launch_bounds__(256)
      push_deposit_kernel(PushArgs A, eb200_prtls_t S, uint32_t npart, FieldView<D> EB,
                          float charge, float inv_dt, FieldView<D> J) {
      const uint32_t p   = blockIdx.x * blockDim.x + threadIdx.x;
      short          tag = 0;
      if (p < npart) {
        tag = S.tag[p];
      }
      bool active = (tag == 1);
      if constexpr (!AGG) {
        if (!active) {
          return;
        }
      }
      Prtl<D> P;
      if (active) {
        load_prtl<D>(S, p, P, false);
        P.tag  = tag;
        auto F = [&](int i, int j, int k, int c) { return EB.ld(i, j, k, c); };
        push_particle<D, O>(A, F, P);
        store_pushed<D>(S, p, P, tag);
        active = (P.tag != 0);
      }
      if constexpr (AGG) {
        deposit_particle_aggregated<D, O>(P, active, charge, inv_dt, A.c.dx, A.ng, J);
      } else {
        if (active) {
          deposit_particle<D, O>(P, charge, inv_dt, A.c.dx, A.ng,
                                 [&](int i, int j, int k, int c, float v, bool guard = true) {
                                   if (guard) atomicAdd(&J.at(i, j, k, c), v);
                                 });
        }
      }
    }

    /* --------------------------------------------------- ordered (serial) deposit */
    // Every particle writes its contributions, in program order, as (key, value) pairs into a
    // fixed-size slot range [p*K, (p+1)*K); unused slots carry key = 0xFFFFFFFF. A stable radix
    // sort by key then lines up, for every J element, its contributions in particle order, and
    // one thread per element adds them one by one onto the current value: exactly the sum a
    // serial loop over the particles produces.
    template <int D, int O>
    struct SlotCount {
      static constexpr int N     = O + 2;
      static constexpr int value = (O == 0) ? (D == 1 ? 10 : (D == 2 ? 16 : 24))
                                            : (D == 1 ? 3 * N : (D == 2 ? 3 * N * N : 3 * N * N * N));
    };

    template <int D, int O>
    __global__ void __launch_bounds__(128)
      deposit_list_kernel(eb200_prtls_t S, uint32_t p0, uint32_t count, float charge,
                          float inv_dt, float dxc, int G, FieldView<D> J, uint32_t* keys,
                          float* vals) {
      constexpr int  K = SlotCount<D, O>::value;
      const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= count) {
        return;
      }
      const uint32_t p    = p0 + q;
      uint32_t*      kout = keys + (size_t)q * K;
      float*         vout = vals + (size_t)q * K;
      int            n    = 0;
      if (S.tag[p] != 0) {
        Prtl<D> P;
        load_prtl<D>(S, p, P, true);
        deposit_particle<D, O>(P, charge, inv_dt, dxc, G,
                               [&](int i, int j, int k, int c, float v, bool guard = true) {
                                 if (guard) {
                                   kout[n] = (uint32_t)(J.idx(i, j, k) + J.plane * c);
                                   vout[n] = v;
                                   ++n;
                                 }
                               });
      }
      for (; n < K; ++n) {
        kout[n] = 0xFFFFFFFFu;
        vout[n] = ZERO;
      }
    }

    __global__ void __launch_bounds__(256)
      ordered_sum_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ vals,
                         size_t n, float* __restrict__ J) {
      const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n) {
        return;
      }
      const uint32_t key = keys[t];
      if (key == 0xFFFFFFFFu || (t > 0 && keys[t - 1] == key)) {
        return; // not the head of a segment
      }
      float acc = J[key];
      for (size_t s = t; s < n && keys[s] == key; ++s) {
        acc += vals[s];
      }
      J[key] = acc;
    }

    /* ----------------------------------------------------------------- launchers */
    template <int D, int O>
    cudaError_t launch_push(const PushArgs& A, const eb200_prtls_t& S, uint32_t npart,
                            const eb200_grid_t& g, const float* em, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> EB(g, const_cast<float*>(em));
      push_kernel<D, O><<<(npart + 255) / 256, 256, 0, st>>>(A, S, npart, EB);
      count_launch();
      return cudaGetLastError();
    }

    template <int D, int O>
    cudaError_t launch_deposit(const eb200_prtls_t& S, uint32_t npart, const eb200_grid_t& g,
                               float charge, float dt, float dxc, float* cur, int mode,
                               Scratch& scratch, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> J(g, cur);
      const float  inv_dt = ONE / dt;
      if (mode == EB200_DEPOSIT_ATOMIC) {
        deposit_atomic_kernel<D, O, false>
          <<<(npart + 255) / 256, 256, 0, st>>>(S, npart, charge, inv_dt, dxc, g.ng, J);
        count_launch();
        return cudaGetLastError();
      }
      if (mode == EB200_DEPOSIT_AGGREGATED) {
        deposit_atomic_kernel<D, O, true>
          <<<(npart + 255) / 256, 256, 0, st>>>(S, npart, charge, inv_dt, dxc, g.ng, J);
        count_launch();
        return cudaGetLastError();
      }
      // ordered mode, in chunks that bound the scratch footprint
      constexpr int  K     = SlotCount<D, O>::value;
      const uint32_t chunk = (1u << 24) / K; // ~16M slots per chunk
      if ((size_t)J.plane * 3 >= 0xFFFFFFFFull) return cudaErrorInvalidValue;
      const size_t slots = (size_t)chunk * K;
      size_t       tmp_bytes = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                      (float*)nullptr, (float*)nullptr, slots, 0, 32, st);
      cudaError_t err = scratch.reserve(slots * 16 + tmp_bytes + 1024);
      if (err != cudaSuccess) return err;
      uint32_t* k0  = (uint32_t*)scratch.ptr;
      uint32_t* k1  = k0 + slots;
      float*    v0  = (float*)(k1 + slots);
      float*    v1  = v0 + slots;
      void*     tmp = (void*)(v1 + slots);
      for (uint32_t p0 = 0; p0 < npart; p0 += chunk) {
        const uint32_t cnt = (npart - p0 < chunk) ? (npart - p0) : chunk;
        const size_t   ns  = (size_t)cnt * K;
        deposit_list_kernel<D, O>
          <<<(cnt + 127) / 128, 128, 0, st>>>(S, p0, cnt, charge, inv_dt, dxc, g.ng, J, k0, v0);
        count_launch();
        err = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, ns, 0, 32, st);
        if (err != cudaSuccess) return err;
        count_launch(); // library sort (several kernels); counted once
        ordered_sum_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(k1, v1, ns, cur);
        count_launch();
      }
      return cudaGetLastError();
    }

    template <int D, int O>
    cudaError_t launch_push_deposit(const PushArgs& A, const eb200_prtls_t& S, uint32_t npart,
                                    const eb200_grid_t& g, const float* em, float* cur,
                                    int mode, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> EB(g, const_cast<float*>(em));
      FieldView<D> J(g, cur);
      if (mode == EB200_DEPOSIT_AGGREGATED) {
        push_deposit_kernel<D, O, true><<<(npart + 255) / 256, 256, 0, st>>>(
          A, S, npart, EB, A.c.charge, ONE / A.c.dt, J);
      } else {
        push_deposit_kernel<D, O, false><<<(npart + 255) / 256, 256, 0, st>>>(
          A, S, npart, EB, A.c.charge, ONE / A.c.dt, J);
      }
      count_launch();
      return cudaGetLastError();
    }

#define EB200_DISPATCH_DO(dim, order, CALL)                                                   \
  switch ((dim) * 10 + (order)) {                                                              \
    case 10: return CALL(1, 0);                                                                \
    case 11: return CALL(1, 1);                                                                \
    case 12: return CALL(1, 2);                                                                \
    case 13: return CALL(1, 3);                                                                \
    case 20: return CALL(2, 0);                                                                \
    case 21: return CALL(2, 1);                                                                \
    case 22: return CALL(2, 2);                                                                \
    case 23: return CALL(2, 3);                                                                \
    case 30: return CALL(3, 0);                                                                \
    case 31: return CALL(3, 1);                                                                \
    case 32: return CALL(3, 2);                                                                \
    case 33: return CALL(3, 3);                                                                \
    default: return cudaErrorInvalidValue;                                                     \
  }

    cudaError_t push_sr(const eb200_grid_t& g, int order, const eb200_pusher_t& c,
                        const eb200_prtls_t& S, uint32_t npart, const float* em,
                        cudaStream_t st) {
      PushArgs A;
      A.c   = c;
      A.ndh = HALF * (c.charge / c.mass) * c.omegaB0 * c.dt;
      A.ng  = g.ng;
      for (int a = 0; a < 3; ++a) A.ni[a] = g.n[a];
#define CALL(D, O) launch_push<D, O>(A, S, npart, g, em, st)
      EB200_DISPATCH_DO(g.dim, order, CALL)
#undef CALL
    }

    cudaError_t deposit(const eb200_grid_t& g, int order, const eb200_prtls_t& S, uint32_t npart,
                        float charge, float dt, float dxc, float* cur, int mode,
                        Scratch& scratch, cudaStream_t st) {
#define CALL(D, O) launch_deposit<D, O>(S, npart, g, charge, dt, dxc, cur, mode, scratch, st)
      EB200_DISPATCH_DO(g.dim, order, CALL)
#undef CALL
    }

    cudaError_t push_deposit_sr(const eb200_grid_t& g, int order, const eb200_pusher_t& c,
                                const eb200_prtls_t& S, uint32_t npart, const float* em,
                                float* cur, int mode, cudaStream_t st) {
      PushArgs A;
      A.c   = c;
      A.ndh = HALF * (c.charge / c.mass) * c.omegaB0 * c.dt;
      A.ng  = g.ng;
      for (int a = 0; a < 3; ++a) A.ni[a] = g.n[a];
#define CALL(D, O) launch_push_deposit<D, O>(A, S, npart, g, em, cur, mode, st)
      EB200_DISPATCH_DO(g.dim, order, CALL)
#undef CALL
    }

  } // namespace EB200_VARIANT
} // namespace eb200
