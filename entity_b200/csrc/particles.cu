// entity_b200 -- particle kernels: SR push, current deposit, fused push+deposit.
//
// Compiled twice (EB200_STRICT=0/1, see common.cuh). Launch helpers at the bottom are
// what capi.cu calls.
#include <cstdlib>
#include "particle.cuh"
#include "philox.cuh"
#include "launch.h"

#include <mutex>
#include <vector>

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

    // push_kernel with an emission policy (sr.hpp:290-331, 1501-1555; archetypes/emission/*.h)
    template <int D, int O>
    __global__ void __launch_bounds__(256)
      push_emit_kernel(PushArgs A, eb200_prtls_t S, uint32_t npart, FieldView<D> EB, EmitArgs M) {
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
      auto emit = [&](Prtl<D>& Q, const float* um, const float* er, const float* br) {
        float       du[3], energy, gamma;
        const float prob = emission_response(M.E, um, er, br, du, energy, gamma);
        Philox      rng(M.seed, M.step, M.call, p);
        // "should not emit if photon energy is above 20% of (gamma - 1) m c^2"
        const bool should = (rng.uniform() < prob) &&
                            (energy < M.E.species_mass * (gamma - ONE) * 0.2f);
        if (should && M.E.should_drag) {
          Q.u[0] += du[0];
          Q.u[1] += du[1];
          Q.u[2] += du[2];
        }
        if (should && energy >= M.E.energy_min) {
          const uint32_t slot = atomicAdd(M.counter, 1u);
          if (slot < M.cap - M.offset) {
            const size_t q   = (size_t)M.offset + slot;
            const float  mag = sqrtf(nsq(du));
            int*         ii[3]  = { M.ph.i1, M.ph.i2, M.ph.i3 };
            float*       dd[3]  = { M.ph.dx1, M.ph.dx2, M.ph.dx3 };
            int*         iip[3] = { M.ph.i1_prev, M.ph.i2_prev, M.ph.i3_prev };
            float*       ddp[3] = { M.ph.dx1_prev, M.ph.dx2_prev, M.ph.dx3_prev };
  #pragma unroll
            for (int a = 0; a < D; ++a) {
              ii[a][q] = Q.i[a], dd[a][q] = Q.d[a];
              iip[a][q] = Q.i[a], ddp[a][q] = Q.d[a];
            }
            M.ph.ux1[q]    = (-du[0] / mag) * energy;
            M.ph.ux2[q]    = (-du[1] / mag) * energy;
            M.ph.ux3[q]    = (-du[2] / mag) * energy;
            M.ph.weight[q] = M.E.photon_weight * Q.w;
            M.ph.tag[q]    = 1;
          }
        }
      };
      push_particle<D, O, decltype(F), false>(A, F, P, emit);
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
#if !EB200_STRICT
        if constexpr (D == 3 && O == 3) {
          deposit_esirkepov3_rows(P, charge, inv_dt, dxc, G, J);
        } else
#endif
        {
          deposit_particle<D, O>(P, charge, inv_dt, dxc, G,
                                 [&](int i, int j, int k, int c, float v, bool guard = true) {
                                   if (guard) atomicAdd(&J.at(i, j, k, c), v);
                                 });
        }
      }
    }

#ifndef EB200_PD_MINBLOCKS
  #define EB200_PD_MINBLOCKS 1
#endif
    template <int D, int O, bool AGG>
    __global__ void __launch_bounds__(256, EB200_PD_MINBLOCKS)
      push_deposit_kernel(PushArgs A, eb200_prtls_t S, uint32_t p_begin, uint32_t npart,
                          FieldView<D> EB, float charge, float inv_dt, FieldView<D> J) {
      const uint32_t p   = p_begin + blockIdx.x * blockDim.x + threadIdx.x;
      short          tag = 0;
      Prtl<D>        P;
      if (p < npart) {
        // tag and state are requested together: one DRAM round trip, not two
        tag = S.tag[p];
        load_prtl<D>(S, p, P, false);
        if (tag != 1) exc_append(A, p);
      }
      bool active = (tag == 1);
      if constexpr (!AGG) {
        if (!active) {
          return;
        }
      }
      if (active) {
        P.tag  = tag;
#ifdef EB200_X_NOGATHER
        auto F = [&](int i, int j, int k, int c) { return 1e-3f * (float)(c + 1) + 1e-6f * (float)i; };
#else
        auto F = [&](int i, int j, int k, int c) { return EB.ld(i, j, k, c); };
#endif
        push_particle<D, O>(A, F, P);
#ifdef EB200_X_NOSTORE
        if (P.u[0] == 1.2345e-30f)
#endif
        store_pushed<D>(S, p, P, tag);
        if (P.tag != 1) exc_append(A, p);
        active = (P.tag != 0);
      }
#ifdef EB200_X_NODEPOSIT
      if (P.u[1] != 1.2345e-30f) return;
#endif
      if constexpr (AGG) {
        deposit_particle_aggregated<D, O>(P, active, charge, inv_dt, A.c.dx, A.ng, J);
      } else {
        if (active) {
#if !EB200_STRICT
          if constexpr (D == 3 && O == 3) {
            deposit_esirkepov3_rows(P, charge, inv_dt, A.c.dx, A.ng, J);
          } else
#endif
          {
            deposit_particle<D, O>(P, charge, inv_dt, A.c.dx, A.ng,
                                   [&](int i, int j, int k, int c, float v, bool guard = true) {
                                     if (guard) atomicAdd(&J.at(i, j, k, c), v);
                                   });
          }
        }
      }
    }


    /* ------------------------------------------- TMA-staged persistent push + deposit */
    // The throughput kernel. One persistent CTA per SM slot streams chunks of 256 particles:
    //  * the SoA slices of a chunk travel global -> shared with cp.async.bulk (TMA), three
    //    chunks deep, completion tracked by mbarriers: no warp ever waits on DRAM, no
    //    registers hold loads in flight, and there is no per-thread address arithmetic;
    //  * results go shared -> global with cp.async.bulk as well; i_prev / dx_prev are stored
    //    straight from the staged input slices (they ARE the old i / dx);
    //  * the deposit is the warp-aggregated one (one atomic per run of same-cell lanes).
    // Chunks that are not a full 256 particles are left to the per-thread kernel.
    namespace tma {
      __device__ __forceinline__ uint32_t saddr(const void* p) {
        return static_cast<uint32_t>(__cvta_generic_to_shared(p));
      }

      __device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count));
      }

      __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)),
                     "r"(bytes)
                     : "memory");
      }

      __device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "WAIT_%=:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                     "@p bra DONE_%=;\n"
                     "bra WAIT_%=;\n"
                     "DONE_%=:\n"
                     "}" ::"r"(saddr(bar)),
                     "r"(parity)
                     : "memory");
      }

      __device__ __forceinline__ void g2s(void* dst, const void* src, unsigned bytes,
                                          uint64_t* bar) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
                     "[%1], %2, [%3];" ::"r"(saddr(dst)),
                     "l"(src), "r"(bytes), "r"(saddr(bar))
                     : "memory");
      }

      __device__ __forceinline__ void s2g(void* dst, const void* src, unsigned bytes) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                     "r"(saddr(src)), "r"(bytes)
                     : "memory");
      }

      __device__ __forceinline__ void prefetch_l2(const void* src, unsigned bytes) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes)
                     : "memory");
      }

      __device__ __forceinline__ void commit() {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }

      __device__ __forceinline__ void wait_read_all() {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }

      __device__ __forceinline__ void wait_all() {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }

      __device__ __forceinline__ void fence_async_smem() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }

      __device__ __forceinline__ void fence_mbar_init() {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
    } // namespace tma

    constexpr int STREAM_CHUNK  = 256;
    constexpr int STREAM_STAGES = 3;

    template <int D>
    struct alignas(128) StreamIn {
      int   i[D][STREAM_CHUNK];
      float d[D][STREAM_CHUNK];
      float u[3][STREAM_CHUNK];
      float w[STREAM_CHUNK];
      short tag[STREAM_CHUNK];
    };

    template <int D>
    struct alignas(128) StreamOut {
      int   i[D][STREAM_CHUNK];
      float d[D][STREAM_CHUNK];
      float u[3][STREAM_CHUNK];
    };

    template <int D>
    struct StreamSmem {
      StreamIn<D>  in[STREAM_STAGES];
      StreamOut<D> out;
      uint64_t     full[STREAM_STAGES];
    };

    template <int D>
    __device__ __forceinline__ void stream_issue_load(StreamIn<D>& st, uint64_t* bar,
                                                      const eb200_prtls_t& S, size_t p0) {
      constexpr unsigned B4 = STREAM_CHUNK * 4, B2 = STREAM_CHUNK * 2;
      tma::mbar_expect_tx(bar, (2 * D + 4) * B4 + B2);
      const int*   ii[3] = { S.i1, S.i2, S.i3 };
      const float* dd[3] = { S.dx1, S.dx2, S.dx3 };
#pragma unroll
      for (int a = 0; a < D; ++a) {
        tma::g2s(st.i[a], ii[a] + p0, B4, bar);
        tma::g2s(st.d[a], dd[a] + p0, B4, bar);
      }
      tma::g2s(st.u[0], S.ux1 + p0, B4, bar);
      tma::g2s(st.u[1], S.ux2 + p0, B4, bar);
      tma::g2s(st.u[2], S.ux3 + p0, B4, bar);
      tma::g2s(st.w, S.weight + p0, B4, bar);
      tma::g2s(st.tag, S.tag + p0, B2, bar);
    }

#ifndef EB200_STREAM_MINBLOCKS
  #define EB200_STREAM_MINBLOCKS 3
#endif
    // resident CTAs per SM the register allocation aims for: the zig-zag body fits 3 x 256
    // threads; the Esirkepov windows need the registers more than the occupancy
    constexpr int stream_minblocks(int D, int O) {
      return (O == 0) ? EB200_STREAM_MINBLOCKS : ((D == 3 && O >= 2) ? 1 : 2);
    }

    template <int D, int O, bool LEAN>
    __global__ void __launch_bounds__(STREAM_CHUNK, stream_minblocks(D, O))
      push_deposit_stream_kernel(PushArgs A, eb200_prtls_t S, uint32_t nchunks, FieldView<D> EB,
                                 float charge, float inv_dt, FieldView<D> J) {
      extern __shared__ __align__(128) unsigned char smem_raw[];
      StreamSmem<D>& sm  = *reinterpret_cast<StreamSmem<D>*>(smem_raw);
      const int      tid = threadIdx.x;
      if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STREAM_STAGES; ++s) tma::mbar_init(&sm.full[s], 1);
        tma::fence_mbar_init();
      }
      __syncthreads();
      // chunks of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
      const uint32_t first = blockIdx.x, stride = gridDim.x;
      if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STREAM_STAGES - 1; ++s) {
          const uint32_t c = first + s * stride;
          if (c < nchunks) {
            stream_issue_load<D>(sm.in[s], &sm.full[s], S, (size_t)c * STREAM_CHUNK);
          }
        }
      }
      auto F = [&](int i, int j, int k, int c) { return EB.ld(i, j, k, c); };
      int      stage = 0;
      unsigned phase = 0;
      for (uint32_t c = first; c < nchunks; c += stride) {
        StreamIn<D>& st = sm.in[stage];
        tma::mbar_wait(&sm.full[stage], phase);
        const size_t p   = (size_t)c * STREAM_CHUNK + tid;
        const short  tag = st.tag[tid];
        Prtl<D>      P;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          P.i[a] = P.ip[a] = (a < D) ? st.i[a][tid] : 0;
          P.d[a] = P.dp[a] = (a < D) ? st.d[a][tid] : ZERO;
          P.u[a]           = st.u[a][tid];
        }
        P.w         = st.w[tid];
        P.tag       = tag;
        bool active = (tag == 1);
        if (active) {
          push_particle<D, O, decltype(F), LEAN>(A, F, P);
          if (P.tag != tag) {
            S.tag[p] = P.tag;
          }
        }
        deposit_particle_aggregated<D, O>(P, active && P.tag != 0, charge, inv_dt, A.c.dx, A.ng,
                                          J);
        // the previous chunk's bulk stores must have finished reading `out` (and the stage
        // that is refilled below) before anybody overwrites them
        if (tid == 0) {
          tma::wait_read_all();
        }
        const int not_all_alive = __syncthreads_or(!active);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (a < D) {
            sm.out.i[a][tid] = P.i[a];
            sm.out.d[a][tid] = P.d[a];
            if (P.ip[a] != st.i[a][tid]) {
              st.i[a][tid] = P.ip[a]; // periodic wrap shifts i_prev too (sr.hpp:664-677)
            }
          }
          sm.out.u[a][tid] = P.u[a];
        }
        if (not_all_alive && active) {
          // particles that are not pushed keep their i_prev / dx_prev: no bulk store of the
          // staged slices for this chunk, alive particles write theirs one by one
          int*   iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
          float* ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
          for (int a = 0; a < D; ++a) {
            iip[a][p] = P.ip[a];
            ddp[a][p] = P.dp[a];
          }
        }
        tma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          constexpr unsigned B4 = STREAM_CHUNK * 4;
          const size_t       p0 = (size_t)c * STREAM_CHUNK;
          int*   ii[3]  = { S.i1, S.i2, S.i3 };
          float* dd[3]  = { S.dx1, S.dx2, S.dx3 };
          int*   iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
          float* ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
          for (int a = 0; a < D; ++a) {
            tma::s2g(ii[a] + p0, sm.out.i[a], B4);
            tma::s2g(dd[a] + p0, sm.out.d[a], B4);
            if (!not_all_alive) {
              tma::s2g(iip[a] + p0, st.i[a], B4);
              tma::s2g(ddp[a] + p0, st.d[a], B4);
            }
          }
          tma::s2g(S.ux1 + p0, sm.out.u[0], B4);
          tma::s2g(S.ux2 + p0, sm.out.u[1], B4);
          tma::s2g(S.ux3 + p0, sm.out.u[2], B4);
          tma::commit();
          // refill the stage that was consumed one iteration ago (its stores were waited for
          // above) with the chunk two iterations ahead
          const uint32_t cn = c + (STREAM_STAGES - 1) * stride;
          if (cn < nchunks) {
            const int sn = (stage + STREAM_STAGES - 1) % STREAM_STAGES;
            stream_issue_load<D>(sm.in[sn], &sm.full[sn], S, (size_t)cn * STREAM_CHUNK);
          }
        }
        if (++stage == STREAM_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (tid == 0) {
        tma::wait_all();
      }
    }

    /* ------------- 3D third-order push + deposit with a shared-memory J tile (kernel 9) */
#if !EB200_STRICT
    // The per-lane deposit of a 5 x 5 x 5 Esirkepov window sends up to 375 reductions per
    // particle to L2 and is bound by their throughput (0.019 of the HBM roofline, r1w). Here a
    // CTA owns the J nodes around its chunk of 256 cell-sorted particles in shared memory:
    //  * the chunk travels global -> shared -> global with cp.async.bulk (TMA), like the
    //    stream kernel; thread t works on slot (9 t) mod 256, so the lanes of a warp are nine
    //    particles (about one cell) apart: their windows start on consecutive nodes, their
    //    shared-memory accesses fall into different banks and rarely on the same address;
    //  * the tile is FIXED POINT: fp32 atomicAdd on shared memory is a compare-and-swap loop
    //    on sm_100 (ATOMS.CAST.SPIN), integer ATOMS.ADD is one fire-and-forget instruction.
    //    A contribution x becomes round(x * scale) with one FFMA (the 1.5 * 2^23 trick) and one
    //    IADD; scale = 2^22 / max|w q / dt| over the chunk, so that |x * scale| < 2^21 and a
    //    node's sum over the 256 particles stays below 2^29. Resolution = 2.4e-7 of the
    //    largest single contribution, and the sum does not depend on the order of the adds;
    //  * the window is evaluated in its separable form (currents_deposit.hpp:618-752 factored):
    //    jx1(i, j, k) = -Q/3 CS1(i) F23(j, k), F23 = iS2 U3 + fS2 V3, U = iS + fS / 2,
    //    V = fS + iS / 2, CS = running sum of fS - iS (zero from the window's last node on,
    //    the reference's guard): two FFMA per (j, k) and FFMA + IADD + ATOMS per node;
    //  * one flush per chunk: the non-zero tile elements go to J as coalesced fp32 reductions
    //    (<= 56 per particle instead of 375). Particles whose window does not lie inside the
    //    tile (strays of a stale order, chunks that straddle a mesh row) take the per-lane
    //    global path, so the result is correct for any particle order.
    constexpr int T3X = 48, T3Y = 10, T3Z = 10, T3N = T3X * T3Y * T3Z;
    constexpr int T3_STAGES = 2;
    constexpr float T3_MAGIC = 12582912.0f; // 1.5 * 2^23
    constexpr int   T3_MAGIC_BITS = 0x4B400000;

    struct Tile3Smem {
      StreamIn<3>  in[T3_STAGES];
      StreamOut<3> out;
      uint64_t     full[T3_STAGES];
      float        wmax[8];
      int          box[4]; // touched tile rows of the current chunk: min / max of o2, min / max of o3
      int          tile[3 * T3N];
    };

    __device__ __forceinline__ void red_f32(float* p, float v) {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    }

    __device__ __forceinline__ void tile_add(int* p, float a, float f) {
      atomicAdd(p, __float_as_int(fmaf(a, f, T3_MAGIC)) - T3_MAGIC_BITS);
    }

    // the 5 x 5 x 5 window of one particle into the tile whose element (0, 0, 0) is node `org`
    __device__ __forceinline__ void deposit_esirkepov3_tile(const Prtl<3>& P, float qs, int G,
                                                            const int (&org)[3], int* tile,
                                                            bool& done, int& row2, int& row3) {
      constexpr int N = 5;
      float iS1[N], fS1[N], iS2[N], fS2[N], iS3[N], fS3[N];
      int   min1, max1, min2, max2, min3, max3;
      deposit_shapes<3>(P.ip[0], P.dp[0], P.i[0], P.d[0], min1, max1, iS1, fS1);
      deposit_shapes<3>(P.ip[1], P.dp[1], P.i[1], P.d[1], min2, max2, iS2, fS2);
      deposit_shapes<3>(P.ip[2], P.dp[2], P.i[2], P.d[2], min3, max3, iS3, fS3);
      const int o1 = min1 + G - org[0], o2 = min2 + G - org[1], o3 = min3 + G - org[2];
      done = (static_cast<unsigned>(o1) <= static_cast<unsigned>(T3X - N)) &&
             (static_cast<unsigned>(o2) <= static_cast<unsigned>(T3Y - N)) &&
             (static_cast<unsigned>(o3) <= static_cast<unsigned>(T3Z - N));
      if (!done) return;
      row2 = o2;
      row3 = o3;
      const int d1 = max1 - min1, d2 = max2 - min2, d3 = max3 - min3;
      float     A1[N - 1], A2[N - 1], A3[N - 1];
      {
        float c1 = ZERO, c2 = ZERO, c3 = ZERO;
#pragma unroll
        for (int n = 0; n < N - 1; ++n) {
          c1 += fS1[n] - iS1[n];
          c2 += fS2[n] - iS2[n];
          c3 += fS3[n] - iS3[n];
          A1[n] = (n < d1) ? qs * c1 : ZERO;
          A2[n] = (n < d2) ? qs * c2 : ZERO;
          A3[n] = (n < d3) ? qs * c3 : ZERO;
        }
      }
      float U2[N], V2[N], U3[N], V3[N];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        U2[n] = fmaf(HALF, fS2[n], iS2[n]);
        V2[n] = fmaf(HALF, iS2[n], fS2[n]);
        U3[n] = fmaf(HALF, fS3[n], iS3[n]);
        V3[n] = fmaf(HALF, iS3[n], fS3[n]);
      }
      int* t0 = tile + (o3 * T3Y + o2) * T3X + o1;
      // jx1
#pragma unroll
      for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float f = fmaf(iS2[j], U3[k], fS2[j] * V3[k]);
#pragma unroll
          for (int i = 0; i < N - 1; ++i) tile_add(t0 + (k * T3Y + j) * T3X + i, A1[i], f);
        }
      }
      // jx2
#pragma unroll
      for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float f = fmaf(iS1[i], U3[k], fS1[i] * V3[k]);
#pragma unroll
          for (int j = 0; j < N - 1; ++j) tile_add(t0 + T3N + (k * T3Y + j) * T3X + i, A2[j], f);
        }
      }
      // jx3
#pragma unroll
      for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float f = fmaf(iS1[i], U2[j], fS1[i] * V2[j]);
#pragma unroll
          for (int k = 0; k < N - 1; ++k) tile_add(t0 + 2 * T3N + (k * T3Y + j) * T3X + i, A3[k], f);
        }
      }
    }

    // the rare out-of-tile particle: per-lane reductions straight to global memory, kept out of
    // line so that its registers do not weigh on the tile path
    __device__ __noinline__ void deposit_esirkepov3_stray(const Prtl<3>& P, float charge,
                                                          float inv_dt, float dxc, int G,
                                                          const FieldView<3>& J) {
      deposit_esirkepov3_rows(P, charge, inv_dt, dxc, G, J);
    }

    template <bool LEAN>
    __global__ void __launch_bounds__(STREAM_CHUNK, 2)
      push_deposit_tile3_kernel(PushArgs A, eb200_prtls_t S, uint32_t nchunks, FieldView<3> EB,
                                float charge, float inv_dt, FieldView<3> J) {
      constexpr int D = 3;
      extern __shared__ __align__(128) unsigned char smem_raw[];
      Tile3Smem& sm  = *reinterpret_cast<Tile3Smem*>(smem_raw);
      const int  tid = threadIdx.x;
      for (int e = tid; e < 3 * T3N; e += STREAM_CHUNK) sm.tile[e] = 0;
      if (tid == 0) {
        sm.box[0] = T3Y, sm.box[1] = -1, sm.box[2] = T3Z, sm.box[3] = -1;
#pragma unroll
        for (int s = 0; s < T3_STAGES; ++s) tma::mbar_init(&sm.full[s], 1);
        tma::fence_mbar_init();
      }
      __syncthreads();
      const uint32_t first = blockIdx.x, stride = gridDim.x;
      if (tid == 0) {
#pragma unroll
        for (int s = 0; s < T3_STAGES - 1; ++s) {
          const uint32_t c = first + s * stride;
          if (c < nchunks) {
            stream_issue_load<D>(sm.in[s], &sm.full[s], S, (size_t)c * STREAM_CHUNK);
          }
        }
      }
      auto      F     = [&](int i, int j, int k, int c) { return EB.ld(i, j, k, c); };
      const int q     = (9 * tid) & (STREAM_CHUNK - 1); // this thread's slot in every chunk
      const int G     = A.ng;
      int       stage = 0;
      unsigned  phase = 0;
      for (uint32_t c = first; c < nchunks; c += stride) {
        StreamIn<D>& st = sm.in[stage];
        tma::mbar_wait(&sm.full[stage], phase);
        const size_t p   = (size_t)c * STREAM_CHUNK + q;
        const short  tag = st.tag[q];
        Prtl<D>      P;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          P.i[a] = P.ip[a] = st.i[a][q];
          P.d[a] = P.dp[a] = st.d[a][q];
          P.u[a]           = st.u[a][q];
        }
        P.w         = st.w[q];
        P.tag       = tag;
        bool active = (tag == 1);
        // tile origin: centred on the cell of the chunk's middle particle (ghost-inclusive node)
        const int org[3] = { st.i[0][STREAM_CHUNK / 2] + G - T3X / 2,
                             st.i[1][STREAM_CHUNK / 2] + G - 4, st.i[2][STREAM_CHUNK / 2] + G - 4 };
        // fixed-point scale from the chunk's largest |weight|
        float wm = active ? fabsf(P.w) : ZERO;
#pragma unroll
        for (int l = 16; l > 0; l >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, l));
        if ((tid & 31) == 0) sm.wmax[tid >> 5] = wm;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < STREAM_CHUNK / 32; ++w) wm = fmaxf(wm, sm.wmax[w]);
        const float qmax  = wm * fabsf(charge) * inv_dt;
        const float scale = (qmax > ZERO) ? 4194304.0f / qmax : ZERO;
        int         lo2 = T3Y, hi2 = -1, lo3 = T3Z, hi3 = -1; // tile rows this lane's window starts on
        if (active) {
          push_particle<D, 3, decltype(F), LEAN>(A, F, P);
          if (P.tag != tag) {
            S.tag[p] = P.tag;
          }
          if (P.tag != 0) {
            bool done = false;
            int  r2 = 0, r3 = 0;
            deposit_esirkepov3_tile(P, -THIRD * (P.w * charge * inv_dt) * scale, G, org, sm.tile,
                                    done, r2, r3);
            if (!done) {
              deposit_esirkepov3_stray(P, charge, inv_dt, A.c.dx, G, J);
            } else {
              lo2 = hi2 = r2;
              lo3 = hi3 = r3;
            }
          }
        }
        // the tile rows the chunk touched (a window is five rows tall): warp reductions, then
        // one shared-memory min / max per warp
        lo2 = __reduce_min_sync(0xffffffffu, lo2);
        hi2 = __reduce_max_sync(0xffffffffu, hi2);
        lo3 = __reduce_min_sync(0xffffffffu, lo3);
        hi3 = __reduce_max_sync(0xffffffffu, hi3);
        if ((tid & 31) == 0) {
          atomicMin(&sm.box[0], lo2);
          atomicMax(&sm.box[1], hi2);
          atomicMin(&sm.box[2], lo3);
          atomicMax(&sm.box[3], hi3);
        }
        if (tid == 0) {
          tma::wait_read_all();
        }
        const int not_all_alive = __syncthreads_or(!active); // also: every tile add has landed
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          sm.out.i[a][q] = P.i[a];
          sm.out.d[a][q] = P.d[a];
          if (P.ip[a] != st.i[a][q]) {
            st.i[a][q] = P.ip[a]; // periodic wrap shifts i_prev too (sr.hpp:664-677)
          }
          sm.out.u[a][q] = P.u[a];
        }
        if (not_all_alive && active) {
          int*   iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
          float* ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
          for (int a = 0; a < D; ++a) {
            iip[a][p] = P.ip[a];
            ddp[a][p] = P.dp[a];
          }
        }
        // flush: one warp per touched tile row (48 nodes along x1), non-zero elements only
        const int b2 = sm.box[0], e2 = sm.box[1] + 5, b3 = sm.box[2], e3 = sm.box[3] + 5;
        if (qmax > ZERO && e2 > b2 && e3 > b3) {
          const float inv_scale = qmax * (1.0f / 4194304.0f);
          const int   lane = tid & 31;
          const int   n2 = e2 - b2, nrow = 3 * n2 * (e3 - b3);
          const int   N1 = J.N1, N12 = J.N1 * J.N2, plane = (int)J.plane;
          float*      jbase = J.p + ((long)org[2] * J.N2 + org[1]) * J.N1 + org[0] + lane;
          for (int row = tid >> 5; row < nrow; row += STREAM_CHUNK / 32) {
            const int rz = row / n2, jy = b2 + (row - rz * n2);
            const int comp = rz / (e3 - b3), kz = b3 + (rz - comp * (e3 - b3));
            int*      trow = sm.tile + ((comp * T3Z + kz) * T3Y + jy) * T3X + lane;
            float*    jrow = jbase + (comp * plane + kz * N12 + jy * N1);
            const int v0 = trow[0];
            const int v1 = (lane < T3X - 32) ? trow[32] : 0;
            // (explicitly global: a generic atomicAdd carries the shared-memory CAS loop along)
            if (v0 != 0) {
              trow[0] = 0;
              red_f32(jrow, static_cast<float>(v0) * inv_scale);
            }
            if (v1 != 0) {
              trow[32] = 0;
              red_f32(jrow + 32, static_cast<float>(v1) * inv_scale);
            }
          }
        }
        tma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
          sm.box[0] = T3Y, sm.box[1] = -1, sm.box[2] = T3Z, sm.box[3] = -1;
        }
        if (tid == 0) {
          constexpr unsigned B4 = STREAM_CHUNK * 4;
          const size_t       p0 = (size_t)c * STREAM_CHUNK;
          int*   ii[3]  = { S.i1, S.i2, S.i3 };
          float* dd[3]  = { S.dx1, S.dx2, S.dx3 };
          int*   iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
          float* ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
          for (int a = 0; a < D; ++a) {
            tma::s2g(ii[a] + p0, sm.out.i[a], B4);
            tma::s2g(dd[a] + p0, sm.out.d[a], B4);
            if (!not_all_alive) {
              tma::s2g(iip[a] + p0, st.i[a], B4);
              tma::s2g(ddp[a] + p0, st.d[a], B4);
            }
          }
          tma::s2g(S.ux1 + p0, sm.out.u[0], B4);
          tma::s2g(S.ux2 + p0, sm.out.u[1], B4);
          tma::s2g(S.ux3 + p0, sm.out.u[2], B4);
          tma::commit();
          const uint32_t cn = c + (T3_STAGES - 1) * stride;
          if (cn < nchunks) {
            const int sn = (stage + T3_STAGES - 1) % T3_STAGES;
            stream_issue_load<D>(sm.in[sn], &sm.full[sn], S, (size_t)cn * STREAM_CHUNK);
          }
        }
        if (++stage == T3_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (tid == 0) {
        tma::wait_all();
      }
    }
#endif // !EB200_STRICT

    /* ------------------------------------ vectorised push + deposit (zig-zag, O = 0) */
    // Four consecutive particles per thread. Every SoA array is read and written with one
    // 128-bit streaming access per thread (a warp covers 512 contiguous bytes per array), so
    // the kernel needs no staging, no barriers and ~1/4 of the load/store instructions of the
    // one-particle-per-thread form. With cell-sorted particles the four particles of a thread
    // mostly sit in the same cell: their zig-zag contributions are summed in registers and only
    // a change of cell inside the thread issues atomics; what is left per thread goes through
    // the warp's segmented shuffle reduction once per FOUR particles.
    constexpr int VEC = 4;

    // streaming 128-bit load that does not allocate in L1 (the field nodes live there)
    __device__ __forceinline__ int4 ld_stream(const int4* p) {
      int4 t;
      asm volatile("ld.global.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                   : "l"(p));
      return t;
    }
    __device__ __forceinline__ float4 ld_stream(const float4* p) {
      float4 t;
      asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                   : "l"(p));
      return t;
    }
    __device__ __forceinline__ short4 ld_stream(const short4* p) { return __ldcs(p); }

    template <class T4, class T>
    __device__ __forceinline__ void ld4(const T* p, T (&v)[VEC]) {
      const T4 t = ld_stream(reinterpret_cast<const T4*>(p));
      v[0]       = t.x;
      v[1]       = t.y;
      v[2]       = t.z;
      v[3]       = t.w;
    }

    template <class T4, class T>
    __device__ __forceinline__ void st4(T* p, const T (&v)[VEC]) {
      T4 t;
      t.x = v[0];
      t.y = v[1];
      t.z = v[2];
      t.w = v[3];
      __stcs(reinterpret_cast<T4*>(p), t);
    }

#ifndef EB200_VEC_MINBLOCKS
  #define EB200_VEC_MINBLOCKS 3
#endif
    // resident CTAs per SM the register allocation aims for: 1D/2D bodies fit 3 x 256 threads
    // without spilling (80 registers), the 3D body (12 nodes per segment) needs 128
    constexpr int vec_minblocks(int D) { return (D == 3) ? 2 : EB200_VEC_MINBLOCKS; }

    // E/B component planes -> 24-byte nodes {Ex, By | Ey, Bx | Ez, Bz} (PackedEM2)
    __global__ void __launch_bounds__(256)
      pack_em2d_kernel(const float* __restrict__ em, long plane, float2* __restrict__ out) {
      const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
      if (n >= plane) return;
      const float e1 = __ldcs(em + n), e2 = __ldcs(em + plane + n), e3 = __ldcs(em + 2 * plane + n);
      const float b1 = __ldcs(em + 3 * plane + n), b2 = __ldcs(em + 4 * plane + n),
                  b3 = __ldcs(em + 5 * plane + n);
      out[3 * n]     = make_float2(e1, b2);
      out[3 * n + 1] = make_float2(e2, b1);
      out[3 * n + 2] = make_float2(e3, b3);
    }

    template <int D, bool LEAN, class EM = FieldView<D>>
    __global__ void __launch_bounds__(256, vec_minblocks(D))
      push_deposit_vec_kernel(PushArgs A, eb200_prtls_t S, uint32_t ngroups, uint32_t ahead,
                              EM EB, float charge, float inv_dt, FieldView<D> J) {
      constexpr int  NV       = ZigZag<D>::NV;
      const uint32_t g        = blockIdx.x * blockDim.x + threadIdx.x;
      const bool     in_range = g < ngroups;
      const size_t   p0       = (size_t)g * VEC;
      int*           ii[3]    = { S.i1, S.i2, S.i3 };
      float*         dd[3]    = { S.dx1, S.dx2, S.dx3 };
      int*           iip[3]   = { S.i1_prev, S.i2_prev, S.i3_prev };
      float*         ddp[3]   = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
      int            iv[3][VEC];
      float          dv[3][VEC], uv[3][VEC], wv[VEC];
      short          tv[VEC] = { 0, 0, 0, 0 };
      bool           all_pushed = false;
#ifndef EB200_VEC_NOPREFETCH
      // CTAs are dispatched in index order: pull the slices of the CTA that will run
      // `ahead` blocks later (about one resident wave) from DRAM into L2 now, so that its
      // first loads are L2 hits. One bulk prefetch per array, issued by one thread.
      if (threadIdx.x == 0) {
        const size_t q0 = ((size_t)blockIdx.x + ahead) * blockDim.x * VEC;
        if (q0 + (size_t)blockDim.x * VEC <= (size_t)ngroups * VEC) {
          const unsigned b4 = blockDim.x * VEC * 4, b2 = blockDim.x * VEC * 2;
#pragma unroll
          for (int a = 0; a < D; ++a) {
            tma::prefetch_l2(ii[a] + q0, b4);
            tma::prefetch_l2(dd[a] + q0, b4);
          }
          tma::prefetch_l2(S.ux1 + q0, b4);
          tma::prefetch_l2(S.ux2 + q0, b4);
          tma::prefetch_l2(S.ux3 + q0, b4);
          tma::prefetch_l2(S.weight + q0, b4);
          tma::prefetch_l2(S.tag + q0, b2);
        }
      }
#endif
      if (in_range) {
        ld4<short4>(S.tag + p0, tv);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          ld4<int4>(ii[a] + p0, iv[a]);
          ld4<float4>(dd[a] + p0, dv[a]);
        }
        ld4<float4>(S.ux1 + p0, uv[0]);
        ld4<float4>(S.ux2 + p0, uv[1]);
        ld4<float4>(S.ux3 + p0, uv[2]);
        ld4<float4>(S.weight + p0, wv);
        all_pushed = (tv[0] == 1) && (tv[1] == 1) && (tv[2] == 1) && (tv[3] == 1);
        if (all_pushed) {
          // i_prev / dx_prev of a pushed particle ARE its old i / dx: stored right away; a
          // periodic wrap (which shifts i_prev too, sr.hpp:664-677) patches its element below
#pragma unroll
          for (int a = 0; a < D; ++a) {
            st4<int4>(iip[a] + p0, iv[a]);
            st4<float4>(ddp[a] + p0, dv[a]);
          }
        }
      }
      const long N12 = (long)J.N1 * J.N2;
      auto       red = [&](int key, const float (&a)[NV]) {
#pragma unroll
        for (int n = 0; n < NV; ++n) {
#ifdef EB200_X_NOATOM
          if (a[n] == 1.2345e-30f)
#endif
          atomicAdd(J.p + key + zigzag_offset<D>(n, J.N1, N12, J.plane), a[n]);
        }
      };
      float acc[NV];
#pragma unroll
      for (int n = 0; n < NV; ++n) acc[n] = ZERO;
      int cur = -1; // cell whose contributions `acc` holds
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const short tag = tv[k];
        if (tag != 1) {
          continue; // neither pushed nor deposited (tv = 0 for threads past the end)
        }
        Prtl<D> P;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          P.i[a] = P.ip[a] = (a < D) ? iv[a][k] : 0;
          P.d[a] = P.dp[a] = (a < D) ? dv[a][k] : ZERO;
          P.u[a]           = uv[a][k];
        }
        P.w   = wv[k];
        P.tag = tag;
        push_particle<D, 0, EM, LEAN>(A, EB, P);
        if (P.tag != tag) {
          S.tag[p0 + k] = P.tag;
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
          if (!all_pushed || P.ip[a] != iv[a][k]) {
            iip[a][p0 + k] = P.ip[a];
          }
          if (!all_pushed) {
            ddp[a][p0 + k] = P.dp[a];
          }
          iv[a][k] = P.i[a];
          dv[a][k] = P.d[a];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) uv[a][k] = P.u[a];
        if (P.tag == 0) {
          continue; // absorbed by a boundary: no current
        }
#ifdef EB200_X_NODEPOSIT
        if (P.u[0] != 1.2345e-30f) continue;
#endif
        float v[2][NV];
        zigzag_values<D>(P, charge, inv_dt, A.c.dx, v);
        const int G    = A.ng;
        const int key0 = (int)J.idx(P.ip[0] + G, (D > 1) ? P.ip[1] + G : 0,
                                    (D > 2) ? P.ip[2] + G : 0);
        const int key1 = (int)J.idx(P.i[0] + G, (D > 1) ? P.i[1] + G : 0,
                                    (D > 2) ? P.i[2] + G : 0);
        const bool cross = key0 != key1;
        if (cross) {
          red(key1, v[1]);
        } else {
#pragma unroll
          for (int n = 0; n < NV; ++n) v[0][n] += v[1][n];
        }
        if (key0 != cur) {
          if (cur >= 0) {
            red(cur, acc);
          }
          cur = key0;
#pragma unroll
          for (int n = 0; n < NV; ++n) acc[n] = v[0][n];
        } else {
#pragma unroll
          for (int n = 0; n < NV; ++n) acc[n] += v[0][n];
        }
      }
      if (in_range) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
          st4<int4>(ii[a] + p0, iv[a]);
          st4<float4>(dd[a] + p0, dv[a]);
        }
        st4<float4>(S.ux1 + p0, uv[0]);
        st4<float4>(S.ux2 + p0, uv[1]);
        st4<float4>(S.ux3 + p0, uv[2]);
      }
      // what is left in `acc`: one segmented reduction over the warp, keyed by `cur`
      const WarpRun run = warp_runs(cur);
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float s = run_sum(acc[n], run);
#ifdef EB200_X_NOATOM
        if (run.head && cur >= 0 && s == 1.2345e-30f) {
#else
        if (run.head && cur >= 0) {
#endif
          atomicAdd(J.p + cur + zigzag_offset<D>(n, J.N1, N12, J.plane), s);
        }
      }
    }

    /* ------------- moment-accumulating push + deposit (2D zig-zag, packed nodes, plain Boris) */
#if !EB200_STRICT
    // Kernel 8: the organisation of push_deposit_vec_kernel (four consecutive particles per
    // thread, 128-bit streaming accesses, packed E/B nodes, L2 prefetch of the next wave) with
    // the per-particle instruction count cut where the r1p capture put it:
    //  * the zig-zag deposit is linear in eight MOMENTS of a segment -- (A, A wy, B, B wx, F,
    //    F wx, F wy, F wx wy) with A / B the in-plane charge fluxes, F the out-of-plane one and
    //    (wx, wy) the segment's midpoint -- from which the eight node values of
    //    currents_deposit.hpp:171-405 follow by seven additions. Moments are what is summed in
    //    registers and in the warp's segmented reduction; the conversion runs once per flush
    //    instead of once per particle and segment. A particle that stays in its cell (97 % of
    //    the cold background) contributes ONE merged segment (midpoint of the move + the
    //    dx dy / 16 cross term of the two half segments); only crossers take the two-segment
    //    form, under a branch;
    //  * 1 / gamma of the position push is reused for the out-of-plane velocity;
    //  * the staggered gather weights come from a compare + select instead of float -> int ->
    //    float conversions (XU pipe), same values;
    //  * i_prev is re-stored only for the rare particle that went through the boundary block.
    // Fast build only: the rounding differs from the reference's operation order at the 1e-7
    // level (same tolerances as kernel 5 in tests/).
    struct Mom2 {
      float m[8];
    };

    __device__ __forceinline__ void mom_flush(const FieldView<2>& J, int key, const float (&m)[8]) {
      float* jx = J.p + key;
      float* jy = jx + J.plane;
      float* jz = jy + J.plane;
      const int N1 = J.N1;
      atomicAdd(jx, m[0] - m[1]);
      atomicAdd(jx + N1, m[1]);
      atomicAdd(jy, m[2] - m[3]);
      atomicAdd(jy + 1, m[3]);
      atomicAdd(jz, (m[4] - m[5]) - (m[6] - m[7]));
      atomicAdd(jz + 1, m[5] - m[7]);
      atomicAdd(jz + N1, m[6] - m[7]);
      atomicAdd(jz + N1 + 1, m[7]);
    }

    // The same flush into a context-owned array of 16-byte nodes {jx1, jx2, jx3, -}: the three
    // components of a node and of its x1 neighbour share a 32-byte sector, so a flush touches
    // 2-4 sectors with three 128-bit reductions + one scalar instead of 8 scalar reductions on
    // 5 sectors of the component planes (stale-order launches spend their extra time on exactly
    // those sectors, r2c capture). unpack_j4_kernel adds the nodes to the planes afterwards.
    __device__ __forceinline__ void red_v4(float4* p, float a, float b, float c) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                   "f"(0.0f)
                   : "memory");
    }

    __device__ __forceinline__ void mom_flush_v4(float4* J4, int N1, int key, const float (&m)[8]) {
      float4* n00 = J4 + key;
      red_v4(n00, m[0] - m[1], m[2] - m[3], (m[4] - m[5]) - (m[6] - m[7]));
      red_v4(n00 + 1, ZERO, m[3], m[5] - m[7]);
      red_v4(n00 + N1, m[1], ZERO, m[6] - m[7]);
      atomicAdd(&(n00 + N1 + 1)->z, m[7]);
    }

    __global__ void __launch_bounds__(256)
      unpack_j4_kernel(float4* __restrict__ J4, long plane, float* __restrict__ cur) {
      const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
      if (n >= plane) return;
      const float4 v = J4[n];
      if (v.x != ZERO || v.y != ZERO || v.z != ZERO) {
        cur[n] += v.x;
        cur[plane + n] += v.y;
        cur[2 * plane + n] += v.z;
        J4[n] = make_float4(ZERO, ZERO, ZERO, ZERO);
      }
    }

    // staggered / primal bilinear weights without conversions; same values as gather_packed()
    __device__ __forceinline__ void gather_packed_sel(const PackedEM2& F, int ng, const int (&i)[2],
                                                      const float (&d)[2], float* e0, float* b0) {
      float          wp[2][2], wd[2][2];
      unsigned       back[2];
      const unsigned st[2] = { 24u, F.rowb };
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const float h  = d[a] + HALF;
        const bool  up = h >= ONE; // == static_cast<int>(d + 1/2) for d in [0, 1)
        back[a]        = up ? 0u : st[a];
        wp[a][0]       = ONE - d[a];
        wp[a][1]       = d[a];
        wd[a][0]       = (up ? TWO : ONE) - h;
        wd[a][1]       = ONE - wd[a][0];
      }
      const unsigned b00 = static_cast<unsigned>(i[0] + ng) * 24u +
                           static_cast<unsigned>(i[1] + ng) * F.rowb;
      const unsigned b10 = b00 - back[0], b01 = b00 - back[1], b11 = b10 - back[1];
      {
        const char*  q  = F.p + b10;
        const char*  r  = F.p + (b10 + F.rowb);
        const float2 q0 = ld_keep2(q), q1 = ld_keep2(q + 24), r0 = ld_keep2(r), r1 = ld_keep2(r + 24);
        const float* wx = wd[0];
        const float* wy = wp[1];
        e0[0] = (q0.x * wx[0] + q1.x * wx[1]) * wy[0] + (r0.x * wx[0] + r1.x * wx[1]) * wy[1];
        b0[1] = (q0.y * wx[0] + q1.y * wx[1]) * wy[0] + (r0.y * wx[0] + r1.y * wx[1]) * wy[1];
      }
      {
        const char*  q  = F.p + (b01 + 8u);
        const char*  r  = F.p + (b01 + 8u + F.rowb);
        const float2 q0 = ld_keep2(q), q1 = ld_keep2(q + 24), r0 = ld_keep2(r), r1 = ld_keep2(r + 24);
        const float* wx = wp[0];
        const float* wy = wd[1];
        e0[1] = (q0.x * wx[0] + q1.x * wx[1]) * wy[0] + (r0.x * wx[0] + r1.x * wx[1]) * wy[1];
        b0[0] = (q0.y * wx[0] + q1.y * wx[1]) * wy[0] + (r0.y * wx[0] + r1.y * wx[1]) * wy[1];
      }
      {
        const char*  q  = F.p + (b00 + 16u);
        const char*  r  = F.p + (b00 + 16u + F.rowb);
        const float* wx = wp[0];
        const float* wy = wp[1];
        e0[2] = (ld_keep1(q) * wx[0] + ld_keep1(q + 24) * wx[1]) * wy[0] +
                (ld_keep1(r) * wx[0] + ld_keep1(r + 24) * wx[1]) * wy[1];
      }
      {
        const char*  q  = F.p + (b11 + 20u);
        const char*  r  = F.p + (b11 + 20u + F.rowb);
        const float* wx = wd[0];
        const float* wy = wd[1];
        b0[2] = (ld_keep1(q) * wx[0] + ld_keep1(q + 24) * wx[1]) * wy[0] +
                (ld_keep1(r) * wx[0] + ld_keep1(r + 24) * wx[1]) * wy[1];
      }
    }

  #ifndef EB200_MOM_MINBLOCKS
    #define EB200_MOM_MINBLOCKS 3
  #endif

    // KEEP_PREV = false (eb200_set_lean_prev): i*_prev / dx*_prev are not stored. Nothing on the
    // path reads them after this kernel (the next push overwrites them before any read,
    // sr.hpp:137-153): 62 instead of 78 bytes per particle.
    template <bool V4, bool KEEP_PREV = true>
    __global__ void __launch_bounds__(256, EB200_MOM_MINBLOCKS)
      push_deposit_mom_kernel(PushArgs A, eb200_prtls_t S, uint32_t ngroups, uint32_t ahead,
                              PackedEM2 EB, float charge, float inv_dt, FieldView<2> J,
                              float4* J4) {
      auto flush = [&](int key, const float (&m)[8]) {
        if constexpr (V4) {
          mom_flush_v4(J4, J.N1, key, m);
        } else {
          mom_flush(J, key, m);
        }
      };
      constexpr int  D        = 2;
      const uint32_t g        = blockIdx.x * blockDim.x + threadIdx.x;
      const bool     in_range = g < ngroups;
      const size_t   p0       = (size_t)g * VEC;
      int*           ii[2]    = { S.i1, S.i2 };
      float*         dd[2]    = { S.dx1, S.dx2 };
      int*           iip[2]   = { S.i1_prev, S.i2_prev };
      float*         ddp[2]   = { S.dx1_prev, S.dx2_prev };
      int            iv[2][VEC];
      float          dv[2][VEC], uv[3][VEC], wv[VEC];
      short          tv[VEC] = { 0, 0, 0, 0 };
      bool           all_pushed = false;
  #ifndef EB200_VEC_NOPREFETCH
      if (threadIdx.x == 0) {
        const size_t q0 = ((size_t)blockIdx.x + ahead) * blockDim.x * VEC;
        if (q0 + (size_t)blockDim.x * VEC <= (size_t)ngroups * VEC) {
          const unsigned b4 = blockDim.x * VEC * 4, b2 = blockDim.x * VEC * 2;
  #pragma unroll
          for (int a = 0; a < D; ++a) {
            tma::prefetch_l2(ii[a] + q0, b4);
            tma::prefetch_l2(dd[a] + q0, b4);
          }
          tma::prefetch_l2(S.ux1 + q0, b4);
          tma::prefetch_l2(S.ux2 + q0, b4);
          tma::prefetch_l2(S.ux3 + q0, b4);
          tma::prefetch_l2(S.weight + q0, b4);
          tma::prefetch_l2(S.tag + q0, b2);
        }
      }
  #endif
      if (in_range) {
        ld4<short4>(S.tag + p0, tv);
  #pragma unroll
        for (int a = 0; a < D; ++a) {
          ld4<int4>(ii[a] + p0, iv[a]);
          ld4<float4>(dd[a] + p0, dv[a]);
        }
        ld4<float4>(S.ux1 + p0, uv[0]);
        ld4<float4>(S.ux2 + p0, uv[1]);
        ld4<float4>(S.ux3 + p0, uv[2]);
        ld4<float4>(S.weight + p0, wv);
        all_pushed = (tv[0] == 1) && (tv[1] == 1) && (tv[2] == 1) && (tv[3] == 1);
        if constexpr (KEEP_PREV) {
          if (all_pushed) {
  #pragma unroll
            for (int a = 0; a < D; ++a) {
              st4<int4>(iip[a] + p0, iv[a]);
              st4<float4>(ddp[a] + p0, dv[a]);
            }
          }
        }
      }
      const int   G    = A.ng;
      const float cdx  = A.c.dx;
      const float cpos = A.c.dt * A.inv_dx; // dt / dx
      float       acc[8];
  #pragma unroll
      for (int n = 0; n < 8; ++n) acc[n] = ZERO;
      int cur = -1; // cell whose moments `acc` holds
  #pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (tv[k] != 1) {
          if (in_range) exc_append(A, (uint32_t)(p0 + k));
          continue;
        }
        int   ip[2] = { iv[0][k], iv[1][k] };
        float dp[2] = { dv[0][k], dv[1][k] };
        float u[3]  = { uv[0][k], uv[1][k], uv[2][k] };
        float ec[3], bc[3];
        gather_packed_sel(EB, G, ip, dp, ec, bc);
        ec[0] *= cdx;
        ec[1] *= cdx;
        bc[0] *= cdx;
        bc[1] *= cdx;
        boris(A.ndh, u, ec, bc);
        const float ig = rsqrtf(ONE + nsq(u)); // 1 / gamma
        const float cs = cpos * ig;
        int         in[2];
        float       dn[2];
  #pragma unroll
        for (int a = 0; a < D; ++a) {
          const float x    = fmaf(u[a], cs, dp[a]);
          const bool  up   = x >= ONE;
          const bool  down = x < ZERO;
          in[a]            = ip[a] + (up ? 1 : 0) - (down ? 1 : 0);
          dn[a]            = up ? (x - ONE) : (down ? (x + ONE) : x);
        }
        short tag = 1;
        if ((static_cast<unsigned>(in[0]) >= static_cast<unsigned>(A.ni[0])) ||
            (static_cast<unsigned>(in[1]) >= static_cast<unsigned>(A.ni[1]))) {
          Prtl<2> P;
          P.i[0] = in[0], P.i[1] = in[1], P.i[2] = 0;
          P.ip[0] = ip[0], P.ip[1] = ip[1], P.ip[2] = 0;
          P.d[0] = dn[0], P.d[1] = dn[1], P.d[2] = ZERO;
          P.dp[0] = dp[0], P.dp[1] = dp[1], P.dp[2] = ZERO;
          P.u[0] = u[0], P.u[1] = u[1], P.u[2] = u[2];
          P.w   = wv[k];
          P.tag = 1;
          particle_boundaries<2>(A, P);
          in[0] = P.i[0], in[1] = P.i[1];
          dn[0] = P.d[0], dn[1] = P.d[1];
          u[0] = P.u[0], u[1] = P.u[1], u[2] = P.u[2];
          tag = P.tag;
          if (tag != 1) {
            S.tag[p0 + k] = tag;
            exc_append(A, (uint32_t)(p0 + k));
          }
          if constexpr (KEEP_PREV) {
            if (all_pushed) {
              // a periodic wrap shifted i_prev with i (sr.hpp:664-677)
  #pragma unroll
              for (int a = 0; a < D; ++a) {
                if (P.ip[a] != ip[a]) iip[a][p0 + k] = P.ip[a];
              }
            }
          }
          ip[0] = P.ip[0], ip[1] = P.ip[1];
        }
        if constexpr (KEEP_PREV) {
          if (!all_pushed) {
  #pragma unroll
            for (int a = 0; a < D; ++a) {
              iip[a][p0 + k] = ip[a];
              ddp[a][p0 + k] = dp[a];
            }
          }
        }
  #pragma unroll
        for (int a = 0; a < D; ++a) {
          iv[a][k] = in[a];
          dv[a][k] = dn[a];
        }
  #pragma unroll
        for (int a = 0; a < 3; ++a) uv[a][k] = u[a];
        if (tag == 0) {
          continue; // absorbed by a boundary: no current
        }
        // ---- deposit
        const float coeff = wv[k] * charge;
        const float Q     = coeff * inv_dt;
        float       vz    = u[2];
        if (!(fabsf(vz) <= 3.4028235e38f)) vz = ZERO; // nan / inf guard of currents_deposit.hpp
        const float Fz  = coeff * (vz * ig);
        const int   di0 = in[0] - ip[0], di1 = in[1] - ip[1];
        const int   key0 = (ip[0] + G) + (ip[1] + G) * J.N1;
        float       m[8];
        if ((di0 | di1) == 0) {
          // one merged segment: midpoint weights + the cross term of the two half segments
          const float mx = HALF * (dn[0] + dp[0]), my = HALF * (dn[1] + dp[1]);
          const float lx = dn[0] - dp[0], ly = dn[1] - dp[1];
          const float Ax = Q * lx, By = Q * ly;
          m[0] = Ax;
          m[1] = Ax * my;
          m[2] = By;
          m[3] = By * mx;
          m[4] = Fz;
          m[5] = Fz * mx;
          m[6] = Fz * my;
          m[7] = Fz * fmaf(mx, my, INV_16 * lx * ly);
        } else {
          // relay point seen from the old / new cell; two segments, the second in the new cell
          const float r0x = (di0 == 0) ? HALF * (dn[0] + dp[0]) : ((di0 > 0) ? ONE : ZERO);
          const float r1x = (di0 == 0) ? r0x : ((di0 > 0) ? ZERO : ONE);
          const float r0y = (di1 == 0) ? HALF * (dn[1] + dp[1]) : ((di1 > 0) ? ONE : ZERO);
          const float r1y = (di1 == 0) ? r0y : ((di1 > 0) ? ZERO : ONE);
          const float Fh  = HALF * Fz;
          {
            const float wx = HALF * (dn[0] + r1x), wy = HALF * (dn[1] + r1y);
            const float Ax = (dn[0] - r1x) * Q, By = (dn[1] - r1y) * Q;
            const float s[8] = { Ax, Ax * wy, By, By * wx, Fh, Fh * wx, Fh * wy, Fh * wx * wy };
            flush((in[0] + G) + (in[1] + G) * J.N1, s);
          }
          const float wx = HALF * (r0x + dp[0]), wy = HALF * (r0y + dp[1]);
          const float Ax = (r0x - dp[0]) * Q, By = (r0y - dp[1]) * Q;
          m[0] = Ax;
          m[1] = Ax * wy;
          m[2] = By;
          m[3] = By * wx;
          m[4] = Fh;
          m[5] = Fh * wx;
          m[6] = Fh * wy;
          m[7] = Fh * wx * wy;
        }
        if (key0 != cur) {
          if (cur >= 0) {
            flush(cur, acc);
          }
          cur = key0;
  #pragma unroll
          for (int n = 0; n < 8; ++n) acc[n] = m[n];
        } else {
  #pragma unroll
          for (int n = 0; n < 8; ++n) acc[n] += m[n];
        }
      }
      if (in_range) {
  #pragma unroll
        for (int a = 0; a < D; ++a) {
          st4<int4>(ii[a] + p0, iv[a]);
          st4<float4>(dd[a] + p0, dv[a]);
        }
        st4<float4>(S.ux1 + p0, uv[0]);
        st4<float4>(S.ux2 + p0, uv[1]);
        st4<float4>(S.ux3 + p0, uv[2]);
      }
      // what is left in `acc`: one segmented reduction of the moments over the warp
      const WarpRun run = warp_runs(cur);
  #pragma unroll
      for (int n = 0; n < 8; ++n) acc[n] = run_sum(acc[n], run);
      if (run.head && cur >= 0) {
        flush(cur, acc);
      }
    }
#endif // !EB200_STRICT

    /* ------------- pipelined push + deposit (2D zig-zag, packed nodes, persistent CTAs) */
    // The vectorised kernel spends its stall cycles in two dependent waits per thread: the
    // particle slices (DRAM/L2 latency) and then the first particle's E/B nodes (an L1 miss:
    // every CTA works on cells nobody on the SM has touched). Here each CTA owns a CONTIGUOUS
    // range of 1024-particle tiles and walks it: the slices of the next tile are copied
    // global -> shared with per-thread cp.async (16 bytes per array and thread, thread-private
    // slots: no barrier, no register cost) while the current tile is computed, and the E/B
    // lines the next tile will gather from (one tile further along the mesh row for
    // cell-sorted particles) are pulled into L1 with prefetch instructions. Arithmetic, stores
    // and the deposit are those of push_deposit_vec_kernel.
    __device__ __forceinline__ void cp_async16(void* smem, const void* g) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tma::saddr(smem)), "l"(g)
                   : "memory");
    }
    __device__ __forceinline__ void cp_async8(void* smem, const void* g) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tma::saddr(smem)), "l"(g)
                   : "memory");
    }
    __device__ __forceinline__ void cp_async_commit() {
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void cp_async_wait_all() {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __device__ __forceinline__ void prefetch_l1(const void* p) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    }

#ifndef EB200_PIPE_MINBLOCKS
  #define EB200_PIPE_MINBLOCKS 3
#endif
    constexpr int PIPE_THREADS = 256;

    template <bool LEAN>
    __global__ void __launch_bounds__(PIPE_THREADS, EB200_PIPE_MINBLOCKS)
      push_deposit_pipe_kernel(PushArgs A, eb200_prtls_t S, uint32_t ngroups,
                               uint32_t tiles_per_cta, uint32_t tile_stride, unsigned ahead_bytes,
                               unsigned em_bytes,
                               PackedEM2 EB, float charge, float inv_dt, FieldView<2> J) {
      constexpr int D  = 2;
      constexpr int NV = ZigZag<D>::NV;
      __shared__ __align__(16) int4   s_i[D][PIPE_THREADS];
      __shared__ __align__(16) float4 s_d[D][PIPE_THREADS];
      __shared__ __align__(16) float4 s_u[3][PIPE_THREADS];
      __shared__ __align__(16) float4 s_w[PIPE_THREADS];
      __shared__ __align__(8) short4  s_t[PIPE_THREADS];
      const int      tid     = threadIdx.x;
      const uint32_t ntiles  = (ngroups + PIPE_THREADS - 1) / PIPE_THREADS;
      // tile_stride == 1: this CTA owns tiles [b * per, (b + 1) * per); otherwise tiles b,
      // b + stride, ... (all resident CTAs then sweep one contiguous window of the arrays)
      const uint32_t t_begin = (tile_stride == 1) ? blockIdx.x * tiles_per_cta : blockIdx.x;
      const uint32_t t_end   = (tile_stride == 1) ? min(ntiles, t_begin + tiles_per_cta) : ntiles;
      if (t_begin >= t_end) return;
      int*   ii[3]  = { S.i1, S.i2, S.i3 };
      float* dd[3]  = { S.dx1, S.dx2, S.dx3 };
      int*   iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
      float* ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
      auto issue = [&](uint32_t tile) {
        const uint32_t gn = tile * PIPE_THREADS + tid;
        if (gn < ngroups) {
          const size_t q0 = (size_t)gn * VEC;
#pragma unroll
          for (int a = 0; a < D; ++a) {
            cp_async16(&s_i[a][tid], ii[a] + q0);
            cp_async16(&s_d[a][tid], dd[a] + q0);
          }
          cp_async16(&s_u[0][tid], S.ux1 + q0);
          cp_async16(&s_u[1][tid], S.ux2 + q0);
          cp_async16(&s_u[2][tid], S.ux3 + q0);
          cp_async16(&s_w[tid], S.weight + q0);
          cp_async8(&s_t[tid], S.tag + q0);
        }
        cp_async_commit();
      };
      issue(t_begin);
      const long N12 = (long)J.N1 * J.N2;
      auto       red = [&](int key, const float (&a)[NV]) {
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          atomicAdd(J.p + key + zigzag_offset<D>(n, J.N1, N12, J.plane), a[n]);
        }
      };
      for (uint32_t tile = t_begin; tile < t_end; tile += tile_stride) {
        const uint32_t g        = tile * PIPE_THREADS + tid;
        const bool     in_range = g < ngroups;
        const size_t   p0       = (size_t)g * VEC;
        int            iv[D][VEC];
        float          dv[D][VEC], uv[3][VEC], wv[VEC];
        short          tv[VEC]    = { 0, 0, 0, 0 };
        bool           all_pushed = false;
        cp_async_wait_all();
        if (in_range) {
#pragma unroll
          for (int a = 0; a < D; ++a) {
            const int4   t4 = s_i[a][tid];
            const float4 f4 = s_d[a][tid];
            iv[a][0] = t4.x, iv[a][1] = t4.y, iv[a][2] = t4.z, iv[a][3] = t4.w;
            dv[a][0] = f4.x, dv[a][1] = f4.y, dv[a][2] = f4.z, dv[a][3] = f4.w;
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float4 f4 = s_u[a][tid];
            uv[a][0] = f4.x, uv[a][1] = f4.y, uv[a][2] = f4.z, uv[a][3] = f4.w;
          }
          {
            const float4 f4 = s_w[tid];
            wv[0] = f4.x, wv[1] = f4.y, wv[2] = f4.z, wv[3] = f4.w;
            const short4 h4 = s_t[tid];
            tv[0] = h4.x, tv[1] = h4.y, tv[2] = h4.z, tv[3] = h4.w;
          }
          // every staged vector has been read (one element of each is consumed here) before
          // the slots are handed to the next tile's copies
          float chk = uv[0][0] + uv[1][0] + uv[2][0] + wv[0] + dv[0][0] + dv[1][0];
          int   chi = iv[0][0] + iv[1][0] + tv[0];
          asm volatile("" ::"f"(chk), "r"(chi) : "memory");
        }
        if (tile + tile_stride < t_end) {
          issue(tile + tile_stride);
          if (in_range && ahead_bytes) {
            // E/B lines of the next tile: the rows around this thread's first particle, one
            // tile further along the row
            const unsigned b = static_cast<unsigned>(iv[0][0] + A.ng) * 24u +
                               static_cast<unsigned>(iv[1][0] + A.ng) * EB.rowb + ahead_bytes;
            if (b + EB.rowb + 128u < em_bytes && b >= EB.rowb) {
              prefetch_l1(EB.p + (b - EB.rowb));
              prefetch_l1(EB.p + b);
              prefetch_l1(EB.p + (b + EB.rowb));
            }
          }
        }
        if (in_range) {
          all_pushed = (tv[0] == 1) && (tv[1] == 1) && (tv[2] == 1) && (tv[3] == 1);
          if (all_pushed) {
#pragma unroll
            for (int a = 0; a < D; ++a) {
              st4<int4>(iip[a] + p0, iv[a]);
              st4<float4>(ddp[a] + p0, dv[a]);
            }
          }
        }
        float acc[NV];
#pragma unroll
        for (int n = 0; n < NV; ++n) acc[n] = ZERO;
        int cur = -1;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const short tag = tv[k];
          if (tag != 1) {
            continue;
          }
          Prtl<D> P;
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            P.i[a] = P.ip[a] = (a < D) ? iv[a][k] : 0;
            P.d[a] = P.dp[a] = (a < D) ? dv[a][k] : ZERO;
            P.u[a]           = uv[a][k];
          }
          P.w   = wv[k];
          P.tag = tag;
          push_particle<D, 0, PackedEM2, LEAN>(A, EB, P);
          if (P.tag != tag) {
            S.tag[p0 + k] = P.tag;
          }
#pragma unroll
          for (int a = 0; a < D; ++a) {
            if (!all_pushed || P.ip[a] != iv[a][k]) {
              iip[a][p0 + k] = P.ip[a];
            }
            if (!all_pushed) {
              ddp[a][p0 + k] = P.dp[a];
            }
            iv[a][k] = P.i[a];
            dv[a][k] = P.d[a];
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) uv[a][k] = P.u[a];
          if (P.tag == 0) {
            continue;
          }
          float v[2][NV];
          zigzag_values<D>(P, charge, inv_dt, A.c.dx, v);
          const int  G     = A.ng;
          const int  key0  = (int)J.idx(P.ip[0] + G, P.ip[1] + G, 0);
          const int  key1  = (int)J.idx(P.i[0] + G, P.i[1] + G, 0);
          const bool cross = key0 != key1;
          if (cross) {
            red(key1, v[1]);
          } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) v[0][n] += v[1][n];
          }
          if (key0 != cur) {
            if (cur >= 0) {
              red(cur, acc);
            }
            cur = key0;
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[n] = v[0][n];
          } else {
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[n] += v[0][n];
          }
        }
        if (in_range) {
#pragma unroll
          for (int a = 0; a < D; ++a) {
            st4<int4>(ii[a] + p0, iv[a]);
            st4<float4>(dd[a] + p0, dv[a]);
          }
          st4<float4>(S.ux1 + p0, uv[0]);
          st4<float4>(S.ux2 + p0, uv[1]);
          st4<float4>(S.ux3 + p0, uv[2]);
        }
        const WarpRun run = warp_runs(cur);
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          const float sm = run_sum(acc[n], run);
          if (run.head && cur >= 0) {
            atomicAdd(J.p + cur + zigzag_offset<D>(n, J.N1, N12, J.plane), sm);
          }
        }
      }
    }

    /* ------------- shared-memory resident push + deposit (2D zig-zag, packed nodes) */
    // The vectorised kernel holds the four particles of a thread in registers (44 of its 80),
    // which caps the SM at 24 warps; it is latency bound (1.1 eligible warps per scheduler).
    // Here the slices are parked in shared memory, transposed to [k][thread] so that the
    // per-particle 32-bit accesses are conflict free, the particle loop is NOT unrolled (one
    // copy of the push/deposit body: a quarter of the code) and the registers only hold one
    // particle + the cell accumulators: more resident warps for the same work. Global accesses
    // stay 128-bit and coalesced; arithmetic and deposit are those of push_deposit_vec_kernel.
#ifndef EB200_SMEM_MINBLOCKS
  #define EB200_SMEM_MINBLOCKS 4
#endif
    template <bool LEAN>
    __global__ void __launch_bounds__(256, EB200_SMEM_MINBLOCKS)
      push_deposit_smem_kernel(PushArgs A, eb200_prtls_t S, uint32_t ngroups, uint32_t ahead,
                               PackedEM2 EB, float charge, float inv_dt, FieldView<2> J) {
      constexpr int D  = 2;
      constexpr int NV = ZigZag<D>::NV;
      __shared__ int   s_i[D][VEC][256];
      __shared__ float s_d[D][VEC][256];
      __shared__ float s_u[3][VEC][256];
      __shared__ float s_w[VEC][256];
      __shared__ short s_t[VEC][256];
      const int      tid      = threadIdx.x;
      const uint32_t g        = blockIdx.x * blockDim.x + tid;
      const bool     in_range = g < ngroups;
      const size_t   p0       = (size_t)g * VEC;
      int*           ii[3]    = { S.i1, S.i2, S.i3 };
      float*         dd[3]    = { S.dx1, S.dx2, S.dx3 };
      int*           iip[3]   = { S.i1_prev, S.i2_prev, S.i3_prev };
      float*         ddp[3]   = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
      if (tid == 0) {
        const size_t q0 = ((size_t)blockIdx.x + ahead) * blockDim.x * VEC;
        if (q0 + (size_t)blockDim.x * VEC <= (size_t)ngroups * VEC) {
          const unsigned b4 = blockDim.x * VEC * 4, b2 = blockDim.x * VEC * 2;
#pragma unroll
          for (int a = 0; a < D; ++a) {
            tma::prefetch_l2(ii[a] + q0, b4);
            tma::prefetch_l2(dd[a] + q0, b4);
          }
          tma::prefetch_l2(S.ux1 + q0, b4);
          tma::prefetch_l2(S.ux2 + q0, b4);
          tma::prefetch_l2(S.ux3 + q0, b4);
          tma::prefetch_l2(S.weight + q0, b4);
          tma::prefetch_l2(S.tag + q0, b2);
        }
      }
      bool all_pushed = false;
#pragma unroll
      for (int k = 0; k < VEC; ++k) s_t[k][tid] = 0;
      if (in_range) {
        short tv[VEC];
        ld4<short4>(S.tag + p0, tv);
        all_pushed = (tv[0] == 1) && (tv[1] == 1) && (tv[2] == 1) && (tv[3] == 1);
#pragma unroll
        for (int k = 0; k < VEC; ++k) s_t[k][tid] = tv[k];
#pragma unroll
        for (int a = 0; a < D; ++a) {
          int   iv[VEC];
          float dv[VEC];
          ld4<int4>(ii[a] + p0, iv);
          ld4<float4>(dd[a] + p0, dv);
          if (all_pushed) {
            st4<int4>(iip[a] + p0, iv);
            st4<float4>(ddp[a] + p0, dv);
          }
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            s_i[a][k][tid] = iv[k];
            s_d[a][k][tid] = dv[k];
          }
        }
        {
          float v[VEC];
          ld4<float4>(S.ux1 + p0, v);
#pragma unroll
          for (int k = 0; k < VEC; ++k) s_u[0][k][tid] = v[k];
          ld4<float4>(S.ux2 + p0, v);
#pragma unroll
          for (int k = 0; k < VEC; ++k) s_u[1][k][tid] = v[k];
          ld4<float4>(S.ux3 + p0, v);
#pragma unroll
          for (int k = 0; k < VEC; ++k) s_u[2][k][tid] = v[k];
          ld4<float4>(S.weight + p0, v);
#pragma unroll
          for (int k = 0; k < VEC; ++k) s_w[k][tid] = v[k];
        }
      }
      const long N12 = (long)J.N1 * J.N2;
      auto       red = [&](int key, const float (&a)[NV]) {
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          atomicAdd(J.p + key + zigzag_offset<D>(n, J.N1, N12, J.plane), a[n]);
        }
      };
      float acc[NV];
#pragma unroll
      for (int n = 0; n < NV; ++n) acc[n] = ZERO;
      int cur = -1;
#pragma unroll 1
      for (int k = 0; k < VEC; ++k) {
        const short tag = s_t[k][tid];
        if (tag != 1) {
          continue;
        }
        Prtl<D> P;
        int     i_old[D];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          if (a < D) {
            i_old[a] = s_i[a][k][tid];
            P.i[a] = P.ip[a] = i_old[a];
            P.d[a] = P.dp[a] = s_d[a][k][tid];
          } else {
            P.i[a] = P.ip[a] = 0;
            P.d[a] = P.dp[a] = ZERO;
          }
          P.u[a] = s_u[a][k][tid];
        }
        P.w   = s_w[k][tid];
        P.tag = tag;
        push_particle<D, 0, PackedEM2, LEAN>(A, EB, P);
        if (P.tag != tag) {
          S.tag[p0 + k] = P.tag;
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
          if (!all_pushed || P.ip[a] != i_old[a]) {
            iip[a][p0 + k] = P.ip[a];
          }
          if (!all_pushed) {
            ddp[a][p0 + k] = P.dp[a];
          }
          s_i[a][k][tid] = P.i[a];
          s_d[a][k][tid] = P.d[a];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) s_u[a][k][tid] = P.u[a];
        if (P.tag == 0) {
          continue;
        }
        float v[2][NV];
        zigzag_values<D>(P, charge, inv_dt, A.c.dx, v);
        const int  G     = A.ng;
        const int  key0  = (int)J.idx(P.ip[0] + G, P.ip[1] + G, 0);
        const int  key1  = (int)J.idx(P.i[0] + G, P.i[1] + G, 0);
        const bool cross = key0 != key1;
        if (cross) {
          red(key1, v[1]);
        } else {
#pragma unroll
          for (int n = 0; n < NV; ++n) v[0][n] += v[1][n];
        }
        if (key0 != cur) {
          if (cur >= 0) {
            red(cur, acc);
          }
          cur = key0;
#pragma unroll
          for (int n = 0; n < NV; ++n) acc[n] = v[0][n];
        } else {
#pragma unroll
          for (int n = 0; n < NV; ++n) acc[n] += v[0][n];
        }
      }
      if (in_range) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
          int   iv[VEC];
          float dv[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            iv[k] = s_i[a][k][tid];
            dv[k] = s_d[a][k][tid];
          }
          st4<int4>(ii[a] + p0, iv);
          st4<float4>(dd[a] + p0, dv);
        }
        float* uu[3] = { S.ux1, S.ux2, S.ux3 };
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float v[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[k] = s_u[a][k][tid];
          st4<float4>(uu[a] + p0, v);
        }
      }
      const WarpRun run = warp_runs(cur);
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float sm = run_sum(acc[n], run);
        if (run.head && cur >= 0) {
          atomicAdd(J.p + cur + zigzag_offset<D>(n, J.N1, N12, J.plane), sm);
        }
      }
    }

    /* ----------------------- tiled push + deposit (2D zig-zag, cell-sorted particles) */
    // The vectorised kernel above with the field gather moved to shared memory: the CTA's 1024
    // consecutive particles sit (when sorted) in ~1024/ppc consecutive cells of one row; the
    // E/B nodes of TILE_R rows x TILE_C columns around the CTA's first particle are staged once
    // (128-bit loads) and every gather is an LDS with an immediate offset instead of a global
    // load behind 64-bit address arithmetic. Particles outside the tile (strays between sorts,
    // row wrap) gather from global memory. The scatter stays on L2 atomics (RED.ADD.F32):
    // shared-memory fp32 atomics are a CAS loop on sm_100 (ATOMS.CAST.SPIN), which costs more
    // than the register accumulation + segmented shuffle reduction it would replace.
    constexpr int TILE_C = 96; // columns (x1), multiple of 4
    constexpr int TILE_R = 5;  // rows (x2): first particle's row -2 .. +2
    constexpr int TILE_N = TILE_C * TILE_R;

    template <bool LEAN>
    __global__ void __launch_bounds__(256, 3)
      push_deposit_tile_kernel(PushArgs A, eb200_prtls_t S, uint32_t ngroups, uint32_t ahead,
                               FieldView<2> EB, float charge, float inv_dt, FieldView<2> J) {
      constexpr int  NV = ZigZag<2>::NV;
      constexpr int  D  = 2;
      __shared__ __align__(16) float em_t[6 * TILE_N];
      const uint32_t g        = blockIdx.x * blockDim.x + threadIdx.x;
      const bool     in_range = g < ngroups;
      const size_t   p0       = (size_t)g * VEC;
      int*           ii[3]    = { S.i1, S.i2, S.i3 };
      float*         dd[3]    = { S.dx1, S.dx2, S.dx3 };
      int*           iip[3]   = { S.i1_prev, S.i2_prev, S.i3_prev };
      float*         ddp[3]   = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
      // tile origin from the CTA's first particle (one broadcast load per array)
      const size_t pf  = (size_t)blockIdx.x * blockDim.x * VEC;
      const int    gi0 = __ldg(S.i1 + pf) + A.ng, gj0 = __ldg(S.i2 + pf) + A.ng;
      if (threadIdx.x == 0) {
        const size_t q0 = ((size_t)blockIdx.x + ahead) * blockDim.x * VEC;
        if (q0 + (size_t)blockDim.x * VEC <= (size_t)ngroups * VEC) {
          const unsigned b4 = blockDim.x * VEC * 4, b2 = blockDim.x * VEC * 2;
#pragma unroll
          for (int a = 0; a < D; ++a) {
            tma::prefetch_l2(ii[a] + q0, b4);
            tma::prefetch_l2(dd[a] + q0, b4);
          }
          tma::prefetch_l2(S.ux1 + q0, b4);
          tma::prefetch_l2(S.ux2 + q0, b4);
          tma::prefetch_l2(S.ux3 + q0, b4);
          tma::prefetch_l2(S.weight + q0, b4);
          tma::prefetch_l2(S.tag + q0, b2);
        }
      }
      int   iv[2][VEC];
      float dv[2][VEC], uv[3][VEC], wv[VEC];
      short tv[VEC]    = { 0, 0, 0, 0 };
      bool  all_pushed = false;
      if (in_range) {
        ld4<short4>(S.tag + p0, tv);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          ld4<int4>(ii[a] + p0, iv[a]);
          ld4<float4>(dd[a] + p0, dv[a]);
        }
        ld4<float4>(S.ux1 + p0, uv[0]);
        ld4<float4>(S.ux2 + p0, uv[1]);
        ld4<float4>(S.ux3 + p0, uv[2]);
        ld4<float4>(S.weight + p0, wv);
      }
      // stage the E/B tile
      const int c0 = min(max((gi0 - 2) & ~3, 0), EB.N1 - TILE_C);
      const int r0 = min(max(gj0 - 2, 0), EB.N2 - TILE_R);
      {
        constexpr int Q = TILE_C / 4; // float4 per tile row
        for (int e = threadIdx.x; e < 6 * TILE_R * Q; e += blockDim.x) {
          const int    row = e / Q, q = e - row * Q; // row = comp * TILE_R + r
          const int    c = row / TILE_R, r = row - c * TILE_R;
          const float* src = EB.p + EB.plane * c + ((long)(r0 + r) * EB.N1 + c0 + 4 * q);
          reinterpret_cast<float4*>(em_t)[e] = __ldg(reinterpret_cast<const float4*>(src));
        }
      }
      if (in_range) {
        all_pushed = (tv[0] == 1) && (tv[1] == 1) && (tv[2] == 1) && (tv[3] == 1);
        if (all_pushed) {
#pragma unroll
          for (int a = 0; a < D; ++a) {
            st4<int4>(iip[a] + p0, iv[a]);
            st4<float4>(ddp[a] + p0, dv[a]);
          }
        }
      }
      __syncthreads();
      const TileEM<TILE_C, TILE_R> EM { em_t, c0, r0, EB };
      const long                   N12 = (long)J.N1 * J.N2;
      auto                         red = [&](int key, const float (&a)[NV]) {
#pragma unroll
        for (int n = 0; n < NV; ++n) {
          atomicAdd(J.p + key + zigzag_offset<2>(n, J.N1, N12, J.plane), a[n]);
        }
      };
      float acc[NV];
#pragma unroll
      for (int n = 0; n < NV; ++n) acc[n] = ZERO;
      int cur = -1;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const short tag = tv[k];
        if (tag != 1) {
          continue;
        }
        Prtl<2> P;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          P.i[a] = P.ip[a] = (a < D) ? iv[a < D ? a : 0][k] : 0;
          P.d[a] = P.dp[a] = (a < D) ? dv[a < D ? a : 0][k] : ZERO;
          P.u[a]           = uv[a][k];
        }
        P.w   = wv[k];
        P.tag = tag;
        push_particle<2, 0, TileEM<TILE_C, TILE_R>, LEAN>(A, EM, P);
        if (P.tag != tag) {
          S.tag[p0 + k] = P.tag;
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
          if (!all_pushed || P.ip[a] != iv[a][k]) {
            iip[a][p0 + k] = P.ip[a];
          }
          if (!all_pushed) {
            ddp[a][p0 + k] = P.dp[a];
          }
          iv[a][k] = P.i[a];
          dv[a][k] = P.d[a];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) uv[a][k] = P.u[a];
        if (P.tag == 0) {
          continue;
        }
        float v[2][NV];
        zigzag_values<2>(P, charge, inv_dt, A.c.dx, v);
        const int  G     = A.ng;
        const int  key0  = (P.ip[0] + G) + J.N1 * (P.ip[1] + G);
        const int  key1  = (P.i[0] + G) + J.N1 * (P.i[1] + G);
        const bool cross = key0 != key1;
        if (cross) {
          red(key1, v[1]);
        } else {
#pragma unroll
          for (int n = 0; n < NV; ++n) v[0][n] += v[1][n];
        }
        if (key0 != cur) {
          if (cur >= 0) {
            red(cur, acc);
          }
          cur = key0;
#pragma unroll
          for (int n = 0; n < NV; ++n) acc[n] = v[0][n];
        } else {
#pragma unroll
          for (int n = 0; n < NV; ++n) acc[n] += v[0][n];
        }
      }
      if (in_range) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
          st4<int4>(ii[a] + p0, iv[a]);
          st4<float4>(dd[a] + p0, dv[a]);
        }
        st4<float4>(S.ux1 + p0, uv[0]);
        st4<float4>(S.ux2 + p0, uv[1]);
        st4<float4>(S.ux3 + p0, uv[2]);
      }
      const WarpRun run = warp_runs(cur);
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float s = run_sum(acc[n], run);
        if (run.head && cur >= 0) {
          atomicAdd(J.p + cur + zigzag_offset<2>(n, J.N1, N12, J.plane), s);
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
    // 128-bit particle loads need 16-byte aligned arrays (8-byte for the int16 tags)
    static bool aligned16(const eb200_prtls_t& S, int dim) {
      const void* a4[] = { S.i1, S.dx1, S.i1_prev, S.dx1_prev, S.ux1, S.ux2, S.ux3, S.weight,
                           dim > 1 ? (const void*)S.i2 : nullptr, dim > 1 ? S.dx2 : nullptr,
                           dim > 1 ? (const void*)S.i2_prev : nullptr, dim > 1 ? S.dx2_prev : nullptr,
                           dim > 2 ? (const void*)S.i3 : nullptr, dim > 2 ? S.dx3 : nullptr,
                           dim > 2 ? (const void*)S.i3_prev : nullptr, dim > 2 ? S.dx3_prev : nullptr };
      for (const void* q : a4) {
        if (q && (reinterpret_cast<uintptr_t>(q) & 15u)) return false;
      }
      return (reinterpret_cast<uintptr_t>(S.tag) & 7u) == 0;
    }

    template <int D, int O>
    cudaError_t launch_push(const PushArgs& A, const eb200_prtls_t& S, uint32_t npart,
                            const eb200_grid_t& g, const float* em, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> EB(g, const_cast<float*>(em));
      push_kernel<D, O><<<(npart + 255) / 256, 256, 0, st>>>(A, S, npart, EB);
      count_launch();
      return cudaGetLastError();
    }

    // 3D Esirkepov windows (27 .. 125 nodes x 3 components): the segmented shuffle reduction of
    // the AGGREGATED mode costs ten instructions per node and lane, more than the atomics it
    // saves. Measured per step on 1.4e7 cell-sorted particles (scripts/deposit_modes.py), atomic
    // vs aggregated: 3D O=1 1.82 / 2.65 ms, O=2 6.26 / 8.97, O=3 13.8 / 19.5; in 2D the
    // aggregation wins (O=2 3.17 / 2.43, O=3 6.76 / 3.30) and stays. The mode is a hint about
    // the order of the additions, not about the result.
    template <int D, int O>
    constexpr bool wide_window() {
      return D == 3 && O >= 1;
    }

    template <int D, int O>
    cudaError_t launch_deposit(const eb200_prtls_t& S, uint32_t npart, const eb200_grid_t& g,
                               float charge, float dt, float dxc, float* cur, int mode,
                               Scratch& scratch, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> J(g, cur);
      const float  inv_dt = ONE / dt;
      if (wide_window<D, O>() && mode == EB200_DEPOSIT_AGGREGATED) mode = EB200_DEPOSIT_ATOMIC;
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


    // resident CTAs of one wave of `kern` on the current device (occupancy x SM count), cached per
    // (device, kernel): the prefetch distance of the vectorised kernels
    inline uint32_t resident_ctas(const void* kern, int threads) {
      struct Key {
        int         dev;
        const void* f;
        uint32_t    v;
      };
      static std::mutex       mu;
      static std::vector<Key> cache;
      int                     dev = 0;
      cudaGetDevice(&dev);
      std::lock_guard<std::mutex> lk(mu);
      for (const Key& k : cache) {
        if (k.dev == dev && k.f == kern) return k.v;
      }
      int nsm = 0, per_sm = 0;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
      const uint32_t v = (uint32_t)(nsm * (per_sm > 0 ? per_sm : 1));
      cache.push_back({ dev, kern, v });
      return v;
    }

    inline bool tile3_disabled() {
      static const bool off = getenv("EB200_NO_TILE3") != nullptr;
      return off;
    }

    inline bool j4_disabled() {
      static const bool off = getenv("EB200_NO_J4") != nullptr;
      return off;
    }

    inline bool mom_disabled() {
      static const bool off = getenv("EB200_NO_MOM") != nullptr;
      return off;
    }
    template <int D, int O>
    cudaError_t launch_push_deposit(const PushArgs& A, const eb200_prtls_t& S, uint32_t npart,
                                    const eb200_grid_t& g, const float* em, float* cur,
                                    int mode, float* packed, bool do_pack, cudaStream_t st,
                                    float* packed_j = nullptr, bool* packed_j_used = nullptr,
                                    ExcList* exc = nullptr) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> EB(g, const_cast<float*>(em));
      FieldView<D> J(g, cur);
      const float inv_dt = ONE / A.c.dt;
      uint32_t    p_begin = 0;
      // which fused kernel (eb200_set_pd_kernel): 0 auto, 1 one particle per thread,
      // 2 TMA-staged persistent, 3 four particles per thread (zig-zag only), 4 shared-memory
      // field tile, 5 four particles per thread gathering from packed nodes (2D zig-zag)
      const int  which     = (mode >> 8) & 0xff;
      const bool lean_prev = ((mode >> 16) & 1) != 0; // eb200_set_lean_prev
      mode &= 0xff;
      if (wide_window<D, O>() && mode == EB200_DEPOSIT_AGGREGATED && which == 0) {
        mode = EB200_DEPOSIT_ATOMIC;
      }
      const bool want_vec = (O == 0) && (which == 0 || which == 3);
#if !EB200_STRICT
      if constexpr (O == 3 && D == 3) {
        // 9 (what 0 selects): shared-memory fixed-point J tile per chunk of 256 particles
        if ((which == 9 || (which == 0 && !tile3_disabled())) && mode != EB200_DEPOSIT_ORDERED &&
            aligned16(S, D) && npart >= (uint32_t)STREAM_CHUNK && J.plane * 3 < 0x7fffffffL) {
          const uint32_t nchunks = npart / STREAM_CHUNK;
          const bool     lean    = lean_pusher(A.c);
          auto           kern    = lean ? push_deposit_tile3_kernel<true> : push_deposit_tile3_kernel<false>;
          const size_t   smem    = sizeof(Tile3Smem);
          static std::mutex mu;
          static bool       attr[2] = { false, false };
          {
            std::lock_guard<std::mutex> lk(mu);
            if (!attr[lean]) {
              cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
              attr[lean] = true;
            }
          }
          int dev = 0, nsm = 0, per_sm = 0;
          cudaGetDevice(&dev);
          cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STREAM_CHUNK, smem);
          const uint32_t slots = (uint32_t)(nsm * (per_sm > 0 ? per_sm : 1));
          const uint32_t grid  = nchunks < slots ? nchunks : slots;
          kern<<<grid, STREAM_CHUNK, smem, st>>>(A, S, nchunks, EB, A.c.charge, inv_dt, J);
          count_launch();
          p_begin = nchunks * STREAM_CHUNK;
          if (p_begin == npart) return cudaGetLastError();
          // the tail: one particle per thread, per-lane atomics
          push_deposit_kernel<D, O, false><<<(npart - p_begin + 255) / 256, 256, 0, st>>>(
            A, S, p_begin, npart, EB, A.c.charge, inv_dt, J);
          count_launch();
          return cudaGetLastError();
        }
      }
#endif
      if constexpr (O == 0 && D == 2) {
        if (which == 7 && packed != nullptr && mode == EB200_DEPOSIT_AGGREGATED && aligned16(S, D) &&
            npart >= VEC && J.plane < 0x7fffffffL && EB.plane * 24 < 0xffffffffL) {
          if (do_pack) {
            pack_em2d_kernel<<<(unsigned)((EB.plane + 255) / 256), 256, 0, st>>>(
              em, EB.plane, reinterpret_cast<float2*>(packed));
            count_launch();
          }
          PackedEM2 PK;
          PK.p    = reinterpret_cast<const char*>(packed);
          PK.rowb = 24u * (unsigned)EB.N1;
          const uint32_t ngroups = npart / VEC;
          const bool     lean    = lean_pusher(A.c);
          auto           kern    = lean ? push_deposit_smem_kernel<true> : push_deposit_smem_kernel<false>;
          static int wave_tbl[64][2] = {}; // per device: occupancy and function attributes are per device
          int        wave_dev     = 0;
          cudaGetDevice(&wave_dev);
          int* wave = wave_tbl[wave_dev & 63];
          if (wave[lean] == 0) {
            int dev = 0, nsm = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 75);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0);
            wave[lean] = nsm * (per_sm > 0 ? per_sm : 1);
          }
          kern<<<(ngroups + 255) / 256, 256, 0, st>>>(A, S, ngroups, (uint32_t)wave[lean], PK,
                                                       A.c.charge, inv_dt, J);
          count_launch();
          p_begin = ngroups * VEC;
          if (p_begin == npart) return cudaGetLastError();
        }
        if (which == 6 && packed != nullptr && mode == EB200_DEPOSIT_AGGREGATED && aligned16(S, D) &&
            npart >= VEC && J.plane < 0x7fffffffL && EB.plane * 24 < 0xffffffffL) {
          if (do_pack) {
            pack_em2d_kernel<<<(unsigned)((EB.plane + 255) / 256), 256, 0, st>>>(
              em, EB.plane, reinterpret_cast<float2*>(packed));
            count_launch();
          }
          PackedEM2 PK;
          PK.p    = reinterpret_cast<const char*>(packed);
          PK.rowb = 24u * (unsigned)EB.N1;
          const uint32_t ngroups = npart / VEC;
          const bool     lean    = lean_pusher(A.c);
          auto           kern    = lean ? push_deposit_pipe_kernel<true> : push_deposit_pipe_kernel<false>;
          static int slots_tbl[64][2] = {}; // per device: occupancy and function attributes are per device
          int        slots_dev     = 0;
          cudaGetDevice(&slots_dev);
          int* slots = slots_tbl[slots_dev & 63];
          if (slots[lean] == 0) {
            int dev = 0, nsm = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PIPE_THREADS, 0);
            slots[lean] = nsm * (per_sm > 0 ? per_sm : 1);
          }
          const uint32_t ntiles = (ngroups + PIPE_THREADS - 1) / PIPE_THREADS;
          const uint32_t grid   = ntiles < (uint32_t)slots[lean] ? ntiles : (uint32_t)slots[lean];
          const uint32_t per    = (ntiles + grid - 1) / grid;
          // one tile of cell-sorted particles spans about 1024 / (particles per cell) cells
          const double   per_cell = (double)npart / ((double)g.n[0] * g.n[1]);
          static const int ahead_env = getenv("EB200_PIPE_AHEAD") ? atoi(getenv("EB200_PIPE_AHEAD")) : -1;
          unsigned       ahead  = ahead_env >= 0 ? (unsigned)ahead_env * 24u
                                                 : (unsigned)((PIPE_THREADS * VEC) / (per_cell > 1.0 ? per_cell : 1.0)) * 24u;
          static const bool blocked = getenv("EB200_PIPE_BLOCKED") != nullptr;
          if (blocked) {
            kern<<<(ntiles + per - 1) / per, PIPE_THREADS, 0, st>>>(
              A, S, ngroups, per, 1u, ahead, (unsigned)(EB.plane * 24), PK, A.c.charge, inv_dt, J);
          } else {
            if (ahead_env < 0) ahead *= grid;
            kern<<<grid, PIPE_THREADS, 0, st>>>(
              A, S, ngroups, per, grid, ahead, (unsigned)(EB.plane * 24), PK, A.c.charge, inv_dt, J);
          }
          count_launch();
          p_begin = ngroups * VEC;
          if (p_begin == npart) return cudaGetLastError();
        }
#if !EB200_STRICT
        // 8 (what 0 selects for a plain Boris push): moment-accumulating form of kernel 5
        if ((which == 8 || (which == 0 && !mom_disabled())) && lean_pusher(A.c) && p_begin == 0 &&
            packed != nullptr && mode == EB200_DEPOSIT_AGGREGATED && aligned16(S, D) && npart >= VEC &&
            J.plane * 3 < 0x7fffffffL && EB.plane * 24 < 0xffffffffL) {
          if (do_pack) {
            pack_em2d_kernel<<<(unsigned)((EB.plane + 255) / 256), 256, 0, st>>>(
              em, EB.plane, reinterpret_cast<float2*>(packed));
            count_launch();
          }
          PackedEM2 PK;
          PK.p    = reinterpret_cast<const char*>(packed);
          PK.rowb = 24u * (unsigned)EB.N1;
          const uint32_t ngroups = npart / VEC;
          const bool     v4      = packed_j != nullptr && packed_j_used != nullptr && !j4_disabled();
          // this kernel and the tail kernel below both report who is not alive afterwards
          PushArgs AX = A;
          if (exc != nullptr && exc->count != nullptr) {
            AX.exc_count = exc->count, AX.exc_idx = exc->idx, AX.exc_cap = exc->cap;
            exc->tracked = true;
          }
          const PushArgs& A = AX; // (shadows the parameter for the launches of this branch)
          if (v4 && lean_prev) {
            const uint32_t wave = resident_ctas(reinterpret_cast<const void*>(push_deposit_mom_kernel<true, false>), 256);
            push_deposit_mom_kernel<true, false><<<(ngroups + 255) / 256, 256, 0, st>>>(
              A, S, ngroups, wave, PK, A.c.charge, inv_dt, J, reinterpret_cast<float4*>(packed_j));
            *packed_j_used = true;
          } else if (v4) {
            const uint32_t wave = resident_ctas(reinterpret_cast<const void*>(push_deposit_mom_kernel<true>), 256);
            push_deposit_mom_kernel<true><<<(ngroups + 255) / 256, 256, 0, st>>>(
              A, S, ngroups, wave, PK, A.c.charge, inv_dt, J, reinterpret_cast<float4*>(packed_j));
            *packed_j_used = true;
          } else {
            const uint32_t wave = resident_ctas(reinterpret_cast<const void*>(push_deposit_mom_kernel<false>), 256);
            push_deposit_mom_kernel<false><<<(ngroups + 255) / 256, 256, 0, st>>>(
              A, S, ngroups, wave, PK, A.c.charge, inv_dt, J, nullptr);
          }
          count_launch();
          p_begin = ngroups * VEC;
          if (p_begin == npart) return cudaGetLastError();
        }
#endif
        if ((which == 5 || which == 0 || which == 8) && p_begin == 0 && packed != nullptr && mode == EB200_DEPOSIT_AGGREGATED && aligned16(S, D) &&
            npart >= VEC && J.plane < 0x7fffffffL && EB.plane * 24 < 0xffffffffL) {
          if (do_pack) {
            pack_em2d_kernel<<<(unsigned)((EB.plane + 255) / 256), 256, 0, st>>>(
              em, EB.plane, reinterpret_cast<float2*>(packed));
            count_launch();
          }
          PackedEM2 PK;
          PK.p    = reinterpret_cast<const char*>(packed);
          PK.rowb = 24u * (unsigned)EB.N1;
          const uint32_t ngroups = npart / VEC;
          const bool     lean    = lean_pusher(A.c);
          auto           kern    = lean ? push_deposit_vec_kernel<2, true, PackedEM2>
                                        : push_deposit_vec_kernel<2, false, PackedEM2>;
          static int wave_tbl[64][2] = {}; // per device: occupancy and function attributes are per device
          int        wave_dev     = 0;
          cudaGetDevice(&wave_dev);
          int* wave = wave_tbl[wave_dev & 63];
          // CTA size of the packed-node kernel (EB200_VEC_THREADS = 64 / 128 / 256, experiment)
          static const int nthr = [] {
            const char* e = getenv("EB200_VEC_THREADS");
            const int   v = e ? atoi(e) : 256;
            return (v == 64 || v == 128) ? v : 256;
          }();
          if (wave[lean] == 0) {
            int dev = 0, nsm = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthr, 0);
            wave[lean] = nsm * (per_sm > 0 ? per_sm : 1);
          }
          kern<<<(ngroups + nthr - 1) / nthr, nthr, 0, st>>>(A, S, ngroups, (uint32_t)wave[lean], PK,
                                                             A.c.charge, inv_dt, J);
          count_launch();
          p_begin = ngroups * VEC;
          if (p_begin == npart) return cudaGetLastError();
        }
      }
      if constexpr (O == 0 && D == 2) {
        // 4: shared-memory field tile (needs 16-byte aligned rows and a mesh at least one
        // tile wide)
        const bool tile_ok = (EB.N1 % 4 == 0) && EB.N1 >= TILE_C && EB.N2 >= TILE_R &&
                             (reinterpret_cast<uintptr_t>(em) & 15u) == 0 &&
                             J.plane * 3 < 0x7fffffffL;
        // measured 6 % slower than kernel 3 on the 4096x2048x32ppc config (the staging barrier
        // costs more than the cheaper gathers save): kept selectable, not the default
        if (which == 4 && tile_ok && mode == EB200_DEPOSIT_AGGREGATED &&
            aligned16(S, D) && npart >= VEC) {
          const uint32_t ngroups = npart / VEC;
          const bool     lean    = lean_pusher(A.c);
          auto           kern    = lean ? push_deposit_tile_kernel<true>
                                        : push_deposit_tile_kernel<false>;
          static int wave_tbl[64][2] = {}; // per device: occupancy and function attributes are per device
          int        wave_dev     = 0;
          cudaGetDevice(&wave_dev);
          int* wave = wave_tbl[wave_dev & 63];
          if (wave[lean] == 0) {
            int dev = 0, nsm = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0);
            wave[lean] = nsm * (per_sm > 0 ? per_sm : 1);
          }
          kern<<<(ngroups + 255) / 256, 256, 0, st>>>(A, S, ngroups, (uint32_t)wave[lean], EB,
                                                       A.c.charge, inv_dt, J);
          count_launch();
          p_begin = ngroups * VEC;
          if (p_begin == npart) return cudaGetLastError();
        }
      }
      if constexpr (O == 0) {
        if (p_begin == 0 && want_vec && mode == EB200_DEPOSIT_AGGREGATED && aligned16(S, D) && npart >= VEC &&
            J.plane < 0x7fffffffL) {
          const uint32_t ngroups = npart / VEC;
          const bool     lean    = lean_pusher(A.c);
          auto           kern    = lean ? push_deposit_vec_kernel<D, true>
                                        : push_deposit_vec_kernel<D, false>;
          static int wave_tbl[64][2] = {}; // per device: occupancy and function attributes are per device
          int        wave_dev     = 0;
          cudaGetDevice(&wave_dev);
          int* wave = wave_tbl[wave_dev & 63]; // resident CTAs of one wave for (full, lean)
          if (wave[lean] == 0) {
            int dev = 0, nsm = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0);
            wave[lean] = nsm * (per_sm > 0 ? per_sm : 1);
          }
          kern<<<(ngroups + 255) / 256, 256, 0, st>>>(A, S, ngroups, (uint32_t)wave[lean], EB,
                                                       A.c.charge, inv_dt, J);
          count_launch();
          p_begin = ngroups * VEC;
          if (p_begin == npart) return cudaGetLastError();
        }
      }
      if (p_begin == 0 && which != 1 && mode == EB200_DEPOSIT_AGGREGATED && aligned16(S, D) &&
          npart >= (uint32_t)STREAM_CHUNK) {
        // full 256-particle chunks: TMA-staged persistent kernel; the tail goes below
        const uint32_t nchunks = npart / STREAM_CHUNK;
        const bool     lean    = lean_pusher(A.c);
        auto           kern    = lean ? push_deposit_stream_kernel<D, O, true>
                                      : push_deposit_stream_kernel<D, O, false>;
        const size_t   smem    = sizeof(StreamSmem<D>);
        static int slots_tbl[64][2] = {}; // per device: occupancy and function attributes are per device
        int        slots_dev     = 0;
        cudaGetDevice(&slots_dev);
        int* slots = slots_tbl[slots_dev & 63]; // resident CTAs per device for (lean, full)
        if (slots[lean] == 0) {
          int dev = 0, nsm = 0, per_sm = 0;
          cudaGetDevice(&dev);
          cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, STREAM_CHUNK, smem);
          slots[lean] = nsm * (per_sm > 0 ? per_sm : 1);
        }
        const uint32_t grid = nchunks < (uint32_t)slots[lean] ? nchunks : (uint32_t)slots[lean];
        kern<<<grid, STREAM_CHUNK, smem, st>>>(A, S, nchunks, EB, A.c.charge, inv_dt, J);
        count_launch();
        p_begin = nchunks * STREAM_CHUNK;
        if (p_begin == npart) return cudaGetLastError();
      }
      const uint32_t nrest = npart - p_begin;
      PushArgs       AT    = A; // the tail reports to the exception list its head reported to
      if (exc != nullptr && exc->tracked) {
        AT.exc_count = exc->count, AT.exc_idx = exc->idx, AT.exc_cap = exc->cap;
      }
      if (mode == EB200_DEPOSIT_AGGREGATED) {
        push_deposit_kernel<D, O, true><<<(nrest + 255) / 256, 256, 0, st>>>(
          AT, S, p_begin, npart, EB, A.c.charge, inv_dt, J);
      } else {
        push_deposit_kernel<D, O, false><<<(nrest + 255) / 256, 256, 0, st>>>(
          AT, S, p_begin, npart, EB, A.c.charge, inv_dt, J);
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

// shape orders 4..11: the unfused push and the per-lane atomic deposit only
#define EB200_DISPATCH_HI_D(D, order, CALL)                                                    \
  switch (order) {                                                                             \
    case 4: return CALL(D, 4);                                                                 \
    case 5: return CALL(D, 5);                                                                 \
    case 6: return CALL(D, 6);                                                                 \
    case 7: return CALL(D, 7);                                                                 \
    case 8: return CALL(D, 8);                                                                 \
    case 9: return CALL(D, 9);                                                                 \
    case 10: return CALL(D, 10);                                                               \
    case 11: return CALL(D, 11);                                                               \
    default: return cudaErrorInvalidValue;                                                     \
  }
#define EB200_DISPATCH_HI(dim, order, CALL)                                                    \
  switch (dim) {                                                                               \
    case 1: EB200_DISPATCH_HI_D(1, order, CALL)                                                \
    case 2: EB200_DISPATCH_HI_D(2, order, CALL)                                                \
    case 3: EB200_DISPATCH_HI_D(3, order, CALL)                                                \
    default: return cudaErrorInvalidValue;                                                     \
  }

    template <int D, int O>
    cudaError_t launch_deposit_hi(const eb200_prtls_t& S, uint32_t npart, const eb200_grid_t& g,
                                  float charge, float dt, float dxc, float* cur, cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> J(g, cur);
      deposit_atomic_kernel<D, O, false>
        <<<(npart + 255) / 256, 256, 0, st>>>(S, npart, charge, ONE / dt, dxc, g.ng, J);
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t push_sr(const eb200_grid_t& g, int order, const eb200_pusher_t& c,
                        const eb200_prtls_t& S, uint32_t npart, const float* em,
                        cudaStream_t st) {
      PushArgs A;
      A.c   = c;
      A.ndh = HALF * (c.charge / c.mass) * c.omegaB0 * c.dt;
      A.ng  = g.ng;
      A.inv_dx = ONE / c.dx;
      for (int a = 0; a < 3; ++a) A.ni[a] = g.n[a];
#define CALL(D, O) launch_push<D, O>(A, S, npart, g, em, st)
      if (order > 3) {
        EB200_DISPATCH_HI(g.dim, order, CALL)
      }
      EB200_DISPATCH_DO(g.dim, order, CALL)
#undef CALL
    }

    template <int D, int O>
    cudaError_t launch_push_emit(const PushArgs& A, const eb200_prtls_t& S, uint32_t npart,
                                 const eb200_grid_t& g, const float* em, const EmitArgs& M,
                                 cudaStream_t st) {
      if (npart == 0) return cudaSuccess;
      FieldView<D> EB(g, const_cast<float*>(em));
      push_emit_kernel<D, O><<<(npart + 255) / 256, 256, 0, st>>>(A, S, npart, EB, M);
      count_launch();
      return cudaGetLastError();
    }

    cudaError_t push_sr_emission(const eb200_grid_t& g, int order, const eb200_pusher_t& c,
                                 const eb200_prtls_t& S, uint32_t npart, const float* em,
                                 const EmitArgs& M, cudaStream_t st) {
      PushArgs A;
      A.c   = c;
      A.ndh = HALF * (c.charge / c.mass) * c.omegaB0 * c.dt;
      A.ng  = g.ng;
      A.inv_dx = ONE / c.dx;
      for (int a = 0; a < 3; ++a) A.ni[a] = g.n[a];
#define CALL(D, O) launch_push_emit<D, O>(A, S, npart, g, em, M, st)
      if (order > 3) return cudaErrorNotSupported;
      EB200_DISPATCH_DO(g.dim, order, CALL)
#undef CALL
    }

    cudaError_t deposit(const eb200_grid_t& g, int order, const eb200_prtls_t& S, uint32_t npart,
                        float charge, float dt, float dxc, float* cur, int mode,
                        Scratch& scratch, cudaStream_t st) {
      if (order > 3) {
        // ATOMIC and AGGREGATED are hints about the order of the additions; ORDERED is a promise
        if (mode == EB200_DEPOSIT_ORDERED) return cudaErrorNotSupported;
#define CALL(D, O) launch_deposit_hi<D, O>(S, npart, g, charge, dt, dxc, cur, st)
        EB200_DISPATCH_HI(g.dim, order, CALL)
#undef CALL
      }
#define CALL(D, O) launch_deposit<D, O>(S, npart, g, charge, dt, dxc, cur, mode, scratch, st)
      EB200_DISPATCH_DO(g.dim, order, CALL)
#undef CALL
    }

    cudaError_t push_deposit_sr(const eb200_grid_t& g, int order, const eb200_pusher_t& c,
                                const eb200_prtls_t& S, uint32_t npart, const float* em,
                                float* cur, int mode, float* packed, bool do_pack,
                                cudaStream_t st, float* packed_j, bool* packed_j_used,
                                ExcList* exc) {
      PushArgs A;
      A.c   = c;
      A.ndh = HALF * (c.charge / c.mass) * c.omegaB0 * c.dt;
      A.ng  = g.ng;
      A.inv_dx = ONE / c.dx;
      for (int a = 0; a < 3; ++a) A.ni[a] = g.n[a];
      if (order > 3) return cudaErrorNotSupported; // unfused path only (engine.cu falls back)
#define CALL(D, O)                                                                             \
  launch_push_deposit<D, O>(A, S, npart, g, em, cur, mode, packed, do_pack, st, packed_j, packed_j_used, exc)
      EB200_DISPATCH_DO(g.dim, order, CALL)
#undef CALL
    }

    cudaError_t unpack_j4(const eb200_grid_t& g, float* packed_j, float* cur, cudaStream_t st) {
#if !EB200_STRICT
      FieldView<2> J(g, cur);
      unpack_j4_kernel<<<(unsigned)((J.plane + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<float4*>(packed_j), J.plane, cur);
      count_launch();
      return cudaGetLastError();
#else
      (void)g, (void)packed_j, (void)cur, (void)st; // the strict build never fills the nodes
      return cudaSuccess;
#endif
    }

    cudaError_t pack_em2d(const eb200_grid_t& g, const float* em, float* packed, cudaStream_t st) {
      FieldView<2> EB(g, const_cast<float*>(em));
      pack_em2d_kernel<<<(unsigned)((EB.plane + 255) / 256), 256, 0, st>>>(
        em, EB.plane, reinterpret_cast<float2*>(packed));
      count_launch();
      return cudaGetLastError();
    }

  } // namespace EB200_VARIANT
} // namespace eb200
