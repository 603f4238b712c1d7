// entity_b200 -- shared device-side definitions.
//
// Every kernel translation unit is compiled twice by the build (see build.py):
//   * EB200_STRICT=1 with --fmad=false: no FMA contraction, so particle pushes and
//     cell updates are bit-identical to the reference's baseline x86-64 CPU build;
//   * EB200_STRICT=0: nvcc's default contraction (what the reference's own CUDA build
//     would get), the throughput path.
// The two variants live in different namespaces and are selected per context.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/entity_b200.h"

#ifndef EB200_STRICT
  #define EB200_STRICT 0
#endif
#if EB200_STRICT
  #define EB200_VARIANT strict_fp
#else
  #define EB200_VARIANT fast_fp
#endif

namespace eb200 {

  constexpr float ONE = 1.0f, TWO = 2.0f, THREE = 3.0f, FOUR = 4.0f, FIVE = 5.0f;
  constexpr float ZERO = 0.0f, HALF = 0.5f;
  // the reference's single-precision THIRD is 0.333333f, not 1/3 (numeric.h:41)
  constexpr float THIRD = 0.333333f;
  constexpr float THREE_FOURTHS = 0.75f, THREE_HALFS = 1.5f;
  constexpr float INV_2 = 0.5f, INV_4 = 0.25f, INV_8 = 0.125f, INV_16 = 0.0625f;
  constexpr float INV_32 = 0.03125f, INV_64 = 0.015625f;

  enum { ex1 = 0, ex2 = 1, ex3 = 2, bx1 = 3, bx2 = 4, bx3 = 5 };
  enum { jx1 = 0, jx2 = 1, jx3 = 2 };

  __host__ __device__ __forceinline__ float SQR(float x) { return x * x; }
  __host__ __device__ __forceinline__ float CUBE(float x) { return x * x * x; }

  // Field of ncomp component planes, i1 fastest (LayoutLeft).
  template <int D>
  struct FieldView {
    float* p;
    int    N1, N2, N3; // extents incl. ghosts (unused dims = 1)
    long   plane;      // N1*N2*N3

    __host__ __device__ FieldView() {}

    __host__ __device__ FieldView(const eb200_grid_t& g, float* ptr) : p { ptr } {
      N1    = g.n[0] + 2 * g.ng;
      N2    = (D > 1) ? g.n[1] + 2 * g.ng : 1;
      N3    = (D > 2) ? g.n[2] + 2 * g.ng : 1;
      plane = (long)N1 * N2 * N3;
    }

    __device__ __forceinline__ long idx(int i, int j, int k) const {
      if constexpr (D == 1) {
        return i;
      } else if constexpr (D == 2) {
        return i + (long)N1 * j;
      } else {
        return i + (long)N1 * (j + (long)N2 * k);
      }
    }

    __device__ __forceinline__ float& at(int i, int j, int k, int c) const {
      return p[idx(i, j, k) + plane * c];
    }

    // read-only path (ld.global.nc)
    __device__ __forceinline__ float ld(int i, int j, int k, int c) const {
      return __ldg(p + idx(i, j, k) + plane * c);
    }
  };

  struct LaunchCounter; // capi.cu

#ifdef __CUDACC__
  // Runs of consecutive lanes with the same key (particles arrive nearly cell-sorted): the head
  // lane of a run acts for the whole run. All 32 lanes must call.
  struct LaneRun {
    bool head;
    int  first, last; // lanes of this lane's run
  };

  __device__ __forceinline__ LaneRun lane_run(long long key) {
    const unsigned  lane  = threadIdx.x & 31u;
    const long long prev  = __shfl_up_sync(0xffffffffu, key, 1);
    const bool      head  = (lane == 0u) || (prev != key);
    const unsigned  heads = __ballot_sync(0xffffffffu, head);
    const unsigned  above = (lane == 31u) ? 0u : (heads & ~((2u << lane) - 1u));
    LaneRun         r;
    r.head  = head;
    r.last  = above ? (__ffs(above) - 2) : 31;
    r.first = 31 - __clz(heads & ((2u << lane) - 1u));
    return r;
  }

  // sum of v over the lanes lane .. run.last; the run's total on its head lane
  __device__ __forceinline__ float lane_run_sum(float v, const LaneRun& r) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o = __shfl_down_sync(0xffffffffu, v, d);
      if (lane + d <= r.last) v += o;
    }
    return v;
  }
#endif

} // namespace eb200
