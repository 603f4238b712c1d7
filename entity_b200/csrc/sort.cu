// entity_b200 -- particle sorting / compaction (integer + copy work, one variant).
//
// Replaces Particles::SortSpatially and Particles::RemoveDead
// (src/framework/containers/particles_sort.cpp:104-253, key: src/global/utils/sorting.h:120-133)
// with one stable key-index radix sort followed by a gather of every SoA array through the
// permutation. Differences by design: the key runs i1-fastest (the field layout, so that a warp
// of consecutive particles touches consecutive field nodes), and the sort is stable, i.e.
// deterministic, where the reference's atomic compaction is not.
#include "common.cuh"
#include "launch.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdio>
#include <cstdlib>

namespace eb200 {

  template <int D>
  __global__ void __launch_bounds__(256)
    sort_keys_kernel(eb200_prtls_t S, uint32_t npart, int n1, int n2, int n3, uint32_t ncells,
                     uint32_t* keys, uint32_t* idx) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npart) {
      uint32_t key = ncells; // dead (or any non-alive tag) goes last
      if (S.tag[p] == 1) {
        int i = min(max(S.i1[p], 0), n1 - 1);
        key   = (uint32_t)i;
        if constexpr (D > 1) {
          int j = min(max(S.i2[p], 0), n2 - 1);
          key  += (uint32_t)n1 * (uint32_t)j;
          if constexpr (D > 2) {
            int k = min(max(S.i3[p], 0), n3 - 1);
            key  += (uint32_t)n1 * (uint32_t)n2 * (uint32_t)k;
          }
        }
      }
      keys[p] = key;
      idx[p]  = p;
    }
  }

  // number of alive particles = position of the first key == ncells in the sorted keys
  // (a binary search by one thread; a per-warp atomic counter on one address costs ~1 ns per
  // warp, 5 ms for 1.7e8 particles)
  __global__ void count_alive_kernel(const uint32_t* __restrict__ sorted_keys, uint32_t npart,
                                     uint32_t ncells, uint32_t* n_alive) {
    uint32_t lo = 0, hi = npart;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      if (sorted_keys[mid] < ncells) {
        lo = mid + 1;
      } else {
        hi = mid;
      }
    }
    *n_alive = lo;
  }

  // ---- counting sort by cell (EB200_SORT_UNSTABLE) ------------------------------------
  // One histogram pass, one scan over the cells, one slot-assignment pass: 18 B per particle
  // instead of the radix sort's 60 B. Runs of consecutive lanes with the same cell (the array is
  // nearly sorted from the previous sort) share one atomic. The order of the particles INSIDE
  // a cell follows the order in which warps reach the cell's counter: unspecified (the
  // reference's order is unspecified as well: particles_sort.cpp:118-147 uses atomics).
  template <int D>
  __global__ void __launch_bounds__(256)
    cell_hist_kernel(eb200_prtls_t S, uint32_t npart, int n1, int n2, int n3, uint32_t ncells,
                     uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    const uint32_t p   = blockIdx.x * blockDim.x + threadIdx.x;
    long long      key = -1 - (long long)(threadIdx.x & 31u);
    if (p < npart) {
      uint32_t k = ncells;
      if (S.tag[p] == 1) {
        k = (uint32_t)min(max(S.i1[p], 0), n1 - 1);
        if constexpr (D > 1) k += (uint32_t)n1 * (uint32_t)min(max(S.i2[p], 0), n2 - 1);
        if constexpr (D > 2) k += (uint32_t)n1 * (uint32_t)n2 * (uint32_t)min(max(S.i3[p], 0), n3 - 1);
      }
      keys[p] = k;
      key     = k;
    }
    const LaneRun r = lane_run(key);
    if (r.head && key >= 0) atomicAdd(hist + key, (uint32_t)(r.last - (int)(threadIdx.x & 31u) + 1));
  }

  // offs[c] = first slot of cell c (exclusive scan of the histogram), advanced by the atomics
  __global__ void __launch_bounds__(256)
    cell_slots_kernel(const uint32_t* __restrict__ keys, uint32_t npart, uint32_t* __restrict__ offs,
                      uint32_t* __restrict__ perm) {
    const uint32_t p    = blockIdx.x * blockDim.x + threadIdx.x;
    const int      lane = threadIdx.x & 31;
    long long      key  = -1 - (long long)lane;
    if (p < npart) key = keys[p];
    const LaneRun r    = lane_run(key);
    uint32_t      base = 0;
    if (r.head && key >= 0) base = atomicAdd(offs + key, (uint32_t)(r.last - lane + 1));
    base = __shfl_sync(0xffffffffu, base, r.first);
    if (p < npart) perm[base + (uint32_t)(lane - r.first)] = p;
  }

  template <class T>
  __global__ void __launch_bounds__(256)
    gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n,
                  T* __restrict__ dst) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) {
      dst[q] = src[perm[q]];
    }
  }

  template <class T>
  static void permute(T* arr, const uint32_t* perm, uint32_t n, void* tmp, cudaStream_t st) {
    if (arr == nullptr || n == 0) return;
    gather_kernel<T><<<(n + 255) / 256, 256, 0, st>>>(arr, perm, n, (T*)tmp);
    count_launch();
    cudaMemcpyAsync(arr, tmp, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st);
  }

  // up to four 4-byte arrays through the same permutation: the index is read once
  struct Quad {
    uint32_t* p[4];
  };

  // four consecutive destinations per thread: one 128-bit index load, up to sixteen gathers in
  // flight, 128-bit streaming stores (dst is 16-byte aligned scratch)
  // tag_src / tag_dst (optional): the int16 tags ride along with the last group of arrays
  __global__ void __launch_bounds__(256)
    gather4_kernel(Quad src, Quad dst, int na, const uint32_t* __restrict__ perm, uint32_t n,
                   const short* __restrict__ tag_src, short* __restrict__ tag_dst) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (q + 3 < n) {
      const uint4 s = __ldcs(reinterpret_cast<const uint4*>(perm + q));
      if (tag_src) {
        short4 t;
        t.x = __ldg(tag_src + s.x), t.y = __ldg(tag_src + s.y);
        t.z = __ldg(tag_src + s.z), t.w = __ldg(tag_src + s.w);
        __stcs(reinterpret_cast<short4*>(tag_dst + q), t);
      }
      uint4       v[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (a < na) {
          const uint32_t* sp = src.p[a];
          v[a] = make_uint4(__ldg(sp + s.x), __ldg(sp + s.y), __ldg(sp + s.z), __ldg(sp + s.w));
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (a < na) __stcs(reinterpret_cast<uint4*>(dst.p[a] + q), v[a]);
      }
    } else {
      for (uint32_t r = q; r < n; ++r) {
        const uint32_t s = perm[r];
        if (tag_src) tag_dst[r] = tag_src[s];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (a < na) dst.p[a][r] = src.p[a][s];
        }
      }
    }
  }

  // `tag`: permuted together with the first group (bounce buffer tag_tmp); returns whether it was
  static bool permute_words(uint32_t* const* arrs, int na, const uint32_t* perm, uint32_t n,
                            char* tmp, size_t stride, short* tag, short* tag_tmp, cudaStream_t st) {
    bool tag_done = false;
    for (int a0 = 0; a0 < na; a0 += 4) {
      const int m = (na - a0 < 4) ? (na - a0) : 4;
      Quad      src {}, dst {};
      for (int a = 0; a < m; ++a) {
        src.p[a] = arrs[a0 + a];
        dst.p[a] = (uint32_t*)(tmp + (size_t)a * stride);
      }
      const bool with_tag = tag && tag_tmp && a0 == 0;
      gather4_kernel<<<(n / 4 + 1 + 255) / 256, 256, 0, st>>>(src, dst, m, perm, n,
                                                             with_tag ? tag : nullptr, tag_tmp);
      count_launch();
      for (int a = 0; a < m; ++a) {
        cudaMemcpyAsync(src.p[a], dst.p[a], (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
      }
      if (with_tag) {
        cudaMemcpyAsync(tag, tag_tmp, (size_t)n * sizeof(short), cudaMemcpyDeviceToDevice, st);
        tag_done = true;
      }
    }
    return tag_done;
  }

  cudaError_t sort_particles(const eb200_grid_t& g, const eb200_prtls_t& S, uint32_t npart,
                             uint32_t maxnpart, int flags, uint32_t* n_alive_out,
                             Scratch& scratch, cudaStream_t st) {
    const int  remove_dead = flags & 1;
    const bool skip_prev   = (flags & EB200_SORT_SKIP_PREV) != 0;
    bool       unstable    = (flags & EB200_SORT_UNSTABLE) != 0;
    if (n_alive_out) *n_alive_out = npart;
    if (npart == 0) return cudaSuccess;
    if (S.npld_r < 0 || S.npld_i < 0 || S.npld_r > EB200_MAX_PLD || S.npld_i > EB200_MAX_PLD) {
      return cudaErrorInvalidValue;
    }
    const int      n1 = g.n[0], n2 = g.dim > 1 ? g.n[1] : 1, n3 = g.dim > 2 ? g.n[2] : 1;
    const uint64_t nc64 = (uint64_t)n1 * n2 * n3;
    if (nc64 >= 0xFFFFFFFFull) return cudaErrorInvalidValue;
    const uint32_t ncells = (uint32_t)nc64;
    int            bits   = 1;
    while ((1ull << bits) <= ncells) ++bits;

    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (size_t)npart, 0,
                                    bits, st);
    // the scratch layout is sized by the arrays' capacity, not by npart: npart moves from call to
    // call (migration, species of different size) and every growth of a 5 GB buffer is a
    // cudaFree + cudaMalloc (tens of milliseconds)
    const size_t ncap = (maxnpart >= npart) ? (size_t)maxnpart : (size_t)npart + npart / 8;
    const size_t n4   = (ncap * 4 + 255) / 256 * 256;
    {
      size_t cap_tmp = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, cap_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                      (uint32_t*)nullptr, (uint32_t*)nullptr, ncap, 0, bits, st);
      if (cap_tmp > tmp_bytes) tmp_bytes = cap_tmp;
    }
    // the histogram lives in the k1 / i0 quarters (2 * n4 bytes); radix sort when it does not fit
    if (((size_t)ncells + 2) * 4 > 2 * n4) unstable = false;
    if (unstable) {
      size_t scan_tmp = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (int)(ncells + 2), st);
      if (scan_tmp > tmp_bytes) tmp_bytes = scan_tmp;
    }
    cudaError_t  err = scratch.reserve(8 * n4 + tmp_bytes + 512);
    if (err != cudaSuccess) return err;
    char*     base  = (char*)scratch.ptr;
    uint32_t* k0    = (uint32_t*)(base);
    uint32_t* k1    = (uint32_t*)(base + n4);
    uint32_t* i0    = (uint32_t*)(base + 2 * n4);
    uint32_t* perm  = (uint32_t*)(base + 3 * n4);
    void*     tmp   = (void*)(base + 4 * n4);
    uint32_t* count = (uint32_t*)(base + 8 * n4); // tmp holds four arrays: 4 * n4 .. 8 * n4
    void*     cubws = (void*)(base + 8 * n4 + 256);

    // EB200_SORT_TRACE=1: stage times on stderr (synchronises; diagnosis only)
    static const bool trace = getenv("EB200_SORT_TRACE") != nullptr;
    cudaEvent_t       tev[5] = {};
    auto              mark   = [&](int k) {
      if (trace) {
        if (!tev[k]) cudaEventCreate(&tev[k]);
        cudaEventRecord(tev[k], st);
      }
    };
    mark(0);
    const unsigned nb = (npart + 255) / 256;
    if (unstable) {
      // k1 .. : histogram / slot counters of ncells + 1 buckets (the last one = not alive)
      uint32_t* hist = k1;
      err = cudaMemsetAsync(hist, 0, ((size_t)ncells + 2) * 4, st);
      if (err != cudaSuccess) return err;
      switch (g.dim) {
        case 1: cell_hist_kernel<1><<<nb, 256, 0, st>>>(S, npart, n1, n2, n3, ncells, k0, hist); break;
        case 2: cell_hist_kernel<2><<<nb, 256, 0, st>>>(S, npart, n1, n2, n3, ncells, k0, hist); break;
        case 3: cell_hist_kernel<3><<<nb, 256, 0, st>>>(S, npart, n1, n2, n3, ncells, k0, hist); break;
        default: return cudaErrorInvalidValue;
      }
      count_launch();
      mark(1);
      err = cub::DeviceScan::ExclusiveSum(cubws, tmp_bytes, hist, hist, (int)(ncells + 2), st);
      if (err != cudaSuccess) return err;
      count_launch();
      if (remove_dead && n_alive_out) {
        // first slot of the not-alive bucket = number of alive particles
        err = cudaMemcpyAsync(count, hist + ncells, 4, cudaMemcpyDeviceToDevice, st);
        if (err != cudaSuccess) return err;
      }
      cell_slots_kernel<<<nb, 256, 0, st>>>(k0, npart, hist, perm);
      count_launch();
      mark(2);
    } else {
      switch (g.dim) {
        case 1:
          sort_keys_kernel<1><<<nb, 256, 0, st>>>(S, npart, n1, n2, n3, ncells, k0, i0);
          break;
        case 2:
          sort_keys_kernel<2><<<nb, 256, 0, st>>>(S, npart, n1, n2, n3, ncells, k0, i0);
          break;
        case 3:
          sort_keys_kernel<3><<<nb, 256, 0, st>>>(S, npart, n1, n2, n3, ncells, k0, i0);
          break;
        default: return cudaErrorInvalidValue;
      }
      count_launch();
      mark(1);
      err = cub::DeviceRadixSort::SortPairs(cubws, tmp_bytes, k0, k1, i0, perm, (size_t)npart, 0,
                                            bits, st);
      if (err != cudaSuccess) return err;
      count_launch();
      mark(2);
      if (remove_dead && n_alive_out) {
        count_alive_kernel<<<1, 1, 0, st>>>(k1, npart, ncells, count);
        count_launch();
      }
    }

    bool tag_done = false;
    {
      uint32_t* w[20 + 2 * EB200_MAX_PLD];
      int       nw = 0;
      auto      add = [&](void* q) {
        if (q) w[nw++] = (uint32_t*)q;
      };
      add(S.i1), add(S.dx1);
      if (g.dim > 1) add(S.i2), add(S.dx2);
      if (g.dim > 2) add(S.i3), add(S.dx3);
      if (!skip_prev) {
        add(S.i1_prev), add(S.dx1_prev);
        if (g.dim > 1) add(S.i2_prev), add(S.dx2_prev);
        if (g.dim > 2) add(S.i3_prev), add(S.dx3_prev);
      }
      add(S.ux1), add(S.ux2), add(S.ux3), add(S.weight);
      add(S.phi);
      // payload planes travel with their particles (particles_sort.cpp:150-160, 239-247)
      if (S.pld_r) {
        for (int k = 0; k < S.npld_r; ++k) add(S.pld_r + (size_t)k * S.pld_stride);
      }
      if (S.pld_i) {
        for (int k = 0; k < S.npld_i; ++k) add(S.pld_i + (size_t)k * S.pld_stride);
      }
      // the keys (k0) are dead once the permutation exists: their quarter bounces the tags
      tag_done = permute_words(w, nw, perm, npart, (char*)tmp, n4, S.tag, (short*)k0, st);
    }
    mark(3);
    if (!tag_done) permute(S.tag, perm, npart, tmp, st);
    mark(4);
    if (trace) {
      cudaEventSynchronize(tev[4]);
      float t[4];
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], tev[k], tev[k + 1]);
      fprintf(stderr, "[eb200 sort] npart %u cap %zu bits %d: keys %.2f radix %.2f gather %.2f tag %.2f ms\n",
              npart, ncap, bits, t[0], t[1], t[2], t[3]);
      for (auto& e : tev) cudaEventDestroy(e);
    }

    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    if (remove_dead && n_alive_out) {
      // the one host-visible result of this call: the new particle count
      uint32_t h = 0;
      err = cudaMemcpyAsync(&h, count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
      if (err != cudaSuccess) return err;
      err = cudaStreamSynchronize(st);
      if (err != cudaSuccess) return err;
      *n_alive_out = h;
    }
    return cudaSuccess;
  }

} // namespace eb200
