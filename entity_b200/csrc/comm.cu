// entity_b200 -- multi-domain exchange: decomposition tables, field halo exchange, additive
// current synchronisation and particle migration over NCCL (NVLink / NVSwitch).
//
// Replaces, for one domain per GPU,
//   tools::Decompose                         src/global/utils/tools.h:84-275
//   Metadomain::{createEmptyDomains, redefineNeighbors, redefineBoundaries}
//                                            src/framework/domain/metadomain.cpp:101-330
//   Metadomain::CommunicateFields / SynchronizeFields / CommunicateParticles
//                                            src/framework/domain/metadomain_comm.cpp:205-653
//   comm::CommunicateField (MPI_Sendrecv per direction and array)
//                                            src/framework/domain/comm_mpi.hpp:142-366
//   Particles::Communicate + kernel::comm::* src/framework/containers/particles_comm.cpp:180-389,
//                                            src/kernels/comm.hpp
//
// Design (B200): every exchange round is  pack kernel -> ONE grouped ncclSend/ncclRecv per
// peer -> unpack kernel  on the caller's stream. All directions that lead to the same peer
// (with 2 domains per dimension +d and -d are the same GPU) and all components travel in one
// message; buffers are persistent; nothing is allocated per call and the host never touches
// field data. Particle counts travel through a first tiny grouped exchange and are read back
// once (the only synchronisation of a step; the reference has one MPI_Sendrecv of counts per
// direction and species).
//
// Message layout (both ends derive it from the decomposition alone): for peer p, the segments
// of the iteration directions d in ascending index for which neighbour(d) == p, where the
// sender contributes its slab on the +d side and the receiver deposits it on its -d side --
// i.e. exactly the (send_slice, recv_slice) pairs of GetSendRecvParams (metadomain_comm.cpp:
// 118-203) in the order of dir::Directions<D>::all.
#include "common.cuh"
#include "launch.h"

#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <nccl.h> // types only; the library is resolved at run time (no link dependency)

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace eb200 {

  /* ------------------------------------------------------------------ NCCL binding */
  struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                             = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)      = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                = nullptr;
    ncclResult_t (*GroupStart)()                                           = nullptr;
    ncclResult_t (*GroupEnd)()                                             = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)       = nullptr;
    const char* (*GetErrorString)(ncclResult_t)                            = nullptr;

    bool load(std::string& err) {
      if (handle) return true;
      // a copy already mapped by the host process (torch's bundled NCCL) wins
      const char* names[] = { getenv("EB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
      for (const char* nm : names) {
        if (!nm || !*nm) continue;
        handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (handle) break;
      }
      for (const char* nm : names) {
        if (handle) break;
        if (!nm || !*nm) continue;
        handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      }
      if (!handle) {
        err = std::string("cannot load NCCL: ") + dlerror();
        return false;
      }
#define SYM(field, name)                                                                       \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                              \
  if (!field) {                                                                                \
    err = std::string("NCCL symbol missing: ") + name;                                        \
    return false;                                                                              \
  }
      SYM(GetUniqueId, "ncclGetUniqueId")
      SYM(CommInitRank, "ncclCommInitRank")
      SYM(CommDestroy, "ncclCommDestroy")
      SYM(GroupStart, "ncclGroupStart")
      SYM(GroupEnd, "ncclGroupEnd")
      SYM(Send, "ncclSend")
      SYM(Recv, "ncclRecv")
      SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
      return true;
    }
  };

  static NcclApi g_nccl;

  /* --------------------------------------------------------------- decomposition */
  // tools::decompose1D, tools.h:84-103
  static bool decompose1d(unsigned nd, int ncells, std::vector<int>& out) {
    if (nd == 0 || ncells <= 0) return false;
    const int size = (int)((double)ncells / (double)nd);
    out.assign(nd, size);
    for (int i = 0; i < ncells - size * (int)nd; ++i) out[i] += 1;
    for (unsigned d = 0; d < nd; ++d) {
      if (out[d] < 5) return false; // "ncells < 5"
    }
    return true;
  }

  // tools::divideInProportions2D, tools.h:111-126
  static bool proportions2d(unsigned ntot, unsigned s1, unsigned s2, unsigned& n1, unsigned& n2) {
    n1 = (unsigned)(std::sqrt((double)ntot * (double)s1 / (double)s2));
    if (n1 == 0) {
      n1 = 1;
      n2 = ntot;
      return true;
    } else if (n1 > ntot) {
      n1 = ntot;
      n2 = 1;
      return true;
    }
    while (ntot % n1 != 0) {
      n1++;
      if (n1 > ntot) return false;
    }
    n2 = ntot / n1;
    return true;
  }

  // tools::divideInProportions3D, tools.h:135-153
  static bool proportions3d(unsigned ntot, unsigned s1, unsigned s2, unsigned s3, unsigned& n1,
                            unsigned& n2, unsigned& n3) {
    n1 = (unsigned)(std::cbrt((double)ntot * (double)((double)s1 * s1) / (double)((double)s2 * s3)));
    if (n1 > ntot) {
      n1 = ntot;
      n2 = 1;
      n3 = 1;
      return true;
    } else if (n1 == 0) {
      n1 = 1;
      return proportions2d(ntot, s2, s3, n2, n3);
    }
    while (ntot % n1 != 0) {
      n1++;
      if (n1 > ntot) return false;
    }
    return proportions2d(ntot / n1, s2, s3, n2, n3);
  }

  struct Metadomain {
    int              D = 0, ndir = 0;
    int              rank = 0, nranks = 1;
    int              ndoms[3] = { 1, 1, 1 };
    std::vector<int> ext[3];
    eb200_domain_info_t info;
    int              dvec[27][3];
    int              nbr_n[27][3]; // active cells of the neighbour in each direction
  };

  static int ipow3(int d) { return d == 1 ? 3 : (d == 2 ? 9 : 27); }

  static void dir_of(int D, int lin, int* d) {
    d[0] = d[1] = d[2] = 0;
    for (int a = D - 1; a >= 0; --a) {
      d[a] = (lin % 3) - 1;
      lin /= 3;
    }
  }

  static bool is_comm_bc(int b) { return b == EB200_FBC_SYNC || b == EB200_FBC_PERIODIC; }

  static const char* build_metadomain(const eb200_metadomain_t& md, Metadomain& M) {
    if (md.dim < 1 || md.dim > 3) return "dim must be 1..3";
    if (md.nranks < 1 || md.rank < 0 || md.rank >= md.nranks) return "bad rank / nranks";
    M.D      = md.dim;
    M.ndir   = ipow3(md.dim);
    M.rank   = md.rank;
    M.nranks = md.nranks;
    long prod = 1;
    for (int a = 0; a < 3; ++a) {
      M.ndoms[a] = (a < md.dim) ? md.ndoms[a] : 1;
      if (M.ndoms[a] < 1) return "ndoms must be positive";
      prod *= M.ndoms[a];
      M.ext[a].clear();
      if (a < md.dim) {
        if (!md.extents[a]) return "extents missing";
        M.ext[a].assign(md.extents[a], md.extents[a] + M.ndoms[a]);
      } else {
        M.ext[a].assign(1, 1);
      }
    }
    if (prod != md.nranks) return "product of ndoms != nranks";
    eb200_domain_info_t& I = M.info;
    std::memset(&I, 0, sizeof(I));
    // index -> offsets, first dimension fastest (tools::TensorProduct)
    int r = md.rank;
    for (int a = 0; a < 3; ++a) {
      I.offset[a] = r % M.ndoms[a];
      r /= M.ndoms[a];
      I.n[a]           = M.ext[a][I.offset[a]];
      I.cell_offset[a] = 0;
      for (int k = 0; k < I.offset[a]; ++k) I.cell_offset[a] += M.ext[a][k];
    }
    auto rank_of = [&](const int* o) { return o[0] + M.ndoms[0] * (o[1] + M.ndoms[1] * o[2]); };
    // neighbours with rollover (redefineNeighbors, metadomain.cpp:196-232)
    const int centre = (M.ndir - 1) / 2;
    for (int lin = 0; lin < 27; ++lin) {
      I.neighbor[lin] = -1;
      I.enabled[lin]  = 0;
      I.dir_fbc[lin]  = EB200_FBC_NONE;
    }
    for (int lin = 0; lin < M.ndir; ++lin) {
      int* d = M.dvec[lin];
      dir_of(M.D, lin, d);
      int o[3] = { I.offset[0], I.offset[1], I.offset[2] };
      for (int a = 0; a < M.D; ++a) {
        o[a] = (o[a] + d[a] + M.ndoms[a]) % M.ndoms[a];
      }
      I.neighbor[lin] = rank_of(o);
      for (int a = 0; a < 3; ++a) M.nbr_n[lin][a] = M.ext[a][o[a]];
    }
    // faces (redefineBoundaries, metadomain.cpp:234-283)
    for (int a = 0; a < M.D; ++a) {
      for (int side = 0; side < 2; ++side) {
        const bool edge = side == 0 ? (I.offset[a] == 0) : (I.offset[a] == M.ndoms[a] - 1);
        int        fb = edge ? md.fbc[2 * a + side] : EB200_FBC_SYNC;
        int        pb = edge ? md.pbc[2 * a + side] : EB200_PBC_NONE;
        // periodic towards a different domain becomes SYNC
        if (M.ndoms[a] > 1) {
          if (fb == EB200_FBC_PERIODIC) fb = EB200_FBC_SYNC;
          if (pb == EB200_PBC_PERIODIC) pb = EB200_PBC_NONE;
        }
        I.face_fbc[2 * a + side] = fb;
        I.face_pbc[2 * a + side] = pb;
      }
      if ((I.face_fbc[2 * a] == EB200_FBC_PERIODIC) != (I.face_fbc[2 * a + 1] == EB200_FBC_PERIODIC)) {
        return "Periodic boundary conditions must be set in both directions";
      }
    }
    for (int a = M.D; a < 3; ++a) {
      I.face_fbc[2 * a] = I.face_fbc[2 * a + 1] = EB200_FBC_NONE;
      I.face_pbc[2 * a] = I.face_pbc[2 * a + 1] = EB200_PBC_PERIODIC;
    }
    // directions: the first non-periodic associated face decides (metadomain.cpp:284-320)
    for (int lin = 0; lin < M.ndir; ++lin) {
      if (lin == centre) continue;
      const int* d  = M.dvec[lin];
      int        bc = EB200_FBC_PERIODIC;
      for (int a = 0; a < M.D; ++a) {
        if (d[a] == 0) continue;
        const int f = I.face_fbc[2 * a + (d[a] > 0 ? 1 : 0)];
        if (f != EB200_FBC_PERIODIC) {
          bc = f;
          break;
        }
      }
      I.dir_fbc[lin] = bc;
      I.enabled[lin] = is_comm_bc(bc) ? 1 : 0;
      if (bc == EB200_FBC_PERIODIC && I.neighbor[lin] != md.rank) {
        return "Periodic boundaries imply communication within the same domain";
      }
      if (bc == EB200_FBC_SYNC && I.neighbor[lin] == md.rank) {
        return "Sync boundaries imply communication between separate domains";
      }
    }
    return nullptr;
  }

  /* ------------------------------------------------------------- field exchange */
  struct Seg {
    int  lo[3], ext[3];
    long off;  // first element of the segment in the flat buffer
    long size; // (c1 - c0) * ext0 * ext1 * ext2
  };

  struct SegTable {
    int  nseg;
    int  c0, c1;
    long total;
    Seg  s[26];
  };

  // recv table indexed by iteration direction (sync unpack walks them in order)
  struct DirSegTable {
    int  ndir;
    int  c0, c1;
    int  on[27];
    Seg  s[27];
  };

  template <int D>
  __global__ void __launch_bounds__(256)
    pack_kernel(const __grid_constant__ SegTable T, FieldView<D> F, float* __restrict__ buf) {
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < T.total;
         e += (long)gridDim.x * blockDim.x) {
      int k = 0;
      while (k + 1 < T.nseg && e >= T.s[k + 1].off) ++k;
      const Seg& s   = T.s[k];
      long       r   = e - s.off;
      const int  x   = (int)(r % s.ext[0]);
      r             /= s.ext[0];
      const int y    = (int)(r % s.ext[1]);
      r             /= s.ext[1];
      const int z    = (int)(r % s.ext[2]);
      const int c    = (int)(r / s.ext[2]) + T.c0;
      buf[e]         = F.at(s.lo[0] + x, s.lo[1] + y, s.lo[2] + z, c);
    }
  }

  template <int D>
  __global__ void __launch_bounds__(256)
    unpack_copy_kernel(const __grid_constant__ SegTable T, FieldView<D> F,
                       const float* __restrict__ buf) {
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < T.total;
         e += (long)gridDim.x * blockDim.x) {
      int k = 0;
      while (k + 1 < T.nseg && e >= T.s[k + 1].off) ++k;
      const Seg& s   = T.s[k];
      long       r   = e - s.off;
      const int  x   = (int)(r % s.ext[0]);
      r             /= s.ext[0];
      const int y    = (int)(r % s.ext[1]);
      r             /= s.ext[1];
      const int z    = (int)(r % s.ext[2]);
      const int c    = (int)(r / s.ext[2]) + T.c0;
      F.at(s.lo[0] + x, s.lo[1] + y, s.lo[2] + z, c) = buf[e];
    }
  }

  // SynchronizeFields: buff = 0; per direction buff[recv_slice] += recv; cur += buff on the
  // active cells (metadomain_comm.cpp:409-562). One thread per active cell sums what the
  // directions deliver to it, from zero and in direction order, then adds onto the cell: the
  // same additions in the same order.
  template <int D>
  __global__ void __launch_bounds__(256)
    unpack_add_kernel(const __grid_constant__ DirSegTable T, int n0, int n1, int n2, int G,
                      FieldView<D> F, const float* __restrict__ buf) {
    const long t     = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)n0 * n1 * n2;
    if (t >= total) return;
    const int x[3] = { (int)(t % n0) + G, (D > 1) ? (int)((t / n0) % n1) + G : 0,
                       (D > 2) ? (int)(t / ((long)n0 * n1)) + G : 0 };
    const int nn[3] = { n0, n1, n2 };
    bool      near  = false;
#pragma unroll
    for (int a = 0; a < D; ++a) near = near || (x[a] < 2 * G) || (x[a] >= nn[a]);
    if (!near) return;
    float acc[3] = { ZERO, ZERO, ZERO };
    for (int d = 0; d < T.ndir; ++d) {
      if (!T.on[d]) continue;
      const Seg& s = T.s[d];
      int        q[3];
      bool       in = true;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        q[a] = x[a] - s.lo[a];
        in   = in && q[a] >= 0 && q[a] < s.ext[a];
      }
      if (!in) continue;
      const long vol = (long)s.ext[0] * s.ext[1] * s.ext[2];
      const long r   = q[0] + (long)s.ext[0] * (q[1] + (long)s.ext[1] * q[2]);
      for (int c = T.c0; c < T.c1; ++c) {
        acc[c - T.c0] += buf[s.off + (c - T.c0) * vol + r];
      }
    }
    for (int c = T.c0; c < T.c1; ++c) {
      F.at(x[0], x[1], x[2], c) += acc[c - T.c0];
    }
  }

  /* ---------------------------------------------------------- particle migration */
  constexpr int MAXTAG = 28; // 2 + 26 directions

  // stable multi-split of the non-alive particles by class (0: dead, t - 1: send tag t), the
  // deterministic counterpart of NpartsPerTagAndOffsets + PrepareOutgoingPrtls
  // (particles_sort.cpp:22-67, kernels/comm.hpp:75-105): class-major, index order inside a class.
  constexpr int SPLIT_THREADS = 256;

  __device__ __forceinline__ int class_of(short tag) { return tag == 0 ? 0 : (int)tag - 1; }

  // eight consecutive tags per thread: one 16-byte load where the array allows it. Nearly all
  // tags are 1 (alive, staying): the packed compare skips them eight at a time.
  constexpr int SPLIT_VEC = 8;

  __device__ __forceinline__ void load_tags8(const short* __restrict__ tag, uint32_t p8, uint32_t hi,
                                             bool aligned, short (&t)[SPLIT_VEC]) {
    if (aligned && p8 + SPLIT_VEC <= hi) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(tag + p8));
      const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        t[2 * k]     = (short)(w[k] & 0xffffu);
        t[2 * k + 1] = (short)(w[k] >> 16);
      }
    } else {
#pragma unroll
      for (int k = 0; k < SPLIT_VEC; ++k) t[k] = (p8 + k < hi) ? tag[p8 + k] : (short)1;
    }
  }

  __global__ void __launch_bounds__(SPLIT_THREADS)
    split_count_kernel(const short* __restrict__ tag, uint32_t npart, uint32_t per_block,
                       int nclass, bool aligned, uint32_t* __restrict__ counts /* [nclass][gridDim.x] */) {
    __shared__ uint32_t hist[MAXTAG];
    if (threadIdx.x < MAXTAG) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t lo = blockIdx.x * per_block;
    const uint32_t hi = min(npart, lo + per_block);
    for (uint32_t p8 = lo + threadIdx.x * SPLIT_VEC; p8 < hi; p8 += SPLIT_THREADS * SPLIT_VEC) {
      short t[SPLIT_VEC];
      load_tags8(tag, p8, hi, aligned, t);
#pragma unroll
      for (int k = 0; k < SPLIT_VEC; ++k) {
        if (t[k] != 1) {
          const int c = class_of(t[k]);
          if (c >= 0 && c < nclass) atomicAdd(&hist[c], 1u);
        }
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < nclass) counts[threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];
  }

  // exclusive scan of counts[nclass * nblocks] (class-major) by one block; also emits the
  // per-class totals and bases
  __global__ void __launch_bounds__(1024)
    split_scan_kernel(uint32_t* __restrict__ counts, int nclass, int nblocks,
                      uint32_t* __restrict__ class_total, uint32_t* __restrict__ class_base) {
    __shared__ uint32_t part[1024];
    const int           n     = nclass * nblocks;
    const int           chunk = (n + 1023) / 1024;
    const int           lo    = threadIdx.x * chunk;
    const int           hi    = min(n, lo + chunk);
    uint32_t            sum   = 0;
    for (int k = lo; k < hi; ++k) sum += counts[k];
    part[threadIdx.x] = sum;
    __syncthreads();
    // simple Hillis-Steele inclusive scan over 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t v = 0;
      if ((int)threadIdx.x >= off) v = part[threadIdx.x - off];
      __syncthreads();
      part[threadIdx.x] += v;
      __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (int k = lo; k < hi; ++k) {
      const uint32_t c = counts[k];
      counts[k]        = run;
      run             += c;
    }
    __syncthreads();
    if ((int)threadIdx.x < nclass) {
      const int      c     = threadIdx.x;
      const uint32_t base  = counts[c * nblocks];
      const uint32_t next  = (c + 1 < nclass) ? counts[(c + 1) * nblocks] : part[1023];
      class_base[c]        = base;
      class_total[c]       = next - base;
    }
    if (threadIdx.x == 0) {
      class_base[nclass] = part[1023]; // number of holes
    }
  }

  // ---- classify from the exception list of the fused push (launch.h ExcList) ----------------
  // The pusher already met every particle that is not alive after it and wrote its index down;
  // sorting that short list by (class, index) gives exactly what the two scans over all tags
  // give -- out_idx in class-major, index-ascending order and the class totals / bases.
  __global__ void __launch_bounds__(256)
    exc_keys_kernel(const short* __restrict__ tag, const uint32_t* __restrict__ idx, uint32_t cap,
                    const uint32_t* __restrict__ count, uint64_t* __restrict__ keys) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cap) return;
    uint64_t key = ~0ull; // unused slots sort last
    if (j < min(*count, cap)) {
      const uint32_t p = idx[j];
      const int      t = tag[p];
      key = ((uint64_t)(uint32_t)(t == 0 ? 0 : t - 1) << 32) | p;
    }
    keys[j] = key;
  }

  __global__ void __launch_bounds__(256)
    exc_finish_kernel(const uint64_t* __restrict__ keys, uint32_t cap,
                      const uint32_t* __restrict__ count, int nclass,
                      uint32_t* __restrict__ class_total, uint32_t* __restrict__ class_base,
                      uint32_t* __restrict__ out_idx, uint32_t* __restrict__ overflow) {
    const uint32_t n = min(*count, cap);
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out_idx[j] = (uint32_t)keys[j];
    if (blockIdx.x == 0 && (int)threadIdx.x <= nclass) {
      // first key of class c: lower bound of (c << 32)
      auto lower = [&](uint64_t v) {
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
          const uint32_t mid = lo + (hi - lo) / 2;
          if (keys[mid] < v) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      const int c = threadIdx.x;
      if (c < nclass) {
        const uint32_t b = lower((uint64_t)c << 32), e = lower((uint64_t)(c + 1) << 32);
        class_base[c]  = b;
        class_total[c] = e - b;
      } else {
        class_base[nclass] = n; // number of holes
        if (*count > cap) atomicExch(overflow, 1u);
      }
    }
  }

  // the usual case -- a few hundred entries: one block sorts the list in shared memory and writes
  // out_idx, the class totals and bases (same results as keys + radix sort + finish)
  constexpr int EXC_SMALL = 4096;

  __global__ void __launch_bounds__(1024)
    exc_small_kernel(const short* __restrict__ tag, const uint32_t* __restrict__ idx, uint32_t n,
                     int nclass, uint32_t* __restrict__ class_total,
                     uint32_t* __restrict__ class_base, uint32_t* __restrict__ out_idx) {
    using Sort = cub::BlockRadixSort<uint64_t, 1024, EXC_SMALL / 1024>;
    __shared__ union {
      typename Sort::TempStorage tmp;
      uint64_t                   sorted[EXC_SMALL];
    } sm;
    uint64_t keys[EXC_SMALL / 1024];
#pragma unroll
    for (int k = 0; k < EXC_SMALL / 1024; ++k) {
      const uint32_t j = threadIdx.x * (EXC_SMALL / 1024) + k;
      keys[k]          = ~0ull;
      if (j < n) {
        const uint32_t p = idx[j];
        const int      t = tag[p];
        keys[k] = ((uint64_t)(uint32_t)(t == 0 ? 0 : t - 1) << 32) | p;
      }
    }
    Sort(sm.tmp).Sort(keys, 0, 40);
    __syncthreads(); // the sort's scratch is reused for the sorted list
#pragma unroll
    for (int k = 0; k < EXC_SMALL / 1024; ++k) {
      const uint32_t j = threadIdx.x * (EXC_SMALL / 1024) + k;
      sm.sorted[j]     = keys[k];
      if (j < n) out_idx[j] = (uint32_t)keys[k];
    }
    __syncthreads();
    if ((int)threadIdx.x <= nclass) {
      auto lower = [&](uint64_t v) {
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
          const uint32_t mid = lo + (hi - lo) / 2;
          if (sm.sorted[mid] < v) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      const int c = threadIdx.x;
      if (c < nclass) {
        const uint32_t b = lower((uint64_t)c << 32), e = lower((uint64_t)(c + 1) << 32);
        class_base[c]  = b;
        class_total[c] = e - b;
      } else {
        class_base[nclass] = n;
      }
    }
  }

  // count message: slot q takes class total src_class[q] of species src_species[q]
  struct CountGather {
    int            n;
    const uint32_t* tot[8];      // per species: class totals
    unsigned char  species[256];
    unsigned char  cls[256];
  };

  __global__ void gather_counts_kernel(const __grid_constant__ CountGather G, uint32_t* __restrict__ out) {
    const int q = threadIdx.x;
    if (q < G.n) out[q] = G.tot[G.species[q]][G.cls[q]];
  }

  __global__ void __launch_bounds__(SPLIT_THREADS)
    split_scatter_kernel(const short* __restrict__ tag, uint32_t npart, uint32_t per_block,
                         int nclass, bool aligned,
                         const uint32_t* __restrict__ offsets /* scanned counts */,
                         uint32_t* __restrict__ out_idx) {
    __shared__ uint32_t cursor[MAXTAG];
    __shared__ uint32_t wcount[SPLIT_THREADS / 32][MAXTAG];
    __shared__ short    stag[SPLIT_THREADS * SPLIT_VEC];
    if ((int)threadIdx.x < nclass) cursor[threadIdx.x] = offsets[threadIdx.x * gridDim.x + blockIdx.x];
    __syncthreads();
    const uint32_t lo   = blockIdx.x * per_block;
    const uint32_t hi   = min(npart, lo + per_block);
    const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t chunk = lo; chunk < hi; chunk += SPLIT_THREADS * SPLIT_VEC) {
      // fast path: 2048 tags per block iteration, all alive -> one barrier and on
      short t8[SPLIT_VEC];
      load_tags8(tag, chunk + threadIdx.x * SPLIT_VEC, hi, aligned, t8);
      bool any = false;
#pragma unroll
      for (int k = 0; k < SPLIT_VEC; ++k) any = any || (t8[k] != 1);
      if (!__syncthreads_or(any)) continue;
      // slow path: hand the tags round in index order, 256 at a time (the stable split)
#pragma unroll
      for (int k = 0; k < SPLIT_VEC; ++k) stag[threadIdx.x * SPLIT_VEC + k] = t8[k];
      __syncthreads();
      for (int r = 0; r < SPLIT_VEC; ++r) {
        const uint32_t p = chunk + r * SPLIT_THREADS + threadIdx.x;
        int            c = -1;
        if (p < hi) {
          const short t = stag[r * SPLIT_THREADS + threadIdx.x];
          if (t != 1) {
            c = class_of(t);
            if (c < 0 || c >= nclass) c = -1;
          }
        }
        if (!__syncthreads_or(c >= 0)) continue;
        // rank inside the warp among lanes of the same class
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        const unsigned below = peers & ((1u << lane) - 1u);
        for (int k = lane; k < MAXTAG; k += 32) wcount[warp][k] = 0;
        __syncwarp();
        if (c >= 0 && below == 0) wcount[warp][c] = __popc(peers);
        __syncthreads();
        if (c >= 0) {
          uint32_t before = 0;
          for (int w = 0; w < warp; ++w) before += wcount[w][c];
          out_idx[cursor[c] + before + __popc(below)] = p;
        }
        __syncthreads();
        if ((int)threadIdx.x < nclass) {
          uint32_t tot = 0;
          for (int w = 0; w < SPLIT_THREADS / 32; ++w) tot += wcount[w][threadIdx.x];
          cursor[threadIdx.x] += tot;
        }
        __syncthreads();
      }
    }
  }

  // per species and direction: where the outgoing particles go
  struct SendPlan {
    int      nclass;
    uint32_t class_base[MAXTAG + 1]; // into out_idx (host copy of the device scan)
    long     seg_off[MAXTAG];        // word offset of the class's segment in the send buffer (-1: not sent)
    int      shift[MAXTAG][3];       // index shift applied on the way out
  };

  // payload words of a record (pld_r then pld_i; kernels/comm.hpp:75-105 sends them too)
  __host__ __device__ inline int npld_words(const eb200_prtls_t& S) {
    return (S.pld_r ? S.npld_r : 0) + (S.pld_i ? S.npld_i : 0);
  }

  // PopulatePrtlSendBuffer (+ the index shifts of PrepareOutgoingPrtls): one record of NW
  // 32-bit words per particle: [i, i_prev] x D, [dx, dx_prev] x D, ux1..3, weight (, phi)
  // (, payloads)
  template <int D>
  __global__ void __launch_bounds__(256)
    prtl_pack_kernel(const __grid_constant__ SendPlan P, eb200_prtls_t S,
                     const uint32_t* __restrict__ out_idx, uint32_t nholes, int has_phi,
                     uint32_t* __restrict__ buf) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x + P.class_base[1];
    if (q >= nholes) return;
    int c = 1;
    while (c + 1 < P.nclass && q >= P.class_base[c + 1]) ++c;
    const uint32_t p = out_idx[q];
    if (P.seg_off[c] >= 0) {
      const int      NB  = 4 * D + 4 + (has_phi ? 1 : 0);
      const int      NW  = NB + npld_words(S);
      uint32_t*      rec = buf + P.seg_off[c] + (long)(q - P.class_base[c]) * NW;
      const int*     ii[3]  = { S.i1, S.i2, S.i3 };
      const int*     iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
      const float*   dd[3]  = { S.dx1, S.dx2, S.dx3 };
      const float*   ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
      for (int a = 0; a < D; ++a) {
        rec[2 * a]             = (uint32_t)(ii[a][p] + P.shift[c][a]);
        rec[2 * a + 1]         = (uint32_t)(iip[a][p] + P.shift[c][a]);
        rec[2 * D + 2 * a]     = __float_as_uint(dd[a][p]);
        rec[2 * D + 2 * a + 1] = __float_as_uint(ddp[a][p]);
      }
      rec[4 * D + 0] = __float_as_uint(S.ux1[p]);
      rec[4 * D + 1] = __float_as_uint(S.ux2[p]);
      rec[4 * D + 2] = __float_as_uint(S.ux3[p]);
      rec[4 * D + 3] = __float_as_uint(S.weight[p]);
      if (has_phi) rec[4 * D + 4] = __float_as_uint(S.phi[p]);
      int w = NB;
      if (S.pld_r) {
        for (int k = 0; k < S.npld_r; ++k) rec[w++] = __float_as_uint(S.pld_r[p + (size_t)k * S.pld_stride]);
      }
      if (S.pld_i) {
        for (int k = 0; k < S.npld_i; ++k) rec[w++] = S.pld_i[p + (size_t)k * S.pld_stride];
      }
    }
    // sent (or unsendable) particles leave the domain
    S.tag[p] = 0;
  }

  struct RecvPlan {
    int      nseg;
    uint32_t first[MAXTAG + 1]; // first received-particle rank of each segment
    long     seg_off[MAXTAG];   // word offset in the receive buffer
  };

  // ExtractReceivedPrtls: holes first (dead slots, then the slots just vacated), then append
  template <int D>
  __global__ void __launch_bounds__(256)
    prtl_unpack_kernel(const __grid_constant__ RecvPlan R, eb200_prtls_t S,
                       const uint32_t* __restrict__ out_idx, uint32_t nholes, uint32_t npart,
                       uint32_t nrecv, int has_phi, const uint32_t* __restrict__ buf) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrecv) return;
    int k = 0;
    while (k + 1 < R.nseg && r >= R.first[k + 1]) ++k;
    const int       NB  = 4 * D + 4 + (has_phi ? 1 : 0);
    const int       NW  = NB + npld_words(S);
    const uint32_t* rec = buf + R.seg_off[k] + (long)(r - R.first[k]) * NW;
    const uint32_t  p   = (r < nholes) ? out_idx[r] : (npart + r - nholes);
    int*            ii[3]  = { S.i1, S.i2, S.i3 };
    int*            iip[3] = { S.i1_prev, S.i2_prev, S.i3_prev };
    float*          dd[3]  = { S.dx1, S.dx2, S.dx3 };
    float*          ddp[3] = { S.dx1_prev, S.dx2_prev, S.dx3_prev };
#pragma unroll
    for (int a = 0; a < D; ++a) {
      ii[a][p]  = (int)rec[2 * a];
      iip[a][p] = (int)rec[2 * a + 1];
      dd[a][p]  = __uint_as_float(rec[2 * D + 2 * a]);
      ddp[a][p] = __uint_as_float(rec[2 * D + 2 * a + 1]);
    }
    S.ux1[p]    = __uint_as_float(rec[4 * D + 0]);
    S.ux2[p]    = __uint_as_float(rec[4 * D + 1]);
    S.ux3[p]    = __uint_as_float(rec[4 * D + 2]);
    S.weight[p] = __uint_as_float(rec[4 * D + 3]);
    if (has_phi) S.phi[p] = __uint_as_float(rec[4 * D + 4]);
    int w = NB;
    if (S.pld_r) {
      for (int k = 0; k < S.npld_r; ++k) S.pld_r[p + (size_t)k * S.pld_stride] = __uint_as_float(rec[w++]);
    }
    if (S.pld_i) {
      for (int k = 0; k < S.npld_i; ++k) S.pld_i[p + (size_t)k * S.pld_stride] = rec[w++];
    }
    S.tag[p] = 1;
  }

  /* ------------------------------------------------------------------ communicator */
  struct Comm {
    Metadomain       M;
    eb200_grid_t     grid;
    ncclComm_t       nccl = nullptr;
    std::vector<int> peers; // distinct neighbour ranks (self included when periodic onto itself)
    Scratch          sendbuf, recvbuf, work;
    uint32_t*        pinned = nullptr; // host-visible counts
    std::string      err;

    ~Comm() {
      if (nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(nccl);
      sendbuf.release();
      recvbuf.release();
      work.release();
      if (pinned) cudaFreeHost(pinned);
    }
  };

  static int fail(Comm& C, int code, const std::string& msg) {
    C.err = msg;
    return code;
  }

#define CU(C, expr)                                                                            \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) return fail((C), EB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

#define NC(C, expr)                                                                            \
  do {                                                                                         \
    ncclResult_t r_ = (expr);                                                                  \
    if (r_ != ncclSuccess) return fail((C), EB200_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r_)); \
  } while (0)

  // exchange of per-peer messages laid out back to back: offsets in elements of `esize` bytes
  static int exchange(Comm& C, const std::vector<long>& soff, const std::vector<long>& scnt,
                      const std::vector<long>& roff, const std::vector<long>& rcnt,
                      size_t esize, cudaStream_t st) {
    const char* sb = (const char*)C.sendbuf.ptr;
    char*       rb = (char*)C.recvbuf.ptr;
    bool        any_remote = false;
    for (size_t k = 0; k < C.peers.size(); ++k) {
      if (C.peers[k] == C.M.rank) {
        if (scnt[k] != rcnt[k]) return fail(C, EB200_ERR_ARG, "self message size mismatch");
        if (scnt[k] > 0) {
          CU(C, cudaMemcpyAsync(rb + roff[k] * esize, sb + soff[k] * esize, scnt[k] * esize,
                                cudaMemcpyDeviceToDevice, st));
        }
      } else if (scnt[k] > 0 || rcnt[k] > 0) {
        any_remote = true;
      }
    }
    if (!any_remote) return EB200_OK;
    if (!C.nccl) return fail(C, EB200_ERR_NCCL, "no NCCL communicator attached to this context");
    NC(C, g_nccl.GroupStart());
    for (size_t k = 0; k < C.peers.size(); ++k) {
      if (C.peers[k] == C.M.rank) continue;
      if (scnt[k] > 0) {
        NC(C, g_nccl.Send(sb + soff[k] * esize, scnt[k] * esize, ncclChar, C.peers[k], C.nccl, st));
      }
      if (rcnt[k] > 0) {
        NC(C, g_nccl.Recv(rb + roff[k] * esize, rcnt[k] * esize, ncclChar, C.peers[k], C.nccl, st));
      }
    }
    NC(C, g_nccl.GroupEnd());
    return EB200_OK;
  }

  // slabs of GetSendRecvParams (metadomain_comm.cpp:118-203) in ghost-inclusive indices
  static void slab(const eb200_grid_t& g, const int* d, bool sync, bool recv, int* lo, int* ext) {
    for (int a = 0; a < 3; ++a) {
      lo[a]  = 0;
      ext[a] = 1;
    }
    const int G = g.ng;
    for (int a = 0; a < g.dim; ++a) {
      const int n   = g.n[a];
      const int dir = recv ? -d[a] : d[a];
      if (!sync) {
        if (dir == 0) {
          lo[a]  = G;
          ext[a] = n;
        } else if (dir == 1) {
          lo[a]  = recv ? (G + n) : n; // recv: [i_max, i_max + G); send: [i_max - G, i_max)
          ext[a] = G;
        } else {
          lo[a]  = recv ? 0 : G;       // recv: [i_min - G, i_min); send: [i_min, i_min + G)
          ext[a] = G;
        }
      } else {
        if (dir == 0) {
          lo[a]  = 0;
          ext[a] = n + 2 * G;
        } else if (dir == 1) {
          lo[a]  = n; // [i_max - G, i_max + G)
          ext[a] = 2 * G;
        } else {
          lo[a]  = 0; // [i_min - G, i_min + G)
          ext[a] = 2 * G;
        }
      }
    }
  }

  static int halo_round(Comm& C, float* fld, int c0, int c1, bool sync, cudaStream_t st) {
    const Metadomain&   M   = C.M;
    const eb200_grid_t& g   = C.grid;
    const int           nc  = c1 - c0;
    const int           ctr = (M.ndir - 1) / 2;
    SegTable            snd;
    DirSegTable         rcv;
    SegTable            rcv_flat;
    snd.nseg = rcv_flat.nseg = 0;
    snd.c0 = rcv.c0 = rcv_flat.c0 = c0;
    snd.c1 = rcv.c1 = rcv_flat.c1 = c1;
    rcv.ndir                       = M.ndir;
    for (int d = 0; d < 27; ++d) rcv.on[d] = 0;
    std::vector<long> soff(C.peers.size(), 0), scnt(C.peers.size(), 0), roff(C.peers.size(), 0),
      rcnt(C.peers.size(), 0);
    long spos = 0, rpos = 0;
    for (size_t k = 0; k < C.peers.size(); ++k) {
      const int p = C.peers[k];
      soff[k]     = spos;
      roff[k]     = rpos;
      for (int d = 0; d < M.ndir; ++d) {
        if (d == ctr) continue;
        const int md = M.ndir - 1 - d; // index of -d
        if (M.info.enabled[d] && M.info.neighbor[d] == p) {
          Seg& s = snd.s[snd.nseg++];
          slab(g, M.dvec[d], sync, false, s.lo, s.ext);
          s.size = (long)nc * s.ext[0] * s.ext[1] * s.ext[2];
          s.off  = spos;
          spos  += s.size;
        }
        if (M.info.enabled[md] && M.info.neighbor[md] == p) {
          Seg s;
          slab(g, M.dvec[d], sync, true, s.lo, s.ext);
          s.size                      = (long)nc * s.ext[0] * s.ext[1] * s.ext[2];
          s.off                       = rpos;
          rpos                       += s.size;
          rcv.s[d]                    = s;
          rcv.on[d]                   = 1;
          rcv_flat.s[rcv_flat.nseg++] = s;
        }
      }
      scnt[k] = spos - soff[k];
      rcnt[k] = rpos - roff[k];
    }
    snd.total      = spos;
    rcv_flat.total = rpos;
    if (spos == 0 && rpos == 0) return EB200_OK;
    CU(C, C.sendbuf.reserve_grow((size_t)std::max(spos, 1L) * sizeof(float)));
    CU(C, C.recvbuf.reserve_grow((size_t)std::max(rpos, 1L) * sizeof(float)));
    auto nblocks = [](long n) { return (unsigned)std::min<long>((n + 255) / 256, 148L * 16); };
    if (spos > 0) {
      switch (g.dim) {
        case 1: pack_kernel<1><<<nblocks(spos), 256, 0, st>>>(snd, FieldView<1>(g, fld), (float*)C.sendbuf.ptr); break;
        case 2: pack_kernel<2><<<nblocks(spos), 256, 0, st>>>(snd, FieldView<2>(g, fld), (float*)C.sendbuf.ptr); break;
        default: pack_kernel<3><<<nblocks(spos), 256, 0, st>>>(snd, FieldView<3>(g, fld), (float*)C.sendbuf.ptr); break;
      }
      count_launch();
    }
    int rc = exchange(C, soff, scnt, roff, rcnt, sizeof(float), st);
    if (rc != EB200_OK) return rc;
    if (rpos > 0) {
      const float* rb = (const float*)C.recvbuf.ptr;
      if (!sync) {
        switch (g.dim) {
          case 1: unpack_copy_kernel<1><<<nblocks(rpos), 256, 0, st>>>(rcv_flat, FieldView<1>(g, fld), rb); break;
          case 2: unpack_copy_kernel<2><<<nblocks(rpos), 256, 0, st>>>(rcv_flat, FieldView<2>(g, fld), rb); break;
          default: unpack_copy_kernel<3><<<nblocks(rpos), 256, 0, st>>>(rcv_flat, FieldView<3>(g, fld), rb); break;
        }
      } else {
        const int  n0 = g.n[0], n1 = g.dim > 1 ? g.n[1] : 1, n2 = g.dim > 2 ? g.n[2] : 1;
        const long nt = (long)n0 * n1 * n2;
        const unsigned nb = (unsigned)((nt + 255) / 256);
        switch (g.dim) {
          case 1: unpack_add_kernel<1><<<nb, 256, 0, st>>>(rcv, n0, n1, n2, g.ng, FieldView<1>(g, fld), rb); break;
          case 2: unpack_add_kernel<2><<<nb, 256, 0, st>>>(rcv, n0, n1, n2, g.ng, FieldView<2>(g, fld), rb); break;
          default: unpack_add_kernel<3><<<nb, 256, 0, st>>>(rcv, n0, n1, n2, g.ng, FieldView<3>(g, fld), rb); break;
        }
      }
      count_launch();
    }
    CU(C, cudaGetLastError());
    return EB200_OK;
  }

  /* --- entry points used by capi.cu ------------------------------------------------ */
  Comm* comm_new() { return new Comm(); }

  void comm_delete(Comm* c) { delete c; }

  const char* comm_error(const Comm* c) { return c->err.c_str(); }

  bool comm_multi(const Comm* c) { return c != nullptr; }

  const eb200_domain_info_t* comm_info(const Comm* c) { return &c->M.info; }

  int comm_setup(Comm& C, const eb200_grid_t& grid, const eb200_metadomain_t& md, const char* id) {
    const char* e = build_metadomain(md, C.M);
    if (e) return fail(C, EB200_ERR_ARG, std::string("eb200_comm_init: ") + e);
    C.grid = grid;
    if (grid.dim != md.dim) return fail(C, EB200_ERR_ARG, "eb200_comm_init: dimension mismatch");
    for (int a = 0; a < grid.dim; ++a) {
      if (grid.n[a] != C.M.info.n[a]) {
        return fail(C, EB200_ERR_ARG, "eb200_comm_init: the context's grid is not this rank's block");
      }
      // a slab of ghost width is sent from the active region of the neighbour
      for (int k = 0; k < C.M.ndoms[a]; ++k) {
        if (C.M.ext[a][k] < 2 * grid.ng) {
          return fail(C, EB200_ERR_ARG, "eb200_comm_init: a domain is thinner than 2 * N_GHOSTS");
        }
      }
    }
    const int ctr = (C.M.ndir - 1) / 2;
    C.peers.clear();
    for (int d = 0; d < C.M.ndir; ++d) {
      if (d == ctr || !C.M.info.enabled[d]) continue;
      C.peers.push_back(C.M.info.neighbor[d]);
    }
    std::sort(C.peers.begin(), C.peers.end());
    C.peers.erase(std::unique(C.peers.begin(), C.peers.end()), C.peers.end());
    if (id != nullptr && md.nranks > 1) {
      std::string err;
      if (!g_nccl.load(err)) return fail(C, EB200_ERR_NCCL, err);
      ncclUniqueId uid;
      std::memcpy(&uid, id, sizeof(uid));
      NC(C, g_nccl.CommInitRank(&C.nccl, md.nranks, uid, md.rank));
    }
    CU(C, cudaMallocHost((void**)&C.pinned, sizeof(uint32_t) * 4096));
    return EB200_OK;
  }

  int comm_unique_id(char* out, std::string& err) {
    if (!g_nccl.load(err)) return EB200_ERR_NCCL;
    ncclUniqueId uid;
    ncclResult_t r = g_nccl.GetUniqueId(&uid);
    if (r != ncclSuccess) {
      err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
      return EB200_ERR_NCCL;
    }
    std::memcpy(out, &uid, sizeof(uid));
    return EB200_OK;
  }

  int comm_fields(Comm& C, float* fld, int c0, int c1, cudaStream_t st) {
    return halo_round(C, fld, c0, c1, false, st);
  }

  int comm_sync_currents(Comm& C, float* cur, cudaStream_t st) {
    return halo_round(C, cur, 0, 3, true, st);
  }

  // EB200_COMM_TRACE=1: stage times of the migration (stream-synchronised; diagnosis only)
  struct MigTrace {
    bool         on;
    cudaStream_t st;
    int          rank;
    double       t[8];
    int          k = 0;
    std::chrono::steady_clock::time_point last;
    MigTrace(cudaStream_t s, int r) : st(s), rank(r) {
      static const bool env = getenv("EB200_COMM_TRACE") != nullptr;
      on = env;
      if (on) {
        cudaStreamSynchronize(st);
        last = std::chrono::steady_clock::now();
      }
    }
    void mark() {
      if (!on || k >= 8) return;
      cudaStreamSynchronize(st);
      const auto now = std::chrono::steady_clock::now();
      t[k++] = std::chrono::duration<double, std::milli>(now - last).count();
      last   = now;
    }
  };

  // Particles::Communicate for every species in one round
  int comm_particles(Comm& C, eb200_species_t* species, int nspecies, cudaStream_t st,
                     const ExcList* exc) {
    const Metadomain& M   = C.M;
    const int         D   = M.D;
    const int         ctr = (M.ndir - 1) / 2;
    const int         nclass = M.ndir; // dead + (ndir - 1) send tags
    if (nspecies <= 0) return EB200_OK;
    if (nspecies * (2 * MAXTAG + 2) > 4096) return fail(C, EB200_ERR_ARG, "too many species");
    MigTrace trace(st, M.rank);
    // directions in which particles travel: enabled and the particle boundary lets them out
    // (a periodic single-domain dimension wraps inside the pusher and never tags)
    auto tag_of = [&](int d) { return 2 + d - (d > ctr ? 1 : 0); };
    std::vector<int> dirs;
    for (int d = 0; d < M.ndir; ++d) {
      if (d != ctr) dirs.push_back(d);
    }
    // ---- 1. classify: per species out_idx + class totals on the device
    const int      nblk = 148 * 16;
    size_t         work_bytes = 0;
    std::vector<size_t> off_idx(nspecies), off_cnt(nspecies), off_tot(nspecies);
    for (int s = 0; s < nspecies; ++s) {
      off_cnt[s]  = work_bytes;
      work_bytes += sizeof(uint32_t) * (size_t)nclass * nblk;
      off_tot[s]  = work_bytes;
      work_bytes += sizeof(uint32_t) * (2 * MAXTAG + 2);
      work_bytes  = (work_bytes + 255) / 256 * 256;
    }
    // out_idx can hold every particle in the worst case: sized by the species' capacity, which
    // does not change from step to step (npart does: sizing by it reallocated ~1 GB whenever a
    // domain gained a particle, 16 ms per occurrence on the 3.3e8-particle block)
    for (int s = 0; s < nspecies; ++s) {
      off_idx[s]  = work_bytes;
      const uint32_t np  = species[s].npart;
      const uint32_t cap = std::max<uint32_t>(species[s].maxnpart >= np ? species[s].maxnpart : np + np / 8, 1);
      work_bytes += sizeof(uint32_t) * (size_t)cap;
      work_bytes  = (work_bytes + 255) / 256 * 256;
    }
    const size_t off_sendcnt = work_bytes; // counts to peers / from peers (uint32)
    work_bytes += sizeof(uint32_t) * 2 * (size_t)nspecies * MAXTAG * 2;
    work_bytes  = (work_bytes + 255) / 256 * 256;
    // exception-list path: two key arrays + the radix sort's scratch + an overflow flag
    uint32_t exc_cap = 0;
    for (int s = 0; s < nspecies; ++s) {
      if (exc && exc[s].tracked && exc[s].count) exc_cap = std::max(exc_cap, exc[s].cap);
    }
    size_t       exc_tmp  = 0;
    const size_t off_keys = work_bytes;
    if (exc_cap > 0) {
      cub::DeviceRadixSort::SortKeys(nullptr, exc_tmp, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                     (size_t)exc_cap, 0, 40, st);
      work_bytes += 2 * sizeof(uint64_t) * (size_t)exc_cap + exc_tmp + 512;
    }
    const size_t off_ovf = work_bytes;
    work_bytes += 256;
    CU(C, C.work.reserve(work_bytes));
    char* W = (char*)C.work.ptr;
    CU(C, cudaMemsetAsync(W + off_ovf, 0, 4, st));
    // sizes of the exception lists (one read-back; the pushes had to finish before the classify
    // in any case): short lists are sorted by one block, long ones by the device-wide radix sort
    std::vector<uint32_t> exc_n(nspecies, 0);
    {
      bool any = false;
      for (int s = 0; s < nspecies; ++s) {
        if (exc && exc[s].tracked && exc[s].count && species[s].npart > 0) {
          CU(C, cudaMemcpyAsync(C.pinned + 3000 + s, exc[s].count, 4, cudaMemcpyDeviceToHost, st));
          any = true;
        }
      }
      if (any) {
        CU(C, cudaStreamSynchronize(st));
        for (int s = 0; s < nspecies; ++s) {
          if (exc && exc[s].tracked && exc[s].count && species[s].npart > 0) exc_n[s] = C.pinned[3000 + s];
        }
      }
    }
    for (int s = 0; s < nspecies; ++s) {
      const eb200_species_t& sp = species[s];
      uint32_t* cnt = (uint32_t*)(W + off_cnt[s]);
      uint32_t* tot = (uint32_t*)(W + off_tot[s]);
      if (sp.npart == 0) {
        CU(C, cudaMemsetAsync(tot, 0, sizeof(uint32_t) * (2 * MAXTAG + 2), st));
        continue;
      }
      // EB200_EXC_SMALL=0 sends every list through the device-wide sort (tests)
      static const uint32_t small_max = getenv("EB200_EXC_SMALL") ? (uint32_t)atoi(getenv("EB200_EXC_SMALL"))
                                                                  : (uint32_t)EXC_SMALL;
      if (exc && exc[s].tracked && exc[s].count && exc_n[s] <= std::min<uint32_t>(small_max, EXC_SMALL) &&
          small_max > 0) {
        exc_small_kernel<<<1, 1024, 0, st>>>(sp.arrays.tag, exc[s].idx, exc_n[s], nclass, tot, tot + MAXTAG,
                                             (uint32_t*)(W + off_idx[s]));
        count_launch();
        continue;
      }
      if (exc && exc[s].tracked && exc[s].count) {
        uint64_t* k0  = (uint64_t*)(W + off_keys);
        uint64_t* k1  = k0 + exc_cap;
        void*     tmp = (void*)(k1 + exc_cap);
        const uint32_t cap = exc[s].cap;
        exc_keys_kernel<<<(cap + 255) / 256, 256, 0, st>>>(sp.arrays.tag, exc[s].idx, cap, exc[s].count, k0);
        CU(C, cub::DeviceRadixSort::SortKeys(tmp, exc_tmp, k0, k1, (size_t)cap, 0, 40, st));
        exc_finish_kernel<<<(cap + 255) / 256, 256, 0, st>>>(k1, cap, exc[s].count, nclass, tot, tot + MAXTAG,
                                                            (uint32_t*)(W + off_idx[s]),
                                                            (uint32_t*)(W + off_ovf));
        count_launch();
        count_launch();
        count_launch();
        continue;
      }
      constexpr uint32_t CH = SPLIT_THREADS * SPLIT_VEC; // tags per block iteration
      const uint32_t per_block = ((sp.npart + nblk - 1) / nblk + CH - 1) / CH * CH;
      const bool     aligned   = (reinterpret_cast<uintptr_t>(sp.arrays.tag) & 15u) == 0;
      split_count_kernel<<<nblk, SPLIT_THREADS, 0, st>>>(sp.arrays.tag, sp.npart, per_block, nclass,
                                                         aligned, cnt);
      split_scan_kernel<<<1, 1024, 0, st>>>(cnt, nclass, nblk, tot, tot + MAXTAG);
      split_scatter_kernel<<<nblk, SPLIT_THREADS, 0, st>>>(sp.arrays.tag, sp.npart, per_block, nclass,
                                                           aligned, cnt, (uint32_t*)(W + off_idx[s]));
      count_launch();
      count_launch();
      count_launch();
    }
    trace.mark(); // classify
    // ---- 2. counts to the peers (device to device), then one read-back
    // message to peer p: for s, for d ascending with neighbour(d) == p: total of class tag(d)-1
    uint32_t* d_sendcnt = (uint32_t*)(W + off_sendcnt);
    uint32_t* d_recvcnt = d_sendcnt + (size_t)nspecies * MAXTAG * 2;
    std::vector<long> soff(C.peers.size(), 0), scnt(C.peers.size(), 0), roff(C.peers.size(), 0),
      rcnt(C.peers.size(), 0);
    // slot tables: (peer k, species s, direction d) -> slot in the count message
    struct Slot { int s, d; };
    std::vector<std::vector<Slot>> sslots(C.peers.size()), rslots(C.peers.size());
    {
      long spos = 0, rpos = 0;
      for (size_t k = 0; k < C.peers.size(); ++k) {
        soff[k] = spos;
        roff[k] = rpos;
        for (int s = 0; s < nspecies; ++s) {
          for (int d : dirs) {
            const int md = M.ndir - 1 - d;
            if (M.info.enabled[d] && M.info.neighbor[d] == C.peers[k]) {
              sslots[k].push_back({ s, d });
              ++spos;
            }
            if (M.info.enabled[md] && M.info.neighbor[md] == C.peers[k]) {
              rslots[k].push_back({ s, d });
              ++rpos;
            }
          }
        }
        scnt[k] = spos - soff[k];
        rcnt[k] = rpos - roff[k];
      }
      // gather the class totals into the count message with tiny D2D copies
      CU(C, C.sendbuf.reserve_grow(std::max<size_t>((size_t)spos * 4, 256)));
      CU(C, C.recvbuf.reserve_grow(std::max<size_t>((size_t)rpos * 4, 256)));
      long pos = 0;
      if (spos <= 256 && nspecies <= 8) {
        CountGather G {};
        for (int s = 0; s < nspecies; ++s) G.tot[s] = (const uint32_t*)(W + off_tot[s]);
        for (size_t k = 0; k < C.peers.size(); ++k) {
          for (const Slot& sl : sslots[k]) {
            G.species[pos] = (unsigned char)sl.s;
            G.cls[pos]     = (unsigned char)(tag_of(sl.d) - 1);
            ++pos;
          }
        }
        G.n = (int)pos;
        if (pos > 0) {
          gather_counts_kernel<<<1, 256, 0, st>>>(G, (uint32_t*)C.sendbuf.ptr);
          count_launch();
        }
      } else {
        for (size_t k = 0; k < C.peers.size(); ++k) {
          for (const Slot& sl : sslots[k]) {
            const uint32_t* tot = (const uint32_t*)(W + off_tot[sl.s]);
            CU(C, cudaMemcpyAsync((uint32_t*)C.sendbuf.ptr + pos, tot + (tag_of(sl.d) - 1), 4,
                                  cudaMemcpyDeviceToDevice, st));
            ++pos;
          }
        }
      }
      int rc = exchange(C, soff, scnt, roff, rcnt, 4, st);
      if (rc != EB200_OK) return rc;
      // read back: per species class bases (MAXTAG + 1) and the received counts
      uint32_t* h = C.pinned;
      for (int s = 0; s < nspecies; ++s) {
        CU(C, cudaMemcpyAsync(h + (size_t)s * (MAXTAG + 1), (uint32_t*)(W + off_tot[s]) + MAXTAG,
                              sizeof(uint32_t) * (MAXTAG + 1), cudaMemcpyDeviceToHost, st));
      }
      if (rpos > 0) {
        CU(C, cudaMemcpyAsync(h + (size_t)nspecies * (MAXTAG + 1), C.recvbuf.ptr, (size_t)rpos * 4,
                              cudaMemcpyDeviceToHost, st));
      }
      // (the last word of the pinned block: did an exception list overflow its capacity?)
      CU(C, cudaMemcpyAsync(h + 4095, W + off_ovf, 4, cudaMemcpyDeviceToHost, st));
      CU(C, cudaStreamSynchronize(st));
      (void)d_sendcnt;
      (void)d_recvcnt;
    }
    const bool exc_overflow = C.pinned[4095] != 0;
    trace.mark(); // counts
    const uint32_t* hbase = C.pinned;
    const uint32_t* hrecv = C.pinned + (size_t)nspecies * (MAXTAG + 1);
    // ---- 3. plans
    std::vector<SendPlan> splan(nspecies);
    std::vector<RecvPlan> rplan(nspecies);
    std::vector<uint32_t> nrecv(nspecies, 0), nholes(nspecies, 0);
    std::vector<int>      NW(nspecies);
    for (int s = 0; s < nspecies; ++s) {
      SendPlan& P = splan[s];
      P.nclass    = nclass;
      for (int c = 0; c <= MAXTAG; ++c) P.class_base[c] = 0;
      for (int c = 0; c <= nclass; ++c) P.class_base[c] = hbase[(size_t)s * (MAXTAG + 1) + c];
      for (int c = nclass + 1; c <= MAXTAG; ++c) P.class_base[c] = P.class_base[nclass];
      nholes[s] = P.class_base[nclass];
      for (int c = 0; c < MAXTAG; ++c) {
        P.seg_off[c] = -1;
        P.shift[c][0] = P.shift[c][1] = P.shift[c][2] = 0;
      }
      NW[s] = 4 * D + 4 + (species[s].arrays.phi ? 1 : 0) + npld_words(species[s].arrays);
      rplan[s].nseg = 0;
    }
    // data message layout: per peer, per species, per direction ascending
    bool overflow = false;
    {
      long spos = 0, rpos = 0;
      size_t ridx = 0;
      std::vector<std::vector<std::pair<int, uint32_t>>> recv_by_species(nspecies); // (d, count)
      std::vector<std::vector<long>>                       recv_off(nspecies);
      for (size_t k = 0; k < C.peers.size(); ++k) {
        soff[k] = spos;
        roff[k] = rpos;
        for (const Slot& sl : sslots[k]) {
          SendPlan&      P = splan[sl.s];
          const int      c = tag_of(sl.d) - 1;
          const uint32_t n = P.class_base[c + 1] - P.class_base[c];
          P.seg_off[c]     = spos;
          // index shifts (metadomain_comm.cpp:608-640): backwards adds the target's extent,
          // forwards subtracts the source's
          for (int a = 0; a < D; ++a) {
            if (M.dvec[sl.d][a] == -1) P.shift[c][a] = M.nbr_n[sl.d][a];
            else if (M.dvec[sl.d][a] == 1) P.shift[c][a] = -C.grid.n[a];
          }
          spos += (long)n * NW[sl.s];
        }
        for (const Slot& sl : rslots[k]) {
          const uint32_t n = hrecv[ridx++];
          recv_by_species[sl.s].push_back({ sl.d, n });
          recv_off[sl.s].push_back(rpos);
          rpos += (long)n * NW[sl.s];
        }
        scnt[k] = spos - soff[k];
        rcnt[k] = rpos - roff[k];
      }
      // received particles are consumed in iteration-direction order (particles_comm.cpp:268-365)
      for (int s = 0; s < nspecies; ++s) {
        std::vector<size_t> order(recv_by_species[s].size());
        for (size_t q = 0; q < order.size(); ++q) order[q] = q;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
          return recv_by_species[s][a].first < recv_by_species[s][b].first;
        });
        RecvPlan& R  = rplan[s];
        uint32_t  at = 0;
        for (size_t q : order) {
          if (recv_by_species[s][q].second == 0) continue;
          R.first[R.nseg]   = at;
          R.seg_off[R.nseg] = recv_off[s][q];
          ++R.nseg;
          at += recv_by_species[s][q].second;
        }
        R.first[R.nseg] = at;
        nrecv[s]        = at;
        const eb200_species_t& sp = species[s];
        // particles_comm.cpp:219-221. The failure is reported AFTER the data exchange: the
        // neighbours are already committed to their sends, and a rank that returned here would
        // leave them blocked in the grouped ncclSend / ncclRecv. The message lands in the
        // context's receive buffer (sized by the counts), nothing is unpacked.
        if ((uint64_t)sp.npart + nrecv[s] >= (uint64_t)sp.maxnpart && nrecv[s] > 0) overflow = true;
      }
      CU(C, C.sendbuf.reserve_grow(std::max<size_t>((size_t)spos * 4, 256)));
      CU(C, C.recvbuf.reserve_grow(std::max<size_t>((size_t)rpos * 4, 256)));
    }
    // ---- 4. pack, exchange, unpack
    for (int s = 0; s < nspecies; ++s) {
      const eb200_species_t& sp   = species[s];
      const uint32_t         nout = nholes[s] - splan[s].class_base[1];
      if (nout == 0) continue;
      const unsigned nb  = (nout + 255) / 256;
      const uint32_t* oi = (const uint32_t*)(W + off_idx[s]);
      const int       hp = sp.arrays.phi ? 1 : 0;
      switch (D) {
        case 1: prtl_pack_kernel<1><<<nb, 256, 0, st>>>(splan[s], sp.arrays, oi, nholes[s], hp, (uint32_t*)C.sendbuf.ptr); break;
        case 2: prtl_pack_kernel<2><<<nb, 256, 0, st>>>(splan[s], sp.arrays, oi, nholes[s], hp, (uint32_t*)C.sendbuf.ptr); break;
        default: prtl_pack_kernel<3><<<nb, 256, 0, st>>>(splan[s], sp.arrays, oi, nholes[s], hp, (uint32_t*)C.sendbuf.ptr); break;
      }
      count_launch();
    }
    trace.mark(); // pack
    int rc = exchange(C, soff, scnt, roff, rcnt, 4, st);
    if (rc != EB200_OK) return rc;
    trace.mark(); // exchange
    if (overflow) {
      return fail(C, EB200_ERR_CAPACITY, "Too many particles to receive (cannot fit into maxptl)");
    }
    if (exc_overflow) {
      return fail(C, EB200_ERR_CAPACITY,
                  "migration: more dead / outgoing particles than the exception list holds; compact more "
                  "often (clear_interval) or set EB200_MIGRATE_SCAN=1");
    }
    for (int s = 0; s < nspecies; ++s) {
      eb200_species_t& sp = species[s];
      if (nrecv[s] == 0) continue;
      const unsigned  nb = (nrecv[s] + 255) / 256;
      const uint32_t* oi = (const uint32_t*)(W + off_idx[s]);
      const int       hp = sp.arrays.phi ? 1 : 0;
      switch (D) {
        case 1: prtl_unpack_kernel<1><<<nb, 256, 0, st>>>(rplan[s], sp.arrays, oi, nholes[s], sp.npart, nrecv[s], hp, (const uint32_t*)C.recvbuf.ptr); break;
        case 2: prtl_unpack_kernel<2><<<nb, 256, 0, st>>>(rplan[s], sp.arrays, oi, nholes[s], sp.npart, nrecv[s], hp, (const uint32_t*)C.recvbuf.ptr); break;
        default: prtl_unpack_kernel<3><<<nb, 256, 0, st>>>(rplan[s], sp.arrays, oi, nholes[s], sp.npart, nrecv[s], hp, (const uint32_t*)C.recvbuf.ptr); break;
      }
      count_launch();
      if (nrecv[s] > nholes[s]) sp.npart += nrecv[s] - nholes[s]; // particles_comm.cpp:384-387
    }
    CU(C, cudaGetLastError());
    if (trace.on) {
      trace.mark(); // unpack
      static int calls = 0;
      if (++calls % 10 == 0) {
        fprintf(stderr, "[migrate r%d #%d] classify %.3f counts %.3f pack %.3f exchange %.3f unpack %.3f ms;",
                M.rank, calls, trace.t[0], trace.t[1], trace.t[2], trace.t[3], trace.t[4]);
        for (int s = 0; s < nspecies; ++s) {
          fprintf(stderr, " s%d: dead %u out %u recv %u npart %u", s, splan[s].class_base[1],
                  nholes[s] - splan[s].class_base[1], nrecv[s], species[s].npart);
        }
        fprintf(stderr, "\n");
      }
    }
    return EB200_OK;
  }

} // namespace eb200

/* ----------------------------------------------------------------- host-only C ABI */
extern "C" int eb200_decompose(int ndomains, int dim, const int* ncells, const int* decomposition,
                               int* ndoms_out, int* e1, int* e2, int* e3) {
  using namespace eb200;
  if (ndomains < 1 || dim < 1 || dim > 3 || !ncells || !decomposition || !ndoms_out) return EB200_ERR_ARG;
  unsigned n[3] = { 1, 1, 1 };
  const int* dc = decomposition;
  const unsigned nd = (unsigned)ndomains;
  if (dim == 1) {
    n[0] = nd;
  } else if (dim == 2) {
    if (dc[0] > 0 && dc[1] > 0) {
      n[0] = dc[0];
      n[1] = dc[1];
    } else if (dc[0] > 0 && dc[1] <= 0) {
      n[0] = dc[0];
      if (nd % n[0] != 0) return EB200_ERR_ARG;
      n[1] = nd / n[0];
    } else if (dc[0] <= 0 && dc[1] > 0) {
      n[1] = dc[1];
      if (nd % n[1] != 0) return EB200_ERR_ARG;
      n[0] = nd / n[1];
    } else if (!proportions2d(nd, ncells[0], ncells[1], n[0], n[1])) {
      return EB200_ERR_ARG;
    }
    if (n[0] * n[1] != nd) return EB200_ERR_ARG;
  } else {
    const bool f0 = dc[0] > 0, f1 = dc[1] > 0, f2 = dc[2] > 0;
    if (f0) n[0] = dc[0];
    if (f1) n[1] = dc[1];
    if (f2) n[2] = dc[2];
    bool ok = true;
    if (f0 && f1 && f2) {
    } else if (!f0 && f1 && f2) {
      ok = nd % (n[1] * n[2]) == 0;
      if (ok) n[0] = nd / (n[1] * n[2]);
    } else if (f0 && !f1 && f2) {
      ok = nd % (n[0] * n[2]) == 0;
      if (ok) n[1] = nd / (n[0] * n[2]);
    } else if (f0 && f1 && !f2) {
      ok = nd % (n[0] * n[1]) == 0;
      if (ok) n[2] = nd / (n[0] * n[1]);
    } else if (!f0 && !f1 && f2) {
      ok = nd % n[2] == 0 && proportions2d(nd / n[2], ncells[0], ncells[1], n[0], n[1]);
    } else if (!f0 && f1 && !f2) {
      ok = nd % n[1] == 0 && proportions2d(nd / n[1], ncells[0], ncells[2], n[0], n[2]);
    } else if (f0 && !f1 && !f2) {
      ok = nd % n[0] == 0 && proportions2d(nd / n[0], ncells[1], ncells[2], n[1], n[2]);
    } else {
      ok = proportions3d(nd, ncells[0], ncells[1], ncells[2], n[0], n[1], n[2]);
    }
    if (!ok || n[0] * n[1] * n[2] != nd) return EB200_ERR_ARG;
  }
  int* outs[3] = { e1, e2, e3 };
  for (int a = 0; a < 3; ++a) ndoms_out[a] = (int)n[a];
  for (int a = 0; a < dim; ++a) {
    std::vector<int> e;
    if (!decompose1d(n[a], ncells[a], e)) return EB200_ERR_ARG;
    if (!outs[a]) return EB200_ERR_ARG;
    for (unsigned k = 0; k < n[a]; ++k) outs[a][k] = e[k];
  }
  return EB200_OK;
}

extern "C" int eb200_domain_info(const eb200_metadomain_t* md, eb200_domain_info_t* out) {
  if (!md || !out) return EB200_ERR_ARG;
  eb200::Metadomain M;
  const char*       e = eb200::build_metadomain(*md, M);
  if (e) return EB200_ERR_ARG;
  *out = M.info;
  return EB200_OK;
}
