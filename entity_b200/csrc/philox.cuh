// entity_b200 -- counter-based random numbers (Philox4x32-10, Salmon et al. 2011).
//
// The reference draws from a Kokkos random pool (one XorShift stream per thread: the sequence a
// particle or cell sees depends on the backend and the thread count). Here every consumer owns
// a stream keyed by (seed, step, call id, item): key = seed, counter = (block, item, call, step);
// the result is a pure function of the arguments, identical for every launch shape and device.
// oracle/philox.py is the numpy restatement, pinned to the published known answers
// (tests/test_inject.py).
#pragma once
#include <cstdint>

namespace eb200 {
  struct Philox {
    uint32_t key[2], ctr[4], out[4];
    int      have;

    __device__ Philox(uint64_t seed, uint32_t step, uint32_t call, uint32_t cell) {
      key[0] = (uint32_t)seed;
      key[1] = (uint32_t)(seed >> 32);
      ctr[0] = 0, ctr[1] = cell, ctr[2] = call, ctr[3] = step;
      have   = 0;
    }

    __device__ void round(uint32_t (&c)[4], const uint32_t (&k)[2]) const {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
      const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
      c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
    }

    __device__ uint32_t next() {
      if (have == 0) {
        uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
        uint32_t k[2] = { key[0], key[1] };
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          round(c, k);
          k[0] += 0x9E3779B9u;
          k[1] += 0xBB67AE85u;
        }
        out[0] = c[0], out[1] = c[1], out[2] = c[2], out[3] = c[3];
        ++ctr[0];
        have = 4;
      }
      return out[--have];
    }

    // Random<real_t>: uniform in [0, 1)
    __device__ float uniform() { return (float)(next() >> 8) * 5.9604645e-08f; }
  };
} // namespace eb200
