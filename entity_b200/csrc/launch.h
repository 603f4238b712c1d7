// entity_b200 -- host-side declarations shared by the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/entity_b200.h"

namespace eb200 {

  // emission policy of the SR pusher (arch::emission::Synchrotron / Compton), by value
  struct EmitParams {
    int   kind;
    float photon_weight, energy_min, nominal_probability, nominal_photon_energy, species_mass;
    int   should_drag;
  };

  // kernel argument of the pusher with an emission policy
  struct EmitArgs {
    EmitParams    E;
    eb200_prtls_t ph;          // the emitted (photon) species
    uint32_t      offset, cap; // its npart before the launch, its maxnpart
    uint32_t*     counter;     // device: photons emitted by this launch (may exceed cap - offset)
    uint64_t      seed;
    uint32_t      step, call;
  };

  // Exception list of one fused push launch: the indices of the particles that are not alive
  // afterwards (dead before, absorbed, or tagged for migration). The migration reads it instead
  // of scanning every tag twice (comm.cu). `tracked` is set by the launcher when every particle of
  // the launch went through a kernel that appends; count may exceed cap (then the list is unusable).
  struct ExcList {
    uint32_t* count   = nullptr; // device
    uint32_t* idx     = nullptr; // device, cap entries, pre-filled with 0xFFFFFFFF
    uint32_t  cap     = 0;
    bool      tracked = false;
  };

  // number of kernel launches issued by this library (eb200_launch_count)
  void     count_launch();
  uint64_t launches();

  // grow-only device scratch owned by a context
  struct Scratch {
    void*  ptr   = nullptr;
    size_t bytes = 0;

    cudaError_t reserve(size_t n) {
      if (n <= bytes) return cudaSuccess;
      if (ptr) cudaFree(ptr);
      ptr   = nullptr;
      bytes = 0;
      cudaError_t e = cudaMalloc(&ptr, n);
      if (e == cudaSuccess) bytes = n;
      return e;
    }

    // for buffers whose size follows the data (particle counts): grow with headroom so that a
    // slowly rising count does not free and reallocate every step (each cudaFree is a device
    // synchronisation, and with peer mappings it costs milliseconds)
    cudaError_t reserve_grow(size_t n, size_t floor_bytes = size_t(1) << 20) {
      if (n <= bytes) return cudaSuccess;
      size_t want = n + n / 2;
      if (want < floor_bytes) want = floor_bytes;
      cudaError_t e = reserve(want);
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        e = reserve(n);
      }
      return e;
    }

    void release() {
      if (ptr) cudaFree(ptr);
      ptr   = nullptr;
      bytes = 0;
    }
  };

  // bcs.cu (compiled once, --fmad=false): matching field boundaries
  cudaError_t match_fields(const eb200_grid_t& g, float* em, const float* target, int o, float dx,
                           float xmin_o, float xg_edge, float ds, int tags, int mask,
                           const int* rmin, const int* rmax, cudaStream_t st);
  // stats.cu (compiled once): reduced statistics of a Minkowski domain
  cudaError_t stats_fields(const eb200_grid_t& g, const float* em, const float* cur, float dx,
                           int what, int comp, double* out_dev, cudaStream_t st);
  cudaError_t stats_particles(const eb200_grid_t& g, const eb200_prtls_t& S, uint32_t npart,
                              float mass, float charge, int use_weights, float dx, int what, int c1,
                              int c2, double* out_dev, cudaStream_t st);

#define EB200_DECLARE_VARIANT(NS)                                                              \
  namespace NS {                                                                               \
    cudaError_t push_sr(const eb200_grid_t& g, int order, const eb200_pusher_t& c,            \
                        const eb200_prtls_t& S, uint32_t npart, const float* em,               \
                        cudaStream_t st);                                                      \
    cudaError_t push_sr_emission(const eb200_grid_t& g, int order, const eb200_pusher_t& c,   \
                                 const eb200_prtls_t& S, uint32_t npart, const float* em,      \
                                 const EmitArgs& M, cudaStream_t st);                          \
    cudaError_t deposit(const eb200_grid_t& g, int order, const eb200_prtls_t& S,             \
                        uint32_t npart, float charge, float dt, float dxc, float* cur,         \
                        int mode, Scratch& scratch, cudaStream_t st);                          \
    cudaError_t push_deposit_sr(const eb200_grid_t& g, int order, const eb200_pusher_t& c,    \
                                const eb200_prtls_t& S, uint32_t npart, const float* em,       \
                                float* cur, int mode, float* packed, bool do_pack,             \
                                cudaStream_t st, float* packed_j, bool* packed_j_used,         \
                                ExcList* exc);                                                 \
    cudaError_t unpack_j4(const eb200_grid_t& g, float* packed_j, float* cur,                 \
                          cudaStream_t st);                                                    \
    cudaError_t pack_em2d(const eb200_grid_t& g, const float* em, float* packed,              \
                          cudaStream_t st);                                                    \
    cudaError_t faraday(const eb200_grid_t& g, float* em, float c1, float c2,                 \
                        const float* stencil9, cudaStream_t st);                               \
    cudaError_t ampere(const eb200_grid_t& g, float* em, float c1, float c2, cudaStream_t st); \
    cudaError_t currents_ampere(const eb200_grid_t& g, float* em, float* cur, float coeff,    \
                                float ppc0, cudaStream_t st);                                  \
    cudaError_t currents_ampere_ext(const eb200_grid_t& g, float* em, float* cur, float coeff, \
                                    float ppc0, const eb200_ext_current_t& ext, float dx,      \
                                    const float* xmin, cudaStream_t st);                       \
    cudaError_t filter_pass(const eb200_grid_t& g, float* cur, const float* buff,             \
                            const int* fbc, int extend, cudaStream_t st);                      \
    cudaError_t filter_fused(const eb200_grid_t& g, const float* src, float* dst, int passes, \
                             int ghosts, cudaStream_t st);                                     \
    cudaError_t comm_fields_self(const eb200_grid_t& g, float* fld, int c0, int c1,           \
                                 const int* fbc, cudaStream_t st);                             \
    cudaError_t sync_currents_self(const eb200_grid_t& g, float* cur, float* buff,            \
                                   const int* fbc, cudaStream_t st);                           \
  }

  EB200_DECLARE_VARIANT(strict_fp)
  EB200_DECLARE_VARIANT(fast_fp)

  // injection and per-cell moments (inject.cu)
  struct MetricParams;
  cudaError_t inject_nonuniform(const eb200_grid_t& g, const MetricParams* mp, float dx,
                                const float* xmin, const eb200_prtls_t& S1, uint32_t npart1,
                                uint32_t cap1, const eb200_prtls_t& S2, uint32_t npart2,
                                uint32_t cap2, float ppc0, const eb200_spatial_dist_t& sd,
                                const float* field, const eb200_maxwellian_t& e1,
                                const eb200_maxwellian_t& e2, const int* rmin, const int* rmax,
                                uint64_t seed, uint32_t step, uint32_t call, uint32_t* n_inj_host,
                                int* overflow, Scratch& scratch, cudaStream_t st);
  cudaError_t particle_moment(const eb200_grid_t& g, const MetricParams* mp, const eb200_prtls_t& S,
                              uint32_t npart, float coeff, bool use_weights, bool volume,
                              float* plane, cudaStream_t st);

  // reduced statistics on (q)spherical SRPIC meshes (stats.cu)
  cudaError_t stats_fields_curv(const MetricParams& mp, const eb200_grid_t& g, const float* em,
                                const float* cur, int what, int comp, double* out_dev, cudaStream_t st);
  cudaError_t stats_particles_curv(const MetricParams& mp, const eb200_prtls_t& S, uint32_t npart,
                                   float mass, float charge, int use_weights, int what, int c1, int c2,
                                   double* out_dev, cudaStream_t st);

  // output staging (output.cu)
  cudaError_t fields_to_phys(const MetricParams* mp, const eb200_grid_t& g, float dx,
                             const float* from, int ncomp_from, float* to, int ncomp_to,
                             const int* cf, const int* ct, int interp, int conv, cudaStream_t st);
  cudaError_t prtls_to_phys(const MetricParams* mp, const eb200_grid_t& g, float dx,
                            const float* xmin, const eb200_prtls_t& S, uint32_t stride,
                            uint32_t nout, float* x1, float* x2, float* x3, float* u1, float* u2,
                            float* u3, float* w, cudaStream_t st);

  cudaError_t conductor_fields2d(const eb200_grid_t& g, float* em, int o, bool pos, int tags,
                                 cudaStream_t st);

  // curvilinear SR and GR kernels: curv.cu (one build, IEEE division, no FMA contraction)
  namespace curv {
    // field boundaries (bcs.cu)
    cudaError_t axis_fields(const eb200_grid_t& g, float* fld, bool pos, int tags, cudaStream_t st);
    cudaError_t horizon_fields(const eb200_grid_t& g, float* fld, int tags, int nfilter,
                               cudaStream_t st);
    cudaError_t match_fields_curv(const MetricParams& m, const eb200_grid_t& g, float* fld,
                                  const float* target, int o, float xg_edge, float ds, int tags,
                                  int mask, const int* rmin, const int* rmax, const int* fbc,
                                  cudaStream_t st);
    cudaError_t enforce_fields(const eb200_grid_t& g, float* em, const float* target, int o,
                               bool pos, int i_edge, int tags, int mask, const int* rmin,
                               const int* rmax, cudaStream_t st);
    cudaError_t absorb_currents(const MetricParams& m, const eb200_grid_t& g, float* cur,
                                float xg_edge, float ds, const int* rmin, const int* rmax,
                                cudaStream_t st);
    cudaError_t push_sr(const MetricParams& m, const eb200_grid_t& g, int order,
                        const eb200_pusher_t& c, const eb200_prtls_t& S, uint32_t npart,
                        const float* em, cudaStream_t st);
    cudaError_t deposit_sr(const MetricParams& m, const eb200_grid_t& g, int order,
                           const eb200_prtls_t& S, uint32_t npart, float charge, float dt,
                           float* cur, int mode, cudaStream_t st);
    cudaError_t faraday_sr(const MetricParams& m, const eb200_grid_t& g, float* em, float coeff,
                           const int* fbc, cudaStream_t st);
    cudaError_t ampere_sr(const MetricParams& m, const eb200_grid_t& g, float* em, float coeff,
                          const int* fbc, cudaStream_t st);
    cudaError_t currents_ampere_sr(const MetricParams& m, const eb200_grid_t& g, float* em,
                                   float* cur, float coeff, float inv_n0, const int* fbc,
                                   cudaStream_t st);
    cudaError_t filter_sph_pass(const eb200_grid_t& g, float* cur, const float* buff,
                                const int* fbc, cudaStream_t st);
    cudaError_t push_gr(const MetricParams& m, const eb200_grid_t& g, int order,
                        const eb200_pusher_gr_t& c, const eb200_prtls_t& S, uint32_t npart,
                        const float* em, const float* em0, cudaStream_t st);
    cudaError_t deposit_gr(const MetricParams& m, const eb200_grid_t& g, int order,
                           const eb200_prtls_t& S, uint32_t npart, float charge, float dt,
                           float* cur, int mode, cudaStream_t st);
    cudaError_t aux_gr(const MetricParams& m, const eb200_grid_t& g, int which_h, const float* Df,
                       const float* Bf, float* out, const int* fbc, cudaStream_t st);
    cudaError_t faraday_gr(const MetricParams& m, const eb200_grid_t& g, const float* Bin,
                           float* Bout, const float* E, float coeff, const int* fbc,
                           cudaStream_t st);
    cudaError_t ampere_gr(const MetricParams& m, const eb200_grid_t& g, const float* Din,
                          float* Dout, const float* H, float coeff, const int* fbc,
                          cudaStream_t st);
    cudaError_t currents_ampere_gr(const MetricParams& m, const eb200_grid_t& g, float* Df,
                                   const float* cur, float coeff, const int* fbc,
                                   cudaStream_t st);
    cudaError_t time_average(const eb200_grid_t& g, float* a, const float* b, int ncomp,
                             cudaStream_t st);
  } // namespace curv

  // variant-independent (integer / copy work): sort.cu
  cudaError_t sort_particles(const eb200_grid_t& g, const eb200_prtls_t& S, uint32_t npart,
                             uint32_t maxnpart, int remove_dead, uint32_t* n_alive_out,
                             Scratch& scratch, cudaStream_t st);

} // namespace eb200
