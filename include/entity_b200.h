/*
 * entity_b200 -- C ABI of the B200-native PIC timestep hot path.
 *
 * One entry point per dispatcher of the reference's engine layer, i.e. what a
 * maintainer binds in place of the Kokkos functors of src/kernels (see
 * INTEGRATION.md for the shim that replaces the bodies of ntt::srpic::* with
 * these calls). Plain pointers and sizes only; no torch / Kokkos types.
 *
 * Conventions
 *  - All array arguments are DEVICE pointers unless the name ends in `_host`.
 *  - Fields: fp32, extents (n_d + 2*ng) per simulated dimension, i1 fastest,
 *    then i2, i3; the component index is slowest ("component planes"). This is
 *    Kokkos LayoutLeft, the layout of the reference's CUDA build for
 *    ndfield_t<D,N> (src/global/arch/kokkos_aliases.h:83-105,
 *    src/framework/containers/fields.h:38-108). The arrays are dense: the ABI carries no
 *    strides on purpose. Kokkos::View<real_t**[N]> of the CUDA build is contiguous LayoutLeft
 *    (stride(0) = 1, stride(k) = product of the lower extents: no padding unless the host asks
 *    for AllowPadding, which the reference does not), the coalescing and the TMA / vector
 *    accesses of the kernels rely on unit stride along i1; a host-space (LayoutRight) build of
 *    the reference would transpose on its side of eb200_srpic_step_host rather than ask every
 *    kernel for strided access. integration/eb200_shim.hpp hands view.data() over as is.
 *  - Particles: SoA; eb200_prtls_t lists the arrays in the member order of
 *    ntt::ParticleArrays (src/framework/containers/particles.h:47-71).
 *  - `stream` is a cudaStream_t passed as void*; every call only enqueues work
 *    on it and never synchronises the device, except where a host-visible value
 *    is returned (particle counts after sort / migration), mirroring
 *    SURVEY.md section 8b "Threading".
 *  - Return value: 0 on success, non-zero on error; eb200_last_error() gives
 *    the message (reference: raise::Error / Kokkos::abort, src/global/utils/error.h).
 *  - There is no CPU fallback: every call fails with EB200_ERR_NO_DEVICE when
 *    no CUDA device is usable.
 */
#ifndef ENTITY_B200_H
#define ENTITY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EB200_VERSION 100

enum {
  EB200_OK            = 0,
  EB200_ERR_ARG       = 1,
  EB200_ERR_CUDA      = 2,
  EB200_ERR_NO_DEVICE = 3,
  EB200_ERR_CAPACITY  = 4, /* npart + nrecv >= maxnpart: particles_comm.cpp:219-221 */
  EB200_ERR_UNSUPPORTED = 5,
  EB200_ERR_NCCL      = 6
};

/* ntt::ParticlePusher flags, src/global/enums.h:317-324 */
enum { EB200_PUSHER_NONE = 0, EB200_PUSHER_PHOTON = 1, EB200_PUSHER_BORIS = 2, EB200_PUSHER_VAY = 4, EB200_PUSHER_GCA = 8 };
/* ntt::RadiativeDrag flags, src/global/enums.h:351-356 */
enum { EB200_DRAG_NONE = 0, EB200_DRAG_SYNCHROTRON = 1, EB200_DRAG_COMPTON = 2 };
/* particle boundary per face; what kernel::sr::PusherBoundaries distinguishes
 * (src/kernels/pushers/context.h:124-178). NONE = SYNC/other: particle leaves, gets a send tag. */
enum { EB200_PBC_NONE = 0, EB200_PBC_PERIODIC = 1, EB200_PBC_ABSORB = 2, EB200_PBC_REFLECT = 3, EB200_PBC_AXIS = 4 };
/* field boundary per face, as far as filter and ghost exchange need it (ntt::FldsBC) */
enum { EB200_FBC_NONE = 0, EB200_FBC_PERIODIC = 1, EB200_FBC_CONDUCTOR = 2, EB200_FBC_AXIS = 3, EB200_FBC_SYNC = 4,
       /* faces the step mirrors apply a boundary kernel to (srpic / grpic ::FieldBoundaries) */
       EB200_FBC_MATCH = 5, EB200_FBC_HORIZON = 6, EB200_FBC_ATMOSPHERE = 7 };
/* ntt::Metric (src/global/enums.h): the curvilinear metrics are 2D axisymmetric, like the
 * reference's (static_asserts in src/metrics/qspherical.h:34-35 etc.) */
enum {
  EB200_METRIC_MINKOWSKI     = 0,
  EB200_METRIC_SPHERICAL     = 1, /* src/metrics/spherical.h */
  EB200_METRIC_QSPHERICAL    = 2, /* src/metrics/qspherical.h */
  EB200_METRIC_KERR_SCHILD   = 3, /* src/metrics/kerr_schild.h */
  EB200_METRIC_QKERR_SCHILD  = 4, /* src/metrics/qkerr_schild.h */
  EB200_METRIC_KERR_SCHILD_0 = 5  /* src/metrics/kerr_schild_0.h */
};
/* deposit modes */
enum {
  EB200_DEPOSIT_ATOMIC     = 0, /* one atomic add per particle and node (what Kokkos ScatterView does on CUDA) */
  EB200_DEPOSIT_ORDERED    = 1, /* deterministic: every J element summed in particle order */
  EB200_DEPOSIT_AGGREGATED = 2  /* warp-aggregated: runs of same-cell lanes are reduced with shuffles,
                                   one atomic per run and node; meant for cell-sorted particles */
};

typedef struct eb200_ctx eb200_ctx_t;
typedef void*            eb200_stream_t;

/* local (per-domain) mesh: Mesh::n_active + N_GHOSTS (src/global/global.h:130-136) */
typedef struct {
  int dim;  /* 1, 2, 3 */
  int n[3]; /* active cells per dimension; unused dimensions = 1 */
  int ng;   /* ghost cells per side */
} eb200_grid_t;

/* raw device pointers in the order of ntt::ParticleArrays (particles.h:53-70) */
typedef struct {
  int*      i1;
  int*      i2;
  int*      i3;
  float*    dx1;
  float*    dx2;
  float*    dx3;
  float*    ux1;
  float*    ux2;
  float*    ux3;
  float*    weight;
  int*      i1_prev;
  int*      i2_prev;
  int*      i3_prev;
  float*    dx1_prev;
  float*    dx2_prev;
  float*    dx3_prev;
  short*    tag;
  float*    pld_r; /* payload k of particle p at pld_r[p + k * pld_stride], may be NULL */
  uint32_t* pld_i; /* same layout (Kokkos LayoutLeft = the reference's device build), may be NULL */
  float*    phi;   /* 2D non-Cartesian only, may be NULL */
  /* payload planes (ParticleSpecies::npld_r / npld_i, particles.h:62-65): sorted, compacted and
   * migrated with their particles (particles_sort.cpp:104-253, particles_comm.cpp:180-389);
   * the pusher and the deposit do not touch them. pld_stride = extent(0) of the views. */
  int       npld_r, npld_i; /* each <= EB200_MAX_PLD */
  uint32_t  pld_stride;
} eb200_prtls_t;
#define EB200_MAX_PLD 16

/* scalar arguments of kernel::sr::Pusher_kernel: PusherContext + PusherBoundaries
 * (src/kernels/pushers/context.h:74-178), filled by the host exactly like
 * srpic::ParticlePush does (src/engines/srpic/particle_pusher.h:36-150). */
typedef struct {
  int    pusher_flags;
  int    drag_flags;
  float  mass, charge;
  double time;
  float  dt, omegaB0;
  float  gca_larmor_max, gca_e_ovr_b_sqr_max;
  float  sync_coeff, compton_coeff;
  int    has_atmosphere;
  float  atm_gx1, atm_gx2, atm_gx3, atm_x_surf, atm_ds;
  int    pbc[6];       /* EB200_PBC_* for i1min,i1max,i2min,i2max,i3min,i3max */
  int    tag_outgoing; /* 1: tag leaving particles with mpi::SendTag (mpi_tags.h:175-233) */
  float  dx;           /* Minkowski cell size (metric::Minkowski::dx) */
  float  xmin[3];      /* Minkowski x*_min */
} eb200_pusher_t;

typedef struct {
  int          device;      /* CUDA device ordinal */
  int          strict_fp;   /* 1: kernels built without FMA contraction (bit-exact with the
                               reference's baseline CPU build); 0: contraction allowed */
  eb200_grid_t grid;        /* local mesh */
  int          shape_order; /* SHAPE_ORDER of the reference build: 0 (zig-zag), 1..11 (Esirkepov;
                               4..11: Minkowski, unfused kernels, ATOMIC / AGGREGATED modes) */
  int          metric;      /* EB200_METRIC_* */
  float        metric_params[8]; /* Minkowski: dx, x1min, x2min, x3min;
                                    curvilinear / GR: x1min, x1max, x2min, x2max (physical extent of
                                    THIS domain), qsph_r0, qsph_h, ks_a */
  uint32_t     maxnpart;    /* capacity of each particle array handed to this context */
} eb200_config_t;

/* ------------------------------------------------------------------ lifetime */
int         eb200_version(void);
int         eb200_device_count(void);
int         eb200_init(const eb200_config_t* cfg, eb200_ctx_t** out);
void        eb200_finalize(eb200_ctx_t* ctx);
const char* eb200_last_error(const eb200_ctx_t* ctx); /* ctx may be NULL: last global error */
/* number of kernels this library has launched since eb200_init (bench.py's gpu_launches) */
uint64_t    eb200_launch_count(const eb200_ctx_t* ctx);

/* ------------------------------------------------------------- field solvers */
/* replaces kernel::mink::Faraday_kernel as launched by srpic::Faraday
 * (src/engines/srpic/fieldsolvers.h:36-98, src/kernels/faraday_mink.hpp:71-166).
 * stencil9 = {delta_x, delta_y, beta_xy, beta_yx, delta_z, beta_xz, beta_zx, beta_yz, beta_zy}
 * (host pointer, may be NULL = all zero). */
int eb200_faraday(eb200_ctx_t* ctx, float* em, float coeff1, float coeff2,
                  const float* stencil9_host, eb200_stream_t stream);
/* kernel::mink::Ampere_kernel / srpic::Ampere (fieldsolvers.h:101-139, ampere_mink.hpp:48-89) */
int eb200_ampere(eb200_ctx_t* ctx, float* em, float coeff1, float coeff2, eb200_stream_t stream);
/* kernel::mink::CurrentsAmpere_kernel without ext. current / srpic::CurrentsAmpere
 * (fieldsolvers.h:142-199, ampere_mink.hpp:134-215): E += J*coeff; J /= ppc0 */
int eb200_currents_ampere(eb200_ctx_t* ctx, float* em, float* cur, float coeff, float ppc0,
                          eb200_stream_t stream);
/* The `ext_current` of a problem generator (traits::pgen::HasExtCurrent; CurrentsAmpere_kernel<D,
 * ExtCurrent>, ampere_mink.hpp:134-215: J_c += ppc0 * ext_current.jx_c(x) on the component's own
 * node before E += J coeff) cannot cross a C ABI as a functor. It is handed over as a table of
 * Fourier modes, which is what the reference's one ext_current is (the turbulence antenna,
 * pgens/turbulence/pgen.hpp:139-298):
 *   jx_c(x) = sum_m pref[c][m]  * (a_real[m]  cos(k_m . x) - a_imag[m]  sin(k_m . x))
 *           + sum_m pref2[c][m] * (a_real2[m] cos(k_m . x) - a_imag2[m] sin(k_m . x))
 * evaluated mode by mode in this order (the second sum inside the same loop over m, as the 2D
 * antenna's inverse-helicity term). The host advances the amplitudes (the pgen's
 * CustomPostStep) and refills the table; k_m . x = k[0][m] x1 + k[1][m] x2 (+ k[2][m] x3). */
#define EB200_MAX_MODES 16
typedef struct {
  int   nmodes;
  float k[3][EB200_MAX_MODES];
  float pref[3][EB200_MAX_MODES];
  float a_real[EB200_MAX_MODES], a_imag[EB200_MAX_MODES];
  float pref2[3][EB200_MAX_MODES];
  float a_real2[EB200_MAX_MODES], a_imag2[EB200_MAX_MODES];
} eb200_ext_current_t;
int eb200_currents_ampere_ext(eb200_ctx_t* ctx, float* em, float* cur, float coeff, float ppc0,
                              const eb200_ext_current_t* ext_host, eb200_stream_t stream);
/* registers (copies) the table eb200_srpic_step uses in srpic::CurrentsAmpere; NULL clears it */
int eb200_srpic_set_ext_current(eb200_ctx_t* ctx, const eb200_ext_current_t* ext_host);
/* srpic::CurrentsFilter (src/engines/srpic/currents.h:89-119): nfilter x { buff = cur;
 * DigitalFilter_kernel (digital_filter.hpp:99-388, Cartesian); ghost exchange of J }.
 * fbc_host[6] = EB200_FBC_* per face. */
int eb200_filter(eb200_ctx_t* ctx, float* cur, float* buff, int nfilter, const int* fbc_host,
                 eb200_stream_t stream);

/* ------------------------------------------------------------------ particles */
/* srpic::ParticlePush for one species (particle_pusher.h:36-185, sr.hpp:117-332) */
int eb200_push_sr(eb200_ctx_t* ctx, const eb200_pusher_t* pusher, const eb200_prtls_t* prtls,
                  uint32_t npart, const float* em, eb200_stream_t stream);
/* one species of srpic::CurrentsDeposit (currents.h:32-62, currents_deposit.hpp:108-761):
 * accumulates into cur (the caller zeroes it once per step with eb200_zero_currents, as
 * currents.h:67 does). mode = EB200_DEPOSIT_*; ORDERED sums every J element in particle order
 * (the reference's Serial-backend order) and is bit-reproducible. */
int eb200_deposit(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart, float charge,
                  float dt, float* cur, int mode, eb200_stream_t stream);
/* fused ParticlePush + CurrentsDeposit of one species in a single pass over the particles;
 * mode = EB200_DEPOSIT_ATOMIC | EB200_DEPOSIT_AGGREGATED */
int eb200_push_deposit_sr(eb200_ctx_t* ctx, const eb200_pusher_t* pusher,
                          const eb200_prtls_t* prtls, uint32_t npart, const float* em, float* cur,
                          int mode, eb200_stream_t stream);
int eb200_zero_currents(eb200_ctx_t* ctx, float* cur, eb200_stream_t stream);
/* Fused kernel 5 gathers E/B from a context-owned copy of em repacked node by node. A call of
 * eb200_push_deposit_sr rebuilds that copy unless the caller holds it: hold(em) packs once and
 * promises that em is not modified until release (eb200_srpic_step does this around the species
 * loop of srpic::ParticlePush, where the reference does not touch em either). No-ops for the
 * other kernels. */
int eb200_pack_fields_hold(eb200_ctx_t* ctx, const float* em, eb200_stream_t stream);
int eb200_pack_fields_release(eb200_ctx_t* ctx);
/* which fused kernel eb200_push_deposit_sr / eb200_srpic_step launch in AGGREGATED mode:
 * 0 = automatic (default), 1 = one particle per thread, 2 = TMA-staged persistent chunks,
 * 3 = four particles per thread with 128-bit accesses (zig-zag only), 4 = 3 with the E/B
 * nodes of a CTA staged in shared memory, 5 = 3 gathering from the packed E/B nodes (2D; what
 * 0 selects for 2D zig-zag in the strict build), 6 = pipelined persistent CTAs, 7 =
 * shared-memory resident slices, 8 = 5 with the zig-zag deposit accumulated as segment
 * moments (fast build, plain Boris push; what 0 selects there). A tuning knob for
 * measurements; results are the same up to the summation order of J. */
int eb200_set_pd_kernel(eb200_ctx_t* ctx, int which);

/* ------------------------------------------------- single-domain ghost exchange */
/* Metadomain::CommunicateFields for a domain that is its own periodic neighbour
 * (metadomain_comm.cpp:205-367 + comm_nompi.hpp:29-119): ghost fill of components [c0,c1). */
int eb200_comm_fields(eb200_ctx_t* ctx, float* fld, int ncomp, int c0, int c1,
                      const int* fbc_host, eb200_stream_t stream);
/* Metadomain::SynchronizeFields(Comm::J) (metadomain_comm.cpp:409-562): additive sync of the
 * 2*ng wide boundary strips through buff, then cur += buff on active cells. */
int eb200_sync_currents(eb200_ctx_t* ctx, float* cur, float* buff, const int* fbc_host,
                        eb200_stream_t stream);

/* ------------------------------------------------------- multi-domain (NCCL) */
/* The decomposition of the global mesh, as Metadomain builds it
 * (src/framework/domain/metadomain.cpp:101-196, 196-330): domains form a Cartesian product,
 * domain index = rank = o1 + nd1*(o2 + nd2*o3) (tools::TensorProduct, tools.h:62-77),
 * neighbours wrap around (redefineNeighbors), internal faces are SYNC and a periodic face whose
 * neighbour is another domain becomes SYNC (redefineBoundaries). */
typedef struct {
  int        dim;
  int        rank, nranks;  /* one domain per rank */
  int        ndoms[3];      /* domains per dimension; product = nranks */
  const int* extents[3];    /* extents[a][k] = active cells of the k-th domain along a (host) */
  int        fbc[6];        /* EB200_FBC_* of the GLOBAL mesh faces */
  int        pbc[6];        /* EB200_PBC_* of the GLOBAL mesh faces */
} eb200_metadomain_t;

/* what Metadomain derives for one domain; direction index = lexicographic over {-1,0,1}^dim
 * with x1 slowest (dir::Directions<D>::all order incl. the null direction at the centre) */
typedef struct {
  int offset[3];      /* offset_ndomains */
  int n[3];           /* local active cells */
  int cell_offset[3]; /* offset_ncells */
  int face_fbc[6];    /* local EB200_FBC_* per face (SYNC where another domain adjoins) */
  int face_pbc[6];    /* local EB200_PBC_* per face (NONE where particles leave to a neighbour) */
  int dir_fbc[27];    /* per direction: first non-periodic among the associated faces */
  int neighbor[27];   /* rank of the neighbour in each direction (wraps around) */
  int enabled[27];    /* 1: fields/particles are exchanged in this direction */
} eb200_domain_info_t;

/* tools::Decompose (src/global/utils/tools.h:166-275): decomposition[a] <= 0 means "choose".
 * ndoms_out[3]; extents_out[a] must hold ndoms_out[a] ints (pass arrays of >= ndomains).
 * Pure host code. Fails (EB200_ERR_ARG) where the reference raises. */
int eb200_decompose(int ndomains, int dim, const int* ncells, const int* decomposition,
                    int* ndoms_out, int* extents1_out, int* extents2_out, int* extents3_out);
/* pure host code: the tables above for md->rank */
int eb200_domain_info(const eb200_metadomain_t* md, eb200_domain_info_t* out);

#define EB200_UNIQUE_ID_BYTES 128
/* ncclGetUniqueId; rank 0 calls it and the launcher distributes the bytes
 * (the reference bootstraps through MPI_Init, src/global/global.cpp:9-14) */
int eb200_comm_unique_id(char* id_out /* EB200_UNIQUE_ID_BYTES */);
/* attaches a communicator (ncclCommInitRank) and the decomposition to the context; the
 * context's grid must be this rank's block. From here on eb200_comm_fields,
 * eb200_sync_currents, eb200_filter and eb200_srpic_step exchange with the neighbour
 * domains (packed by device kernels, one grouped ncclSend/ncclRecv per peer and round)
 * instead of the self-periodic copy. id = NULL: decomposition tables only (no transport);
 * valid for nranks == 1. */
int eb200_comm_init(eb200_ctx_t* ctx, const eb200_metadomain_t* md, const char* id);
/* ------------------------------------------------------------- sort / compaction */
/* Particles::SortSpatially (particles_sort.cpp:197-253): stable sort by cell index
 * (i1 fastest, matching the field layout), dead particles moved to the end.
 * Particles::RemoveDead (particles_sort.cpp:104-194) is the same call with
 * remove_dead = 1: *npart_inout_host becomes the number of alive particles.
 * remove_dead | EB200_SORT_SKIP_PREV: i*_prev / dx*_prev are left where they are. Between the
 * deposit of one step and the pusher of the next they are dead values (the pusher overwrites
 * them before anything reads them, sr.hpp:137-153), which is where SortParticles runs
 * (srpic.hpp:184-186): eb200_srpic_step sorts this way after eb200_set_lean_prev(ctx, 1). */
#define EB200_SORT_SKIP_PREV 2
/* remove_dead | EB200_SORT_UNSTABLE: counting sort by cell (histogram, scan, slot assignment: 18 B
 * per particle for the permutation instead of the radix sort's 60 B). Particles are grouped by
 * cell as before and not-alive ones go last, but the order INSIDE a cell is unspecified and not
 * reproducible from run to run -- as in the reference, whose sort assigns slots with atomics
 * (particles_sort.cpp:118-147). eb200_srpic_step sorts this way on fast-build contexts
 * (strict_fp = 0), whose deposit sums are unordered anyway; strict contexts keep the stable
 * radix sort (bit-reproducible ORDERED deposit). eb200_set_sort_mode overrides. */
#define EB200_SORT_UNSTABLE 4
int eb200_sort_particles(eb200_ctx_t* ctx, const eb200_prtls_t* prtls,
                         uint32_t* npart_inout_host, int remove_dead, eb200_stream_t stream);

/* on = 1: the caller declares that it does not read i*_prev / dx*_prev between calls. They are
 * scratch of one step -- written by the pusher, read by the deposit of the same step, and
 * overwritten by the next push before anything reads them (sr.hpp:137-153) -- so the fused
 * push+deposit kernel then does not store them (62 instead of 78 bytes per particle-step in
 * 2D), the step mirrors' sort does not permute them and eb200_srpic_step_host moves them in
 * neither direction. Their content is unspecified afterwards. Default 0: the arrays are left
 * exactly as the reference leaves them. */
int eb200_set_lean_prev(eb200_ctx_t* ctx, int on);
/* Sort used by the step mirrors: -1 = by build (default: counting on fast, radix on strict),
 * 0 = stable radix sort, 1 = counting sort. */
int eb200_set_sort_mode(eb200_ctx_t* ctx, int mode);

/* ------------------------------------------------------------- field boundaries */
/* kernel::bc::MatchBoundaries_kernel<SRPIC, Minkowski<D>, FS, o> (src/kernels/fields_bcs.hpp:
 * 42-560) as launched by srpic::MatchFieldsIn (src/engines/srpic/fields_bcs.h:38-215) over the
 * ghost-inclusive index range [range_min, range_max) that Mesh::ExtentToRange gives it:
 *   F = s F + (1 - s) transform<T->U>(target),  s = tanh(|x_o - xg_edge| 4 / ds)
 * per component, at that component's staggered node. `target` (6 component planes, layout of em)
 * holds the values of the problem generator's MatchFields functor at those nodes (tetrad basis):
 * the functor itself cannot cross a C ABI, the host evaluates it (once, or per step when it
 * depends on time). components_mask: bit c set = the functor defines component c (ex1, ex2, ex3,
 * bx1, bx2, bx3): the reference skips what the functor lacks. tags: EB200_BC_E | EB200_BC_B
 * (BC::E / BC::B, src/global/enums.h). o = matching direction (0, 1, 2). */
enum { EB200_BC_E = 1, EB200_BC_B = 2 };
int eb200_match_fields(eb200_ctx_t* ctx, float* em, const float* target, int o, float xg_edge,
                       float ds, int tags, int components_mask, const int* range_min,
                       const int* range_max, eb200_stream_t stream);

/* ------------------------------------------------------------- reduced statistics */
/* kernel::ReducedFields_kernel (src/kernels/reduced_stats.hpp:25-386) through ReduceFields
 * (src/framework/domain/metadomain_stats.cpp:128-183), Minkowski 1D/2D/3D and 2D (q)spherical
 * SRPIC meshes: the sum over the
 * active cells of this domain of B_I^2, E_I^2, (E x B)_I (I = comp, 1..3) or J.E, each term
 * evaluated as the reference does (cell-centred averages, tetrad components, sqrt(det h)).
 * Returns the LOCAL sum (before the MPI reduction and the division by totVolume the reference's
 * stats writer applies). Synchronises the stream: the result is a host value, as the
 * reference's parallel_reduce. cur may be NULL unless what == EB200_STATS_JDOTE. */
enum { EB200_STATS_B2 = 0, EB200_STATS_E2 = 1, EB200_STATS_EXB = 2, EB200_STATS_JDOTE = 3 };
int eb200_stats_fields(eb200_ctx_t* ctx, const float* em, const float* cur, int what, int comp,
                       double* out_host, eb200_stream_t stream);
/* kernel::ReducedParticleMoments_kernel (reduced_stats.hpp:400-536) through ComputeMoments
 * (metadomain_stats.cpp:72-126) for ONE species of a Minkowski or 2D (q)spherical SRPIC domain
 * (dV = sqrt_det_h at the particle; momenta to the tetrad basis at (x, phi)): Npart (alive count), N,
 * Rho, Charge (sum of dV * (use_weights ? weight : 1 | mass | charge)) or the stress-energy
 * component T^{c1 c2} (c = 0: energy, 1..3: u_c; sum of dV * coeff / energy -- the reference
 * applies no weight there). LOCAL sum of one species; the caller adds species and normalises by
 * totVolume * ppc0 as ComputeMoments does. Synchronises the stream. */
enum { EB200_STATS_NPART = 0, EB200_STATS_N = 1, EB200_STATS_RHO = 2, EB200_STATS_CHARGE = 3,
       EB200_STATS_T = 4 };
int eb200_stats_particles(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart, float mass,
                          float charge, int use_weights, int what, int c1, int c2,
                          double* out_host, eb200_stream_t stream);

/* ------------------------------------------------------------- whole SRPIC step */
/* One species as the engine sees it: ntt::ParticleSpecies (src/framework/containers/species.h)
 * + the SoA arrays + the live particle count. */
typedef struct {
  float         mass, charge;
  int           pusher_flags; /* EB200_PUSHER_NONE: species is neither pushed nor deposited */
  int           drag_flags;
  uint32_t      npart;        /* in/out */
  uint32_t      maxnpart;
  eb200_prtls_t arrays;
} eb200_species_t;

/* The SimulationParams entries the SRPIC dispatchers read (src/engines/srpic/*.h), by value. */
typedef struct {
  float dt;           /* algorithms.timestep.dt = CFL * dx0 (algorithms.cpp:21-22) */
  float correction;   /* algorithms.timestep.correction */
  float omegaB0;      /* scales.omegaB0 = 1 / larmor0 */
  float q0, B0, V0;   /* scales.* (parameters.cpp:47-78) */
  float ppc0;         /* particles.ppc0 */
  int   nfilter;      /* algorithms.current_filters */
  int   fieldsolver_enabled, deposit_enabled;
  float stencil[9];   /* algorithms.fieldsolver.{delta_x,delta_y,beta_xy,beta_yx,delta_z,beta_xz,beta_zx,beta_yz,beta_zy} */
  int   fbc[6];       /* EB200_FBC_* per face */
  int   pbc[6];       /* EB200_PBC_* per face */
  float gca_larmor_max, gca_e_ovr_b_max;         /* algorithms.gca.* */
  float sync_gamma_rad, compton_gamma_rad;       /* radiation.drag.*.gamma_rad */
  int   fuse_push_deposit; /* 1: one pass over the particles for push + deposit */
  int   deposit_mode;      /* EB200_DEPOSIT_* */
  int   sort_interval;     /* particles.spatial_sorting_interval (0: never) */
  int   clear_interval;    /* particles.clear_interval (0: never) */
  /* curvilinear (spherical / qspherical) domains */
  float n0;                /* scales.n0: CurrentsAmpere uses coeff = -dt q0 n0 / B0, 1 / n0 */
  /* the atmosphere as srpic::ParticlePush hands it to the pusher (particle_pusher.h:45-80):
   * g along the boundary's direction with its sign, the surface coordinate and
   * grid.boundaries.atmosphere.ds; has_atmosphere = a face carries PrtlBC::ATMOSPHERE */
  int   has_atmosphere;
  float atm_g[3], atm_x_surf, atm_ds;
} eb200_srpic_params_t;

/* SRPICEngine::step_forward (src/engines/srpic/srpic.hpp:65-188) for one domain:
 * Faraday(1/2) -> comm B -> ParticlePush -> CurrentsDeposit -> sync J, comm J -> CurrentsFilter ->
 * [particle migration] -> Faraday(1/2) -> comm B -> Ampere -> CurrentsAmpere -> comm E|J ->
 * SortParticles, with srpic::FieldBoundaries after every field exchange. Minkowski contexts:
 * 1D / 2D / 3D, MATCH faces registered with eb200_srpic_set_match. Spherical / qspherical
 * contexts (2D): the curvilinear dispatchers (eb200_faraday_sr, ...), AXIS faces from
 * prm->fbc, MATCH and ATMOSPHERE faces registered with eb200_srpic_set_field_bcs, the
 * atmosphere's gravity in the pusher; push and deposit run as two passes there. The injectors
 * (srpic::ParticleInjector, the pgen's CustomPostStep) stay the host's. species[s].npart is
 * updated in place. */
int eb200_srpic_step(eb200_ctx_t* ctx, const eb200_srpic_params_t* prm, float* em, float* cur,
                     float* buff, eb200_species_t* species, int nspecies, uint32_t step,
                     double time, eb200_stream_t stream);

/* srpic::FieldBoundaries with MATCH faces inside eb200_srpic_step (src/engines/srpic/srpic.hpp:
 * 74-80, 93-101, 144-152, 168-176; src/engines/srpic/fields_bcs.h:38-215): the faces registered
 * here are matched (eb200_match_fields, in the order given) for B after each Faraday half-step's
 * exchange, for E after the Ampere / CurrentsAmpere exchange, and for both at step 0. `target`
 * and the face table are borrowed until the next call (nfaces = 0 clears them); faces of a
 * matched dimension carry EB200_FBC_NONE in eb200_srpic_params_t.fbc. */
/* Functor-driven field boundaries of a curvilinear SRPIC domain inside eb200_srpic_step
 * (srpic::MatchFieldsIn / AtmosphereFieldsIn, src/engines/srpic/fields_bcs.h:39-215, 470-600):
 * applied in the order of dir::Directions<D>::orth (-x1, -x2, +x2, +x1) together with the AXIS
 * faces of prm->fbc. `target` (device pointer, layout of em, contravariant components on each
 * component's node) stands for the pgen's MatchFields / AtmFields functor and is borrowed until
 * the next call; n = 0 clears the table. */
typedef struct {
  int          kind;    /* EB200_FBC_MATCH or EB200_FBC_ATMOSPHERE */
  int          o, sign; /* face: dimension 0 / 1 and side */
  float        xg_edge, ds;                /* MATCH: edge of the global box, layer thickness */
  int          i_edge;                     /* ATMOSPHERE: ghost-inclusive cell index of the edge */
  int          range_min[2], range_max[2]; /* ghost-inclusive cell range (Mesh::ExtentToRange) */
  const float* target;
  int          mask;                       /* components the functor defines */
} eb200_field_bc_t;
int eb200_srpic_set_field_bcs(eb200_ctx_t* ctx, const eb200_field_bc_t* bcs, int n);

typedef struct {
  int   o;                          /* matching direction */
  float xg_edge, ds;                /* edge of the global box on that side, layer thickness */
  int   range_min[3], range_max[3]; /* ghost-inclusive cell range of the layer in this domain */
} eb200_match_face_t;
/* The geometry srpic::MatchFieldsIn hands to the kernel (src/engines/srpic/fields_bcs.h:72-114
 * with Mesh::Intersects / Mesh::ExtentToRange, src/framework/domain/mesh.h:69-203) for the
 * face (o, sign) of the GLOBAL box and a Minkowski domain with the given local extent: the
 * layer [edge - ds, edge] (sign > 0) or [edge, edge + ds] (sign < 0), ghosts included on the
 * outer side and over the whole transverse extent. Returns 1 and fills *face when the layer
 * intersects the domain, 0 when it does not (MatchFieldsIn then returns without a launch),
 * < 0 on a bad argument. Pure host code, fp32 like the reference (the cell size used for the
 * index range is the domain metric's own, (local_xmax[0] - local_xmin[0]) / n[0]). */
int eb200_match_layer(const eb200_grid_t* local_grid, float dx, const float* local_xmin,
                      const float* local_xmax, float global_xmin_o, float global_xmax_o, int o,
                      int sign, float ds, eb200_match_face_t* face);
int eb200_srpic_set_match(eb200_ctx_t* ctx, const eb200_match_face_t* faces, int nfaces,
                          const float* target, int components_mask);

/* Metadomain::CommunicateParticles + Particles::Communicate for all species in one round
 * (metadomain_comm.cpp:565-653, particles_comm.cpp:180-389, kernels/comm.hpp): particles
 * tagged by the pusher with a send tag move to the neighbour in that direction with their
 * cell indices shifted; received particles fill dead/sent slots first, then extend npart.
 * Synchronises the stream once (particle counts are host-visible results). */
int eb200_comm_particles(eb200_ctx_t* ctx, eb200_species_t* species, int nspecies,
                         eb200_stream_t stream);

/* ------------------------------------------------------------------- profiling */
/* Phases of eb200_srpic_step, bracketed by CUDA events on the caller's stream when profiling
 * is enabled (names follow the reference's timers, src/engines/engine.hpp:245-257). */
enum {
  EB200_PHASE_FIELDSOLVER = 0, /* Faraday x2 + Ampere + CurrentsAmpere */
  EB200_PHASE_PUSH_DEPOSIT = 1, /* ParticlePusher + CurrentDeposit (incl. zeroing J) */
  EB200_PHASE_FILTER = 2,       /* CurrentFiltering */
  EB200_PHASE_COMM = 3,         /* Communications: field ghost fill, J sync */
  EB200_PHASE_SORT = 4,         /* ParticleSort */
  EB200_PHASE_MIGRATION = 5,    /* Communications: particle migration (multi-domain only) */
  EB200_NPHASES = 6
};
int eb200_profile_enable(eb200_ctx_t* ctx, int on);
/* synchronises the recorded events; ms_host[EB200_NPHASES], calls_host[EB200_NPHASES]
 * (accumulated since enable; reading resets the accumulators) */
int eb200_profile_read(eb200_ctx_t* ctx, float* ms_host, int* calls_host);

/* ---------------------------------------------------------- host-buffer entry */
/* The same step for a caller whose state lives in HOST memory (pinned or pageable): copies
 * em, cur and every species array to context-owned device mirrors, runs eb200_srpic_step,
 * copies everything back, and returns after the stream has drained. Pointers inside
 * species_host[s].arrays are host pointers. bytes_h2d/bytes_d2h (may be NULL) receive the
 * bytes moved. */
int eb200_srpic_step_host(eb200_ctx_t* ctx, const eb200_srpic_params_t* prm, float* em_host,
                          float* cur_host, eb200_species_t* species_host, int nspecies,
                          uint32_t step, double time, uint64_t* bytes_h2d, uint64_t* bytes_d2h);

/* ============================================================ curvilinear SR (2D) */
/* For EB200_METRIC_SPHERICAL / QSPHERICAL contexts eb200_push_sr and eb200_deposit run the
 * curvilinear branches of the same reference kernels (sr.hpp:573-657 position push through
 * Cartesian space, phi carried in prtls->phi; currents_deposit.hpp:133-139 velocity transform).
 * pusher->dx / xmin are ignored; atm_gx2 / atm_gx3 must be zero (the reference raises). The
 * field solvers of the curvilinear path have their own argument lists: */
/* kernel::sr::Faraday_kernel (src/kernels/faraday_sr.hpp:53-89), coeff = dT;
 * fbc_host[6]: only the AXIS flags of the x2 faces are read */
int eb200_faraday_sr(eb200_ctx_t* ctx, float* em, float coeff, const int* fbc_host,
                     eb200_stream_t stream);
/* kernel::sr::Ampere_kernel (src/kernels/ampere_sr.hpp:58-109) over srpic::RangeWithAxisBCs */
int eb200_ampere_sr(eb200_ctx_t* ctx, float* em, float coeff, const int* fbc_host,
                    eb200_stream_t stream);
/* kernel::sr::CurrentsAmpere_kernel (ampere_sr.hpp:163-209): coeff = -dt q0 n0 / B0 */
int eb200_currents_ampere_sr(eb200_ctx_t* ctx, float* em, float* cur, float coeff, float inv_n0,
                             const int* fbc_host, eb200_stream_t stream);
/* eb200_filter on such a context runs DigitalFilter_kernel<Dim::_2D, Coord::Spherical>
 * (digital_filter.hpp:196-281) with the axis rows treated as the reference does. */

/* ====================================================================== GRPIC (2D) */
/* kernel::gr::PusherContext + PusherBoundaries (src/kernels/pushers/context.h:182-228) */
typedef struct {
  int   pusher_flags; /* EB200_PUSHER_BORIS (massive) or EB200_PUSHER_PHOTON (massless) */
  float mass, charge;
  float dt, omegaB0;
  float epsilon;      /* algorithms.gr.pusher_eps (unused by the analytic-derivative pusher) */
  int   niter;        /* algorithms.gr.pusher_niter */
  int   pbc[6];       /* EB200_PBC_ABSORB (incl. HORIZON) on x1 faces, EB200_PBC_AXIS on x2 faces */
  int   tag_outgoing;
} eb200_pusher_gr_t;
/* grpic::ParticlePush for one species (src/engines/grpic/particle_pusher.h:27-89,
 * src/kernels/pushers/gr.hpp:668-865): em holds D (components 0..2), em0 holds B (3..5);
 * ux1..3 are the covariant momentum components. */
int eb200_push_gr(eb200_ctx_t* ctx, const eb200_pusher_gr_t* pusher, const eb200_prtls_t* prtls,
                  uint32_t npart, const float* em, const float* em0, eb200_stream_t stream);
/* kernel::gr::ComputeAuxE_kernel / ComputeAuxH_kernel (src/kernels/aux_fields_gr.hpp:49-226)
 * over grpic::range_with_axis_BCs; D / B / out are 6-component fields (the reference passes
 * em or em0 for either, src/engines/grpic/fieldsolvers.h:79-128) */
int eb200_gr_aux_e(eb200_ctx_t* ctx, const float* d_fld, const float* b_fld, float* e_out,
                   const int* fbc_host, eb200_stream_t stream);
int eb200_gr_aux_h(eb200_ctx_t* ctx, const float* d_fld, const float* b_fld, float* h_out,
                   const int* fbc_host, eb200_stream_t stream);
/* kernel::gr::Faraday_kernel (src/kernels/faraday_gr.hpp:63-93): b_out = b_in + dT curl E */
int eb200_faraday_gr(eb200_ctx_t* ctx, const float* b_in, float* b_out, const float* e_aux,
                     float coeff, const int* fbc_host, eb200_stream_t stream);
/* kernel::gr::Ampere_kernel (src/kernels/ampere_gr.hpp:65-102): d_out = d_in + dT curl H. In
 * place (d_in == d_out) the axis rows follow the reference's serial row order. */
int eb200_ampere_gr(eb200_ctx_t* ctx, const float* d_in, float* d_out, const float* h_aux,
                    float coeff, const int* fbc_host, eb200_stream_t stream);
/* kernel::gr::CurrentsAmpere_kernel (ampere_gr.hpp:143-176): coeff = -dt q0 / B0 */
int eb200_currents_ampere_gr(eb200_ctx_t* ctx, float* d_fld, const float* cur, float coeff,
                             const int* fbc_host, eb200_stream_t stream);
/* kernel::gr::TimeAverageDB_kernel / TimeAverageJ_kernel (aux_fields_gr.hpp:253-302):
 * a = (a + b) / 2 on the active cells of an ncomp-component field */
int eb200_time_average(eb200_ctx_t* ctx, float* a, const float* b, int ncomp,
                       eb200_stream_t stream);

/* ============ field boundaries of 2D curvilinear SRPIC and GRPIC domains ==========
 * What srpic::FieldBoundaries / grpic::FieldBoundaries launch (src/engines/srpic/fields_bcs.h:
 * 39-672, src/engines/grpic/fields_bcs.h:42-270). tags: EB200_BC_E = the first three components
 * of the field (E, D, aux E), EB200_BC_B = the last three (B, aux H) -- what the reference's
 * BC::E | BC::D and BC::B | BC::H select. Ranges are ghost-inclusive [min, max) per dimension,
 * as Mesh::ExtentToRange returns them. Functor arguments of the reference's kernels (the pgen's
 * MatchFields / AtmFields / init_flds) arrive as `target`: six component planes in the layout
 * of the field, every value taken at the component's own node, in the basis the reference's
 * kernel blends with (SRPIC: metric.transform<c, Idx::T, Idx::U>(node, f(x_Ph)); GRPIC: f(x_Ph)
 * as returned); `mask` bit c = the functor defines component c. */
/* kernel::bc::AxisBoundaries_kernel<Dim::_2D, P> (fields_bcs.hpp:824-868) on the x2 face given
 * by sign (< 0: i2min, > 0: i2max); fld = em, em0 (6 components) */
int eb200_axis_fields(eb200_ctx_t* ctx, float* fld, int sign, int tags, eb200_stream_t stream);
/* kernel::bc::gr::HorizonBoundaries_kernel<Dim::_2D> (fields_bcs.hpp:1187-1240) at i1min;
 * nfilter = algorithms.current_filters; fld = em, em0 or aux */
int eb200_horizon_fields(eb200_ctx_t* ctx, float* fld, int tags, int nfilter,
                         eb200_stream_t stream);
/* kernel::bc::MatchBoundaries_kernel<S, M, FS, o> for a non-Cartesian metric (fields_bcs.hpp:
 * 176-340): F = s F + (1 - s) target, s = tanh(|x_o - xg_edge| 4 / ds) at the component's node;
 * fbc_host[6]: the AXIS flags of the x2 faces switch off the target of the third E / D
 * component on the axis rows. o = 0 (x1) or 1 (x2). */
int eb200_match_fields_curv(eb200_ctx_t* ctx, float* fld, const float* target, int o,
                            float xg_edge, float ds, int tags, int components_mask,
                            const int* range_min, const int* range_max, const int* fbc_host,
                            eb200_stream_t stream);
/* kernel::bc::EnforcedBoundaries_kernel<M, FS, P, O> (fields_bcs.hpp:870-1185; the ATMOSPHERE
 * field boundary of srpic::AtmosphereFieldsIn, fields_bcs.h:470-600): inside the range the
 * defined components are SET to target, the normal E and tangential B components only on the
 * far side of i_edge (ghost-inclusive cell index; i >= i_edge for sign > 0, i < i_edge else) */
int eb200_enforce_fields(eb200_ctx_t* ctx, float* em, const float* target, int o, int sign,
                         int i_edge, int tags, int components_mask, const int* range_min,
                         const int* range_max, eb200_stream_t stream);
/* kernel::bc::gr::AbsorbCurrents_kernel<M, 1> (fields_bcs.hpp:1242-1283; grpic gr_bc::curr):
 * J *= tanh(|r - xg_edge| / (ds / 4)) inside the range; cur = the 3-component cur0 */
int eb200_absorb_currents_gr(eb200_ctx_t* ctx, float* cur, float xg_edge, float ds,
                             const int* range_min, const int* range_max, eb200_stream_t stream);
/* kernel::bc::ConductorBoundaries_kernel<Dim::_2D, o, P> over the range of
 * srpic::PerfectConductorFieldsIn (fields_bcs.h:384-470), Minkowski 2D domains */
int eb200_conductor_fields(eb200_ctx_t* ctx, float* em, int o, int sign, int tags,
                           eb200_stream_t stream);

/* ------------------------------------------------------------ the GRPIC step (2D) */
/* What GRPICEngine::step_forward reads from SimulationParams / Domain, by value. */
typedef struct {
  float dt;           /* algorithms.timestep.dt */
  float correction;   /* algorithms.timestep.correction */
  float omegaB0;      /* scales.omegaB0 */
  float q0, B0;       /* scales.* */
  int   nfilter;      /* algorithms.current_filters */
  int   fieldsolver_enabled, deposit_enabled;
  int   fbc[6];       /* EB200_FBC_* per face: x1 = {HORIZON, MATCH}, x2 = {AXIS, AXIS} */
  int   pbc[6];       /* EB200_PBC_* per face: x1 = {ABSORB (horizon), ABSORB}, x2 = {AXIS, AXIS} */
  float pusher_eps;   /* algorithms.gr.pusher_eps */
  int   pusher_niter; /* algorithms.gr.pusher_niter */
  int   deposit_mode; /* EB200_DEPOSIT_* */
  int   sort_interval, clear_interval;
  /* the +x1 MATCH layer of grpic::MatchFieldsIn (fields_bcs.h:42-133): edge of the global box,
   * grid.boundaries.match.ds and the ghost-inclusive cell range Mesh::ExtentToRange gives;
   * currents are absorbed in the same layer (gr_bc::curr) where pbc[1] is ABSORB */
  float match_xg_edge, match_ds;
  int   match_range_min[2], match_range_max[2];
  int   match_mask;   /* components pgen.init_flds defines */
} eb200_grpic_params_t;
/* One call = GRPICEngine::step_forward (src/engines/grpic/grpic.hpp:66-634) for a single 2D
 * Kerr-Schild type domain, including the start-up sequence of step 0: time averages, auxiliary
 * fields, the two Faraday / Ampere sub-steps with currents, pusher, deposit into cur0, filter,
 * every grpic::FieldBoundaries call (MATCH towards match_target = pgen.init_flds tabulated on
 * each component's node, AXIS, HORIZON, AbsorbCurrents), SwapFields and SortParticles.
 * SwapFields exchanges the caller's POINTERS (the reference swaps its views, utils.h:31-35):
 * *em <-> *em0 and *cur <-> *cur0 on return. aux and buff are 6- / 3-component scratch fields
 * of the host (Fields::aux, Fields::buff). */
int eb200_grpic_step(eb200_ctx_t* ctx, const eb200_grpic_params_t* prm, float** em, float** em0,
                     float** cur, float** cur0, float* aux, float* buff,
                     const float* match_target, eb200_species_t* species, int nspecies,
                     uint32_t step, double time, eb200_stream_t stream);

/* --------------------------------- injection and particle moments */
/* arch::energy_dist::Maxwellian (src/archetypes/energy_dist.h:170-290): temperature in m c^2
 * (0 = cold), drift four-velocity (Cartesian meshes only; any direction) */
typedef struct {
  float temperature;
  float drift_u[3];
} eb200_maxwellian_t;
/* spatial distributions: the functor of arch::InjectNonUniform cannot cross a C ABI */
enum {
  EB200_SDIST_UNIFORM   = 0, /* spatial_dist = 1 (what arch::InjectUniform draws in expectation) */
  EB200_SDIST_TABLE     = 1, /* field = the functor's value at every cell centre */
  EB200_SDIST_REPLENISH = 2, /* arch::spatial_dist::ReplenishUniform (spatial_dist.h:87-125):
                                field = density moment, refilled up to target_density where it
                                fell below 0.9 of it */
  EB200_SDIST_REPLENISH_TABLE = 3, /* arch::spatial_dist::Replenish<M, N, T> (spatial_dist.h:
                                28-81): (target - density) / target_max where density < 0.9
                                target; target_field = the target functor at every cell centre */
  EB200_SDIST_ATMOSPHERE = 4 /* Replenish with arch::AtmosphereDensityProfile (particle_injector.h:
                                141-190) evaluated in the kernel: atm_* below, target_max =
                                atm_nmax */
};
typedef struct {
  int          kind;
  const float* field; /* device: component plane(s) in the mesh layout (ghost-inclusive) */
  int          comp;  /* component of `field` to read */
  float        target_density;
  const float* target_field; /* REPLENISH_TABLE: one plane, mesh layout */
  float        target_max;
  int          atm_dim, atm_sign; /* ATMOSPHERE: direction of the boundary (dim 0..2, sign -1 / +1) */
  float        atm_nmax, atm_height, atm_xsurf, atm_ds;
  float        inv_V0; /* 1 / scales.V0: weight = sqrt_det_h(cell centre) * inv_V0 on curvilinear
                          meshes (injectors.hpp:746-748); unused on Minkowski meshes */
} eb200_spatial_dist_t;
/* arch::InjectNonUniform (src/archetypes/particle_injector.h:296-387) with
 * kernel::NonUniformInjector_kernel (src/kernels/injectors.hpp:526-859) on a Minkowski or a 2D
 * (q)spherical SRPIC domain: in every cell of the ghost-inclusive range, ppc = number_density *
 * ppc0 / 2 * spatial_dist pairs (the fraction rounded stochastically), both species of a pair at
 * the same position, velocities from ed1 / ed2 (curvilinear: drawn in the tetrad basis at the
 * cell centre and stored Cartesian, transform_xyz<T, XYZ> with phi = 0; phi = 0), weight 1
 * (curvilinear: sqrt_det_h(cell centre) / V0), appended at npart of either species (both npart
 * grow by the same number; returns EB200_ERR_CAPACITY and injects nothing when maxnpart would be
 * exceeded).
 * Every cell draws from its own counter-based Philox4x32-10 stream keyed by (seed, step, call,
 * cell): the result is a pure function of the arguments, identical on every device and for
 * every launch shape; `call` distinguishes several injections of one step. Synchronises the
 * stream once (the new particle count). */
int eb200_inject_nonuniform(eb200_ctx_t* ctx, eb200_species_t* species1, eb200_species_t* species2,
                            float ppc, const eb200_spatial_dist_t* sdist,
                            const eb200_maxwellian_t* ed1, const eb200_maxwellian_t* ed2,
                            const int* range_min, const int* range_max, uint64_t seed,
                            uint32_t step, uint32_t call, eb200_stream_t stream);
/* arch::ComputeMomentWithSpecies / kernel::ParticleMoments_kernel (src/archetypes/utils.h:
 * 136-169, src/kernels/particle_moments.hpp:37-419) for the scalar moments what =
 * EB200_STATS_N | RHO | CHARGE | NPART (Nppc), smoothing order 0 (the archetypes' default), any
 * metric of the context (the volume element is sqrt_det_h at the particle's cell centre,
 * particle_moments.hpp:306-327): adds one species into component comp of buff (ncomp planes; the
 * host zeroes it once, as ComputeMomentWithSpecies does before its species loop). */
int eb200_particle_moment(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart, float mass,
                          float charge, int use_weights, int what, float inv_n0, float* buff,
                          int ncomp, int comp, eb200_stream_t stream);

/* srpic::AtmosphereParticlesIn (src/engines/srpic/particles_bcs.h:33-153), what
 * srpic::ParticleInjector runs for a face whose particle boundary is ATMOSPHERE: the Rho moment
 * of the two species into one plane, then InjectNonUniform<Replenish<AtmosphereDensityProfile>>
 * with two Maxwellians of the atmosphere's temperature over the active cells. */
typedef struct {
  int      dim, sign;        /* the face: direction.get_dim(), direction.get_sign() */
  float    x_surf, ds;       /* GetAtmosphereExtent (src/engines/srpic/utils.h:47-103): x_surf =
                                xg_min (sign > 0) or xg_max; grid.boundaries.atmosphere.ds */
  float    height, temperature, density; /* grid.boundaries.atmosphere.{height, temperature, density} */
  int      species[2];       /* 0-based indices of grid.boundaries.atmosphere.species */
  float    inv_n0, inv_V0, ppc0; /* 1 / scales.n0, 1 / scales.V0, particles.ppc0 */
  uint64_t seed;             /* of the counter-based streams (eb200_inject_nonuniform) */
} eb200_atmosphere_t;
/* plane: scratch of one component plane (the reference uses bckp component 0). assume_empty =
 * Inj::AssumeEmpty (density taken as zero). npart of both species updated in place. */
int eb200_atmosphere_particles(eb200_ctx_t* ctx, const eb200_atmosphere_t* atm,
                               eb200_species_t* species, int nspecies, float* plane,
                               int assume_empty, uint32_t step, eb200_stream_t stream);
/* Registers (atm != NULL) or clears the atmosphere injector of eb200_srpic_step: it then runs
 * where SRPICEngine::step_forward calls srpic::ParticleInjector (srpic.hpp:81, 179-183: at step 0
 * after the first FieldBoundaries, and at the end of every step before the sort), with buff
 * component 0 as the density plane. */
int eb200_srpic_set_atmosphere_injector(eb200_ctx_t* ctx, const eb200_atmosphere_t* atm);

/* ------------------------------------------------ emission policies of the SR pusher */
/* arch::emission::Synchrotron / Compton (src/archetypes/emission/synchrotron.h:29-287,
 * compton.h:28-235), the policies kernel::sr::MakePusherPolicyEmission builds from
 * radiation.emission.* (src/kernels/pushers/sr_policies.h:53-111) and the pusher runs through
 * processEmission (sr.hpp:290-331, 1501-1555): after the velocity update of a massive particle,
 * with u' = (u_before + u_after) / 2 and the interpolated Cartesian E, B,
 *   synchrotron: p = nominal_probability (-kappaR . u' / (gamma^2 |u'|) + beta chiR^2),
 *   Compton:     p = nominal_probability beta,
 * photon energy gamma^2 nominal_photon_energy; a photon is emitted when a uniform draw < p and the
 * energy is below 20 % of (gamma - 1) m and not below photon_energy_min; the emitter recoils
 * (drag) when should_drag and the draw succeeded. With an emission policy the continuous drag of
 * pusher->drag_flags is NOT applied (sr.hpp:311-328). The photon is appended to `photons` at its
 * emitter's position BEFORE the position push, momentum = energy along -delta_u, weight =
 * photon_weight * emitter weight.
 * nominal_probability = |q / m| * radiation.emission.<kind>.nominal_probability and
 * nominal_photon_energy = m * radiation.emission.<kind>.nominal_photon_energy (the policies'
 * constructors). The draw is the first number of the Philox stream (seed, step, call, particle
 * index): reproducible for any launch shape, where the reference's pool is not.
 * Minkowski domains, massive emitters; the unfused pusher. */
enum { EB200_EMISSION_NONE = 0, EB200_EMISSION_SYNCHROTRON = 1, EB200_EMISSION_COMPTON = 2 };
typedef struct {
  int           kind;
  float         photon_weight;         /* radiation.emission.<kind>.photon_weight */
  float         photon_energy_min;     /* radiation.emission.<kind>.photon_energy_min */
  float         nominal_probability;   /* see above */
  float         nominal_photon_energy; /* see above */
  int           should_drag;           /* radiative_drag_flags & SYNCHROTRON / COMPTON */
  eb200_prtls_t photons;               /* arrays of the emitted (photon) species */
  uint32_t      photon_npart;          /* in: append position; out: += number emitted */
  uint32_t      photon_maxnpart;
  uint64_t      seed;
  uint32_t      step, call;
} eb200_emission_t;
/* eb200_push_sr with an emission policy. emission->photon_npart is updated (one stream
 * synchronisation: the count is a host-visible result, as Particles::set_npart after the
 * reference's pusher, particle_pusher.h:161-183). EB200_ERR_CAPACITY when the photons do not
 * fit: the emitters are pushed, the photons beyond the capacity are dropped. */
int eb200_push_sr_emission(eb200_ctx_t* ctx, const eb200_pusher_t* pusher, const eb200_prtls_t* prtls,
                           uint32_t npart, const float* em, eb200_emission_t* emission,
                           eb200_stream_t stream);

/* Registers (policy != NULL) or clears the emission policy of species `species` for
 * eb200_srpic_step: its pusher then runs eb200_push_sr_emission with the arrays of species
 * `photon_species` as the emitted species (photons / photon_npart / photon_maxnpart / step / call of
 * `policy` are filled per step; call = the emitter's index), followed by its deposit; the emitted
 * species' npart grows before its own turn in the species loop, as in srpic::ParticlePush
 * (particle_pusher.h:161-183). */
int eb200_srpic_set_emission(eb200_ctx_t* ctx, int species, int photon_species,
                             const eb200_emission_t* policy);

/* ------------------------------------------------ output staging */
/* kernel::FieldsToPhys_kernel<M, N1, N2> over Mesh::rangeActiveCells (src/kernels/
 * fields_to_phys.hpp:33-239; what the writer launches per output field, src/output/
 * fields.cpp): three components comps_from[3] of `from` (ncomp_from planes) are interpolated to
 * the cell centre and / or converted to another basis and stored in components comps_to[3] of
 * `to`. interp: 0 none, 1 InterpToCellCenterFromEdges (E, D, J), 2 ...FromFaces (B, H);
 * convert: 0 none, 1 ConvertToHat (U -> T), 2 ConvertToPhysCntrv (U -> PU), 3 ConvertToPhysCov
 * (D -> PD). Any metric of the context; 1D / 2D / 3D for Minkowski. */
int eb200_fields_to_phys(eb200_ctx_t* ctx, const float* from, int ncomp_from, float* to,
                         int ncomp_to, const int* comps_from, const int* comps_to, int interp,
                         int convert, eb200_stream_t stream);
/* kernel::PrtlToPhys_kernel<S, M, false> (src/kernels/prtls_to_phys.hpp:30-218): every stride-th
 * particle (nout = ceil(npart / stride) samples) -> physical coordinates (x3 = phi on 2D
 * curvilinear meshes; unused buffers may be NULL), velocities in the orthonormal frame (SRPIC)
 * or physical covariant components (GRPIC), and the weight. Payload columns are plain strided
 * copies and stay with the host's writer. */
int eb200_prtls_to_phys(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart,
                        uint32_t stride, float* x1, float* x2, float* x3, float* u1, float* u2,
                        float* u3, float* weight, eb200_stream_t stream);

/* Host-side evaluation of the metric functions the kernels use (same source, compiled for the
 * host): what the reference's setup code gets from metric.h_<i,j>() etc. (src/metrics/*.h).
 * n_active[2], metric_params[8] as in eb200_config_t. out[nq][16] for the SR metrics:
 * h_11 h_22 h_33 sqrt_h_11 sqrt_h_22 sqrt_h_33 sqrt_det_h polar_area r theta x1(r) x2(theta);
 * out[nq][32] for the GR metrics: h_11 h_22 h_33 h_13 h^11 h^22 h^33 h^13 alpha beta^1
 * sqrt_det_h sqrt_det_h_tilde polar_area dr_alpha dt_alpha dr_beta1 dt_beta1 dr_h11 dr_h22
 * dr_h33 dr_h13 dt_h11 dt_h22 dt_h33 dt_h13 theta x2(theta). Needs no device. */
int eb200_metric_eval(int metric, const int* n_active, const float* metric_params, int nq,
                      const float* x1_host, const float* x2_host, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* ENTITY_B200_H */
