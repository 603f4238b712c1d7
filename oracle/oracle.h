/* TEST INFRASTRUCTURE ONLY -- C interface of the CPU oracle (liborc.so).
 *
 * The oracle restates, in scalar fp32 C++ and in program (serial particle)
 * order, the reference kernels of the PIC hot path. It is the checker for the
 * CUDA library, never a fallback: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Layouts (identical to the product ABI, include/entity_b200.h):
 *  - fields: fp32, extents (n_d + 2*ng) per simulated dimension, i1 fastest,
 *    then i2, i3, component slowest (Kokkos LayoutLeft = what the reference's
 *    CUDA build holds: src/framework/containers/fields.h:38-108).
 *  - particles: SoA, pointer order of ParticleArrays
 *    (src/framework/containers/particles.h:47-71).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int dim;  /* 1,2,3 */
  int n[3]; /* active cells per dimension (unused dims: 1) */
  int ng;   /* N_GHOSTS: src/global/global.h:130-136 */
} orc_grid_t;

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
  float*    pld_r;
  uint32_t* pld_i;
  float*    phi;
} orc_prtls_t;

/* pusher flags: src/global/enums.h:317-324 */
enum { ORC_PUSHER_NONE = 0, ORC_PUSHER_PHOTON = 1, ORC_PUSHER_BORIS = 2, ORC_PUSHER_VAY = 4, ORC_PUSHER_GCA = 8 };
/* radiative drag flags: src/global/enums.h (RadiativeDrag) */
enum { ORC_DRAG_NONE = 0, ORC_DRAG_SYNCHROTRON = 1, ORC_DRAG_COMPTON = 2 };
/* particle boundary kinds per face (min/max of each dim) */
enum { ORC_PBC_NONE = 0, ORC_PBC_PERIODIC = 1, ORC_PBC_ABSORB = 2, ORC_PBC_REFLECT = 3, ORC_PBC_AXIS = 4 };
/* field boundary kinds per face, as far as the filter / comm need them */
enum { ORC_FBC_NONE = 0, ORC_FBC_PERIODIC = 1, ORC_FBC_CONDUCTOR = 2, ORC_FBC_AXIS = 3, ORC_FBC_SYNC = 4 };

/* scalar arguments of sr::Pusher_kernel: src/kernels/pushers/context.h:74-122 */
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
  int    pbc[6];       /* ORC_PBC_* for i1min,i1max,i2min,i2max,i3min,i3max */
  int    tag_outgoing; /* 1: write mpi::SendTag (src/global/arch/mpi_tags.h:175-233) */
  float  dx;           /* Minkowski cell size */
  float  xmin[3];      /* Minkowski x*_min (only used by the atmosphere force) */
} orc_pusher_t;

void orc_faraday_mink(const orc_grid_t* g, float* em, float coeff1, float coeff2,
                      const float* stencil9 /* dx,dy,bxy,byx,dz,bxz,bzx,byz,bzy or NULL */);
void orc_ampere_mink(const orc_grid_t* g, float* em, float coeff1, float coeff2);
void orc_currents_ampere_mink(const orc_grid_t* g, float* em, float* cur, float coeff, float ppc0);
/* one filter pass: cur = filter(buff); fbc[6] = ORC_FBC_* per face; cartesian only */
void orc_filter_pass(const orc_grid_t* g, float* cur, const float* buff, const int* fbc);
/* serial-order push of particles [0,npart) */
void orc_push_sr_mink(const orc_grid_t* g, int order, const orc_pusher_t* ctx,
                      const orc_prtls_t* p, uint32_t npart, const float* em);
/* serial-order deposit into cur (accumulates; caller zeroes) */
void orc_deposit_mink(const orc_grid_t* g, int order, const orc_prtls_t* p, uint32_t npart,
                      float charge, float dt, float dx, float* cur);
/* single-domain periodic/none ghost exchange, restating comm_nompi.hpp:29-119 driven by
   metadomain_comm.cpp:122-195 (slices) and :276 (direction order). comp range [c0,c1). */
void orc_comm_fields_self(const orc_grid_t* g, float* fld, int ncomp, int c0, int c1, const int* fbc);
/* additive sync of currents into buff (zeroed inside) followed by cur += buff on active cells:
   metadomain_comm.cpp:409-562 */
void orc_sync_currents_self(const orc_grid_t* g, float* cur, float* buff, const int* fbc);

#ifdef __cplusplus
}
#endif
#endif
