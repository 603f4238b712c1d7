// entity_b200 -- the C ABI (include/entity_b200.h): context, argument checks, dispatch to
// the strict / fast kernel variants. No compute happens on the host and there is no CPU
// fallback: without a usable CUDA device every entry point fails.
#include "launch.h"
#include "metrics.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <limits>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>

namespace eb200 {
  static std::atomic<uint64_t> g_launches { 0 };

  void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

  uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }
} // namespace eb200

namespace eb200 {
  struct EngineState;
  EngineState* engine_state_new();
  void         engine_state_delete(EngineState*);

  // comm.cu
  struct Comm;
  Comm*       comm_new();
  void        comm_delete(Comm*);
  const char* comm_error(const Comm*);
  int         comm_setup(Comm&, const eb200_grid_t&, const eb200_metadomain_t&, const char* id);
  int         comm_unique_id(char* out, std::string& err);
  int         comm_fields(Comm&, float* fld, int c0, int c1, cudaStream_t st);
  int         comm_sync_currents(Comm&, float* cur, cudaStream_t st);
  int         comm_particles(Comm&, eb200_species_t* species, int nspecies, cudaStream_t st,
                             const ExcList* exc);
  const eb200_domain_info_t* comm_info(const Comm*);
} // namespace eb200

struct eb200_ctx {
  eb200_config_t cfg;
  eb200::EngineState* engine = nullptr;
  eb200::Comm*   comm = nullptr; // attached by eb200_comm_init (multi-domain exchange)
  eb200::Scratch scratch;
  uint32_t       sort_cap = 0; // high-water mark of the sort's capacity (see eb200_sort_particles)
  std::string    err;
  uint64_t       launches_at_init;
  int            pd_kernel = 0; // eb200_set_pd_kernel
  int            sort_mode = -1; // eb200_set_sort_mode
  int            lean_prev = 0;  // eb200_set_lean_prev
  // exception lists of the fused push launches of the current step, by the species' tag array
  // (consumed and invalidated by eb200_comm_particles)
  struct ExcTrack {
    const short*   tag = nullptr;
    uint32_t       npart = 0;
    eb200::Scratch buf; // [count: 256 B][idx: cap x 4 B]
    eb200::ExcList list;
  };
  std::vector<ExcTrack*> exc;
  eb200::Scratch packed;        // E/B repacked node by node for the fused 2D zig-zag kernel
  eb200::Scratch packed_j;      // J as 16-byte nodes {jx1, jx2, jx3, -}: target of kernel 8's flushes
  void*          packed_j_zeroed = nullptr; // the allocation that has been cleared
  float*         packed_j_cur    = nullptr; // planes the pending nodes belong to (while held)
  cudaStream_t   packed_j_stream = nullptr;
  eb200::Scratch stats;         // one double: the device-side accumulator of the reductions
  eb200::Scratch emit_count;    // one uint32: photons emitted by one eb200_push_sr_emission launch
  const float*   packed_hold = nullptr; // em the packed copy is guaranteed current for
  bool           no_filter_fusion = false; // EB200_NO_FILTER_FUSION=1: pass-by-pass filter
  eb200::MetricParams metric {}; // curvilinear / GR contexts
};

static std::string g_last_error;
static std::mutex  g_err_mutex;

static int fail(eb200_ctx* ctx, int code, const std::string& msg) {
  {
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_last_error = msg;
  }
  if (ctx) ctx->err = msg;
  return code;
}

static int check_cuda(eb200_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return EB200_OK;
  return fail(ctx, EB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define REQUIRE(ctx, cond, msg)                                                                \
  do {                                                                                         \
    if (!(cond)) return fail((ctx), EB200_ERR_ARG, std::string(__func__) + ": " + (msg));      \
  } while (0)

#define ENTER(ctx)                                                                             \
  do {                                                                                         \
    if ((ctx) == nullptr) return fail(nullptr, EB200_ERR_ARG, std::string(__func__) + ": null context"); \
    cudaError_t e_ = cudaSetDevice((ctx)->cfg.device);                                         \
    if (e_ != cudaSuccess) return check_cuda((ctx), e_, "cudaSetDevice");                      \
  } while (0)

// run the variant the context was created for
#define VARIANT_CALL(ctx, expr)                                                                \
  ((ctx)->cfg.strict_fp ? eb200::strict_fp::expr : eb200::fast_fp::expr)

/* ------------------------------------------------------------- metric evaluation (host) */
namespace {
  template <class M>
  void eval_sr_metric(const eb200::MetricParams& m, int nq, const float* x1, const float* x2,
                      float* out) {
    for (int q = 0; q < nq; ++q) {
      float* o = out + 16 * q;
      o[0]  = M::h11(m, x1[q], x2[q]);
      o[1]  = M::h22(m, x1[q], x2[q]);
      o[2]  = M::h33(m, x1[q], x2[q]);
      o[3]  = M::sqrt_h11(m, x1[q], x2[q]);
      o[4]  = M::sqrt_h22(m, x1[q], x2[q]);
      o[5]  = M::sqrt_h33(m, x1[q], x2[q]);
      o[6]  = M::sqrt_det_h(m, x1[q], x2[q]);
      o[7]  = M::polar_area(m, x1[q]);
      o[8]  = M::r(m, x1[q]);
      o[9]  = M::theta(m, x2[q]);
      o[10] = M::x1_of_r(m, o[8]);
      o[11] = M::x2_of_theta(m, o[9]);
      for (int k = 12; k < 16; ++k) o[k] = 0.0f;
    }
  }

  template <class M>
  void eval_gr_metric(const eb200::MetricParams& m, int nq, const float* x1, const float* x2,
                      float* out) {
    for (int q = 0; q < nq; ++q) {
      float*      o = out + 32 * q;
      const float a = x1[q], b = x2[q];
      o[0]  = M::h_11(m, a, b);
      o[1]  = M::h_22(m, a, b);
      o[2]  = M::h_33(m, a, b);
      o[3]  = M::h_13(m, a, b);
      o[4]  = M::h11(m, a, b);
      o[5]  = M::h22(m, a, b);
      o[6]  = M::h33(m, a, b);
      o[7]  = M::h13(m, a, b);
      o[8]  = M::alpha(m, a, b);
      o[9]  = M::beta1(m, a, b);
      o[10] = M::sqrt_det_h(m, a, b);
      o[11] = M::sqrt_det_h_tilde(m, a, b);
      o[12] = M::polar_area(m, a);
      o[13] = M::dr_alpha(m, a, b);
      o[14] = M::dt_alpha(m, a, b);
      o[15] = M::dr_beta1(m, a, b);
      o[16] = M::dt_beta1(m, a, b);
      o[17] = M::dr_h11(m, a, b);
      o[18] = M::dr_h22(m, a, b);
      o[19] = M::dr_h33(m, a, b);
      o[20] = M::dr_h13(m, a, b);
      o[21] = M::dt_h11(m, a, b);
      o[22] = M::dt_h22(m, a, b);
      o[23] = M::dt_h33(m, a, b);
      o[24] = M::dt_h13(m, a, b);
      o[25] = M::theta(m, b);
      o[26] = M::x2_of_theta(m, o[25]);
      for (int k = 27; k < 32; ++k) o[k] = 0.0f;
    }
  }
} // namespace


extern "C" {

int eb200_version(void) { return EB200_VERSION; }

int eb200_device_count(void) {
  int         n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* eb200_last_error(const eb200_ctx_t* ctx) {
  if (ctx) return ctx->err.c_str();
  return g_last_error.c_str();
}

uint64_t eb200_launch_count(const eb200_ctx_t* ctx) {
  return eb200::launches() - (ctx ? ctx->launches_at_init : 0);
}

int eb200_init(const eb200_config_t* cfg, eb200_ctx_t** out) {
  if (cfg == nullptr || out == nullptr) return fail(nullptr, EB200_ERR_ARG, "eb200_init: null argument");
  *out = nullptr;
  if (eb200_device_count() <= 0) {
    return fail(nullptr, EB200_ERR_NO_DEVICE,
                "eb200_init: no CUDA device available (this library has no CPU fallback)");
  }
  if (cfg->grid.dim < 1 || cfg->grid.dim > 3) return fail(nullptr, EB200_ERR_ARG, "eb200_init: dim must be 1..3");
  if (cfg->shape_order < 0 || cfg->shape_order > 11) {
    return fail(nullptr, EB200_ERR_UNSUPPORTED, "eb200_init: shape_order must be 0..11");
  }
  if (cfg->shape_order > 3 && cfg->metric != EB200_METRIC_MINKOWSKI) {
    return fail(nullptr, EB200_ERR_UNSUPPORTED,
                "eb200_init: shape orders 4..11 are built for Minkowski domains");
  }
  if (cfg->metric < EB200_METRIC_MINKOWSKI || cfg->metric > EB200_METRIC_KERR_SCHILD_0) {
    return fail(nullptr, EB200_ERR_UNSUPPORTED, "eb200_init: unknown metric");
  }
  if (cfg->metric != EB200_METRIC_MINKOWSKI) {
    // the reference's curvilinear metrics are 2D only (static_asserts in src/metrics/*.h)
    if (cfg->grid.dim != 2) {
      return fail(nullptr, EB200_ERR_UNSUPPORTED, "eb200_init: curvilinear metrics are 2D only");
    }
    const float* mp = cfg->metric_params;
    if (!(mp[1] > mp[0]) || !(mp[3] > mp[2])) {
      return fail(nullptr, EB200_ERR_ARG, "eb200_init: metric_params must hold x1min < x1max, x2min < x2max");
    }
    const bool quasi = cfg->metric == EB200_METRIC_QSPHERICAL || cfg->metric == EB200_METRIC_QKERR_SCHILD;
    if (quasi && !(mp[0] - mp[4] > 0.0f)) {
      return fail(nullptr, EB200_ERR_ARG, "eb200_init: x1min must exceed qsph_r0");
    }
  }
  const int need_ng = cfg->shape_order == 0 ? 2 : (cfg->shape_order + 1) / 2 + 1;
  if (cfg->grid.ng < need_ng) {
    return fail(nullptr, EB200_ERR_ARG, "eb200_init: too few ghost cells for this shape order");
  }
  for (int a = 0; a < cfg->grid.dim; ++a) {
    if (cfg->grid.n[a] < cfg->grid.ng) return fail(nullptr, EB200_ERR_ARG, "eb200_init: n[a] < ng");
  }
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return check_cuda(nullptr, e, "cudaSetDevice");
  eb200_ctx* ctx        = new eb200_ctx();
  ctx->cfg              = *cfg;
  ctx->launches_at_init = eb200::launches();
  ctx->engine           = eb200::engine_state_new();
  {
    const char* nf = getenv("EB200_NO_FILTER_FUSION");
    ctx->no_filter_fusion = nf && nf[0] == '1';
  }
  for (int a = cfg->grid.dim; a < 3; ++a) ctx->cfg.grid.n[a] = 1;
  if (cfg->metric != EB200_METRIC_MINKOWSKI) {
    const float* mp = cfg->metric_params;
    ctx->metric = eb200::make_metric(cfg->metric, cfg->grid.n[0], cfg->grid.n[1], mp[0], mp[1],
                                     mp[2], mp[3], mp[4], mp[5], mp[6]);
  }
  *out = ctx;
  return EB200_OK;
}

void eb200_finalize(eb200_ctx_t* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  ctx->scratch.release();
  ctx->packed.release();
  ctx->packed_j.release();
  ctx->stats.release();
  ctx->emit_count.release();
  for (auto* q : ctx->exc) {
    q->buf.release();
    delete q;
  }
  if (ctx->comm) eb200::comm_delete(ctx->comm);
  eb200::engine_state_delete(ctx->engine);
  delete ctx;
}

eb200::EngineState* eb200_ctx_engine_state(eb200_ctx_t* ctx) { return ctx->engine; }

// internal accessor used by engine.cu
int eb200_ctx_grid(const eb200_ctx_t* ctx, eb200_grid_t* grid, float* dx, float* xmin3) {
  if (!ctx || !grid || !dx || !xmin3) return EB200_ERR_ARG;
  *grid = ctx->cfg.grid;
  *dx   = ctx->cfg.metric_params[0];
  for (int a = 0; a < 3; ++a) xmin3[a] = ctx->cfg.metric_params[1 + a];
  return EB200_OK;
}

int eb200_ctx_metric(const eb200_ctx_t* ctx) { return ctx ? ctx->cfg.metric : -1; }

static bool is_sr_curv(const eb200_ctx* ctx) {
  return ctx->cfg.metric == EB200_METRIC_SPHERICAL || ctx->cfg.metric == EB200_METRIC_QSPHERICAL;
}

static bool is_gr(const eb200_ctx* ctx) { return ctx->cfg.metric >= EB200_METRIC_KERR_SCHILD; }

#define REQUIRE_MINK(ctx, what)                                                                \
  REQUIRE(ctx, (ctx)->cfg.metric == EB200_METRIC_MINKOWSKI, what ": Minkowski contexts only")

int eb200_faraday(eb200_ctx_t* ctx, float* em, float coeff1, float coeff2,
                  const float* stencil9_host, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr, "em is null");
  REQUIRE_MINK(ctx, "eb200_faraday");
  return check_cuda(ctx,
                    VARIANT_CALL(ctx, faraday(ctx->cfg.grid, em, coeff1, coeff2, stencil9_host,
                                              (cudaStream_t)stream)),
                    "faraday");
}

int eb200_ampere(eb200_ctx_t* ctx, float* em, float coeff1, float coeff2, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr, "em is null");
  REQUIRE_MINK(ctx, "eb200_ampere");
  return check_cuda(ctx, VARIANT_CALL(ctx, ampere(ctx->cfg.grid, em, coeff1, coeff2, (cudaStream_t)stream)),
                    "ampere");
}

int eb200_currents_ampere(eb200_ctx_t* ctx, float* em, float* cur, float coeff, float ppc0,
                          eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr && cur != nullptr, "null field");
  REQUIRE_MINK(ctx, "eb200_currents_ampere");
  return check_cuda(ctx,
                    VARIANT_CALL(ctx, currents_ampere(ctx->cfg.grid, em, cur, coeff, ppc0,
                                                      (cudaStream_t)stream)),
                    "currents_ampere");
}

int eb200_currents_ampere_ext(eb200_ctx_t* ctx, float* em, float* cur, float coeff, float ppc0,
                              const eb200_ext_current_t* ext, eb200_stream_t stream) {
  ENTER(ctx);
  if (ext == nullptr) return eb200_currents_ampere(ctx, em, cur, coeff, ppc0, stream);
  REQUIRE(ctx, em != nullptr && cur != nullptr, "null field");
  REQUIRE_MINK(ctx, "eb200_currents_ampere_ext");
  REQUIRE(ctx, ext->nmodes >= 0 && ext->nmodes <= EB200_MAX_MODES, "ext_current: bad number of modes");
  const float* mp = ctx->cfg.metric_params; // dx, x1min, x2min, x3min
  return check_cuda(ctx,
                    VARIANT_CALL(ctx, currents_ampere_ext(ctx->cfg.grid, em, cur, coeff, ppc0, *ext,
                                                          mp[0], mp + 1, (cudaStream_t)stream)),
                    "currents_ampere_ext");
}

static size_t field_bytes(const eb200_grid_t& g, int ncomp) {
  size_t n = ncomp;
  for (int a = 0; a < g.dim; ++a) n *= (size_t)(g.n[a] + 2 * g.ng);
  return n * sizeof(float);
}

int eb200_filter(eb200_ctx_t* ctx, float* cur, float* buff, int nfilter, const int* fbc,
                 eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, cur != nullptr && buff != nullptr && fbc != nullptr, "null argument");
  REQUIRE(ctx, nfilter >= 0, "nfilter < 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->cfg.metric != EB200_METRIC_MINKOWSKI) {
    // spherical-type coordinates: axis-aware stencil; a single domain has no periodic
    // neighbour, so the per-pass J exchange of currents.h:117 has nothing to do
    REQUIRE(ctx, ctx->comm == nullptr, "curvilinear filter: multi-domain exchange not built");
    const size_t nb = field_bytes(ctx->cfg.grid, 3);
    for (int pass = 0; pass < nfilter; ++pass) {
      cudaError_t e = cudaMemcpyAsync(buff, cur, nb, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return check_cuda(ctx, e, "filter copy");
      e = eb200::curv::filter_sph_pass(ctx->cfg.grid, cur, buff, fbc, st);
      if (e != cudaSuccess) return check_cuda(ctx, e, "filter pass");
    }
    return EB200_OK;
  }
  for (int a = 0; a < 2 * ctx->cfg.grid.dim; ++a) {
    if (fbc[a] == EB200_FBC_AXIS) {
      return fail(ctx, EB200_ERR_UNSUPPORTED, "eb200_filter: axis boundaries need a spherical metric");
    }
  }
  const size_t bytes = field_bytes(ctx->cfg.grid, 3);
  // single doubly periodic 2D domain: several passes per sweep (fields.cu, temporal blocking)
  // ... or a 2D domain whose non-periodic dimensions carry EB200_FBC_NONE on both faces (MATCH /
  // open boundaries: their ghost cells are static during the filter, as in the reference, which
  // exchanges nothing there): the sweeps read those ghost cells and carry them
  bool fuse = ctx->cfg.grid.dim == 2 && ctx->comm == nullptr && nfilter > 0 && !ctx->no_filter_fusion;
  int  static_dims = 0;
  for (int a = 0; a < 2; ++a) {
    const bool per = fbc[2 * a] == EB200_FBC_PERIODIC && fbc[2 * a + 1] == EB200_FBC_PERIODIC;
    const bool non = fbc[2 * a] == EB200_FBC_NONE && fbc[2 * a + 1] == EB200_FBC_NONE;
    fuse           = fuse && (per || non);
    if (non) static_dims |= 2 << a;
  }
  if (static_dims && nfilter < 2) fuse = false; // an odd single sweep would land in buff
  if (fuse) {
    if (static_dims) {
      // the second array must show the same static ghost cells
      cudaError_t e = cudaMemcpyAsync(buff, cur, bytes, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return check_cuda(ctx, e, "filter copy");
    }
    // an even number of sweeps of <= 4 passes each, so that the result lands in `cur`
    int sweeps = (nfilter + 3) / 4;
    if (sweeps % 2 == 1 && nfilter >= 2) ++sweeps;
    float* a = cur;
    float* b = buff;
    int    left = nfilter;
    for (int s = 0; s < sweeps; ++s) {
      const int   p = (left + (sweeps - s) - 1) / (sweeps - s);
      cudaError_t e = VARIANT_CALL(ctx, filter_fused(ctx->cfg.grid, a, b, p, static_dims, st));
      if (e != cudaSuccess) return check_cuda(ctx, e, "fused filter");
      left -= p;
      float* t = a;
      a        = b;
      b        = t;
    }
    if (a != cur) {
      cudaError_t e = cudaMemcpyAsync(cur, buff, bytes, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return check_cuda(ctx, e, "filter copy");
    }
    cudaError_t e = VARIANT_CALL(ctx, comm_fields_self(ctx->cfg.grid, cur, 0, 3, fbc, st));
    return check_cuda(ctx, e, "filter ghost exchange");
  }
  // two passes per halo exchange where every face is periodic or adjoins another domain: with
  // N_GHOSTS >= 2 exchanged layers the first pass of a pair can also produce the first ghost
  // layer, which is all the second pass needs (halves the exchange rounds of a multi-domain run)
  bool pair_ok = ctx->cfg.grid.ng >= 2 && !ctx->no_filter_fusion;
  for (int a = 0; a < 2 * ctx->cfg.grid.dim; ++a) {
    pair_ok = pair_ok && (fbc[a] == EB200_FBC_PERIODIC || fbc[a] == EB200_FBC_SYNC);
  }
  auto exchange = [&](float* fld) -> int {
    if (ctx->comm) {
      int rc = eb200::comm_fields(*ctx->comm, fld, 0, 3, st);
      return rc == EB200_OK ? rc : fail(ctx, rc, eb200::comm_error(ctx->comm));
    }
    return check_cuda(ctx, VARIANT_CALL(ctx, comm_fields_self(ctx->cfg.grid, fld, 0, 3, fbc, st)),
                      "filter ghost exchange");
  };
  if (pair_ok && ctx->cfg.grid.dim == 2 && nfilter > 0) {
    // 2D: each pair is ONE sweep of the temporally blocked kernel reading the exchanged ghost
    // cells as its halo (two passes in shared memory), ping-ponging between the two arrays;
    // the exchange follows the array that holds the result
    float* a    = cur;
    float* b    = buff;
    int    left = nfilter;
    while (left > 0) {
      const int   p = left >= 2 ? 2 : 1;
      cudaError_t e = VARIANT_CALL(ctx, filter_fused(ctx->cfg.grid, a, b, p, 1, st));
      if (e != cudaSuccess) return check_cuda(ctx, e, "fused filter (ghost halo)");
      int rc = exchange(b);
      if (rc != EB200_OK) return rc;
      left -= p;
      float* t = a;
      a        = b;
      b        = t;
    }
    if (a != cur) {
      cudaError_t e = cudaMemcpyAsync(cur, a, bytes, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return check_cuda(ctx, e, "filter copy");
    }
    return EB200_OK;
  }
  for (int pass = 0; pass < nfilter; ++pass) {
    cudaError_t e;
    if (pair_ok && pass + 1 < nfilter) {
      // a pair ping-pongs between the two arrays: cur -> buff over the active cells + one ghost
      // layer, then buff -> cur over the active cells; no copy (each filter pass writes every
      // cell the next one reads)
      e = VARIANT_CALL(ctx, filter_pass(ctx->cfg.grid, buff, cur, fbc, 1, st));
      if (e != cudaSuccess) return check_cuda(ctx, e, "filter pass (extended)");
      ++pass;
    } else {
      // buff <- cur (currents.h:108), filter into cur (:109-116), ghost exchange (:117)
      e = cudaMemcpyAsync(buff, cur, bytes, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return check_cuda(ctx, e, "filter copy");
    }
    e = VARIANT_CALL(ctx, filter_pass(ctx->cfg.grid, cur, buff, fbc, 0, st));
    if (e != cudaSuccess) return check_cuda(ctx, e, "filter pass");
    int rc = exchange(cur);
    if (rc != EB200_OK) return rc;
  }
  return EB200_OK;
}

static int check_prtls(eb200_ctx* ctx, const eb200_prtls_t* p, uint32_t npart) {
  if (p == nullptr) return fail(ctx, EB200_ERR_ARG, "particle struct is null");
  if (npart == 0) return EB200_OK;
  const int d = ctx->cfg.grid.dim;
  bool      ok = p->i1 && p->dx1 && p->i1_prev && p->dx1_prev && p->ux1 && p->ux2 && p->ux3 &&
            p->weight && p->tag;
  if (d > 1) ok = ok && p->i2 && p->dx2 && p->i2_prev && p->dx2_prev;
  if (d > 2) ok = ok && p->i3 && p->dx3 && p->i3_prev && p->dx3_prev;
  if (!ok) return fail(ctx, EB200_ERR_ARG, "a required particle array is null");
  if (ctx->cfg.maxnpart && npart > ctx->cfg.maxnpart) {
    return fail(ctx, EB200_ERR_CAPACITY, "npart exceeds maxnpart");
  }
  return EB200_OK;
}

static int check_pusher(eb200_ctx* ctx, const eb200_pusher_t* c) {
  if (c == nullptr) return fail(ctx, EB200_ERR_ARG, "pusher context is null");
  // kernel::sr::Pusher_kernel ctor: "No particle pusher specified" (sr.hpp:112-114)
  if (c->pusher_flags == EB200_PUSHER_NONE) return fail(ctx, EB200_ERR_ARG, "No particle pusher specified");
  if (c->pusher_flags != EB200_PUSHER_PHOTON &&
      !(c->pusher_flags & (EB200_PUSHER_BORIS | EB200_PUSHER_VAY))) {
    return fail(ctx, EB200_ERR_ARG, "Invalid pusher algorithm");
  }
  if (ctx->cfg.metric == EB200_METRIC_MINKOWSKI && !(c->dx > 0.0f)) {
    return fail(ctx, EB200_ERR_ARG, "pusher.dx must be positive");
  }
  return EB200_OK;
}

int eb200_push_sr(eb200_ctx_t* ctx, const eb200_pusher_t* pusher, const eb200_prtls_t* prtls,
                  uint32_t npart, const float* em, eb200_stream_t stream) {
  ENTER(ctx);
  int rc = check_pusher(ctx, pusher);
  if (rc) return rc;
  rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  REQUIRE(ctx, em != nullptr, "em is null");
  REQUIRE(ctx, !is_gr(ctx), "eb200_push_sr on a GR context: use eb200_push_gr");
  if (is_sr_curv(ctx)) {
    REQUIRE(ctx, npart == 0 || prtls->phi != nullptr, "curvilinear 2D pusher needs prtls->phi");
    REQUIRE(ctx, !pusher->has_atmosphere || (pusher->atm_gx2 == 0.0f && pusher->atm_gx3 == 0.0f),
            "Invalid force for coordinate system"); // sr.hpp:1466, 1480
    return check_cuda(ctx,
                      eb200::curv::push_sr(ctx->metric, ctx->cfg.grid, ctx->cfg.shape_order,
                                           *pusher, *prtls, npart, em, (cudaStream_t)stream),
                      "push_sr (curvilinear)");
  }
  return check_cuda(ctx,
                    VARIANT_CALL(ctx, push_sr(ctx->cfg.grid, ctx->cfg.shape_order, *pusher,
                                              *prtls, npart, em, (cudaStream_t)stream)),
                    "push_sr");
}

int eb200_push_sr_emission(eb200_ctx_t* ctx, const eb200_pusher_t* pusher, const eb200_prtls_t* prtls,
                           uint32_t npart, const float* em, eb200_emission_t* em_policy,
                           eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em_policy != nullptr, "emission policy is null");
  REQUIRE_MINK(ctx, "eb200_push_sr_emission");
  REQUIRE(ctx, ctx->cfg.shape_order <= 3, "eb200_push_sr_emission: shape orders 0..3");
  REQUIRE(ctx, em_policy->kind == EB200_EMISSION_SYNCHROTRON || em_policy->kind == EB200_EMISSION_COMPTON,
          "unknown emission policy");
  int rc = check_pusher(ctx, pusher);
  if (rc) return rc;
  REQUIRE(ctx, pusher->pusher_flags != EB200_PUSHER_PHOTON && pusher->mass > 0.0f,
          "emission: massive emitters only");
  rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  REQUIRE(ctx, em != nullptr, "em is null");
  REQUIRE(ctx, em_policy->photon_npart <= em_policy->photon_maxnpart, "photon npart > maxnpart");
  {
    // the photon arrays are written, not read: check them as a full particle struct
    const eb200_prtls_t& q = em_policy->photons;
    const int            d = ctx->cfg.grid.dim;
    bool ok = q.i1 && q.dx1 && q.i1_prev && q.dx1_prev && q.ux1 && q.ux2 && q.ux3 && q.weight && q.tag;
    if (d > 1) ok = ok && q.i2 && q.dx2 && q.i2_prev && q.dx2_prev;
    if (d > 2) ok = ok && q.i3 && q.dx3 && q.i3_prev && q.dx3_prev;
    REQUIRE(ctx, ok || em_policy->photon_maxnpart == em_policy->photon_npart,
            "a required array of the emitted species is null");
  }
  rc = check_cuda(ctx, ctx->emit_count.reserve(256), "emission counter");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = check_cuda(ctx, cudaMemsetAsync(ctx->emit_count.ptr, 0, 4, st), "emission counter");
  if (rc) return rc;
  eb200::EmitArgs M {};
  M.E.kind                  = em_policy->kind;
  M.E.photon_weight         = em_policy->photon_weight;
  M.E.energy_min            = em_policy->photon_energy_min;
  M.E.nominal_probability   = em_policy->nominal_probability;
  M.E.nominal_photon_energy = em_policy->nominal_photon_energy;
  M.E.species_mass          = pusher->mass;
  M.E.should_drag           = em_policy->should_drag;
  M.ph      = em_policy->photons;
  M.offset  = em_policy->photon_npart;
  M.cap     = em_policy->photon_maxnpart;
  M.counter = (uint32_t*)ctx->emit_count.ptr;
  M.seed = em_policy->seed, M.step = em_policy->step, M.call = em_policy->call;
  rc = check_cuda(ctx,
                  VARIANT_CALL(ctx, push_sr_emission(ctx->cfg.grid, ctx->cfg.shape_order, *pusher, *prtls,
                                                     npart, em, M, st)),
                  "push_sr_emission");
  if (rc) return rc;
  uint32_t n = 0;
  rc = check_cuda(ctx, cudaMemcpyAsync(&n, ctx->emit_count.ptr, 4, cudaMemcpyDeviceToHost, st), "emission count");
  if (rc) return rc;
  rc = check_cuda(ctx, cudaStreamSynchronize(st), "emission count");
  if (rc) return rc;
  const uint32_t room = em_policy->photon_maxnpart - em_policy->photon_npart;
  em_policy->photon_npart += n < room ? n : room;
  if (n > room) return fail(ctx, EB200_ERR_CAPACITY, "emission: photons do not fit into maxnpart of the emitted species");
  return EB200_OK;
}

int eb200_deposit(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart, float charge,
                  float dt, float* cur, int mode, eb200_stream_t stream) {
  ENTER(ctx);
  int rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  REQUIRE(ctx, cur != nullptr, "cur is null");
  REQUIRE(ctx, dt > 0.0f, "dt must be positive");
  REQUIRE(ctx, mode == EB200_DEPOSIT_ATOMIC || mode == EB200_DEPOSIT_ORDERED ||
                 mode == EB200_DEPOSIT_AGGREGATED, "bad deposit mode");
  if (ctx->cfg.metric != EB200_METRIC_MINKOWSKI) {
    REQUIRE(ctx, mode != EB200_DEPOSIT_ORDERED,
            "curvilinear deposit: ATOMIC and AGGREGATED modes are built");
    REQUIRE(ctx, npart == 0 || is_gr(ctx) || prtls->phi != nullptr,
            "curvilinear 2D deposit needs prtls->phi");
    cudaError_t e = is_gr(ctx)
                      ? eb200::curv::deposit_gr(ctx->metric, ctx->cfg.grid, ctx->cfg.shape_order,
                                                *prtls, npart, charge, dt, cur, mode,
                                                (cudaStream_t)stream)
                      : eb200::curv::deposit_sr(ctx->metric, ctx->cfg.grid, ctx->cfg.shape_order,
                                                *prtls, npart, charge, dt, cur, mode,
                                                (cudaStream_t)stream);
    return check_cuda(ctx, e, "deposit (curvilinear)");
  }
  const float dx = ctx->cfg.metric_params[0];
  REQUIRE(ctx, dx > 0.0f, "metric_params[0] (dx) must be positive");
  return check_cuda(ctx,
                    VARIANT_CALL(ctx, deposit(ctx->cfg.grid, ctx->cfg.shape_order, *prtls, npart,
                                              charge, dt, dx, cur, mode, ctx->scratch,
                                              (cudaStream_t)stream)),
                    "deposit");
}

// the fused 2D zig-zag kernel that gathers from packed nodes (kernel 5)
static bool uses_packed_nodes(const eb200_ctx* ctx, int mode) {
  return (ctx->pd_kernel == 5 || ctx->pd_kernel == 6 || ctx->pd_kernel == 7 || ctx->pd_kernel == 8 || ctx->pd_kernel == 0) && ctx->cfg.metric == EB200_METRIC_MINKOWSKI && ctx->cfg.grid.dim == 2 &&
         ctx->cfg.shape_order == 0 && mode == EB200_DEPOSIT_AGGREGATED;
}

static size_t packed_bytes(const eb200_ctx* ctx) {
  const eb200_grid_t& g = ctx->cfg.grid;
  return (size_t)(g.n[0] + 2 * g.ng) * (size_t)(g.n[1] + 2 * g.ng) * 24;
}

int eb200_push_deposit_sr(eb200_ctx_t* ctx, const eb200_pusher_t* pusher,
                          const eb200_prtls_t* prtls, uint32_t npart, const float* em, float* cur,
                          int mode, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, mode == EB200_DEPOSIT_ATOMIC || mode == EB200_DEPOSIT_AGGREGATED,
          "fused push+deposit supports the ATOMIC and AGGREGATED modes");
  REQUIRE_MINK(ctx, "eb200_push_deposit_sr");
  int rc = check_pusher(ctx, pusher);
  if (rc) return rc;
  rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  REQUIRE(ctx, em != nullptr && cur != nullptr, "null field");
  if (ctx->cfg.shape_order > 3) {
    // shape orders 4..11 have no fused kernel: the same two kernels the unfused step runs
    rc = eb200_push_sr(ctx, pusher, prtls, npart, em, stream);
    if (rc) return rc;
    return eb200_deposit(ctx, prtls, npart, pusher->charge, pusher->dt, cur, mode, stream);
  }
  float* packed  = nullptr;
  bool   do_pack = false;
  if (uses_packed_nodes(ctx, mode)) {
    // the packed copy is rebuilt on every call unless the engine holds it for this em
    // (eb200_pack_fields_hold: nothing touches em between the species of one step)
    if (check_cuda(ctx, ctx->packed.reserve(packed_bytes(ctx)), "packed E/B") == EB200_OK) {
      packed  = (float*)ctx->packed.ptr;
      do_pack = ctx->packed_hold != em;
    }
  }
  // kernel 8 flushes into 16-byte J nodes; they are added to the planes right away, or when the
  // engine releases its hold (one unpack per step instead of one per species)
  float* pj      = nullptr;
  bool   pj_used = false;
  if (packed != nullptr && !ctx->cfg.strict_fp) {
    const size_t nb = packed_bytes(ctx) / 24 * 16;
    if (check_cuda(ctx, ctx->packed_j.reserve(nb), "packed J") == EB200_OK) {
      if (ctx->packed_j_zeroed != ctx->packed_j.ptr) {
        rc = check_cuda(ctx, cudaMemsetAsync(ctx->packed_j.ptr, 0, nb, (cudaStream_t)stream), "packed J");
        if (rc) return rc;
        ctx->packed_j_zeroed = ctx->packed_j.ptr;
      }
      pj = (float*)ctx->packed_j.ptr;
      if (ctx->packed_j_cur != nullptr && ctx->packed_j_cur != cur) {
        // pending nodes of another array: settle them first
        rc = check_cuda(ctx, VARIANT_CALL(ctx, unpack_j4(ctx->cfg.grid, pj, ctx->packed_j_cur,
                                                         ctx->packed_j_stream)), "unpack_j4");
        if (rc) return rc;
        ctx->packed_j_cur = nullptr;
      }
    }
  }
  // multi-domain: let the kernel write down who is not alive afterwards (the migration reads
  // that list instead of scanning every tag twice)
  eb200::ExcList* exc = nullptr;
  static const bool exc_off = getenv("EB200_MIGRATE_SCAN") != nullptr;
  if (ctx->comm && !exc_off && npart > 0) {
    eb200_ctx::ExcTrack* t = nullptr;
    for (auto* q : ctx->exc) {
      if (q->tag == prtls->tag) t = q;
    }
    if (!t) {
      t      = new eb200_ctx::ExcTrack();
      t->tag = prtls->tag;
      ctx->exc.push_back(t);
    }
    const uint32_t cap = 1u << 20;
    if (check_cuda(ctx, t->buf.reserve(256 + (size_t)cap * 4), "exception list") == EB200_OK &&
        cudaMemsetAsync(t->buf.ptr, 0, 4, (cudaStream_t)stream) == cudaSuccess &&
        cudaMemsetAsync((char*)t->buf.ptr + 256, 0xFF, (size_t)cap * 4, (cudaStream_t)stream) == cudaSuccess) {
      t->npart        = npart;
      t->list.count   = (uint32_t*)t->buf.ptr;
      t->list.idx     = (uint32_t*)((char*)t->buf.ptr + 256);
      t->list.cap     = cap;
      t->list.tracked = false;
      exc             = &t->list;
    }
  }
  rc = check_cuda(ctx,
                  VARIANT_CALL(ctx, push_deposit_sr(ctx->cfg.grid, ctx->cfg.shape_order, *pusher,
                                                    *prtls, npart, em, cur,
                                                    mode | (ctx->pd_kernel << 8) | (ctx->lean_prev << 16),
                                                    packed, do_pack,
                                                    (cudaStream_t)stream, pj, &pj_used, exc)),
                  "push_deposit_sr");
  if (rc) return rc;
  if (pj_used) {
    if (ctx->packed_hold == em) {
      ctx->packed_j_cur    = cur;
      ctx->packed_j_stream = (cudaStream_t)stream;
    } else {
      rc = check_cuda(ctx, VARIANT_CALL(ctx, unpack_j4(ctx->cfg.grid, pj, cur, (cudaStream_t)stream)),
                      "unpack_j4");
    }
  }
  return rc;
}

int eb200_pack_fields_hold(eb200_ctx_t* ctx, const float* em, eb200_stream_t stream) {
  ENTER(ctx);
  ctx->packed_hold = nullptr;
  if (em == nullptr || !uses_packed_nodes(ctx, EB200_DEPOSIT_AGGREGATED)) return EB200_OK;
  int rc = check_cuda(ctx, ctx->packed.reserve(packed_bytes(ctx)), "packed E/B");
  if (rc) return rc;
  rc = check_cuda(ctx, VARIANT_CALL(ctx, pack_em2d(ctx->cfg.grid, em, (float*)ctx->packed.ptr,
                                                   (cudaStream_t)stream)), "pack_em2d");
  if (rc) return rc;
  ctx->packed_hold = em;
  return EB200_OK;
}

int eb200_match_fields(eb200_ctx_t* ctx, float* em, const float* target, int o, float xg_edge,
                       float ds, int tags, int components_mask, const int* range_min,
                       const int* range_max, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE_MINK(ctx, "eb200_match_fields");
  REQUIRE(ctx, em != nullptr && target != nullptr && range_min != nullptr && range_max != nullptr,
          "null argument");
  const eb200_grid_t& g = ctx->cfg.grid;
  REQUIRE(ctx, o >= 0 && o < g.dim, "matching direction outside the simulated dimensions");
  REQUIRE(ctx, ds > 0.0f, "match: ds must be positive");
  REQUIRE(ctx, (tags & ~(EB200_BC_E | EB200_BC_B)) == 0, "match: tags = EB200_BC_E | EB200_BC_B");
  for (int a = 0; a < g.dim; ++a) {
    REQUIRE(ctx, range_min[a] >= 0 && range_max[a] <= g.n[a] + 2 * g.ng,
            "match: range outside the array");
  }
  const float dx = ctx->cfg.metric_params[0];
  REQUIRE(ctx, dx > 0.0f, "metric_params[0] (dx) must be positive");
  return check_cuda(ctx,
                    eb200::match_fields(g, em, target, o, dx, ctx->cfg.metric_params[1 + o], xg_edge,
                                        ds, tags, components_mask & 63, range_min, range_max,
                                        (cudaStream_t)stream),
                    "match_fields");
}

// Mesh::Intersects + Mesh::ExtentToRange for the one box srpic::MatchFieldsIn builds
int eb200_match_layer(const eb200_grid_t* g, float dx, const float* xmin, const float* xmax,
                      float gx_lo, float gx_hi, int o, int sign, float ds,
                      eb200_match_face_t* face) {
  if (!g || !xmin || !xmax || !face || g->dim < 1 || g->dim > 3 || o < 0 || o >= g->dim ||
      sign == 0 || !(dx > 0.0f) || !(ds > 0.0f)) {
    return -EB200_ERR_ARG;
  }
  float bmin, bmax, edge;
  if (sign > 0) { // fields_bcs.h:78-88
    bmax = gx_hi;
    bmin = bmax - ds;
    edge = bmax;
  } else {
    bmin = gx_lo;
    bmax = bmin + ds;
    edge = bmin;
  }
  const float lo = xmin[o], hi = xmax[o];
  // Mesh::Intersection, mesh.h:84-98 (both bounds finite here)
  const float x_min = std::min(hi, std::max(lo, bmin));
  const float x_max = std::max(lo, std::min(hi, bmax));
  const float eps   = std::numeric_limits<float>::epsilon();
  if (x_min > x_max || x_min == x_max ||
      std::fabs(x_min - x_max) <= std::min(std::fabs(x_min), std::fabs(x_max)) * eps) {
    return 0; // mesh.h:112-120
  }
  const int  G = g->ng;
  const bool incl_lo = sign < 0, incl_hi = sign > 0; // fields_bcs.h:93-98
  face->o       = o;
  face->xg_edge = edge;
  face->ds      = ds;
  for (int d = 0; d < 3; ++d) {
    face->range_min[d] = 0;
    face->range_max[d] = (d < g->dim) ? g->n[d] + 2 * G : 1; // Range::All with ghosts, mesh.h:149-153
  }
  // mesh.h:155-197; metric.convert<Ph, Cd>(x) = (x - x_min) * dx_inv (minkowski.h:175-181) with
  // the domain metric's own dx = (x1_max - x1_min) / nx1, dx_inv = 1 / dx (minkowski.h:54-55)
  const float dx_m   = (xmax[0] - xmin[0]) / (float)g->n[0];
  const float dx_inv = 1.0f / dx_m;
  (void)dx;
  float       c_min  = std::floor((x_min - lo) * dx_inv);
  float       c_max  = std::ceil((x_max - lo) * dx_inv);
  if (!incl_lo) c_min = std::max(c_min, 0.0f);
  if (!incl_hi) c_max = std::min(c_max, (float)g->n[o]);
  face->range_min[o] = (int)c_min + (incl_lo ? 0 : G);
  face->range_max[o] = (int)c_max + (incl_hi ? 2 * G : G);
  return 1;
}

static int stats_finish(eb200_ctx_t* ctx, cudaError_t e, double* out_host, cudaStream_t st,
                        const char* what) {
  if (e != cudaSuccess) return check_cuda(ctx, e, what);
  e = cudaMemcpyAsync(out_host, ctx->stats.ptr, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return check_cuda(ctx, e, what);
  return check_cuda(ctx, cudaStreamSynchronize(st), what);
}

int eb200_stats_fields(eb200_ctx_t* ctx, const float* em, const float* cur, int what, int comp,
                       double* out_host, eb200_stream_t stream) {
  ENTER(ctx);
  const bool sph = ctx->cfg.metric == EB200_METRIC_SPHERICAL || ctx->cfg.metric == EB200_METRIC_QSPHERICAL;
  REQUIRE(ctx, sph || ctx->cfg.metric == EB200_METRIC_MINKOWSKI,
          "eb200_stats_fields: Minkowski and (q)spherical SRPIC meshes");
  REQUIRE(ctx, em != nullptr && out_host != nullptr, "null argument");
  REQUIRE(ctx, what >= EB200_STATS_B2 && what <= EB200_STATS_JDOTE, "unknown field statistic");
  REQUIRE(ctx, what == EB200_STATS_JDOTE || (comp >= 1 && comp <= 3), "component must be 1..3");
  REQUIRE(ctx, what != EB200_STATS_JDOTE || cur != nullptr, "J.E needs cur");
  if (sph) {
    int rc = check_cuda(ctx, ctx->stats.reserve(256), "stats scratch");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return stats_finish(ctx, eb200::stats_fields_curv(ctx->metric, ctx->cfg.grid, em, cur, what, comp,
                                                      (double*)ctx->stats.ptr, st),
                        out_host, st, "stats_fields");
  }
  const float dx = ctx->cfg.metric_params[0];
  REQUIRE(ctx, dx > 0.0f, "metric_params[0] (dx) must be positive");
  int rc = check_cuda(ctx, ctx->stats.reserve(256), "stats scratch");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return stats_finish(ctx, eb200::stats_fields(ctx->cfg.grid, em, cur, dx, what, comp,
                                               (double*)ctx->stats.ptr, st),
                      out_host, st, "stats_fields");
}

int eb200_stats_particles(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart, float mass,
                          float charge, int use_weights, int what, int c1, int c2,
                          double* out_host, eb200_stream_t stream) {
  ENTER(ctx);
  const bool sph = ctx->cfg.metric == EB200_METRIC_SPHERICAL || ctx->cfg.metric == EB200_METRIC_QSPHERICAL;
  REQUIRE(ctx, sph || ctx->cfg.metric == EB200_METRIC_MINKOWSKI,
          "eb200_stats_particles: Minkowski and (q)spherical SRPIC meshes");
  REQUIRE(ctx, out_host != nullptr, "null argument");
  REQUIRE(ctx, what >= EB200_STATS_NPART && what <= EB200_STATS_T, "unknown particle statistic");
  REQUIRE(ctx, what != EB200_STATS_T || (c1 >= 0 && c1 <= 3 && c2 >= 0 && c2 <= 3),
          "stress-energy components must be 0..3");
  // reduced_stats.hpp:428-431
  REQUIRE(ctx, !((what == EB200_STATS_RHO || what == EB200_STATS_CHARGE) && mass == 0.0f),
          "Rho & Charge for massless particles not defined");
  int rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  if (sph) {
    REQUIRE(ctx, npart == 0 || what != EB200_STATS_T || prtls->phi != nullptr, "curvilinear T needs prtls->phi");
    REQUIRE(ctx, npart == 0 || prtls->phi != nullptr, "curvilinear particle statistics need prtls->phi");
    rc = check_cuda(ctx, ctx->stats.reserve(256), "stats scratch");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return stats_finish(ctx, eb200::stats_particles_curv(ctx->metric, *prtls, npart, mass, charge, use_weights,
                                                         what, c1, c2, (double*)ctx->stats.ptr, st),
                        out_host, st, "stats_particles");
  }
  const float dx = ctx->cfg.metric_params[0];
  REQUIRE(ctx, dx > 0.0f, "metric_params[0] (dx) must be positive");
  rc = check_cuda(ctx, ctx->stats.reserve(256), "stats scratch");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return stats_finish(ctx, eb200::stats_particles(ctx->cfg.grid, *prtls, npart, mass, charge,
                                                  use_weights, dx, what, c1, c2,
                                                  (double*)ctx->stats.ptr, st),
                      out_host, st, "stats_particles");
}

int eb200_pack_fields_release(eb200_ctx_t* ctx) {
  if (!ctx) return EB200_OK;
  ctx->packed_hold = nullptr;
  if (ctx->packed_j_cur != nullptr) {
    float* cur        = ctx->packed_j_cur;
    ctx->packed_j_cur = nullptr;
    return check_cuda(ctx, VARIANT_CALL(ctx, unpack_j4(ctx->cfg.grid, (float*)ctx->packed_j.ptr, cur,
                                                       ctx->packed_j_stream)), "unpack_j4");
  }
  return EB200_OK;
}

/* ------------------------------------------------------ curvilinear SR field solvers */
int eb200_faraday_sr(eb200_ctx_t* ctx, float* em, float coeff, const int* fbc,
                     eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr && fbc != nullptr, "null argument");
  REQUIRE(ctx, is_sr_curv(ctx), "eb200_faraday_sr: spherical / qspherical contexts only");
  return check_cuda(ctx,
                    eb200::curv::faraday_sr(ctx->metric, ctx->cfg.grid, em, coeff, fbc,
                                            (cudaStream_t)stream),
                    "faraday_sr");
}

int eb200_ampere_sr(eb200_ctx_t* ctx, float* em, float coeff, const int* fbc,
                    eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr && fbc != nullptr, "null argument");
  REQUIRE(ctx, is_sr_curv(ctx), "eb200_ampere_sr: spherical / qspherical contexts only");
  return check_cuda(ctx,
                    eb200::curv::ampere_sr(ctx->metric, ctx->cfg.grid, em, coeff, fbc,
                                           (cudaStream_t)stream),
                    "ampere_sr");
}

int eb200_currents_ampere_sr(eb200_ctx_t* ctx, float* em, float* cur, float coeff, float inv_n0,
                             const int* fbc, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr && cur != nullptr && fbc != nullptr, "null argument");
  REQUIRE(ctx, is_sr_curv(ctx), "eb200_currents_ampere_sr: spherical / qspherical contexts only");
  return check_cuda(ctx,
                    eb200::curv::currents_ampere_sr(ctx->metric, ctx->cfg.grid, em, cur, coeff,
                                                    inv_n0, fbc, (cudaStream_t)stream),
                    "currents_ampere_sr");
}

/* --------------------------------------------------------------------------- GRPIC */
int eb200_push_gr(eb200_ctx_t* ctx, const eb200_pusher_gr_t* pusher, const eb200_prtls_t* prtls,
                  uint32_t npart, const float* em, const float* em0, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, is_gr(ctx), "eb200_push_gr: Kerr-Schild type contexts only");
  REQUIRE(ctx, pusher != nullptr, "pusher context is null");
  // grpic::ParticlePush: PHOTON or BORIS, anything else raises "not implemented"
  REQUIRE(ctx, pusher->pusher_flags == EB200_PUSHER_PHOTON || pusher->pusher_flags == EB200_PUSHER_BORIS,
          "not implemented");
  REQUIRE(ctx, pusher->niter >= 0, "niter < 0");
  int rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  REQUIRE(ctx, em != nullptr && em0 != nullptr, "null field");
  return check_cuda(ctx,
                    eb200::curv::push_gr(ctx->metric, ctx->cfg.grid, ctx->cfg.shape_order, *pusher,
                                         *prtls, npart, em, em0, (cudaStream_t)stream),
                    "push_gr");
}

static int gr_aux(eb200_ctx_t* ctx, int which_h, const float* d, const float* b, float* out,
                  const int* fbc, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, is_gr(ctx), "GR aux fields: Kerr-Schild type contexts only");
  REQUIRE(ctx, d != nullptr && b != nullptr && out != nullptr && fbc != nullptr, "null argument");
  return check_cuda(ctx,
                    eb200::curv::aux_gr(ctx->metric, ctx->cfg.grid, which_h, d, b, out, fbc,
                                        (cudaStream_t)stream),
                    which_h ? "gr_aux_h" : "gr_aux_e");
}

int eb200_gr_aux_e(eb200_ctx_t* ctx, const float* d, const float* b, float* e_out,
                   const int* fbc, eb200_stream_t stream) {
  return gr_aux(ctx, 0, d, b, e_out, fbc, stream);
}

int eb200_gr_aux_h(eb200_ctx_t* ctx, const float* d, const float* b, float* h_out,
                   const int* fbc, eb200_stream_t stream) {
  return gr_aux(ctx, 1, d, b, h_out, fbc, stream);
}

int eb200_faraday_gr(eb200_ctx_t* ctx, const float* b_in, float* b_out, const float* e_aux,
                     float coeff, const int* fbc, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, is_gr(ctx), "eb200_faraday_gr: Kerr-Schild type contexts only");
  REQUIRE(ctx, b_in && b_out && e_aux && fbc, "null argument");
  return check_cuda(ctx,
                    eb200::curv::faraday_gr(ctx->metric, ctx->cfg.grid, b_in, b_out, e_aux, coeff,
                                            fbc, (cudaStream_t)stream),
                    "faraday_gr");
}

int eb200_ampere_gr(eb200_ctx_t* ctx, const float* d_in, float* d_out, const float* h_aux,
                    float coeff, const int* fbc, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, is_gr(ctx), "eb200_ampere_gr: Kerr-Schild type contexts only");
  REQUIRE(ctx, d_in && d_out && h_aux && fbc, "null argument");
  return check_cuda(ctx,
                    eb200::curv::ampere_gr(ctx->metric, ctx->cfg.grid, d_in, d_out, h_aux, coeff,
                                           fbc, (cudaStream_t)stream),
                    "ampere_gr");
}

int eb200_currents_ampere_gr(eb200_ctx_t* ctx, float* d_fld, const float* cur, float coeff,
                             const int* fbc, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, is_gr(ctx), "eb200_currents_ampere_gr: Kerr-Schild type contexts only");
  REQUIRE(ctx, d_fld && cur && fbc, "null argument");
  return check_cuda(ctx,
                    eb200::curv::currents_ampere_gr(ctx->metric, ctx->cfg.grid, d_fld, cur, coeff,
                                                    fbc, (cudaStream_t)stream),
                    "currents_ampere_gr");
}

/* ------------------------------------------- curvilinear / GR field boundaries */
static int check_range2(eb200_ctx* ctx, const int* rmin, const int* rmax) {
  REQUIRE(ctx, rmin != nullptr && rmax != nullptr, "null range");
  const eb200_grid_t& g = ctx->cfg.grid;
  for (int a = 0; a < 2; ++a) {
    REQUIRE(ctx, rmin[a] >= 0 && rmax[a] <= g.n[a] + 2 * g.ng, "range outside the array");
  }
  return EB200_OK;
}

int eb200_axis_fields(eb200_ctx_t* ctx, float* fld, int sign, int tags, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, fld != nullptr, "null field");
  REQUIRE(ctx, ctx->cfg.grid.dim == 2 && (is_sr_curv(ctx) || is_gr(ctx)),
          "Invalid coordinate type for axis BCs");
  REQUIRE(ctx, sign != 0, "axis: sign must be -1 or +1");
  return check_cuda(ctx, eb200::curv::axis_fields(ctx->cfg.grid, fld, sign > 0, tags, (cudaStream_t)stream),
                    "axis_fields");
}

int eb200_horizon_fields(eb200_ctx_t* ctx, float* fld, int tags, int nfilter,
                         eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, fld != nullptr, "null field");
  REQUIRE(ctx, is_gr(ctx), "HORIZON BCs only applicable for GR");
  REQUIRE(ctx, nfilter >= 0 && 3 + nfilter <= ctx->cfg.grid.ng + ctx->cfg.grid.n[0],
          "horizon: nfilter out of range");
  return check_cuda(ctx, eb200::curv::horizon_fields(ctx->cfg.grid, fld, tags, nfilter, (cudaStream_t)stream),
                    "horizon_fields");
}

int eb200_match_fields_curv(eb200_ctx_t* ctx, float* fld, const float* target, int o,
                            float xg_edge, float ds, int tags, int mask, const int* rmin,
                            const int* rmax, const int* fbc, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, fld != nullptr && target != nullptr && fbc != nullptr, "null argument");
  REQUIRE(ctx, ctx->cfg.grid.dim == 2 && (is_sr_curv(ctx) || is_gr(ctx)),
          "eb200_match_fields_curv: 2D spherical / Kerr-Schild type contexts only");
  REQUIRE(ctx, o == 0 || o == 1, "matching direction outside the simulated dimensions");
  REQUIRE(ctx, ds > 0.0f, "match: ds must be positive");
  int rc = check_range2(ctx, rmin, rmax);
  if (rc) return rc;
  return check_cuda(ctx,
                    eb200::curv::match_fields_curv(ctx->metric, ctx->cfg.grid, fld, target, o, xg_edge,
                                                   ds, tags, mask & 63, rmin, rmax, fbc,
                                                   (cudaStream_t)stream),
                    "match_fields_curv");
}

int eb200_enforce_fields(eb200_ctx_t* ctx, float* em, const float* target, int o, int sign,
                         int i_edge, int tags, int mask, const int* rmin, const int* rmax,
                         eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr && target != nullptr, "null argument");
  REQUIRE(ctx, ctx->cfg.grid.dim == 2, "eb200_enforce_fields: 2D domains");
  REQUIRE(ctx, (o == 0 || o == 1) && sign != 0, "Invalid Orientation");
  int rc = check_range2(ctx, rmin, rmax);
  if (rc) return rc;
  return check_cuda(ctx,
                    eb200::curv::enforce_fields(ctx->cfg.grid, em, target, o, sign > 0, i_edge, tags,
                                                mask & 63, rmin, rmax, (cudaStream_t)stream),
                    "enforce_fields");
}

int eb200_absorb_currents_gr(eb200_ctx_t* ctx, float* cur, float xg_edge, float ds, const int* rmin,
                             const int* rmax, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, cur != nullptr, "null field");
  REQUIRE(ctx, is_gr(ctx) && ctx->cfg.grid.dim == 2, "eb200_absorb_currents_gr: 2D Kerr-Schild type contexts only");
  REQUIRE(ctx, ds > 0.0f, "absorb: ds must be positive");
  int rc = check_range2(ctx, rmin, rmax);
  if (rc) return rc;
  return check_cuda(ctx,
                    eb200::curv::absorb_currents(ctx->metric, ctx->cfg.grid, cur, xg_edge, ds, rmin, rmax,
                                                 (cudaStream_t)stream),
                    "absorb_currents");
}

int eb200_conductor_fields(eb200_ctx_t* ctx, float* em, int o, int sign, int tags,
                           eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, em != nullptr, "null field");
  REQUIRE(ctx, ctx->cfg.metric == EB200_METRIC_MINKOWSKI,
          "Perfect conductor BCs only applicable to cartesian coordinates");
  REQUIRE(ctx, ctx->cfg.grid.dim == 2, "eb200_conductor_fields: 2D domains");
  REQUIRE(ctx, (o == 0 || o == 1) && sign != 0, "Invalid dimension");
  return check_cuda(ctx, eb200::conductor_fields2d(ctx->cfg.grid, em, o, sign > 0, tags, (cudaStream_t)stream),
                    "conductor_fields");
}

/* ------------------------------------------------ injection and particle moments */
// metric argument of the injection / moment kernels: null on Minkowski contexts
static const eb200::MetricParams* metric_or_null(const eb200_ctx_t* ctx) {
  return ctx->cfg.metric == EB200_METRIC_MINKOWSKI ? nullptr : &ctx->metric;
}

int eb200_inject_nonuniform(eb200_ctx_t* ctx, eb200_species_t* sp1, eb200_species_t* sp2, float ppc,
                            const eb200_spatial_dist_t* sd, const eb200_maxwellian_t* ed1,
                            const eb200_maxwellian_t* ed2, const int* rmin, const int* rmax,
                            uint64_t seed, uint32_t step, uint32_t call, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, sp1 && sp2 && sd && ed1 && ed2 && rmin && rmax, "null argument");
  const int  metric = ctx->cfg.metric;
  const bool mink   = metric == EB200_METRIC_MINKOWSKI;
  REQUIRE(ctx, mink || metric == EB200_METRIC_SPHERICAL || metric == EB200_METRIC_QSPHERICAL,
          "eb200_inject_nonuniform: Minkowski and (q)spherical SRPIC meshes");
  REQUIRE(ctx, ppc >= 0.0f, "ppc < 0");
  REQUIRE(ctx, ed1->temperature >= 0.0f && ed2->temperature >= 0.0f,
          "Maxwellian: Temperature must be non-negative");
  REQUIRE(ctx, sd->kind >= EB200_SDIST_UNIFORM && sd->kind <= EB200_SDIST_ATMOSPHERE,
          "unknown spatial distribution");
  REQUIRE(ctx, sd->kind == EB200_SDIST_UNIFORM || (sd->field != nullptr && sd->comp >= 0),
          "spatial distribution: null field");
  REQUIRE(ctx, sd->kind != EB200_SDIST_REPLENISH_TABLE || (sd->target_field && sd->target_max > 0.0f),
          "Replenish: null target table or target_max <= 0");
  const eb200_grid_t& g = ctx->cfg.grid;
  if (sd->kind == EB200_SDIST_ATMOSPHERE) {
    REQUIRE(ctx, sd->atm_dim >= 0 && sd->atm_dim < g.dim && sd->atm_sign != 0 &&
                   sd->atm_nmax > 0.0f && sd->atm_height > 0.0f,
            "atmosphere: bad direction, density or height");
    // utils.h:59-63, particle_injector.h:161-168
    REQUIRE(ctx, mink || (sd->atm_dim == 0 && sd->atm_sign < 0),
            "For non-cartesian coordinates atmosphere BCs is possible only in -x1 (@ rmin)");
  }
  if (!mink) {
    REQUIRE(ctx, g.dim == 2, "curvilinear injection: 2D");
    REQUIRE(ctx, sd->inv_V0 > 0.0f, "curvilinear injection: inv_V0 must be set (weights)");
    for (const eb200_maxwellian_t* e : { ed1, ed2 }) {
      REQUIRE(ctx, e->drift_u[0] == 0.0f && e->drift_u[1] == 0.0f && e->drift_u[2] == 0.0f,
              "Maxwellian: drift on Cartesian meshes only");
    }
    REQUIRE(ctx, sp1->arrays.weight && sp2->arrays.weight,
            "Weights must be used for non-Cartesian coordinates");
  }
  for (int a = 0; a < g.dim; ++a) {
    REQUIRE(ctx, rmin[a] >= 0 && rmax[a] <= g.n[a] + 2 * g.ng, "inject: range outside the array");
  }
  int rc = check_prtls(ctx, &sp1->arrays, sp1->npart);
  if (rc) return rc;
  rc = check_prtls(ctx, &sp2->arrays, sp2->npart);
  if (rc) return rc;
  size_t plane = 1;
  for (int a = 0; a < g.dim; ++a) plane *= (size_t)(g.n[a] + 2 * g.ng);
  const float* field = sd->field ? sd->field + (size_t)sd->comp * plane : nullptr;
  uint32_t     n_inj = 0;
  int          overflow = 0;
  const float  xmin[3] = { ctx->cfg.metric_params[1], ctx->cfg.metric_params[2], ctx->cfg.metric_params[3] };
  rc = check_cuda(ctx,
                  eb200::inject_nonuniform(g, metric_or_null(ctx), ctx->cfg.metric_params[0], xmin,
                                           sp1->arrays, sp1->npart, sp1->maxnpart, sp2->arrays,
                                           sp2->npart, sp2->maxnpart, ppc, *sd, field, *ed1, *ed2,
                                           rmin, rmax, seed, step, call, &n_inj, &overflow,
                                           ctx->scratch, (cudaStream_t)stream),
                  "inject_nonuniform");
  if (rc) return rc;
  if (overflow) return fail(ctx, EB200_ERR_CAPACITY, "inject: npart + injected > maxnpart");
  sp1->npart += n_inj;
  sp2->npart += n_inj;
  return EB200_OK;
}

int eb200_particle_moment(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart, float mass,
                          float charge, int use_weights, int what, float inv_n0, float* buff,
                          int ncomp, int comp, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, buff != nullptr && comp >= 0 && comp < ncomp, "Invalid buffer index");
  REQUIRE(ctx, what == EB200_STATS_N || what == EB200_STATS_RHO || what == EB200_STATS_CHARGE ||
                 what == EB200_STATS_NPART, "eb200_particle_moment: N, Rho, Charge or Nppc");
  REQUIRE(ctx, !((what == EB200_STATS_RHO || what == EB200_STATS_CHARGE) && mass == 0.0f),
          "Rho & Charge for massless particles not defined");
  int rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  const eb200_grid_t& g    = ctx->cfg.grid;
  const bool          mink = ctx->cfg.metric == EB200_METRIC_MINKOWSKI;
  REQUIRE(ctx, mink || g.dim == 2, "curvilinear moments: 2D");
  const float         dx = ctx->cfg.metric_params[0];
  float               sqrt_det_h = 1.0f; // minkowski.h: dx^D
  size_t              plane      = 1;
  for (int a = 0; a < g.dim; ++a) {
    sqrt_det_h *= dx;
    plane *= (size_t)(g.n[a] + 2 * g.ng);
  }
  float coeff  = (what == EB200_STATS_RHO) ? mass : ((what == EB200_STATS_CHARGE) ? charge : 1.0f);
  bool  uw     = use_weights != 0;
  bool  volume = true;
  if (what != EB200_STATS_NPART) {
    coeff *= inv_n0;
    if (mink) coeff /= sqrt_det_h;
  } else {
    uw = false, volume = false; // Nppc: no volume, weights or smoothing (particle_moments.hpp:306-309)
  }
  return check_cuda(ctx,
                    eb200::particle_moment(g, metric_or_null(ctx), *prtls, npart, coeff, uw, volume,
                                           buff + (size_t)comp * plane, (cudaStream_t)stream),
                    "particle_moment");
}

int eb200_atmosphere_particles(eb200_ctx_t* ctx, const eb200_atmosphere_t* atm,
                               eb200_species_t* species, int nspecies, float* plane,
                               int assume_empty, uint32_t step, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, atm && species && plane, "null argument");
  REQUIRE(ctx, atm->species[0] >= 0 && atm->species[0] < nspecies && atm->species[1] >= 0 &&
                 atm->species[1] < nspecies && atm->species[0] != atm->species[1],
          "atmosphere: bad species pair");
  const eb200_grid_t& g = ctx->cfg.grid;
  size_t              n = 1;
  for (int a = 0; a < g.dim; ++a) n *= (size_t)(g.n[a] + 2 * g.ng);
  // particles_bcs.h:56-119: the density of the two species (Rho, weights on curvilinear meshes)
  if (cudaMemsetAsync(plane, 0, n * sizeof(float), (cudaStream_t)stream) != cudaSuccess) {
    return fail(ctx, EB200_ERR_CUDA, "atmosphere: memset");
  }
  const int use_weights = ctx->cfg.metric != EB200_METRIC_MINKOWSKI;
  if (!assume_empty) {
    for (int k = 0; k < 2; ++k) {
      eb200_species_t& sp = species[atm->species[k]];
      if (sp.npart == 0) continue;
      int rc = eb200_particle_moment(ctx, &sp.arrays, sp.npart, sp.mass, sp.charge, use_weights,
                                     EB200_STATS_RHO, atm->inv_n0, plane, 1, 0, stream);
      if (rc) return rc;
    }
  }
  // :121-152: InjectNonUniform<Replenish<AtmosphereDensityProfile>>, number_density = nmax
  eb200_spatial_dist_t sd {};
  sd.kind = EB200_SDIST_ATMOSPHERE, sd.field = plane, sd.comp = 0;
  sd.atm_dim = atm->dim, sd.atm_sign = atm->sign;
  sd.atm_nmax = atm->density, sd.atm_height = atm->height;
  sd.atm_xsurf = atm->x_surf, sd.atm_ds = atm->ds;
  sd.inv_V0 = atm->inv_V0;
  const eb200_maxwellian_t mw { atm->temperature, { 0.0f, 0.0f, 0.0f } };
  int rmin[3] = { 0, 0, 0 }, rmax[3] = { 1, 1, 1 };
  for (int a = 0; a < g.dim; ++a) rmin[a] = g.ng, rmax[a] = g.ng + g.n[a];
  return eb200_inject_nonuniform(ctx, &species[atm->species[0]], &species[atm->species[1]],
                                 atm->density * atm->ppc0 * 0.5f, &sd, &mw, &mw, rmin, rmax,
                                 atm->seed, step, /*call*/ 0x41544du, stream);
}

/* ------------------------------------------------------------------ output staging */
int eb200_fields_to_phys(eb200_ctx_t* ctx, const float* from, int ncomp_from, float* to,
                         int ncomp_to, const int* cf, const int* ct, int interp, int convert,
                         eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, from && to && cf && ct, "null argument");
  REQUIRE(ctx, ncomp_from >= 3 && ncomp_to >= 3, "FieldsToPhys_kernel: at least 3 components");
  for (int c = 0; c < 3; ++c) {
    REQUIRE(ctx, cf[c] >= 0 && cf[c] < ncomp_from && ct[c] >= 0 && ct[c] < ncomp_to,
            "FieldsToPhys_kernel: Invalid component index");
  }
  REQUIRE(ctx, interp >= 0 && interp <= 2 && convert >= 0 && convert <= 3, "bad flags");
  const bool mink = ctx->cfg.metric == EB200_METRIC_MINKOWSKI;
  return check_cuda(ctx,
                    eb200::fields_to_phys(mink ? nullptr : &ctx->metric, ctx->cfg.grid,
                                          ctx->cfg.metric_params[0], from, ncomp_from, to, ncomp_to,
                                          cf, ct, interp, convert, (cudaStream_t)stream),
                    "fields_to_phys");
}

int eb200_prtls_to_phys(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t npart,
                        uint32_t stride, float* x1, float* x2, float* x3, float* u1, float* u2,
                        float* u3, float* weight, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, stride >= 1, "stride must be >= 1");
  int rc = check_prtls(ctx, prtls, npart);
  if (rc) return rc;
  const bool mink = ctx->cfg.metric == EB200_METRIC_MINKOWSKI;
  const int  dim  = ctx->cfg.grid.dim;
  REQUIRE(ctx, x1 && u1 && u2 && u3 && weight, "Invalid buffer size");
  REQUIRE(ctx, dim < 2 || x2, "Invalid buffer size");
  REQUIRE(ctx, !((dim == 2 && !mink) || dim == 3) || x3, "Invalid buffer size");
  REQUIRE(ctx, mink || prtls->phi != nullptr, "phi is required on curvilinear meshes");
  const uint32_t nout = (npart + stride - 1) / stride;
  return check_cuda(ctx,
                    eb200::prtls_to_phys(mink ? nullptr : &ctx->metric, ctx->cfg.grid,
                                         ctx->cfg.metric_params[0], ctx->cfg.metric_params + 1, *prtls,
                                         stride, nout, x1, x2, x3, u1, u2, u3, weight,
                                         (cudaStream_t)stream),
                    "prtls_to_phys");
}

int eb200_time_average(eb200_ctx_t* ctx, float* a, const float* b, int ncomp,
                       eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, ctx->cfg.grid.dim == 2, "eb200_time_average: 2D only");
  REQUIRE(ctx, a && b && ncomp > 0, "bad argument");
  return check_cuda(ctx, eb200::curv::time_average(ctx->cfg.grid, a, b, ncomp, (cudaStream_t)stream),
                    "time_average");
}

int eb200_metric_eval(int metric, const int* n_active, const float* metric_params, int nq,
                      const float* x1, const float* x2, float* out) {
  if (!n_active || !metric_params || !x1 || !x2 || !out || nq < 0) {
    return fail(nullptr, EB200_ERR_ARG, "eb200_metric_eval: bad argument");
  }
  const float*              mp = metric_params;
  const eb200::MetricParams m  = eb200::make_metric(metric, n_active[0], n_active[1], mp[0], mp[1],
                                                    mp[2], mp[3], mp[4], mp[5], mp[6]);
  switch (metric) {
    case EB200_METRIC_SPHERICAL: eval_sr_metric<eb200::Spherical>(m, nq, x1, x2, out); break;
    case EB200_METRIC_QSPHERICAL: eval_sr_metric<eb200::QSpherical>(m, nq, x1, x2, out); break;
    case EB200_METRIC_KERR_SCHILD: eval_gr_metric<eb200::KerrSchild>(m, nq, x1, x2, out); break;
    case EB200_METRIC_QKERR_SCHILD: eval_gr_metric<eb200::QKerrSchild>(m, nq, x1, x2, out); break;
    case EB200_METRIC_KERR_SCHILD_0: eval_gr_metric<eb200::KerrSchild0>(m, nq, x1, x2, out); break;
    default: return fail(nullptr, EB200_ERR_UNSUPPORTED, "eb200_metric_eval: unknown metric");
  }
  return EB200_OK;
}

int eb200_set_pd_kernel(eb200_ctx_t* ctx, int which) {
  ENTER(ctx);
  REQUIRE(ctx, which >= 0 && which <= 9,
          "pd kernel: 0 auto, 1 per-thread, 2 TMA stream, 3 vec4, 4 shared-memory tiles, 5 vec4 + packed nodes, 6 pipelined, 7 shared-memory resident, 8 vec4 + packed nodes + moment deposit, 9 3D O=3 shared-memory J tile");
  ctx->pd_kernel = which;
  return EB200_OK;
}

int eb200_set_sort_mode(eb200_ctx_t* ctx, int mode) {
  ENTER(ctx);
  REQUIRE(ctx, mode >= -1 && mode <= 1, "sort mode: -1 by build, 0 radix (stable), 1 counting");
  ctx->sort_mode = mode;
  return EB200_OK;
}

int eb200_set_lean_prev(eb200_ctx_t* ctx, int on) {
  ENTER(ctx);
  ctx->lean_prev = on != 0;
  return EB200_OK;
}

// the sort flags of the step mirrors (engine.cu)
int eb200_ctx_sort_flags(const eb200_ctx_t* ctx) {
  const int mode = ctx->sort_mode >= 0 ? ctx->sort_mode : (ctx->cfg.strict_fp ? 0 : 1);
  return (mode == 1 ? EB200_SORT_UNSTABLE : 0) | (ctx->lean_prev ? EB200_SORT_SKIP_PREV : 0);
}

int eb200_ctx_lean_prev(const eb200_ctx_t* ctx) { return ctx->lean_prev; }

int eb200_zero_currents(eb200_ctx_t* ctx, float* cur, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, cur != nullptr, "cur is null");
  return check_cuda(ctx,
                    cudaMemsetAsync(cur, 0, field_bytes(ctx->cfg.grid, 3), (cudaStream_t)stream),
                    "zero_currents");
}

int eb200_comm_fields(eb200_ctx_t* ctx, float* fld, int ncomp, int c0, int c1, const int* fbc,
                      eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, fld != nullptr && fbc != nullptr, "null argument");
  REQUIRE(ctx, 0 <= c0 && c0 < c1 && c1 <= ncomp, "bad component range");
  if (ctx->comm) {
    int rc = eb200::comm_fields(*ctx->comm, fld, c0, c1, (cudaStream_t)stream);
    return rc == EB200_OK ? rc : fail(ctx, rc, eb200::comm_error(ctx->comm));
  }
  return check_cuda(ctx,
                    VARIANT_CALL(ctx, comm_fields_self(ctx->cfg.grid, fld, c0, c1, fbc,
                                                       (cudaStream_t)stream)),
                    "comm_fields");
}

int eb200_sync_currents(eb200_ctx_t* ctx, float* cur, float* buff, const int* fbc,
                        eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, cur != nullptr && fbc != nullptr, "null argument");
  if (ctx->comm) {
    int rc = eb200::comm_sync_currents(*ctx->comm, cur, (cudaStream_t)stream);
    return rc == EB200_OK ? rc : fail(ctx, rc, eb200::comm_error(ctx->comm));
  }
  return check_cuda(ctx,
                    VARIANT_CALL(ctx, sync_currents_self(ctx->cfg.grid, cur, buff, fbc,
                                                         (cudaStream_t)stream)),
                    "sync_currents");
}

int eb200_sort_particles(eb200_ctx_t* ctx, const eb200_prtls_t* prtls, uint32_t* npart_inout,
                         int remove_dead, eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, npart_inout != nullptr, "npart pointer is null");
  int rc = check_prtls(ctx, prtls, *npart_inout);
  if (rc) return rc;
  uint32_t    n_alive = *npart_inout;
  // The scratch layout follows a capacity, not npart: with migration npart moves from sort to
  // sort, and every growth of the 5 GB scratch is a cudaFree + cudaMalloc (13 ms per sort measured
  // at 4 x B200). Without a configured capacity keep a high-water mark with 1/8 headroom.
  {
    const uint32_t n    = *npart_inout;
    const uint64_t want = (uint64_t)n + n / 8;
    if (ctx->cfg.maxnpart > ctx->sort_cap) ctx->sort_cap = ctx->cfg.maxnpart;
    if (n > ctx->sort_cap) ctx->sort_cap = (uint32_t)(want < 0xFFFFFFFFull ? want : 0xFFFFFFFFull);
  }
  cudaError_t e = eb200::sort_particles(ctx->cfg.grid, *prtls, *npart_inout, ctx->sort_cap,
                                        remove_dead, &n_alive, ctx->scratch, (cudaStream_t)stream);
  if (e != cudaSuccess) return check_cuda(ctx, e, "sort_particles");
  if (remove_dead & 1) *npart_inout = n_alive;
  return EB200_OK;
}

int eb200_comm_unique_id(char* id_out) {
  if (!id_out) return fail(nullptr, EB200_ERR_ARG, "eb200_comm_unique_id: null argument");
  std::string err;
  int rc = eb200::comm_unique_id(id_out, err);
  return rc == EB200_OK ? rc : fail(nullptr, rc, err);
}

int eb200_comm_init(eb200_ctx_t* ctx, const eb200_metadomain_t* md, const char* id) {
  ENTER(ctx);
  REQUIRE(ctx, md != nullptr, "metadomain is null");
  REQUIRE(ctx, id != nullptr || md->nranks == 1, "a unique id is required for nranks > 1");
  if (ctx->comm) {
    eb200::comm_delete(ctx->comm);
    ctx->comm = nullptr;
  }
  eb200::Comm* c = eb200::comm_new();
  int rc = eb200::comm_setup(*c, ctx->cfg.grid, *md, id);
  if (rc != EB200_OK) {
    fail(ctx, rc, eb200::comm_error(c));
    eb200::comm_delete(c);
    return rc;
  }
  ctx->comm = c;
  return EB200_OK;
}

int eb200_comm_particles(eb200_ctx_t* ctx, eb200_species_t* species, int nspecies,
                         eb200_stream_t stream) {
  ENTER(ctx);
  REQUIRE(ctx, nspecies == 0 || species != nullptr, "species is null");
  if (!ctx->comm) return EB200_OK; // a single self-periodic domain migrates nothing
  for (int s = 0; s < nspecies; ++s) {
    int rc = check_prtls(ctx, &species[s].arrays, species[s].npart);
    if (rc) return rc;
  }
  // exception lists left by this step's fused push launches, matched by tag array and count
  std::vector<eb200::ExcList> lists(nspecies > 0 ? nspecies : 1);
  for (int s = 0; s < nspecies; ++s) {
    for (auto* q : ctx->exc) {
      if (q->tag == species[s].arrays.tag && q->npart == species[s].npart && q->list.tracked) {
        lists[s] = q->list;
      }
    }
  }
  int rc = eb200::comm_particles(*ctx->comm, species, nspecies, (cudaStream_t)stream, lists.data());
  for (auto* q : ctx->exc) q->list.tracked = false; // the arrays have changed: lists are stale
  return rc == EB200_OK ? rc : fail(ctx, rc, eb200::comm_error(ctx->comm));
}

// internal: 1 when a communicator / decomposition is attached (engine.cu)
int eb200_ctx_has_comm(const eb200_ctx_t* ctx) { return ctx && ctx->comm ? 1 : 0; }

} // extern "C"
