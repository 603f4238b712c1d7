// TEST INFRASTRUCTURE ONLY.
// extern "C" driver around the REFERENCE's own Mesh::Intersects / Mesh::ExtentToRange
// (src/framework/domain/mesh.h:69-203) and the box construction of srpic::MatchFieldsIn
// (src/engines/srpic/fields_bcs.h:72-114, restated in the twenty lines below because that
// function needs a whole Domain), compiled in place from $(REF)/src against ref_shim/.
#include "oracle.h"

#include "enums.h"
#include "global.h"

#include "arch/kokkos_aliases.h"
#include "utils/numeric.h"

#include "metrics/minkowski.h"

#include "framework/domain/mesh.h"

#include <map>
#include <string>
#include <vector>

using namespace ntt;

namespace {
  template <Dimension D>
  int layer(const orc_grid_t* g, const float* xmin, const float* xmax, float gx_lo, float gx_hi,
            int o, int sign, float ds, float* edge_out, int* rmin, int* rmax) {
    using M = metric::Minkowski<D>;
    std::vector<ncells_t> res;
    boundaries_t<real_t>  ext;
    for (int a = 0; a < (int)D; ++a) {
      res.push_back((ncells_t)g->n[a]);
      ext.push_back({ xmin[a], xmax[a] });
    }
    Mesh<M> mesh(res, ext, std::map<std::string, real_t> {});
    // fields_bcs.h:75-103
    real_t xg_min, xg_max, xg_edge;
    if (sign > 0) {
      xg_max  = gx_hi;
      xg_min  = xg_max - ds;
      xg_edge = xg_max;
    } else {
      xg_min  = gx_lo;
      xg_max  = xg_min + ds;
      xg_edge = xg_min;
    }
    boundaries_t<real_t> box;
    boundaries_t<bool>   incl_ghosts;
    for (dim_t d { 0 }; d < (dim_t)D; ++d) {
      if (d == (dim_t)o) {
        box.emplace_back(xg_min, xg_max);
        if (sign > 0) {
          incl_ghosts.emplace_back(false, true);
        } else {
          incl_ghosts.emplace_back(true, false);
        }
      } else {
        box.push_back(Range::All);
        incl_ghosts.emplace_back(true, true);
      }
    }
    *edge_out = xg_edge;
    if (not mesh.Intersects(box)) {
      return 0;
    }
    const auto r = mesh.ExtentToRange(box, incl_ghosts);
    for (int d = 0; d < (int)D; ++d) {
      rmin[d] = (int)r[d].first;
      rmax[d] = (int)r[d].second;
    }
    return 1;
  }
} // namespace

extern "C" int ref_match_layer(const orc_grid_t* g, const float* xmin, const float* xmax,
                               float gx_lo, float gx_hi, int o, int sign, float ds,
                               float* edge_out, int* rmin, int* rmax) {
  switch (g->dim) {
    case 1: return layer<Dim::_1D>(g, xmin, xmax, gx_lo, gx_hi, o, sign, ds, edge_out, rmin, rmax);
    case 2: return layer<Dim::_2D>(g, xmin, xmax, gx_lo, gx_hi, o, sign, ds, edge_out, rmin, rmax);
    default: return layer<Dim::_3D>(g, xmin, xmax, gx_lo, gx_hi, o, sign, ds, edge_out, rmin, rmax);
  }
}
