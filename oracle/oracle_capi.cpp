// TEST INFRASTRUCTURE - see oracle.hpp.  Flat C API over the oracle so tests and
// bench.py's cpu_baseline leg can drive it through ctypes with the same plain
// structs as include/formgpu.h.
#include "formgpu.h"
#include "oracle_hotpath.hpp"

#include <cstring>

using namespace form_oracle;

static_assert(sizeof(formgpu_scan_pose) == sizeof(form::ScanPose), "ScanPose layout");
static_assert(sizeof(formgpu_pair) == sizeof(form::PairKey), "PairKey layout");
static_assert(sizeof(formgpu_pair_count) == sizeof(form::PairCount), "PairCount layout");
static_assert(sizeof(formgpu_planar_feat) == sizeof(PlanarFeat), "PlanarFeat layout");
static_assert(sizeof(formgpu_point_feat) == sizeof(PointFeat), "PointFeat layout");

namespace {
form::HotPathParams to_host_params(const formgpu_params *p, int threads) {
  form::HotPathParams h;
  h.neighbor_points = (size_t)p->neighbor_points;
  h.num_sectors = (size_t)p->num_sectors;
  h.planar_threshold = p->planar_threshold;
  h.planar_feats_per_sector = (size_t)p->planar_feats_per_sector;
  h.point_feats_per_sector = (size_t)p->point_feats_per_sector;
  h.radius = p->radius;
  h.min_points = (size_t)p->min_points;
  h.min_norm_squared = p->min_norm_squared;
  h.max_norm_squared = p->max_norm_squared;
  h.num_columns = p->num_columns;
  h.num_rows = p->num_rows;
  h.max_dist_matching = p->max_dist_matching;
  h.min_dist_map = p->min_dist_map;
  h.sigma = p->sigma;
  h.num_threads = (size_t)threads;
  return h;
}
template <typename T> void copy_out(const std::vector<T> &v, T *out) {
  if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(T));
}
} // namespace

extern "C" {

void *oracle_create(const formgpu_params *p, int threads) {
  return new OracleHotPath(to_host_params(p, threads));
}
void oracle_destroy(void *h) { delete static_cast<OracleHotPath *>(h); }

int oracle_extract(void *h, const formgpu_point4f *scan, size_t n, uint64_t scan_idx,
                   size_t *n_planar, size_t *n_point) {
  auto *o = static_cast<OracleHotPath *>(h);
  std::vector<PlanarFeat> pl;
  std::vector<PointFeat> pt;
  try {
    o->extract(reinterpret_cast<const PointXYZf *>(scan), n, scan_idx, pl, pt);
  } catch (const form::HotPathError &) {
    return FORMGPU_ERR_BAD_SCAN_SIZE;
  }
  *n_planar = pl.size();
  *n_point = pt.size();
  return FORMGPU_OK;
}

void oracle_get_features(void *h, formgpu_planar_feat *planar, formgpu_point_feat *point) {
  auto *o = static_cast<OracleHotPath *>(h);
  copy_out(o->cur_planar, reinterpret_cast<PlanarFeat *>(planar));
  copy_out(o->cur_point, reinterpret_cast<PointFeat *>(point));
}

void oracle_extract_debug(void *h, uint8_t *valid, uint8_t *pvalid, float *curv,
                          uint32_t *planar_idx, uint8_t *keep, int32_t *cprev, int32_t *cnext,
                          size_t *n_planar_picks, uint32_t *point_idx, size_t *n_point_picks) {
  auto *o = static_cast<OracleHotPath *>(h);
  const ExtractResult &r = o->last_extract;
  copy_out(r.valid_mask, valid);
  copy_out(r.point_valid_mask, pvalid);
  copy_out(r.curvature, curv);
  copy_out(r.planar_indices, planar_idx);
  copy_out(r.planar_keep, keep);
  copy_out(r.closest_prev, cprev);
  copy_out(r.closest_next, cnext);
  copy_out(r.point_indices, point_idx);
  if (n_planar_picks) *n_planar_picks = r.planar_indices.size();
  if (n_point_picks) *n_point_picks = r.point_indices.size();
}

void oracle_map_rebuild(void *h, const formgpu_scan_pose *poses, size_t n) {
  static_cast<OracleHotPath *>(h)->map_rebuild(reinterpret_cast<const form::ScanPose *>(poses), n);
}

size_t oracle_num_voxels(void *h, int type) {
  auto *o = static_cast<OracleHotPath *>(h);
  return type == 0 ? o->planar_map.num_voxels() : o->point_map.num_voxels();
}

int oracle_associate(void *h, const formgpu_pose *pose_k, formgpu_pair_count *out, size_t cap,
                     size_t *n) {
  auto *o = static_cast<OracleHotPath *>(h);
  std::vector<form::PairCount> c;
  o->associate(*reinterpret_cast<const Pose3 *>(pose_k), c);
  *n = c.size();
  if (c.size() > cap) return FORMGPU_ERR_CAPACITY;
  copy_out(c, reinterpret_cast<form::PairCount *>(out));
  return FORMGPU_OK;
}

int oracle_get_matches(void *h, int type, formgpu_match *out, size_t cap, size_t *n) {
  auto *o = static_cast<OracleHotPath *>(h);
  auto fill = [&](const auto &matches) {
    *n = matches.size();
    if (matches.size() > cap) return FORMGPU_ERR_CAPACITY;
    for (size_t j = 0; j < matches.size(); ++j) {
      out[j].scan = matches[j].found ? matches[j].scan : 0;
      out[j].k = matches[j].found ? matches[j].k : 0;
      out[j].found = matches[j].found ? 1u : 0u;
      out[j].dist_sqrd = matches[j].dist_sqrd;
    }
    return FORMGPU_OK;
  };
  return type == 0 ? fill(o->planar_matches) : fill(o->point_matches);
}

/// Correspondences of pair (i, j) as the reference stores them (factor.hpp:52-54,
/// :99-100): planar p_i, n_i, p_j (3 doubles each) and point p_i, p_j.
int oracle_get_pair(void *h, uint64_t i, uint64_t j, double *pl_pi, double *pl_ni,
                    double *pl_pj, size_t *n_planar, double *pt_pi, double *pt_pj,
                    size_t *n_point) {
  auto *o = static_cast<OracleHotPath *>(h);
  *n_planar = *n_point = 0;
  auto jt = o->constraints.find(j);
  if (jt == o->constraints.end()) return FORMGPU_OK;
  auto it = jt->second.find(i);
  if (it == jt->second.end()) return FORMGPU_OK;
  const PairConstraints &c = it->second;
  *n_planar = c.plane.num_constraints();
  *n_point = c.point.num_constraints();
  copy_out(c.plane.p_i, pl_pi);
  copy_out(c.plane.n_i, pl_ni);
  copy_out(c.plane.p_j, pl_pj);
  copy_out(c.point.p_i, pt_pi);
  copy_out(c.point.p_j, pt_pj);
  return FORMGPU_OK;
}

void oracle_linearize(void *h, const formgpu_pair *pairs, size_t n_pairs,
                      const formgpu_scan_pose *poses, size_t n_poses, double *out91) {
  static_cast<OracleHotPath *>(h)->linearize(reinterpret_cast<const form::PairKey *>(pairs),
                                             n_pairs,
                                             reinterpret_cast<const form::ScanPose *>(poses),
                                             n_poses, out91);
}

void oracle_error(void *h, const formgpu_pair *pairs, size_t n_pairs,
                  const formgpu_scan_pose *poses, size_t n_poses, double *out) {
  static_cast<OracleHotPath *>(h)->error(reinterpret_cast<const form::PairKey *>(pairs), n_pairs,
                                         reinterpret_cast<const form::ScanPose *>(poses), n_poses,
                                         out);
}

void oracle_commit_scan(void *h, size_t *n_planar, size_t *n_point) {
  static_cast<OracleHotPath *>(h)->commit_scan(*n_planar, *n_point);
}

void oracle_remove_scans(void *h, const uint64_t *scans, size_t n) {
  static_cast<OracleHotPath *>(h)->remove_scans(scans, n);
}

int oracle_get_keypoints(void *h, int type, uint64_t scan, void *out, size_t cap, size_t *n) {
  auto *o = static_cast<OracleHotPath *>(h);
  if (type == 0) {
    const auto &m = o->planar_map.scans();
    auto it = m.find(scan);
    *n = it == m.end() ? 0 : it->second.size();
    if (out && it != m.end()) {
      if (*n > cap) return FORMGPU_ERR_CAPACITY;
      copy_out(it->second, static_cast<PlanarFeat *>(out));
    }
  } else {
    const auto &m = o->point_map.scans();
    auto it = m.find(scan);
    *n = it == m.end() ? 0 : it->second.size();
    if (out && it != m.end()) {
      if (*n > cap) return FORMGPU_ERR_CAPACITY;
      copy_out(it->second, static_cast<PointFeat *>(out));
    }
  }
  return FORMGPU_OK;
}

// ---- stand-alone kernels of the oracle, for unit pins ----

void oracle_eigen3f(const float cov[9], float evals[3], float evecs[9]) {
  self_adjoint_eigen3f(cov, evals, evecs);
}

void oracle_compute_coords(double x, double y, double z, double w, int32_t out[3]) {
  const VoxelKey k = compute_coords(x, y, z, w);
  out[0] = k.x;
  out[1] = k.y;
  out[2] = k.z;
}

void oracle_voxel_shifts(int32_t out[81]) {
  for (int s = 0; s < 27; ++s)
    for (int a = 0; a < 3; ++a) out[3 * s + a] = kVoxelShifts[s][a];
}

/// PlanePoint::evaluateError on raw arrays: r[n], H1[n*6], H2[n*6].
void oracle_plane_point(const double *p_i, const double *n_i, const double *p_j, size_t n,
                        const formgpu_pose *Ti, const formgpu_pose *Tj, double *r, double *H1,
                        double *H2) {
  PlanePoint pp;
  pp.p_i.assign(p_i, p_i + 3 * n);
  pp.n_i.assign(n_i, n_i + 3 * n);
  pp.p_j.assign(p_j, p_j + 3 * n);
  pp.evaluate(*reinterpret_cast<const Pose3 *>(Ti), *reinterpret_cast<const Pose3 *>(Tj), r, H1,
              H2);
}

/// PointPoint::evaluateError on raw arrays: r[3m], H1[3m*6], H2[3m*6].
void oracle_point_point(const double *p_i, const double *p_j, size_t m, const formgpu_pose *Ti,
                        const formgpu_pose *Tj, double *r, double *H1, double *H2) {
  PointPoint pp;
  pp.p_i.assign(p_i, p_i + 3 * m);
  pp.p_j.assign(p_j, p_j + 3 * m);
  pp.evaluate(*reinterpret_cast<const Pose3 *>(Ti), *reinterpret_cast<const Pose3 *>(Tj), r, H1,
              H2);
}

/// linearize_pair / error_pair on raw arrays.
void oracle_linearize_raw(const double *pl_pi, const double *pl_ni, const double *pl_pj, size_t n,
                          const double *pt_pi, const double *pt_pj, size_t m,
                          const formgpu_pose *Ti, const formgpu_pose *Tj, double sigma,
                          double out91[91], double *err) {
  PairConstraints c;
  c.plane.p_i.assign(pl_pi, pl_pi + 3 * n);
  c.plane.n_i.assign(pl_ni, pl_ni + 3 * n);
  c.plane.p_j.assign(pl_pj, pl_pj + 3 * n);
  c.point.p_i.assign(pt_pi, pt_pi + 3 * m);
  c.point.p_j.assign(pt_pj, pt_pj + 3 * m);
  const Pose3 &A = *reinterpret_cast<const Pose3 *>(Ti), &B = *reinterpret_cast<const Pose3 *>(Tj);
  if (out91) linearize_pair(c, A, B, sigma, out91);
  if (err) *err = error_pair(c, A, B, sigma);
}

} // extern "C"
