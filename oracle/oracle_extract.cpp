// TEST INFRASTRUCTURE - see oracle.hpp. Stage 1 oracle: feature extraction.
// Follows /root/reference/form/feature/extraction.tpp:29-448 step by step.
#include "oracle.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <thread>

namespace form_oracle {

namespace {

// 4-lane float squared norm of a difference, Eigen packet order (A.2):
// (d0^2 + d2^2) + (d1^2 + d3^2).  extraction.tpp:413,430,441 via utils.hpp:57-60.
inline float diff_sqnorm4(const PointXYZf &a, const PointXYZf &b) {
  const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a._ - b._;
  return (d0 * d0 + d2 * d2) + (d1 * d1 + d3 * d3);
}

// compute_valid_points, extraction.tpp:136-180 (dilate = true) and
// compute_point_valid_points, :182-222 (dilate = false).
std::vector<uint8_t> valid_points(const ExtractParams &P, const PointXYZf *scan, bool dilate) {
  const size_t rows = (size_t)P.num_rows, cols = (size_t)P.num_columns;
  std::vector<uint8_t> mask(rows * cols, 1);
  for (size_t r = 0; r < rows; ++r) {
    for (size_t c = 0; c < cols; ++c) {
      const size_t idx = r * cols + c;
      // CHECK 1 (:159-163): row ends are invalid and never dilate
      if (c < P.neighbor_points || c >= cols - P.neighbor_points) {
        mask[idx] = 0;
        continue;
      }
      // CHECK 2 (:166-175): float norm promoted to double, strict compares
      const double range2 = (double)scan[idx].squaredNorm();
      if (range2 < P.min_norm_squared || range2 > P.max_norm_squared) {
        mask[idx] = 0;
        if (dilate) {
          for (size_t i = 1; i <= P.neighbor_points; ++i) {
            mask[idx - i] = 0;
            mask[idx + i] = 0;
          }
        }
      }
    }
  }
  return mask;
}

// compute_curvature, extraction.tpp:226-261; stored as float (extraction.hpp:45,48).
std::vector<float> curvature_of(const ExtractParams &P, const PointXYZf *scan,
                                const std::vector<uint8_t> &mask) {
  const size_t rows = (size_t)P.num_rows, cols = (size_t)P.num_columns;
  std::vector<float> curv(rows * cols);
  for (size_t idx = 0; idx < rows * cols; ++idx) {
    if (!mask[idx]) {
      curv[idx] = FLT_MAX;
      continue;
    }
    const double k = -(2.0 * (double)P.neighbor_points);
    double dx = k * (double)scan[idx].x;
    double dy = k * (double)scan[idx].y;
    double dz = k * (double)scan[idx].z;
    for (size_t n = 1; n <= P.neighbor_points; ++n) {
      dx = (dx + (double)scan[idx - n].x) + (double)scan[idx + n].x;
      dy = (dy + (double)scan[idx - n].y) + (double)scan[idx + n].y;
      dz = (dz + (double)scan[idx - n].z) + (double)scan[idx + n].z;
    }
    curv[idx] = (float)((dx * dx + dy * dy) + dz * dz);
  }
  return curv;
}

// find_neighbors, extraction.tpp:422-448: + direction then - direction, stop at
// the first out-of-radius neighbour, masks ignored.
void find_neighbors(const ExtractParams &P, size_t idx, const PointXYZf *scan,
                    std::vector<PointXYZf> &out) {
  const PointXYZf &p = scan[idx];
  const double r2 = P.radius * P.radius;
  for (size_t i = 1; i <= P.neighbor_points; ++i) {
    const PointXYZf &q = scan[idx + i];
    if ((double)diff_sqnorm4(q, p) < r2) out.push_back(q);
    else break;
  }
  for (size_t i = 1; i <= P.neighbor_points; ++i) {
    const PointXYZf &q = scan[idx - i];
    if ((double)diff_sqnorm4(q, p) < r2) out.push_back(q);
    else break;
  }
}

// find_closest, extraction.tpp:402-420 with rule R2 (strict < over ascending idx
// == min (dist2, idx)).
int64_t find_closest(const PointXYZf &p, size_t start, size_t end, const PointXYZf *scan,
                     const std::vector<uint8_t> &valid) {
  int64_t best = -1;
  double min_d2 = std::numeric_limits<double>::max();
  for (size_t idx = start; idx < end; ++idx) {
    if (!valid[idx]) continue;
    const double d2 = (double)diff_sqnorm4(scan[idx], p);
    if (d2 < min_d2) {
      min_d2 = d2;
      best = (int64_t)idx;
    }
  }
  return best;
}

// compute_normal, extraction.tpp:263-329.
bool compute_normal(const ExtractParams &P, size_t idx, const PointXYZf *scan,
                    const std::vector<uint8_t> &valid, float normal[3], int32_t &cprev,
                    int32_t &cnext) {
  const size_t cols = (size_t)P.num_columns, rows = (size_t)P.num_rows;
  const size_t row = idx / cols;
  const PointXYZf &p = scan[idx];
  std::vector<PointXYZf> nbrs;
  nbrs.reserve(4 * P.neighbor_points + 2);
  find_neighbors(P, idx, scan, nbrs);
  bool other = false;
  cprev = cnext = -1;
  if (row > 0) {
    const int64_t c = find_closest(p, cols * (row - 1), cols * row, scan, valid);
    if (c >= 0) {
      other = true;
      cprev = (int32_t)c;
      nbrs.push_back(scan[c]);
      find_neighbors(P, (size_t)c, scan, nbrs);
    }
  }
  if (row < rows - 1) {
    const int64_t c = find_closest(p, cols * (row + 1), cols * (row + 2), scan, valid);
    if (c >= 0) {
      other = true;
      cnext = (int32_t)c;
      nbrs.push_back(scan[c]);
      find_neighbors(P, (size_t)c, scan, nbrs);
    }
  }
  if (!other || nbrs.size() < P.min_points) return false;

  // A = (nbr - p) / n ; Cov = A^T A, float (:314-319). Sequential sum over
  // neighbours in list order.
  const float nf = (float)nbrs.size();
  float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t j = 0; j < nbrs.size(); ++j) {
    const float a0 = (nbrs[j].x - p.x) / nf;
    const float a1 = (nbrs[j].y - p.y) / nf;
    const float a2 = (nbrs[j].z - p.z) / nf;
    cov[0] = cov[0] + a0 * a0;
    cov[3] = cov[3] + a1 * a0;
    cov[4] = cov[4] + a1 * a1;
    cov[6] = cov[6] + a2 * a0;
    cov[7] = cov[7] + a2 * a1;
    cov[8] = cov[8] + a2 * a2;
  }
  cov[1] = cov[3];
  cov[2] = cov[6];
  cov[5] = cov[7];
  float evals[3], evecs[9];
  self_adjoint_eigen3f(cov, evals, evecs);
  // eigenvector of the smallest eigenvalue = column 0, then normalize() (:325-326)
  float n0 = evecs[0], n1 = evecs[3], n2 = evecs[6];
  const float z = n0 * n0 + (n1 * n1 + n2 * n2);
  if (z > 0.0f) {
    const float s = std::sqrt(z);
    n0 = n0 / s;
    n1 = n1 / s;
    n2 = n2 / s;
  }
  normal[0] = n0;
  normal[1] = n1;
  normal[2] = n2;
  return true;
}

} // namespace

// ---------------------------------------------------------------------------
// Eigen::SelfAdjointEigenSolver<Matrix3f>::compute (iterative path, Eigen 3.4
// [external], restated from the published algorithm): scale to [-1,1],
// closed-form 3x3 Householder tridiagonalisation, implicit symmetric QR steps
// with Wilkinson shift, eigenvalue sort.  Tolerance-class output (the normal),
// but evaluated identically on CPU and GPU so both sides agree bit-for-bit.
// ---------------------------------------------------------------------------
namespace {
inline float hypot_pos(float x, float y) {
  const float ax = std::fabs(x), ay = std::fabs(y);
  const float p = ax > ay ? ax : ay;
  if (p == 0.0f) return 0.0f;
  const float q = ax > ay ? ay : ax;
  const float qp = q / p;
  return p * std::sqrt(1.0f + qp * qp);
}

inline void make_givens(float p, float q, float &c, float &s) {
  if (q == 0.0f) {
    c = p < 0.0f ? -1.0f : 1.0f;
    s = 0.0f;
  } else if (p == 0.0f) {
    c = 0.0f;
    s = q < 0.0f ? 1.0f : -1.0f;
  } else if (std::fabs(p) > std::fabs(q)) {
    const float t = q / p;
    float u = std::sqrt(1.0f + t * t);
    if (p < 0.0f) u = -u;
    c = 1.0f / u;
    s = -t * c;
  } else {
    const float t = p / q;
    float u = std::sqrt(1.0f + t * t);
    if (q < 0.0f) u = -u;
    s = -1.0f / u;
    c = -t * s;
  }
}
} // namespace

void self_adjoint_eigen3f(const float cov[9], float evals[3], float evecs[9]) {
  // lower triangle, scaled by the max |coeff|
  float m00 = cov[0], m10 = cov[3], m11 = cov[4], m20 = cov[6], m21 = cov[7], m22 = cov[8];
  float scale = std::fabs(m00);
  scale = std::fmax(scale, std::fabs(m10));
  scale = std::fmax(scale, std::fabs(m11));
  scale = std::fmax(scale, std::fabs(m20));
  scale = std::fmax(scale, std::fabs(m21));
  scale = std::fmax(scale, std::fabs(m22));
  if (scale == 0.0f) scale = 1.0f;
  m00 = m00 / scale; m10 = m10 / scale; m11 = m11 / scale;
  m20 = m20 / scale; m21 = m21 / scale; m22 = m22 / scale;

  float diag[3], sub[2];
  float Q[9]; // row-major
  const float tol = FLT_MIN;
  diag[0] = m00;
  const float v1norm2 = m20 * m20;
  if (v1norm2 <= tol) {
    diag[1] = m11;
    diag[2] = m22;
    sub[0] = m10;
    sub[1] = m21;
    Q[0] = 1; Q[1] = 0; Q[2] = 0; Q[3] = 0; Q[4] = 1; Q[5] = 0; Q[6] = 0; Q[7] = 0; Q[8] = 1;
  } else {
    const float beta = std::sqrt(m10 * m10 + v1norm2);
    const float invBeta = 1.0f / beta;
    const float m01 = m10 * invBeta;
    const float m02 = m20 * invBeta;
    const float q = 2.0f * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    sub[0] = beta;
    sub[1] = m21 - m01 * q;
    Q[0] = 1; Q[1] = 0; Q[2] = 0; Q[3] = 0; Q[4] = m01; Q[5] = m02; Q[6] = 0; Q[7] = m02; Q[8] = -m01;
  }

  const int n = 3;
  int end = n - 1, start = 0, iter = 0;
  const int maxIterations = 30;
  const float considerAsZero = FLT_MIN;
  const float precision_inv = 1.0f / FLT_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (std::fabs(sub[i]) < considerAsZero) {
        sub[i] = 0.0f;
      } else {
        const float scaled = precision_inv * sub[i];
        if (scaled * scaled <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.0f;
      }
    }
    while (end > 0 && sub[end - 1] == 0.0f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0f) start--;

    // tridiagonal_qr_step
    const float td = (diag[end - 1] - diag[end]) * 0.5f;
    const float e = sub[end - 1];
    float mu = diag[end];
    if (td == 0.0f) {
      mu = mu - std::fabs(e);
    } else if (e != 0.0f) {
      const float e2 = e * e;
      const float h = hypot_pos(td, e);
      if (e2 == 0.0f) mu = mu - e / ((td + (td > 0.0f ? h : -h)) / e);
      else mu = mu - e2 / (td + (td > 0.0f ? h : -h));
    }
    float x = diag[start] - mu;
    float z = sub[start];
    for (int k = start; k < end && z != 0.0f; ++k) {
      float c, s;
      make_givens(x, z, c, s);
      const float sdk = s * diag[k] + c * sub[k];
      const float dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) {
        z = -s * sub[k + 1];
        sub[k + 1] = c * sub[k + 1];
      }
      // Q = Q * G on columns k, k+1
      for (int r = 0; r < 3; ++r) {
        const float xi = Q[3 * r + k], yi = Q[3 * r + k + 1];
        Q[3 * r + k] = c * xi - s * yi;
        Q[3 * r + k + 1] = s * xi + c * yi;
      }
    }
  }
  // selection sort ascending, swapping eigenvector columns
  for (int i = 0; i < n - 1; ++i) {
    int k = 0;
    float mn = diag[i];
    for (int j = 1; j < n - i; ++j)
      if (diag[i + j] < mn) {
        mn = diag[i + j];
        k = j;
      }
    if (k > 0) {
      std::swap(diag[i], diag[k + i]);
      for (int r = 0; r < 3; ++r) std::swap(Q[3 * r + i], Q[3 * r + k + i]);
    }
  }
  for (int i = 0; i < 3; ++i) evals[i] = diag[i] * scale;
  for (int i = 0; i < 9; ++i) evecs[i] = Q[i];
}

// ---------------------------------------------------------------------------
// FeatureExtractor::extract, extraction.tpp:29-132
// ---------------------------------------------------------------------------
bool extract(const ExtractParams &P, const PointXYZf *scan, size_t n, size_t scan_idx,
             int num_threads, ExtractResult &out) {
  const size_t rows = (size_t)P.num_rows, cols = (size_t)P.num_columns;
  if (n != rows * cols) return false; // :141-145 throws
  const size_t pps = cols / P.num_sectors; // :33

  out = ExtractResult();
  out.valid_mask = valid_points(P, scan, true);        // :36
  out.curvature = curvature_of(P, scan, out.valid_mask); // :39

  // ---- planar features (:42-68) ----
  std::vector<uint8_t> used = out.valid_mask; // :43
  std::vector<uint32_t> order;                // per-sector sorted indices (rule R1)
  for (size_t r = 0; r < rows; ++r) {
    for (size_t s = 0; s < P.num_sectors; ++s) {
      const size_t start = r * cols + s * pps;
      const size_t end = (s == P.num_sectors - 1) ? (r + 1) * cols : start + pps;
      order.resize(end - start);
      for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)(start + i);
      // :57-58 std::sort by curvature; R1: ties by ascending index
      std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return out.curvature[a] < out.curvature[b];
      });
      // extract_planar, :332-358
      size_t count = 0;
      for (size_t i = 0; i < order.size(); ++i) {
        const size_t idx = order[i];
        if (used[idx] && (double)out.curvature[idx] < P.planar_threshold) {
          out.planar_indices.push_back((uint32_t)idx);
          for (size_t k = 0; k < P.neighbor_points; ++k) {
            used[idx + k] = 0;
            used[idx - k] = 0;
          }
          count++;
        }
        if (count > P.planar_feats_per_sector) break; // '>' => up to 51
      }
    }
  }

  // ---- point features (:70-96) ----
  out.point_valid_mask = valid_points(P, scan, false); // :72
  std::vector<uint8_t> pmask(rows * cols);
  for (size_t i = 0; i < rows * cols; ++i)
    pmask[i] = (used[i] == out.valid_mask[i]) && out.point_valid_mask[i]; // :76-80
  for (size_t r = 0; r < rows; ++r) {
    for (size_t s = 0; s < P.num_sectors; ++s) {
      const size_t start = r * cols + s * pps;
      const size_t end = (s == P.num_sectors - 1) ? (r + 1) * cols : start + pps;
      // extract_point, :360-399
      if (P.point_feats_per_sector == 0) continue;
      std::vector<size_t> unused;
      for (size_t idx = start; idx < end; ++idx)
        if (pmask[idx]) unused.push_back(idx);
      const size_t factor = 1 + unused.size() / P.point_feats_per_sector;
      size_t count = 0;
      for (size_t offset = 0; offset < factor; ++offset) {
        for (size_t u = offset; u < unused.size(); u += factor) {
          const size_t idx = unused[u];
          if (pmask[idx]) {
            out.point_indices.push_back((uint32_t)idx);
            for (size_t k = 0; k < P.neighbor_points; ++k) {
              pmask[idx + k] = 0;
              pmask[idx - k] = 0;
            }
            count++;
          }
          if (count > P.point_feats_per_sector) break; // leaves the inner loop only
        }
      }
    }
  }

  // ---- normals (:98-119), parallel over planar indices; rule R3 order ----
  const size_t np = out.planar_indices.size();
  out.planar_keep.assign(np, 0);
  out.closest_prev.assign(np, -1);
  out.closest_next.assign(np, -1);
  std::vector<float> normals(3 * np);
  auto work = [&](size_t a, size_t b) {
    for (size_t i = a; i < b; ++i) {
      out.planar_keep[i] = compute_normal(P, out.planar_indices[i], scan, out.valid_mask,
                                          &normals[3 * i], out.closest_prev[i],
                                          out.closest_next[i]);
    }
  };
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt == 1 || np < 64) {
    work(0, np);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, np * t / nt, np * (t + 1) / nt);
    for (auto &t : th) t.join();
  }
  out.planar.reserve(np);
  for (size_t i = 0; i < np; ++i) {
    if (!out.planar_keep[i]) continue;
    const PointXYZf &p = scan[out.planar_indices[i]];
    out.planar.emplace_back((double)p.x, (double)p.y, (double)p.z, (double)normals[3 * i],
                            (double)normals[3 * i + 1], (double)normals[3 * i + 2], scan_idx);
  }
  out.point.reserve(out.point_indices.size());
  for (uint32_t idx : out.point_indices) {
    const PointXYZf &p = scan[idx];
    out.point.emplace_back((double)p.x, (double)p.y, (double)p.z, scan_idx);
  }
  return true;
}

} // namespace form_oracle
