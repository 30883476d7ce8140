// Host-side value types of the FORM hot path (B200 build).
//
// These mirror the reference's public PODs so that code written against
// form.hpp keeps compiling:
//   PointXYZf  <- /root/reference/form/utils.hpp:38-91      (16 B input point)
//   PointFeat  <- /root/reference/form/feature/features.hpp:31-86   (40 B)
//   PlanarFeat <- /root/reference/form/feature/features.hpp:89-163  (72 B)
// Field names, order and sizes are the contract (they cross the C-ABI as raw
// bytes, see include/formgpu.h).  No Eigen / GTSAM here: the vec3()/vec4()
// Eigen::Map accessors of the reference are replaced by plain members.
#pragma once

#include <cstddef>
#include <cstdint>

namespace form {

/// Input LiDAR point, float4-shaped. Row-major organised scans:
/// idx = row * num_columns + col (extraction.tpp:155).
struct PointXYZf {
  using Scalar = float;
  float x;
  float y;
  float z;
  float _ = 0.0f;

  PointXYZf() : x(0), y(0), z(0), _(0) {}
  PointXYZf(float x_, float y_, float z_) : x(x_), y(y_), z(z_), _(0) {}

  /// 4-lane squared norm with the packet reduction order the oracle fixes
  /// (SURVEY A.2): (x^2 + z^2) + (y^2 + _^2).
  [[nodiscard]] float squaredNorm() const noexcept {
    return (x * x + z * z) + (y * y + _ * _);
  }
  [[nodiscard]] constexpr bool operator==(const PointXYZf &o) const noexcept {
    return x == o.x && y == o.y && z == o.z;
  }
};
static_assert(sizeof(PointXYZf) == 16, "PointXYZf must be 16 bytes");

/// Point keypoint, scan-local frame, double precision.
struct PointFeat {
  double x;
  double y;
  double z;
  double _ = 0.0;
  size_t scan;

  PointFeat() = default;
  PointFeat(double x_, double y_, double z_, size_t scan_)
      : x(x_), y(y_), z(z_), _(0), scan(scan_) {}

  [[nodiscard]] constexpr bool operator==(const PointFeat &o) const noexcept {
    return x == o.x && y == o.y && z == o.z && scan == o.scan;
  }
};
static_assert(sizeof(PointFeat) == 40, "PointFeat must be 40 bytes");

/// Planar keypoint with unit normal, scan-local frame, double precision.
struct PlanarFeat {
  double x;
  double y;
  double z;
  double _ = 0.0;
  double nx;
  double ny;
  double nz;
  double _n = 0.0;
  size_t scan;

  PlanarFeat() = default;
  PlanarFeat(double x_, double y_, double z_, double nx_, double ny_, double nz_,
             size_t scan_)
      : x(x_), y(y_), z(z_), _(0), nx(nx_), ny(ny_), nz(nz_), _n(0), scan(scan_) {}

  [[nodiscard]] constexpr bool operator==(const PlanarFeat &o) const noexcept {
    return x == o.x && y == o.y && z == o.z && nx == o.nx && ny == o.ny &&
           nz == o.nz && scan == o.scan;
  }
};
static_assert(sizeof(PlanarFeat) == 72, "PlanarFeat must be 72 bytes");

using ScanIndex = size_t;

} // namespace form
