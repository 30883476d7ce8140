// C entry points of the synthetic scan generator (form/synth.hpp) for the
// Python tests and bench.py.
#include "form/synth.hpp"
#include "formgpu.h"

extern "C" {

/// sensor: 0 = OS1-64 (64x1024), 1 = OS0-128 (128x1024), 2 = VLP-16 (16x1800),
/// 3 = stress 128x2048.  Returns rows*cols, or 0 for an unknown sensor.
size_t formhost_synth_shape(int sensor, int *rows, int *cols) {
  using form::synth::SensorModel;
  SensorModel sm;
  switch (sensor) {
  case 0: sm = SensorModel::OS1_64(); break;
  case 1: sm = SensorModel::OS0_128(); break;
  case 2: sm = SensorModel::VLP_16(); break;
  case 3: sm = SensorModel::Stress_128x2048(); break;
  default: return 0;
  }
  if (rows) *rows = sm.rows;
  if (cols) *cols = sm.cols;
  return (size_t)sm.rows * sm.cols;
}

/// Fill `out` (rows*cols points) with scan k of `sequence_id`.
int formhost_synth_scan(int sensor, uint64_t sequence_id, uint64_t k, formgpu_point4f *out,
                        int threads) {
  using form::synth::SensorModel;
  SensorModel sm;
  switch (sensor) {
  case 0: sm = SensorModel::OS1_64(); break;
  case 1: sm = SensorModel::OS0_128(); break;
  case 2: sm = SensorModel::VLP_16(); break;
  case 3: sm = SensorModel::Stress_128x2048(); break;
  default: return 1;
  }
  form::synth::generate_scan(sm, sequence_id, (size_t)k, reinterpret_cast<form::PointXYZf *>(out),
                             threads);
  return 0;
}

/// Stress configuration (BASELINE.json configs[4]): 128x2048 scan k of tile `tile` of the tiled
/// hall, and its world pose.
void formhost_synth_stress_scan(uint64_t tile, uint64_t k, formgpu_point4f *out, int threads) {
  form::synth::generate_stress_scan(tile, (size_t)k, reinterpret_cast<form::PointXYZf *>(out), threads);
}
void formhost_synth_stress_pose(uint64_t tile, uint64_t k, formgpu_pose *out) {
  const form::Pose3 T = form::synth::stress_pose(tile, (size_t)k);
  for (int i = 0; i < 9; ++i) out->R[i] = T.R[i];
  for (int i = 0; i < 3; ++i) out->t[i] = T.t[i];
}

/// Ground-truth sensor pose of scan k.
void formhost_synth_gt_pose(uint64_t sequence_id, uint64_t k, formgpu_pose *out) {
  const form::Pose3 T = form::synth::gt_pose(sequence_id, (size_t)k);
  for (int i = 0; i < 9; ++i) out->R[i] = T.R[i];
  for (int i = 0; i < 3; ++i) out->t[i] = T.t[i];
}

} // extern "C"

// ---- small hooks so the Python tests can exercise the host-side SE(3) math ----
extern "C" {
void formhost_pose_expmap(const double xi[6], formgpu_pose *out) {
  const form::Pose3 T = form::Pose3::Expmap({xi[0], xi[1], xi[2], xi[3], xi[4], xi[5]});
  for (int i = 0; i < 9; ++i) out->R[i] = T.R[i];
  for (int i = 0; i < 3; ++i) out->t[i] = T.t[i];
}
void formhost_pose_logmap(const formgpu_pose *p, double xi[6]) {
  const form::Vec6 v = form::Pose3::Logmap(*reinterpret_cast<const form::Pose3 *>(p));
  for (int i = 0; i < 6; ++i) xi[i] = v[i];
}
void formhost_pose_logmap_derivative(const formgpu_pose *p, double J[36]) {
  const form::Mat6 M = form::Pose3::LogmapDerivative(*reinterpret_cast<const form::Pose3 *>(p));
  for (int i = 0; i < 36; ++i) J[i] = M[i];
}
void formhost_pose_compose(const formgpu_pose *a, const formgpu_pose *b, formgpu_pose *out) {
  const form::Pose3 T = *reinterpret_cast<const form::Pose3 *>(a) * *reinterpret_cast<const form::Pose3 *>(b);
  *reinterpret_cast<form::Pose3 *>(out) = T;
}
void formhost_pose_inverse(const formgpu_pose *a, formgpu_pose *out) {
  *reinterpret_cast<form::Pose3 *>(out) = reinterpret_cast<const form::Pose3 *>(a)->inverse();
}
void formhost_pose_rzryrx(double rx, double ry, double rz, const double t[3], formgpu_pose *out) {
  *reinterpret_cast<form::Pose3 *>(out) = form::Pose3::RzRyRx(rx, ry, rz, {t[0], t[1], t[2]});
}
void formhost_pose_normalized(const formgpu_pose *a, formgpu_pose *out) {
  *reinterpret_cast<form::Pose3 *>(out) = reinterpret_cast<const form::Pose3 *>(a)->normalized();
}
}
