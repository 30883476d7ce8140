// Seeded synthetic organised LiDAR scans (SURVEY.md 8(d) "Synthetic inputs").
//
// The reference ships no data and no generator; its inputs are organised
// row-major scans of 16-byte points handed to Estimator::register_scan
// (/root/reference/form/form.hpp:82-83, python/bindings.cpp:150-156).  This
// header produces scans of that exact layout by ray-casting a closed room:
//   * box room 40 x 30 x 6 m (floor = ground plane), 12 vertical cylinders
//     r = 0.3 m, 6 interior wall slabs,
//   * row r elevation linearly spaced over the sensor FOV, column c azimuth
//     2*pi*c/cols, idx = r*cols + c,
//   * trajectory: 10 Hz, 1 m/s along +x, yaw 0.2 sin(0.05 k), z 0.05 sin(0.1 k),
//   * range noise N(0, 0.01 m) along the ray, dropout p = 0.01 -> (0,0,0),
//     returns beyond 100 m -> (0,0,0); never NaN/Inf,
//   * RNG xoshiro256** seeded through splitmix64 with
//     0xF0A30000 + sequence_id*1000003 + k, one stream per scan.
// Stress configuration (BASELINE.json configs[4]: 128x2048 scans against a ~1 M-voxel map): the
// world is a TILED hall - the same 160 x 160 x 40 m hall repeated every 400 m in x and y, each
// tile with its own jittered pillars and wall slabs (World::hall(tile)).  A 128x2048 scan of one
// tile leaves ~58 k keypoints spread over tens of thousands of 0.8 m voxels; seeding the map with
// one scan per tile (stress_pose(tile, k) = tile offset + a short local trajectory) reaches a
// million occupied voxels with a few dozen map scans.
#pragma once

#include "form/pose3.hpp"
#include "form/types.hpp"

#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace form {
namespace synth {

struct SensorModel {
  int rows = 64;
  int cols = 1024;
  double fov_up_deg = 16.6;
  double fov_down_deg = -16.6;
  double max_range = 100.0;
  double noise_sigma = 0.01;
  double dropout = 0.01;

  static SensorModel OS1_64() { return {64, 1024, 16.6, -16.6, 100.0, 0.01, 0.01}; }
  static SensorModel OS0_128() { return {128, 1024, 45.0, -45.0, 100.0, 0.01, 0.01}; }
  static SensorModel VLP_16() { return {16, 1800, 15.0, -15.0, 100.0, 0.01, 0.01}; }
  static SensorModel Stress_128x2048() { return {128, 2048, 45.0, -45.0, 100.0, 0.01, 0.01}; }
};

struct Rng {
  uint64_t s[4];
  static uint64_t splitmix64(uint64_t &x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) {
    for (auto &v : s) v = splitmix64(seed);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double gauss() { // Box-Muller, one value per call
    double u1 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    const double u2 = uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

struct Box {
  double lo[3], hi[3];
};
struct Cyl {
  double cx, cy, r, z0, z1;
};

struct World {
  Box room{{-20, -15, -1.5}, {20, 15, 4.5}};
  std::vector<Box> slabs;
  std::vector<Cyl> cyls;
  World() {
    // 6 interior wall slabs (thin boxes standing on the floor)
    slabs = {{{-12.0, 5.0, -1.5}, {-4.0, 5.2, 2.0}},   {{2.0, -9.2, -1.5}, {10.0, -9.0, 2.5}},
             {{-6.0, -6.2, -1.5}, {-5.8, 1.0, 1.8}},   {{8.0, 3.0, -1.5}, {8.2, 11.0, 3.0}},
             {{-16.0, -10.0, -1.5}, {-15.8, -2.0, 2.2}}, {{13.0, -4.0, -1.5}, {17.0, -3.8, 2.0}}};
    // 12 vertical cylinders r = 0.3
    const double cxy[12][2] = {{-15, 8},  {-10, -8}, {-7, 10}, {-3, -4}, {0, 6},   {3, -11},
                               {5, 2},    {7, -6},   {11, 9},  {12, -1}, {15, 6},  {17, -10}};
    for (auto &c : cxy) cyls.push_back({c[0], c[1], 0.3, -1.5, 4.5});
  }

  /// Tile `tile` of the stress world: a 160 x 160 x 40 m hall with a jittered 6 x 6 grid of
  /// pillars (r = 1.5 m) and 8 jittered wall slabs, in tile-local coordinates.
  static World hall(uint64_t tile) {
    World w;
    w.room = Box{{-80, -80, -20}, {80, 80, 20}};
    w.slabs.clear();
    w.cyls.clear();
    uint64_t seed = 0x5EED7113ull + tile * 7919ull;
    Rng rng(seed);
    for (int gx = 0; gx < 6; ++gx)
      for (int gy = 0; gy < 6; ++gy) {
        const double cx = -62.5 + 25.0 * gx + 6.0 * (rng.uniform() - 0.5);
        const double cy = -62.5 + 25.0 * gy + 6.0 * (rng.uniform() - 0.5);
        if (cx * cx + cy * cy < 100.0) continue; // keep the sensor's neighbourhood free
        w.cyls.push_back({cx, cy, 1.5, -20.0, 20.0});
      }
    for (int k = 0; k < 8; ++k) {
      const double ang = 0.785398163397448 * k + 0.3 * (rng.uniform() - 0.5);
      const double r = 35.0 + 25.0 * rng.uniform();
      const double cx = r * std::cos(ang), cy = r * std::sin(ang);
      const double half = 6.0 + 6.0 * rng.uniform(), top = -20.0 + 12.0 + 20.0 * rng.uniform();
      if (k % 2 == 0) w.slabs.push_back({{cx - half, cy - 0.2, -20.0}, {cx + half, cy + 0.2, top}});
      else w.slabs.push_back({{cx - 0.2, cy - half, -20.0}, {cx + 0.2, cy + half, top}});
    }
    return w;
  }

  // distance along the ray to the inside of the room box (origin is inside)
  static double exit_box(const Box &b, const double o[3], const double d[3]) {
    double tmax = 1e300;
    for (int a = 0; a < 3; ++a) {
      if (d[a] > 1e-15) tmax = std::fmin(tmax, (b.hi[a] - o[a]) / d[a]);
      else if (d[a] < -1e-15) tmax = std::fmin(tmax, (b.lo[a] - o[a]) / d[a]);
    }
    return tmax;
  }
  static double enter_box(const Box &b, const double o[3], const double d[3]) {
    double t0 = 0.0, t1 = 1e300;
    for (int a = 0; a < 3; ++a) {
      if (std::fabs(d[a]) < 1e-15) {
        if (o[a] < b.lo[a] || o[a] > b.hi[a]) return 1e300;
      } else {
        double ta = (b.lo[a] - o[a]) / d[a], tb = (b.hi[a] - o[a]) / d[a];
        if (ta > tb) { const double tmp = ta; ta = tb; tb = tmp; }
        t0 = std::fmax(t0, ta);
        t1 = std::fmin(t1, tb);
        if (t0 > t1) return 1e300;
      }
    }
    return t0 > 0.0 ? t0 : 1e300;
  }
  static double hit_cyl(const Cyl &c, const double o[3], const double d[3]) {
    const double ox = o[0] - c.cx, oy = o[1] - c.cy;
    const double A = d[0] * d[0] + d[1] * d[1];
    if (A < 1e-18) return 1e300;
    const double B = ox * d[0] + oy * d[1];
    const double C = ox * ox + oy * oy - c.r * c.r;
    const double disc = B * B - A * C;
    if (disc < 0) return 1e300;
    const double t = (-B - std::sqrt(disc)) / A;
    if (t <= 0) return 1e300;
    const double z = o[2] + t * d[2];
    if (z < c.z0 || z > c.z1) return 1e300;
    return t;
  }
  double cast(const double o[3], const double d[3]) const {
    double t = exit_box(room, o, d);
    for (const auto &s : slabs) t = std::fmin(t, enter_box(s, o, d));
    for (const auto &c : cyls) t = std::fmin(t, hit_cyl(c, o, d));
    return t;
  }
};

/// Ground-truth sensor pose of scan k of a sequence.
inline Pose3 gt_pose(uint64_t sequence_id, size_t k) {
  const double kk = (double)k;
  const double yaw = 0.2 * std::sin(0.05 * kk);
  const double x = -10.0 + 0.1 * kk;
  const double y = -2.0 + 0.45 * (double)(sequence_id % 8);
  const double z = 0.05 * std::sin(0.1 * kk);
  const double c = std::cos(yaw), s = std::sin(yaw);
  return Pose3({c, -s, 0, s, c, 0, 0, 0, 1}, {x, y, z});
}

inline uint64_t scan_seed(uint64_t sequence_id, size_t k) {
  return 0xF0A30000ull + sequence_id * 1000003ull + (uint64_t)k;
}

/// Fill `out` (rows*cols points, row-major) with the scan of `world` seen from pose T (in the
/// world's own frame); `seed` drives the range noise and the dropouts.
inline void generate_scan_in(const World &world, const SensorModel &sm, const Pose3 &T, uint64_t seed,
                             PointXYZf *out, int num_threads = 0);

/// Fill `out` (rows*cols points, row-major) with scan k of sequence_id.
inline void generate_scan(const SensorModel &sm, uint64_t sequence_id, size_t k,
                          PointXYZf *out, int num_threads = 0) {
  static const World world;
  generate_scan_in(world, sm, gt_pose(sequence_id, k), scan_seed(sequence_id, k), out, num_threads);
}

// ---- stress configuration: tiled hall ----
/// Sensor pose of scan k inside its tile (tile-local frame).
inline Pose3 stress_local_pose(uint64_t tile, size_t k) {
  const double kk = (double)k;
  const double yaw = 0.15 * std::sin(0.07 * kk) + 0.05 * (double)(tile % 7);
  const double c = std::cos(yaw), s = std::sin(yaw);
  return Pose3({c, -s, 0, s, c, 0, 0, 0, 1},
               {-3.0 + 0.1 * kk, 0.35 * (double)(tile % 5), 0.04 * std::sin(0.1 * kk)});
}
/// World pose of scan k of tile `tile`: tiles repeat every 400 m on an 8-wide grid.
inline Pose3 stress_pose(uint64_t tile, size_t k) {
  Pose3 T = stress_local_pose(tile, k);
  T.t[0] += 400.0 * (double)(tile % 8);
  T.t[1] += 400.0 * (double)(tile / 8);
  return T;
}
inline void generate_stress_scan(uint64_t tile, size_t k, PointXYZf *out, int num_threads = 0) {
  const World world = World::hall(tile);
  generate_scan_in(world, SensorModel::Stress_128x2048(), stress_local_pose(tile, k),
                   0x57E55000ull + tile * 1000003ull + (uint64_t)k, out, num_threads);
}

inline void generate_scan_in(const World &world, const SensorModel &sm, const Pose3 &T, uint64_t seed,
                             PointXYZf *out, int num_threads) {
  const double o[3] = {T.t[0], T.t[1], T.t[2]};
  const size_t n = (size_t)sm.rows * sm.cols;
  // the noise / dropout stream is consumed in index order so the scan does not
  // depend on the thread count: draw it first, serially.
  std::vector<float> noise(n);
  std::vector<uint8_t> drop(n);
  {
    Rng rng(seed);
    for (size_t i = 0; i < n; ++i) {
      noise[i] = (float)(sm.noise_sigma * rng.gauss());
      drop[i] = rng.uniform() < sm.dropout;
    }
  }
  const double DEG = 3.14159265358979323846 / 180.0;
  auto work = [&](int r0, int r1) {
    for (int r = r0; r < r1; ++r) {
      const double el =
          (sm.rows > 1)
              ? (sm.fov_down_deg + (sm.fov_up_deg - sm.fov_down_deg) * r / (sm.rows - 1)) * DEG
              : 0.0;
      const double ce = std::cos(el), se = std::sin(el);
      for (int c = 0; c < sm.cols; ++c) {
        const size_t idx = (size_t)r * sm.cols + c;
        const double az = 6.283185307179586 * c / sm.cols;
        const Vec3 dl{ce * std::cos(az), ce * std::sin(az), se};
        const Vec3 dw = T.rotate(dl);
        const double d[3] = {dw[0], dw[1], dw[2]};
        const double range = world.cast(o, d) + (double)noise[idx];
        if (drop[idx] || !(range < sm.max_range) || !(range > 0.0)) {
          out[idx] = PointXYZf(0.f, 0.f, 0.f);
        } else {
          out[idx] = PointXYZf((float)(dl[0] * range), (float)(dl[1] * range),
                               (float)(dl[2] * range));
        }
      }
    }
  };
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > sm.rows) nt = sm.rows;
  if (nt == 1) {
    work(0, sm.rows);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nt; ++i)
      th.emplace_back(work, sm.rows * i / nt, sm.rows * (i + 1) / nt);
    for (auto &t : th) t.join();
  }
}

} // namespace synth
} // namespace form
