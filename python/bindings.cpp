// form._core - the Python surface of FORM over the B200 hot path.
//
// Mirrors /root/reference/python/bindings.cpp:22-241: class FORM (an evalio Pipeline) with
// name() / url() / default_params() / set_params(), pose(), map(), set_imu_params(),
// set_lidar_params(), set_imu_T_lidar(), initialize(), add_imu(), add_lidar();
// KeypointExtractionParams (read-write fields); extract_keypoints(points, params,
// lidar_params) -> (planar_points, normals, point_points).  The reference binds with
// nanobind against evalio's C++ types; neither exists in this image, so the module is
// built with pybind11 and carries minimal stand-ins for the evalio value types it touches
// (same field names: SE3{rot, trans}, SO3{qx,qy,qz,qw}, Point{x,y,z,intensity,t,row,col},
// LidarMeasurement{stamp, points}, LidarParams{num_rows, num_columns, min_range, max_range,
// rate}).  form::Estimator is the facade of form_b200/host/form/form.hpp, i.e. the CUDA
// library behind FORM's own API; there is no CPU fallback.
#include "form/form.hpp"

#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <array>
#include <cmath>
#include <map>
#include <string>
#include <variant>
#include <vector>

namespace py = pybind11;

namespace evalio { // stand-ins for the evalio types of the reference binding
struct Duration {
  double sec = 0.0;
  static Duration from_sec(double s) { return Duration{s}; }
  double to_sec() const { return sec; }
};
struct Stamp {
  double sec = 0.0;
  static Stamp from_sec(double s) { return Stamp{s}; }
  double to_sec() const { return sec; }
};
struct SO3 {
  double qx = 0, qy = 0, qz = 0, qw = 1;
  static SO3 identity() { return SO3(); }
};
struct SE3 {
  SO3 rot;
  std::array<double, 3> trans{0, 0, 0};
  static SE3 identity() { return SE3(); }
};
struct Point {
  double x = 0, y = 0, z = 0, intensity = 0;
  Duration t;
  uint8_t row = 0;
  uint16_t col = 0;
};
struct LidarMeasurement {
  Stamp stamp;
  std::vector<Point> points;
};
struct LidarParams {
  int num_rows = 64, num_columns = 1024;
  double min_range = 1.0, max_range = 100.0, rate = 10.0;
  Duration delta_time() const { return Duration::from_sec(1.0 / rate); }
};
struct ImuParams {};
struct ImuMeasurement {};
using Param = std::variant<bool, int, double, std::string>;
} // namespace evalio

namespace {

form::Pose3 pose_to_form(const evalio::SE3 &p) { // bindings.cpp:22-24
  const double x = p.rot.qx, y = p.rot.qy, z = p.rot.qz, w = p.rot.qw;
  form::Pose3 T;
  const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w),     2 * (x * z + y * w),
                       2 * (x * y + z * w),     1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                       2 * (x * z - y * w),     2 * (y * z + x * w),     1 - 2 * (x * x + y * y)};
  for (int i = 0; i < 9; ++i) T.R[i] = R[i];
  for (int i = 0; i < 3; ++i) T.t[i] = p.trans[i];
  return T;
}

evalio::SE3 pose_to_evalio(const form::Pose3 &T) { // bindings.cpp:26-29 (Rot3::toQuaternion)
  const double *R = T.R.data();
  evalio::SE3 out;
  const double tr = R[0] + R[4] + R[8];
  double qw, qx, qy, qz;
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    qw = 0.25 * s; qx = (R[7] - R[5]) / s; qy = (R[2] - R[6]) / s; qz = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    qw = (R[7] - R[5]) / s; qx = 0.25 * s; qy = (R[1] + R[3]) / s; qz = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    qw = (R[2] - R[6]) / s; qx = (R[1] + R[3]) / s; qy = 0.25 * s; qz = (R[5] + R[7]) / s;
  } else {
    const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    qw = (R[3] - R[1]) / s; qx = (R[2] + R[6]) / s; qy = (R[5] + R[7]) / s; qz = 0.25 * s;
  }
  out.rot = evalio::SO3{qx, qy, qz, qw};
  out.trans = {T.t[0], T.t[1], T.t[2]};
  return out;
}

template <typename Feat> evalio::Point point_to_evalio(const Feat &f) { // bindings.cpp:31-41
  evalio::Point p;
  p.x = f.x;
  p.y = f.y;
  p.z = f.z;
  p.intensity = 0.0;
  p.t = evalio::Duration::from_sec(0);
  p.row = 0;
  p.col = static_cast<uint16_t>(f.scan);
  return p;
}

using PointMap = std::map<std::string, std::vector<evalio::Point>>;

// The evalio Pipeline protocol, bindings.cpp:48-180.
class FORM {
public:
  FORM() = default;

  static std::string name() { return "form"; }
  static std::string url() { return "https://github.com/rpl-cmu/form"; }

  // EVALIO_SETUP_PARAMS, bindings.cpp:66-88
  static std::map<std::string, evalio::Param> default_params() {
    return {
        {"neighbor_points", 5},        {"num_sectors", 6},          {"planar_threshold", 1.0},
        {"planar_feats_per_sector", 50}, {"point_feats_per_sector", 3}, {"radius", 1.0},
        {"min_points", 5},             {"max_dist_matching", 0.8},  {"new_pose_threshold", 1e-4},
        {"max_num_rematches", 30},     {"disable_smoothing", false}, {"max_num_keyscans", 50},
        {"max_num_recent_scans", 10},  {"max_steps_unused_keyscan", 10}, {"keyscan_match_ratio", 0.1},
        {"max_dist_map", 0.1},         {"num_threads", 0},
    };
  }

  /// evalio::Pipeline::set_params: applies the known keys, returns the unused ones.
  std::map<std::string, evalio::Param> set_params(const std::map<std::string, evalio::Param> &params) {
    std::map<std::string, evalio::Param> unused;
    auto as_d = [](const evalio::Param &p) {
      return std::holds_alternative<double>(p) ? std::get<double>(p)
             : std::holds_alternative<int>(p)  ? (double)std::get<int>(p)
                                               : (double)std::get<bool>(p);
    };
    auto as_i = [&](const evalio::Param &p) { return (long long)as_d(p); };
    for (const auto &[k, v] : params) {
      if (std::holds_alternative<std::string>(v)) {
        unused.insert({k, v});
        continue;
      }
      auto &P = params_;
      if (k == "neighbor_points") P.extraction.neighbor_points = (size_t)as_i(v);
      else if (k == "num_sectors") P.extraction.num_sectors = (size_t)as_i(v);
      else if (k == "planar_threshold") P.extraction.planar_threshold = as_d(v);
      else if (k == "planar_feats_per_sector") P.extraction.planar_feats_per_sector = (size_t)as_i(v);
      else if (k == "point_feats_per_sector") P.extraction.point_feats_per_sector = (size_t)as_i(v);
      else if (k == "radius") P.extraction.radius = as_d(v);
      else if (k == "min_points") P.extraction.min_points = (size_t)as_i(v);
      else if (k == "max_dist_matching") P.matcher.max_dist_matching = as_d(v);
      else if (k == "new_pose_threshold") P.matcher.new_pose_threshold = as_d(v);
      else if (k == "max_num_rematches") P.matcher.max_num_rematches = (size_t)as_i(v);
      else if (k == "disable_smoothing") P.constraints.disable_smoothing = as_d(v) != 0.0;
      else if (k == "max_num_keyscans") P.scans.max_num_keyscans = (int)as_i(v);
      else if (k == "max_num_recent_scans") P.scans.max_num_recent_scans = (size_t)as_i(v);
      else if (k == "max_steps_unused_keyscan") P.scans.max_steps_unused_keyscan = (int)as_i(v);
      else if (k == "keyscan_match_ratio") P.scans.keyscan_match_ratio = as_d(v);
      else if (k == "max_dist_map") P.map.min_dist_map = as_d(v); // bindings.cpp:85: name vs field
      else if (k == "num_threads") P.num_threads = (size_t)as_i(v);
      else unused.insert({k, v});
    }
    return unused;
  }

  evalio::SE3 pose() const { return current_pose_; } // bindings.cpp:93

  PointMap map() { // bindings.cpp:96-119: every stored keypoint in the world frame
    PointMap points = {{"planar", {}}, {"point", {}}};
    if (!estimator_) return points;
    std::vector<form::PlanarFeat> pl;
    std::vector<form::PointFeat> pt;
    estimator_->m_keypoint_map.world_keypoints(estimator_->m_constraints.get_values(), pl, pt);
    for (const auto &p : pl) points["planar"].push_back(point_to_evalio(p));
    for (const auto &p : pt) points["point"].push_back(point_to_evalio(p));
    return points;
  }

  void set_imu_params(const evalio::ImuParams &) {} // bindings.cpp:123
  void set_lidar_params(const evalio::LidarParams &p) { // bindings.cpp:126-132
    params_.extraction.min_norm_squared = p.min_range * p.min_range;
    params_.extraction.max_norm_squared = p.max_range * p.max_range;
    params_.extraction.num_columns = p.num_columns;
    params_.extraction.num_rows = p.num_rows;
    delta_time_ = p.delta_time();
  }
  void set_imu_T_lidar(const evalio::SE3 &T) { lidar_T_imu_ = pose_to_form(T).inverse(); } // :135-137

  void initialize() { estimator_ = std::make_unique<form::Estimator>(params_); } // bindings.cpp:141
  void add_imu(const evalio::ImuMeasurement &) {}                                // bindings.cpp:144

  PointMap add_lidar(const evalio::LidarMeasurement &mm) { // bindings.cpp:147-179
    if (!estimator_) throw std::runtime_error("FORM: initialize() has not been called");
    std::vector<form::PointXYZf> scan;
    scan.reserve(mm.points.size());
    for (const auto &p : mm.points) scan.emplace_back((float)p.x, (float)p.y, (float)p.z);
    auto [planar_kp, point_kp] = estimator_->register_scan(scan);
    if (planar_kp.empty() && point_kp.empty() && !estimator_->last_error().empty())
      throw std::runtime_error("FORM: " + estimator_->last_error());
    current_pose_ = pose_to_evalio(estimator_->current_lidar_estimate() * lidar_T_imu_);
    PointMap points = {{"planar", {}}, {"point", {}}};
    for (const auto &p : planar_kp) points["planar"].push_back(point_to_evalio(p));
    for (const auto &p : point_kp) points["point"].push_back(point_to_evalio(p));
    return points;
  }

private:
  std::unique_ptr<form::Estimator> estimator_;
  form::Estimator::Params params_;
  form::Pose3 lidar_T_imu_;
  evalio::Duration delta_time_;
  evalio::SE3 current_pose_ = evalio::SE3::identity();
};

} // namespace

PYBIND11_MODULE(_core, m) {
  m.doc() = "FORM pipeline over the B200 hot path (mirror of the reference's form._core)";

  py::class_<evalio::Duration>(m, "Duration")
      .def(py::init<>())
      .def_static("from_sec", &evalio::Duration::from_sec)
      .def("to_sec", &evalio::Duration::to_sec);
  py::class_<evalio::Stamp>(m, "Stamp")
      .def(py::init<>())
      .def_static("from_sec", &evalio::Stamp::from_sec)
      .def("to_sec", &evalio::Stamp::to_sec);
  py::class_<evalio::SO3>(m, "SO3")
      .def(py::init<>())
      .def(py::init([](double qx, double qy, double qz, double qw) { return evalio::SO3{qx, qy, qz, qw}; }),
           py::arg("qx"), py::arg("qy"), py::arg("qz"), py::arg("qw"))
      .def_static("identity", &evalio::SO3::identity)
      .def_readwrite("qx", &evalio::SO3::qx)
      .def_readwrite("qy", &evalio::SO3::qy)
      .def_readwrite("qz", &evalio::SO3::qz)
      .def_readwrite("qw", &evalio::SO3::qw);
  py::class_<evalio::SE3>(m, "SE3")
      .def(py::init<>())
      .def(py::init([](const evalio::SO3 &r, const std::array<double, 3> &t) { return evalio::SE3{r, t}; }),
           py::arg("rot"), py::arg("trans"))
      .def_static("identity", &evalio::SE3::identity)
      .def_readwrite("rot", &evalio::SE3::rot)
      .def_readwrite("trans", &evalio::SE3::trans);
  py::class_<evalio::Point>(m, "Point")
      .def(py::init<>())
      .def(py::init([](double x, double y, double z) {
             evalio::Point p;
             p.x = x; p.y = y; p.z = z;
             return p;
           }),
           py::arg("x"), py::arg("y"), py::arg("z"))
      .def_readwrite("x", &evalio::Point::x)
      .def_readwrite("y", &evalio::Point::y)
      .def_readwrite("z", &evalio::Point::z)
      .def_readwrite("intensity", &evalio::Point::intensity)
      .def_readwrite("t", &evalio::Point::t)
      .def_readwrite("row", &evalio::Point::row)
      .def_readwrite("col", &evalio::Point::col);
  py::class_<evalio::LidarMeasurement>(m, "LidarMeasurement")
      .def(py::init<>())
      .def(py::init([](const evalio::Stamp &s, std::vector<evalio::Point> pts) {
             return evalio::LidarMeasurement{s, std::move(pts)};
           }),
           py::arg("stamp"), py::arg("points"))
      .def_static(
          "from_xyz",
          [](const evalio::Stamp &s, py::buffer xyz) { // (n, >=3) float32/float64 array, row-major scan
            py::buffer_info info = xyz.request();
            if (info.ndim != 2 || info.shape[1] < 3) throw std::runtime_error("expected an (n, >=3) array");
            evalio::LidarMeasurement mm;
            mm.stamp = s;
            mm.points.resize((size_t)info.shape[0]);
            const bool f32 = info.format == py::format_descriptor<float>::format();
            const bool f64 = info.format == py::format_descriptor<double>::format();
            if (!f32 && !f64) throw std::runtime_error("expected float32 or float64");
            for (py::ssize_t i = 0; i < info.shape[0]; ++i) {
              const char *row = static_cast<const char *>(info.ptr) + i * info.strides[0];
              for (int c = 0; c < 3; ++c) {
                const char *e = row + c * info.strides[1];
                const double v = f32 ? (double)*reinterpret_cast<const float *>(e) : *reinterpret_cast<const double *>(e);
                (c == 0 ? mm.points[i].x : c == 1 ? mm.points[i].y : mm.points[i].z) = v;
              }
            }
            return mm;
          },
          py::arg("stamp"), py::arg("xyz"))
      .def_readwrite("stamp", &evalio::LidarMeasurement::stamp)
      .def_readwrite("points", &evalio::LidarMeasurement::points);
  py::class_<evalio::LidarParams>(m, "LidarParams")
      .def(py::init<>())
      .def(py::init([](int rows, int cols, double min_range, double max_range, double rate) {
             return evalio::LidarParams{rows, cols, min_range, max_range, rate};
           }),
           py::arg("num_rows"), py::arg("num_columns"), py::arg("min_range"), py::arg("max_range"),
           py::arg("rate") = 10.0)
      .def_readwrite("num_rows", &evalio::LidarParams::num_rows)
      .def_readwrite("num_columns", &evalio::LidarParams::num_columns)
      .def_readwrite("min_range", &evalio::LidarParams::min_range)
      .def_readwrite("max_range", &evalio::LidarParams::max_range)
      .def_readwrite("rate", &evalio::LidarParams::rate)
      .def("delta_time", &evalio::LidarParams::delta_time);
  py::class_<evalio::ImuParams>(m, "ImuParams").def(py::init<>());
  py::class_<evalio::ImuMeasurement>(m, "ImuMeasurement").def(py::init<>());

  // bindings.cpp:189-193 (+ the methods the reference inherits from evalio::Pipeline)
  py::class_<FORM>(m, "FORM")
      .def(py::init<>())
      .def_static("name", &FORM::name)
      .def_static("url", &FORM::url)
      .def_static("default_params", &FORM::default_params)
      .def("set_params", &FORM::set_params)
      .def("pose", &FORM::pose)
      .def("map", &FORM::map)
      .def("set_imu_params", &FORM::set_imu_params)
      .def("set_lidar_params", &FORM::set_lidar_params)
      .def("set_imu_T_lidar", &FORM::set_imu_T_lidar)
      .def("initialize", &FORM::initialize)
      .def("add_imu", &FORM::add_imu)
      .def("add_lidar", &FORM::add_lidar, py::call_guard<py::gil_scoped_release>());

  // bindings.cpp:196-212
  using EP = form::FeatureExtractor::Params;
  py::class_<EP>(m, "KeypointExtractionParams")
      .def(py::init<>())
      .def_readwrite("neighbor_points", &EP::neighbor_points)
      .def_readwrite("num_sectors", &EP::num_sectors)
      .def_readwrite("planar_feats_per_sector", &EP::planar_feats_per_sector)
      .def_readwrite("planar_threshold", &EP::planar_threshold)
      .def_readwrite("point_feats_per_sector", &EP::point_feats_per_sector)
      .def_readwrite("radius", &EP::radius)
      .def_readwrite("min_points", &EP::min_points)
      .def_readwrite("min_norm_squared", &EP::min_norm_squared)
      .def_readwrite("max_norm_squared", &EP::max_norm_squared)
      .def_readwrite("num_rows", &EP::num_rows)
      .def_readwrite("num_columns", &EP::num_columns);

  // bindings.cpp:214-240
  m.def(
      "extract_keypoints",
      [](const std::vector<std::array<double, 3>> &points, const EP &params, evalio::LidarParams &) {
        std::vector<form::PointXYZf> pts;
        pts.reserve(points.size());
        for (const auto &p : points) pts.emplace_back((float)p[0], (float)p[1], (float)p[2]);
        auto [planar, point] = form::FeatureExtractor(params, 0).extract(pts, 0);
        std::vector<std::array<double, 3>> planar_points, normals, point_points;
        for (const auto &k : planar) {
          planar_points.push_back({k.x, k.y, k.z});
          normals.push_back({k.nx, k.ny, k.nz});
        }
        for (const auto &k : point) point_points.push_back({k.x, k.y, k.z});
        return std::make_tuple(planar_points, normals, point_points);
      },
      py::arg("points"), py::arg("params"), py::arg("lidar_params"));
}
