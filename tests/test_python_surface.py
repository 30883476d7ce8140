"""The Python surface of the reference (python/bindings.cpp:182-241, python/form/__init__.py)
as shipped here in python/form: class FORM with the evalio Pipeline methods, the 17-entry
parameter table with the reference's defaults, KeypointExtractionParams, extract_keypoints.
CPU part: the surface itself; GPU part: the pipeline produces what form::Estimator produces."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python"))

from form_b200 import _capi, synth  # noqa: E402


def test_module_surface_and_reference_defaults():
    import form

    assert form.FORM.name() == "form"                       # bindings.cpp:61
    assert form.FORM.url() == "https://github.com/rpl-cmu/form"
    d = form.FORM.default_params()                          # bindings.cpp:66-88
    assert d == {
        "neighbor_points": 5, "num_sectors": 6, "planar_threshold": 1.0, "planar_feats_per_sector": 50,
        "point_feats_per_sector": 3, "radius": 1.0, "min_points": 5, "max_dist_matching": 0.8,
        "new_pose_threshold": 1e-4, "max_num_rematches": 30, "disable_smoothing": False,
        "max_num_keyscans": 50, "max_num_recent_scans": 10, "max_steps_unused_keyscan": 10,
        "keyscan_match_ratio": 0.1, "max_dist_map": 0.1, "num_threads": 0}
    f = form.FORM()
    for m in ("pose", "map", "set_imu_params", "set_lidar_params", "set_imu_T_lidar", "initialize", "add_imu",
              "add_lidar", "set_params"):
        assert callable(getattr(f, m)), m                   # bindings.cpp:93-179
    assert f.set_params({"radius": 2.0, "point_feats_per_sector": 0, "unknown_key": 3}) == {"unknown_key": 3}
    p = form.KeypointExtractionParams()                     # bindings.cpp:196-212
    for name, default in (("neighbor_points", 5), ("num_sectors", 6), ("planar_feats_per_sector", 50),
                          ("planar_threshold", 1.0), ("point_feats_per_sector", 3), ("radius", 1.0),
                          ("min_points", 5), ("min_norm_squared", 1.0), ("max_norm_squared", 1e4),
                          ("num_rows", 64), ("num_columns", 1024)):
        assert getattr(p, name) == default, name
        setattr(p, name, type(default)(2))
        assert getattr(p, name) == 2
    pose = f.pose()
    assert (pose.rot.qw, pose.trans) == (1.0, [0.0, 0.0, 0.0])
    assert f.map() == {"planar": [], "point": []}


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_error_not_fallback():
    import form

    f = form.FORM()
    with pytest.raises(RuntimeError, match="no CUDA device"):
        f.initialize()


@pytest.mark.gpu
def test_form_pipeline_matches_estimator():
    import form
    from form_b200.pipeline import Estimator

    sensor, n_scans = "vlp-16", 6
    rows, cols = synth.shape(sensor)
    scans = [synth.scan(sensor, 0, k) for k in range(n_scans)]
    lp = form.LidarParams(num_rows=rows, num_columns=cols, min_range=1.0, max_range=100.0)
    pipe = form.FORM()
    assert pipe.set_params(form.FORM.default_params()) == {}
    pipe.set_lidar_params(lp)
    pipe.set_imu_T_lidar(form.SE3.identity())
    pipe.initialize()
    est = Estimator(_capi.default_est_params(rows, cols))
    for k, scan in enumerate(scans):
        xyz = np.stack([scan["x"], scan["y"], scan["z"]], 1)
        out = pipe.add_lidar(form.LidarMeasurement.from_xyz(form.Stamp.from_sec(0.1 * k), xyz))
        pl, pt = est.register_scan(scan)
        assert len(out["planar"]) == len(pl) and len(out["point"]) == len(pt)
        assert np.array_equal([p.x for p in out["planar"]], pl["x"])
        assert all(p.col == k for p in out["planar"][:5])            # col = scan index (bindings.cpp:39)
        a, b = pipe.pose(), est.pose()
        assert np.allclose(a.trans, b["t"], atol=1e-9)
    m = pipe.map()
    epl, ept = est.map()
    assert len(m["planar"]) == len(epl) and len(m["point"]) == len(ept)
    # extract_keypoints helper (bindings.cpp:214-240)
    kp = form.KeypointExtractionParams()
    kp.num_rows, kp.num_columns = rows, cols
    xyz = np.stack([scans[0]["x"], scans[0]["y"], scans[0]["z"]], 1).astype(np.float64)
    planar_pts, normals, point_pts = form.extract_keypoints([list(r) for r in xyz], kp, lp)
    with Estimator(_capi.default_est_params(rows, cols)) as e0:
        pl0, pt0 = e0.register_scan(scans[0])
    assert np.array_equal(np.array(planar_pts)[:, 0], pl0["x"]) and np.array_equal(np.array(normals)[:, 2], pl0["nz"])
    assert len(point_pts) == len(pt0)
