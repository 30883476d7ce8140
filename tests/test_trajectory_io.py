"""Trajectory CSV in evalio's layout (form_b200/trajectory.py): round trip and conventions."""
import io

import numpy as np

from form_b200 import _capi, trajectory
from test_oracle_map_factor import expmap


def test_quaternion_round_trip_and_sign():
    rng = np.random.default_rng(0)
    for _ in range(200):
        T = expmap(rng.normal(size=6) * 2.0)
        R = T["R"].reshape(3, 3)
        q = trajectory.quat_from_rotation(R)
        assert abs(np.linalg.norm(q) - 1.0) < 1e-14 and q[3] >= 0.0
        assert np.max(np.abs(trajectory.rotation_from_quat(q) - R)) < 1e-13
    # half-turns (trace = -1) take the other branch of Shepperd's method
    for axis in range(3):
        R = -np.eye(3)
        R[axis, axis] = 1.0
        assert np.max(np.abs(trajectory.rotation_from_quat(trajectory.quat_from_rotation(R)) - R)) < 1e-13


def test_csv_round_trip_keeps_every_bit():
    rng = np.random.default_rng(1)
    poses = np.zeros(50, dtype=_capi.POSE)
    for k in range(50):
        T = expmap(rng.normal(size=6))
        poses[k]["R"], poses[k]["t"] = T["R"], T["t"] * 37.0
    stamps = 1.7e9 + 0.1 * np.arange(50)
    text = trajectory.to_string(stamps, poses, params={"max_dist_matching": 0.8, "disable_smoothing": False},
                                total_elapsed=1.25, max_step_elapsed=0.031, sequence="synthetic/os1-64/0")
    lines = text.splitlines()
    assert lines[0] == "# name: form" and "# timestamp, x, y, z, qx, qy, qz, qw" in lines
    meta, st, tr, rot = trajectory.read_evalio_csv(io.StringIO(text))
    assert meta["status"] == "complete" and meta["max_dist_matching"] == "0.8"
    assert float(meta["total_elapsed"]) == 1.25
    assert np.allclose(st, stamps, atol=1e-6)
    assert np.array_equal(tr, np.stack([p["t"] for p in poses]))  # repr() round-trips doubles exactly
    assert np.max(np.abs(rot - np.stack([p["R"].reshape(3, 3) for p in poses]))) < 1e-13
