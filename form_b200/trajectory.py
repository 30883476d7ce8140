"""Trajectory files in evalio's CSV layout.

FORM itself writes no files: `evalio run -M form` records the pose stream of FORM::add_lidar
(/root/reference/python/bindings.cpp:147-179) and FORM's experiment scripts read those files back
with evalio's Trajectory.from_file (/root/reference/experiments/window_size.py:66-74,
experiments/env.py:157-176 consumes the per-run `hz` and `status`).  evalio is not installable
here, so this module restates its layout [external, evalio 0.4]: a block of `# key: value`
metadata lines (name, pipeline, its parameters, status, total_elapsed, max_step_elapsed), a
`# timestamp, x, y, z, qx, qy, qz, qw` header and one row per pose (seconds, metres, unit
quaternion, scalar last)."""
from __future__ import annotations

import io
from typing import Iterable, Mapping

import numpy as np

COLUMNS = ("timestamp", "x", "y", "z", "qx", "qy", "qz", "qw")


def quat_from_rotation(R: np.ndarray) -> np.ndarray:
    """Unit quaternion (qx, qy, qz, qw), qw >= 0, of a 3x3 rotation (Shepperd's method)."""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    t = np.trace(R)
    if t > 0.0:
        s = np.sqrt(t + 1.0) * 2.0
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2.0
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    q /= np.linalg.norm(q)
    return q if q[3] >= 0.0 else -q


def rotation_from_quat(q: Iterable[float]) -> np.ndarray:
    x, y, z, w = (float(v) for v in q)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def write_evalio_csv(path_or_file, stamps: Iterable[float], poses, *, name: str = "form",
                     pipeline: str = "form", params: Mapping[str, object] | None = None,
                     status: str = "complete", total_elapsed: float | None = None,
                     max_step_elapsed: float | None = None, sequence: str | None = None) -> None:
    """poses: records with fields R (9, row-major) and t (3) - formgpu_pose / Estimator.pose()."""
    own = isinstance(path_or_file, (str, bytes)) or hasattr(path_or_file, "__fspath__")
    f = open(path_or_file, "w") if own else path_or_file
    try:
        f.write(f"# name: {name}\n# pipeline: {pipeline}\n")
        if sequence is not None:
            f.write(f"# sequence: {sequence}\n")
        for k, v in (params or {}).items():
            f.write(f"# {k}: {v}\n")
        f.write(f"# status: {status}\n")
        if total_elapsed is not None:
            f.write(f"# total_elapsed: {total_elapsed:.6f}\n")
        if max_step_elapsed is not None:
            f.write(f"# max_step_elapsed: {max_step_elapsed:.6f}\n")
        f.write("#\n# " + ", ".join(COLUMNS) + "\n")
        for stamp, pose in zip(stamps, poses):
            q = quat_from_rotation(np.asarray(pose["R"]))
            t = np.asarray(pose["t"], dtype=np.float64)
            f.write(f"{float(stamp):.9f}, {t[0]!r}, {t[1]!r}, {t[2]!r}, {q[0]!r}, {q[1]!r}, {q[2]!r}, {q[3]!r}\n"
                    .replace("np.float64(", "").replace(")", ""))
    finally:
        if own:
            f.close()


def read_evalio_csv(path_or_file):
    """-> (metadata dict, stamps (n,), translations (n, 3), rotations (n, 3, 3))."""
    own = isinstance(path_or_file, (str, bytes)) or hasattr(path_or_file, "__fspath__")
    f = open(path_or_file) if own else path_or_file
    try:
        meta, rows = {}, []
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith("#"):
                body = line[1:].strip()
                if ":" in body:
                    k, v = body.split(":", 1)
                    meta[k.strip()] = v.strip()
                continue
            rows.append([float(v) for v in line.split(",")])
    finally:
        if own:
            f.close()
    a = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
    rot = np.stack([rotation_from_quat(r[4:8]) for r in a]) if len(a) else np.zeros((0, 3, 3))
    return meta, a[:, 0], a[:, 1:4], rot


def to_string(stamps, poses, **kw) -> str:
    buf = io.StringIO()
    write_evalio_csv(buf, stamps, poses, **kw)
    return buf.getvalue()
