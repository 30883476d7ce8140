"""Seeded synthetic organised scans (host C++ generator, form/synth.hpp)."""
from __future__ import annotations

import numpy as np

from . import _capi

SENSORS = {"os1-64": 0, "os0-128": 1, "vlp-16": 2, "stress-128x2048": 3}


def shape(sensor: str) -> tuple[int, int]:
    import ctypes as C

    r, c = C.c_int(), C.c_int()
    n = _capi.synth_lib().formhost_synth_shape(SENSORS[sensor], C.byref(r), C.byref(c))
    assert n == r.value * c.value
    return r.value, c.value


def scan(sensor: str, sequence_id: int, k: int, threads: int = 0) -> np.ndarray:
    """Scan k of a sequence as a (rows*cols,) array of POINT4F, row-major."""
    rows, cols = shape(sensor)
    out = np.zeros(rows * cols, dtype=_capi.POINT4F)
    rc = _capi.synth_lib().formhost_synth_scan(SENSORS[sensor], sequence_id, k, _capi.ptr(out), threads)
    if rc != 0:
        raise ValueError(f"unknown sensor {sensor}")
    return out


def gt_pose(sequence_id: int, k: int) -> np.ndarray:
    """Ground-truth pose of scan k as a POSE record."""
    out = np.zeros(1, dtype=_capi.POSE)
    _capi.synth_lib().formhost_synth_gt_pose(sequence_id, k, _capi.ptr(out))
    return out[0]


def stress_scan(tile: int, k: int, threads: int = 0) -> np.ndarray:
    """128x2048 scan k of tile `tile` of the tiled-hall stress world (BASELINE.json configs[4])."""
    rows, cols = shape("stress-128x2048")
    out = np.zeros(rows * cols, dtype=_capi.POINT4F)
    _capi.synth_lib().formhost_synth_stress_scan(tile, k, _capi.ptr(out), threads)
    return out


def stress_pose(tile: int, k: int) -> np.ndarray:
    """World pose of scan k of tile `tile` (tiles repeat every 400 m)."""
    out = np.zeros(1, dtype=_capi.POSE)
    _capi.synth_lib().formhost_synth_stress_pose(tile, k, _capi.ptr(out))
    return out[0]
