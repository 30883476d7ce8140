"""ctypes view of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

from form_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
# FORM_ORACLE_LIB: another build of the same sources (tests/sanitize_host.sh: ASan / UBSan / TSan)
LIB = os.environ.get("FORM_ORACLE_LIB") or os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_vp, _sz, _u64, _i, _d = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_double
_psz = C.POINTER(C.c_size_t)

_SYMS = {
    "oracle_create": (_vp, [C.POINTER(_capi.Params), _i]),
    "oracle_destroy": (None, [_vp]),
    "oracle_extract": (_i, [_vp, _vp, _sz, _u64, _psz, _psz]),
    "oracle_get_features": (None, [_vp, _vp, _vp]),
    "oracle_extract_debug": (None, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _psz, _vp, _psz]),
    "oracle_map_rebuild": (None, [_vp, _vp, _sz]),
    "oracle_num_voxels": (_sz, [_vp, _i]),
    "oracle_associate": (_i, [_vp, _vp, _vp, _sz, _psz]),
    "oracle_get_matches": (_i, [_vp, _i, _vp, _sz, _psz]),
    "oracle_get_pair": (_i, [_vp, _u64, _u64, _vp, _vp, _vp, _psz, _vp, _vp, _psz]),
    "oracle_linearize": (None, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "oracle_error": (None, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "oracle_commit_scan": (None, [_vp, _psz, _psz]),
    "oracle_remove_scans": (None, [_vp, _vp, _sz]),
    "oracle_get_keypoints": (_i, [_vp, _i, _u64, _vp, _sz, _psz]),
    "oracle_eigen3f": (None, [_vp, _vp, _vp]),
    "oracle_compute_coords": (None, [_d, _d, _d, _d, _vp]),
    "oracle_voxel_shifts": (None, [_vp]),
    "oracle_plane_point": (None, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "oracle_point_point": (None, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "oracle_linearize_raw": (None, [_vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _d, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()  # bench.py's CPU legs create their estimators from worker threads


def lib() -> C.CDLL:
    """The library with every prototype set.  Published only once it is fully configured: a
    thread that used a function before its restype was set would truncate the returned handle."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB):
                    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])
                loaded = C.CDLL(LIB)
                for name, (res, args) in _SYMS.items():
                    fn = getattr(loaded, name)
                    fn.restype, fn.argtypes = res, args
                _lib = loaded
    return _lib


class Oracle:
    """The oracle behind the same call sequence as the C-ABI context."""

    def __init__(self, params: _capi.Params, threads: int = 0):
        self.params = params
        self.rows, self.cols = params.num_rows, params.num_columns
        self._h = lib().oracle_create(C.byref(params), threads)

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def extract(self, scan: np.ndarray, scan_idx: int):
        npl, npt = C.c_size_t(), C.c_size_t()
        rc = lib().oracle_extract(self._h, _capi.ptr(scan), scan.shape[0], scan_idx, C.byref(npl), C.byref(npt))
        if rc != 0:
            raise ValueError(f"oracle_extract rc={rc}")
        planar = np.zeros(npl.value, dtype=_capi.PLANAR_FEAT)
        point = np.zeros(npt.value, dtype=_capi.POINT_FEAT)
        lib().oracle_get_features(self._h, _capi.ptr(planar), _capi.ptr(point))
        return planar, point

    def extract_debug(self):
        n = self.rows * self.cols
        valid = np.zeros(n, np.uint8)
        pvalid = np.zeros(n, np.uint8)
        curv = np.zeros(n, np.float32)
        pidx = np.zeros(n, np.uint32)
        keep = np.zeros(n, np.uint8)
        cprev = np.zeros(n, np.int32)
        cnext = np.zeros(n, np.int32)
        qidx = np.zeros(n, np.uint32)
        npk, nqk = C.c_size_t(), C.c_size_t()
        lib().oracle_extract_debug(self._h, _capi.ptr(valid), _capi.ptr(pvalid), _capi.ptr(curv),
                                   _capi.ptr(pidx), _capi.ptr(keep), _capi.ptr(cprev), _capi.ptr(cnext),
                                   C.byref(npk), _capi.ptr(qidx), C.byref(nqk))
        k, q = npk.value, nqk.value
        return dict(valid=valid, point_valid=pvalid, curvature=curv, planar_indices=pidx[:k],
                    planar_keep=keep[:k], closest_prev=cprev[:k], closest_next=cnext[:k],
                    point_indices=qidx[:q])

    def map_rebuild(self, poses: np.ndarray):
        lib().oracle_map_rebuild(self._h, _capi.ptr(poses), poses.shape[0])

    def num_voxels(self, type_: int) -> int:
        return lib().oracle_num_voxels(self._h, type_)

    def associate(self, pose: np.ndarray) -> np.ndarray:
        out = np.zeros(256, dtype=_capi.PAIR_COUNT)
        n = C.c_size_t()
        pose = np.ascontiguousarray(pose)
        rc = lib().oracle_associate(self._h, _capi.ptr(pose), _capi.ptr(out), out.shape[0], C.byref(n))
        assert rc == 0
        return out[: n.value].copy()

    def matches(self, type_: int) -> np.ndarray:
        cap = self.rows * self.cols
        out = np.zeros(cap, dtype=_capi.MATCH)
        n = C.c_size_t()
        rc = lib().oracle_get_matches(self._h, type_, _capi.ptr(out), cap, C.byref(n))
        assert rc == 0
        return out[: n.value].copy()

    def pair(self, i: int, j: int):
        cap = self.rows * self.cols
        a = [np.zeros((cap, 3)) for _ in range(5)]
        npl, npt = C.c_size_t(), C.c_size_t()
        lib().oracle_get_pair(self._h, i, j, _capi.ptr(a[0]), _capi.ptr(a[1]), _capi.ptr(a[2]), C.byref(npl),
                              _capi.ptr(a[3]), _capi.ptr(a[4]), C.byref(npt))
        n, m = npl.value, npt.value
        return dict(pl_pi=a[0][:n].copy(), pl_ni=a[1][:n].copy(), pl_pj=a[2][:n].copy(),
                    pt_pi=a[3][:m].copy(), pt_pj=a[4][:m].copy())

    def linearize(self, pairs: np.ndarray, poses: np.ndarray) -> np.ndarray:
        out = np.zeros((pairs.shape[0], 91))
        lib().oracle_linearize(self._h, _capi.ptr(pairs), pairs.shape[0], _capi.ptr(poses), poses.shape[0],
                               _capi.ptr(out))
        return out

    def error(self, pairs: np.ndarray, poses: np.ndarray) -> np.ndarray:
        out = np.zeros(pairs.shape[0])
        lib().oracle_error(self._h, _capi.ptr(pairs), pairs.shape[0], _capi.ptr(poses), poses.shape[0],
                           _capi.ptr(out))
        return out

    def commit_scan(self):
        a, b = C.c_size_t(), C.c_size_t()
        lib().oracle_commit_scan(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def remove_scans(self, scans):
        s = np.asarray(scans, dtype=np.uint64)
        lib().oracle_remove_scans(self._h, _capi.ptr(s), s.shape[0])

    def keypoints(self, type_: int, scan: int) -> np.ndarray:
        n = C.c_size_t()
        lib().oracle_get_keypoints(self._h, type_, scan, None, 0, C.byref(n))
        out = np.zeros(n.value, dtype=_capi.PLANAR_FEAT if type_ == 0 else _capi.POINT_FEAT)
        if n.value:
            lib().oracle_get_keypoints(self._h, type_, scan, _capi.ptr(out), n.value, C.byref(n))
        return out


# ---- the shared host pipeline over the oracle (oracle_pipeline_capi.cpp) ----
from form_b200 import pipeline as _pipeline  # noqa: E402

_pipe_ready = False


def _pipe_lib():
    global _pipe_ready
    l = lib()
    if not _pipe_ready:
        with _lock:
            if not _pipe_ready:
                for name, (res, args) in _capi.estimator_symbols("oracle_").items():
                    fn = getattr(l, name)
                    fn.restype, fn.argtypes = res, args
                l.oracle_est_last_error.restype = C.c_char_p
                l.oracle_replay_create.restype = _vp
                l.oracle_replay_create.argtypes = [_vp, C.POINTER(_capi.EstParams)]
                _pipe_ready = True
    return l


class OracleEstimator(_pipeline.EstimatorBase):
    """form::Estimator host logic running over the CPU oracle."""

    _prefix = "oracle_"

    def _lib(self):
        return _pipe_lib()

    def _last_error(self):
        return (self._lib().oracle_est_last_error() or b"").decode()


class OracleReplay(_pipeline.ReplayBase):
    """Replays a recorded hot-path trace on the CPU oracle (cpu_baseline leg)."""

    _prefix = "oracle_"

    def _lib(self):
        return _pipe_lib()

    def __init__(self, trace, params: _capi.EstParams):
        self._h = self._lib().oracle_replay_create(trace, C.byref(params))
