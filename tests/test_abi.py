"""The C-ABI library loads on a CPU-only box and exports exactly what include/formgpu.h
declares; plain structs have the documented sizes; with no GPU every compute entry point
fails loudly (there is no CPU fallback).  No compute is attempted here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from form_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def header_symbols():
    text = open(os.path.join(ROOT, "include", "formgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(formgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert header_symbols() == sorted(_capi.FORMGPU_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _capi.gpu_lib()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.formgpu_abi_version() == 1


def test_struct_sizes_match_reference_pods():
    # PointXYZf 16 B (utils.hpp:38-91), PointFeat 40 B, PlanarFeat 72 B (features.hpp)
    assert _capi.POINT4F.itemsize == 16
    assert _capi.POINT_FEAT.itemsize == 40
    assert _capi.PLANAR_FEAT.itemsize == 72
    assert _capi.POSE.itemsize == 96 and _capi.SCAN_POSE.itemsize == 104
    assert _capi.PAIR.itemsize == 16 and _capi.PAIR_COUNT.itemsize == 16 and _capi.MATCH.itemsize == 24
    assert C.sizeof(_capi.Params) == 96
    assert C.sizeof(_capi.Request) == 176  # formgpu_request


def test_batch_entry_points_fail_loudly_without_a_gpu_or_with_bad_arguments():
    lib = _capi.gpu_lib()
    h = C.c_void_p()
    p = _capi.default_params(16, 1800)
    assert lib.formgpu_batch_create(C.byref(p), 0, None, 0, C.byref(h)) == _capi.ERR_INVALID_ARG
    assert lib.formgpu_batch_create(None, 0, None, 2, C.byref(h)) == _capi.ERR_INVALID_ARG
    if not _has_gpu():
        assert lib.formgpu_batch_create(C.byref(p), 0, None, 2, C.byref(h)) == _capi.ERR_CUDA
        assert b"no CUDA device" in lib.formgpu_batch_last_error(None)
    assert lib.formgpu_batch_submit(None, None, 0) == _capi.ERR_INVALID_ARG
    assert lib.formgpu_batch_size(None) == 0 and not lib.formgpu_batch_ctx(None, 0)


def test_default_params_are_the_reference_defaults():
    p = _capi.Params()
    _capi.gpu_lib().formgpu_default_params(C.byref(p))
    q = _capi.default_params()
    for name, _ in _capi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    # python/bindings.cpp:66-88 of FORM
    assert (p.neighbor_points, p.num_sectors, p.planar_feats_per_sector, p.point_feats_per_sector) == (5, 6, 50, 3)
    assert (p.planar_threshold, p.radius, p.min_points) == (1.0, 1.0, 5)
    assert (p.max_dist_matching, p.min_dist_map, p.sigma) == (0.8, 0.1, 0.1)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_error_not_fallback():
    lib = _capi.gpu_lib()
    h = C.c_void_p()
    p = _capi.default_params(16, 1800)
    rc = lib.formgpu_create(C.byref(p), 0, None, C.byref(h))
    assert rc == _capi.ERR_CUDA and not h.value
    assert b"no CUDA device" in lib.formgpu_last_error(None)
    from form_b200.pipeline import Estimator

    with pytest.raises(RuntimeError):
        Estimator(_capi.default_est_params(16, 1800))


def test_invalid_parameters_are_rejected_before_touching_the_gpu():
    lib = _capi.gpu_lib()
    h = C.c_void_p()
    for bad in (dict(neighbor_points=0), dict(num_columns=8), dict(num_columns=5000), dict(max_window_scans=1),
                dict(sigma=0.0)):
        p = _capi.default_params(16, 1800, **bad)
        rc = lib.formgpu_create(C.byref(p), 0, None, C.byref(h))
        assert rc in (_capi.ERR_UNSUPPORTED, _capi.ERR_INVALID_ARG), bad
    assert lib.formgpu_create(None, 0, None, C.byref(h)) == _capi.ERR_INVALID_ARG
    # null context is an error code, not a crash
    assert lib.formgpu_synchronize(None) == _capi.ERR_INVALID_ARG
    assert lib.formgpu_launch_count(None) == 0


def test_host_library_exports():
    lib = _capi.host_lib()
    for name in _capi.FORMHOST_SYMBOLS:
        assert hasattr(lib, name), name
    p = _capi.EstParams()
    lib.formhost_default_est_params(C.byref(p))
    q = _capi.default_est_params()
    assert (p.max_num_rematches, p.max_num_keyscans, p.max_num_recent_scans) == (30, 50, 10)
    assert p.new_pose_threshold == q.new_pose_threshold == 1e-4
