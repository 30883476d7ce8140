"""Stage 1 parity: CUDA extraction vs the CPU oracle, through the C-ABI (bit-exact)."""
import zlib

import numpy as np
import pytest

from form_b200 import _capi, synth

pytestmark = pytest.mark.gpu


def _ctx(params):
    from form_b200.context import Context

    return Context(params)


def _compare(params, scan, scan_idx=7):
    import oracle_lib

    ref = oracle_lib.Oracle(params)
    rpl, rpt = ref.extract(scan, scan_idx)
    rd = ref.extract_debug()
    with _ctx(params) as ctx:
        pl, pt = ctx.extract(scan, scan_idx)
        d = ctx.extract_debug()
    for k in ("valid", "point_valid"):
        assert np.array_equal(d[k], rd[k]), k
    assert np.array_equal(d["curvature"].view(np.uint32), rd["curvature"].view(np.uint32)), "curvature bits"
    assert np.array_equal(d["planar_indices"], rd["planar_indices"]), "planar picks"
    assert np.array_equal(d["planar_keep"], rd["planar_keep"]), "normal keep flags"
    assert np.array_equal(d["closest_prev"], rd["closest_prev"]), "closest prev"
    assert np.array_equal(d["closest_next"], rd["closest_next"]), "closest next"
    assert np.array_equal(d["point_indices"], rd["point_indices"]), "point picks"
    assert pl.tobytes() == rpl.tobytes(), "planar features (incl. normals) not bit-exact"
    assert pt.tobytes() == rpt.tobytes(), "point features not bit-exact"
    return pl, pt


@pytest.mark.parametrize("sensor", ["os1-64", "os0-128", "vlp-16", "stress-128x2048"])
def test_extract_matches_oracle_on_synthetic(sensor):
    rows, cols = synth.shape(sensor)
    params = _capi.default_params(rows, cols)
    for k in (0, 57):
        pl, pt = _compare(params, synth.scan(sensor, 3, k), scan_idx=k)
        assert len(pl) > 100 and len(pt) > 10


@pytest.mark.parametrize("overrides", [
    dict(point_feats_per_sector=0),
    dict(planar_feats_per_sector=5, point_feats_per_sector=10),
    dict(neighbor_points=3, num_sectors=4, min_points=8),
    dict(neighbor_points=8, num_sectors=7, planar_threshold=0.05, radius=0.3),
    dict(min_norm_squared=0.01, radius=5.0),
])
def test_extract_parameter_variants(overrides):
    rows, cols = synth.shape("os1-64")
    params = _capi.default_params(rows, cols, **overrides)
    _compare(params, synth.scan("os1-64", 1, 11))


def _random_scan(rng, rows, cols, kind):
    n = rows * cols
    az = np.tile(np.linspace(0, 2 * np.pi, cols, endpoint=False), rows)
    el = np.repeat(np.linspace(-0.3, 0.3, rows), cols)
    if kind == "noise":
        r = rng.uniform(0.2, 120.0, n)
    elif kind == "ties":
        r = np.round(rng.uniform(2, 6, n))  # few distinct ranges -> exact curvature ties
    elif kind == "dropouts":
        r = 5.0 + 0.01 * rng.standard_normal(n)
        r[rng.uniform(size=n) < 0.3] = 0.0
    elif kind == "all_invalid":
        r = np.zeros(n)
    else:  # smooth
        r = 8.0 + np.sin(3 * az) + 0.002 * rng.standard_normal(n)
    scan = np.zeros(n, dtype=_capi.POINT4F)
    scan["x"] = (r * np.cos(el) * np.cos(az)).astype(np.float32)
    scan["y"] = (r * np.cos(el) * np.sin(az)).astype(np.float32)
    scan["z"] = (r * np.sin(el)).astype(np.float32)
    return scan


@pytest.mark.parametrize("kind", ["noise", "ties", "dropouts", "all_invalid", "smooth"])
@pytest.mark.parametrize("shape", [(4, 64), (7, 333), (16, 1800), (3, 2048)])
def test_extract_edge_cases(kind, shape):
    rows, cols = shape
    rng = np.random.default_rng(zlib.crc32(f"{kind}-{rows}-{cols}".encode()))
    params = _capi.default_params(rows, cols)
    _compare(params, _random_scan(rng, rows, cols, kind))


def test_extract_wrong_size_is_an_error():
    from form_b200.context import FormGpuError

    params = _capi.default_params(16, 1800)
    with _ctx(params) as ctx:
        with pytest.raises(FormGpuError) as e:
            ctx.extract(np.zeros(100, dtype=_capi.POINT4F), 0)
        assert e.value.code == _capi.ERR_BAD_SCAN_SIZE


def test_single_row_scan_drops_all_planar():
    # no adjacent row -> compute_normal fails for every pick (extraction.tpp:308-310)
    params = _capi.default_params(1, 512)
    rng = np.random.default_rng(5)
    pl, pt = _compare(params, _random_scan(rng, 1, 512, "smooth"))
    assert len(pl) == 0


def test_extract_into_pinned_buffers_matches_pageable():
    """Page-locked caller buffers are filled by the kernel itself (no staging copy): same bytes."""
    import ctypes as C

    lib = _capi.gpu_lib()
    rows, cols = synth.shape("os1-64")
    params = _capi.default_params(rows, cols)
    scan = synth.scan("os1-64", 2, 5)
    with _ctx(params) as ctx:
        pl, pt = ctx.extract(scan, 9)  # pageable numpy buffers
        capp, capq = ctx.max_planar, ctx.max_point
        bp = lib.formgpu_alloc_pinned(capp * 72)
        bq = lib.formgpu_alloc_pinned(capq * 40)
        assert bp and bq
        try:
            a, b = C.c_size_t(), C.c_size_t()
            rc = lib.formgpu_extract(ctx._h, _capi.ptr(scan), scan.shape[0], 9, C.c_void_p(bp), capp, C.byref(a),
                                     C.c_void_p(bq), capq, C.byref(b))
            assert rc == 0 and a.value == len(pl) and b.value == len(pt)
            got_pl = np.ctypeslib.as_array((C.c_uint8 * (72 * a.value)).from_address(bp)).view(_capi.PLANAR_FEAT)
            got_pt = np.ctypeslib.as_array((C.c_uint8 * (40 * b.value)).from_address(bq)).view(_capi.POINT_FEAT)
            assert got_pl.tobytes() == pl.tobytes()
            assert got_pt.tobytes() == pt.tobytes()
        finally:
            lib.formgpu_free_pinned(bp)
            lib.formgpu_free_pinned(bq)
