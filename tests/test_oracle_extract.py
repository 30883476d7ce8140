"""Pins for the stage-1 oracle: numpy restatement (bit-exact indices), numpy eigh
(normals, tolerance), edge cases.  CPU only."""
import zlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import np_restatement as npr
import oracle_lib
from form_b200 import _capi


def make_scan(rng, rows, cols, kind):
    n = rows * cols
    az = np.tile(np.linspace(0, 2 * np.pi, cols, endpoint=False), rows)
    el = np.repeat(np.linspace(-0.3, 0.3, rows), cols)
    if kind == "noise":
        r = rng.uniform(0.2, 120.0, n)
    elif kind == "ties":
        r = np.round(rng.uniform(2, 6, n))
    elif kind == "dropouts":
        r = 5.0 + 0.01 * rng.standard_normal(n)
        r[rng.uniform(size=n) < 0.3] = 0.0
    elif kind == "all_invalid":
        r = np.zeros(n)
    elif kind == "edge_dropouts":
        r = 5.0 + 0.01 * rng.standard_normal(n)
        r.reshape(rows, cols)[:, :8] = 0.0
        r.reshape(rows, cols)[:, -7:] = 500.0
    else:
        r = 8.0 + np.sin(3 * az) + 0.002 * rng.standard_normal(n)
    scan = np.zeros(n, dtype=_capi.POINT4F)
    scan["x"] = (r * np.cos(el) * np.cos(az)).astype(np.float32)
    scan["y"] = (r * np.cos(el) * np.sin(az)).astype(np.float32)
    scan["z"] = (r * np.sin(el)).astype(np.float32)
    return scan


def as4(scan):
    return np.stack([scan["x"], scan["y"], scan["z"], scan["w"]], axis=1)


def check_against_numpy(params, scan):
    rows, cols = params.num_rows, params.num_columns
    o = oracle_lib.Oracle(params, threads=2)
    pl, pt = o.extract(scan, 3)
    d = o.extract_debug()
    s4 = as4(scan)
    np_ = params.neighbor_points
    valid, pvalid = npr.masks(s4, rows, cols, np_, params.min_norm_squared, params.max_norm_squared)
    assert np.array_equal(valid, d["valid"].astype(bool))
    assert np.array_equal(pvalid, d["point_valid"].astype(bool))
    curv = npr.curvature(s4, rows, cols, np_, valid)
    assert np.array_equal(curv.view(np.uint32), d["curvature"].view(np.uint32))
    picks, used = npr.planar_picks(curv, valid, rows, cols, np_, params.num_sectors,
                                   params.planar_threshold, params.planar_feats_per_sector)
    assert np.array_equal(picks, d["planar_indices"])
    qpicks = npr.point_picks(used, valid, pvalid, rows, cols, np_, params.num_sectors,
                             params.point_feats_per_sector)
    assert np.array_equal(qpicks, d["point_indices"])
    # find_closest + keep flags + normals
    kept = 0
    for j, idx in enumerate(picks[:: max(1, len(picks) // 60)]):
        jj = j * max(1, len(picks) // 60)
        row = idx // cols
        cp = npr.closest_in_row(s4, valid, s4[idx], (row - 1) * cols, row * cols) if row > 0 else -1
        cn = npr.closest_in_row(s4, valid, s4[idx], (row + 1) * cols, (row + 2) * cols) if row < rows - 1 else -1
        assert cp == d["closest_prev"][jj] and cn == d["closest_next"][jj]
        ok, nrm, w = npr.normal_f64(s4, valid, int(idx), rows, cols, np_, params.radius, params.min_points)
        assert ok == bool(d["planar_keep"][jj])
        if ok:
            k = int(np.sum(d["planar_keep"][:jj]))
            got = np.array([pl["nx"][k], pl["ny"][k], pl["nz"][k]])
            assert abs(np.linalg.norm(got) - 1.0) < 1e-5
            if w[2] > 0 and (w[1] - w[0]) / w[2] > 1e-3:   # well-conditioned (SURVEY 8c iii)
                ang = np.arccos(min(1.0, abs(float(got @ nrm))))
                assert ang < 2e-3, (ang, w)
                kept += 1
    # final features = picks in order (rule R3), doubles of the float scan values
    keep = d["planar_keep"].astype(bool)
    assert np.array_equal(pl["x"], scan["x"][picks[keep]].astype(np.float64))
    assert np.array_equal(pt["z"], scan["z"][qpicks].astype(np.float64))
    assert np.all(pl["scan"] == 3) and np.all(pt["scan"] == 3)
    return len(picks), len(qpicks)


@pytest.mark.parametrize("kind", ["noise", "ties", "dropouts", "all_invalid", "edge_dropouts", "smooth"])
@pytest.mark.parametrize("shape", [(3, 64), (5, 333), (4, 600)])
def test_oracle_matches_numpy_restatement(kind, shape):
    rows, cols = shape
    rng = np.random.default_rng(zlib.crc32(f"{kind}{rows}{cols}".encode()))
    params = _capi.default_params(rows, cols)
    check_against_numpy(params, make_scan(rng, rows, cols, kind))


@pytest.mark.parametrize("overrides", [
    dict(point_feats_per_sector=0),
    dict(planar_feats_per_sector=3, point_feats_per_sector=7),
    dict(neighbor_points=2, num_sectors=5),
    dict(neighbor_points=7, num_sectors=1, planar_threshold=0.02),
    dict(min_norm_squared=0.01, radius=4.0, min_points=9),
])
def test_oracle_parameter_variants(overrides):
    rng = np.random.default_rng(7)
    params = _capi.default_params(4, 420, **overrides)
    check_against_numpy(params, make_scan(rng, 4, 420, "smooth"))
    check_against_numpy(params, make_scan(rng, 4, 420, "dropouts"))


def test_planar_cap_is_51_and_point_quirk():
    """'>' after the increment gives 51 planar picks per sector; the
    point loop's break leaves only the inner loop, so more than point_feats_per_sector+1
    features come out of a sector."""
    rows, cols = 2, 3000  # sector_len 500 -> up to 100 spaced picks > 51
    rng = np.random.default_rng(11)
    params = _capi.default_params(rows, cols)
    scan = make_scan(rng, rows, cols, "smooth")
    o = oracle_lib.Oracle(params, threads=1)
    o.extract(scan, 0)
    d = o.extract_debug()
    per_sector = np.bincount((d["planar_indices"] % cols) // 500 + 6 * (d["planar_indices"] // cols), minlength=12)
    assert per_sector.max() == 51
    params = _capi.default_params(rows, cols, planar_threshold=0.0)  # no planar picks at all
    o = oracle_lib.Oracle(params, threads=1)
    o.extract(scan, 0)
    d = o.extract_debug()
    assert len(d["planar_indices"]) == 0
    per_sector = np.bincount((d["point_indices"] % cols) // 500 + 6 * (d["point_indices"] // cols), minlength=12)
    assert per_sector.max() > params.point_feats_per_sector + 1


def test_wrong_scan_size_rejected():
    params = _capi.default_params(4, 64)
    o = oracle_lib.Oracle(params)
    with pytest.raises(ValueError):
        o.extract(np.zeros(10, dtype=_capi.POINT4F), 0)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2**32 - 1))
def test_eigen_solver_vs_numpy(seed):
    """Restated SelfAdjointEigenSolver<Matrix3f> against numpy.linalg.eigh (f64)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((rng.integers(3, 12), 3)) * rng.uniform(1e-3, 10)
    if seed % 5 == 0:
        A[:, 2] *= 1e-3  # nearly planar neighbourhood, the common case
    cov = (A.T @ A).astype(np.float32)
    evals = np.zeros(3, np.float32)
    evecs = np.zeros(9, np.float32)
    cov_c = np.ascontiguousarray(cov.reshape(9))
    oracle_lib.lib().oracle_eigen3f(_capi.ptr(cov_c), _capi.ptr(evals), _capi.ptr(evecs))
    w, v = np.linalg.eigh(cov.astype(np.float64))
    scale = max(abs(w).max(), 1e-30)
    assert np.all(np.diff(evals) >= 0)
    assert np.max(np.abs(evals - w)) < 2e-5 * scale
    V = evecs.reshape(3, 3).astype(np.float64)
    assert np.max(np.abs(V.T @ V - np.eye(3))) < 1e-5
    # residual of the eigen-decomposition
    assert np.max(np.abs(cov.astype(np.float64) @ V - V * evals[None, :])) < 5e-5 * scale


def test_eigen_solver_degenerate_inputs():
    for cov in (np.zeros(9), np.eye(3).reshape(9), np.diag([3.0, 1.0, 2.0]).reshape(9)):
        c = cov.astype(np.float32)
        evals = np.zeros(3, np.float32)
        evecs = np.zeros(9, np.float32)
        oracle_lib.lib().oracle_eigen3f(_capi.ptr(c), _capi.ptr(evals), _capi.ptr(evecs))
        assert np.allclose(np.sort(np.diag(cov.reshape(3, 3))), evals)
        assert np.allclose(np.abs(np.linalg.det(evecs.reshape(3, 3))), 1.0, atol=1e-6)
