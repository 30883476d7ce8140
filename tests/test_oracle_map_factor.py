"""Pins for the stage-2 / stage-3 oracle (CPU only): hand-computed voxel keys, brute-force
nearest neighbour with rule R5, finite-difference Jacobians at the reference's stale test
seeds (tests/test_SeparateFactor.cpp:26-31,53-59), explicit A^T A in numpy."""
import ctypes as C

import numpy as np
import pytest

import np_restatement as npr
import oracle_lib
from form_b200 import _capi
from helpers import perturbed, scan_poses, unpack91


def pose_rzryrx(rx, ry, rz, t):
    out = np.zeros(1, dtype=_capi.POSE)
    tt = np.asarray(t, dtype=np.float64)
    _capi.synth_lib().formhost_pose_rzryrx(rx, ry, rz, _capi.ptr(tt), _capi.ptr(out))
    return out[0]


def expmap(xi):
    out = np.zeros(1, dtype=_capi.POSE)
    x = np.asarray(xi, dtype=np.float64)
    _capi.synth_lib().formhost_pose_expmap(_capi.ptr(x), _capi.ptr(out))
    return out[0]


def compose(a, b):
    out = np.zeros(1, dtype=_capi.POSE)
    aa, bb = np.array([a], dtype=_capi.POSE), np.array([b], dtype=_capi.POSE)
    _capi.synth_lib().formhost_pose_compose(_capi.ptr(aa), _capi.ptr(bb), _capi.ptr(out))
    return out[0]


# ---------------------------------------------------------------- voxel keys
@pytest.mark.parametrize("p,w,expect", [
    ((0.0, 0.0, 0.0), 0.8, (0, 0, 0)),
    ((-0.0, 0.79999, 0.8), 0.8, (0, 0, 1)),
    ((-1e-12, -0.8, -0.80000001), 0.8, (-1, -1, -2)),
    ((1.6, 2.4000000000000004, 2.3999999999999995), 0.8, (2, 3, 2)),
    ((-37.123, 12.0, 99.99), 0.8, (-47, 15, 124)),
    ((0.1, 0.2, 0.30000000000000004), 0.1, (1, 2, 3)),
])
def test_voxel_keys_hand_computed(p, w, expect):
    out = np.zeros(3, np.int32)
    oracle_lib.lib().oracle_compute_coords(p[0], p[1], p[2], w, _capi.ptr(out))
    assert tuple(out) == expect
    assert tuple(npr.voxel_key(p, w)) == expect


def test_voxel_key_boundaries_one_ulp():
    w = 0.8
    for k in (-5, -1, 0, 1, 7, 123):
        x = k * w
        for v in (np.nextafter(x, -np.inf), x, np.nextafter(x, np.inf)):
            out = np.zeros(3, np.int32)
            oracle_lib.lib().oracle_compute_coords(v, v, v, w, _capi.ptr(out))
            assert out[0] == int(np.floor(np.float64(v) / np.float64(w)))


def test_shift_table_order():
    out = np.zeros(81, np.int32)
    oracle_lib.lib().oracle_voxel_shifts(_capi.ptr(out))
    assert np.array_equal(out.reshape(27, 3), npr.SHIFTS)
    assert len({tuple(s) for s in npr.SHIFTS}) == 27


# ---------------------------------------------------------------- nearest neighbour
def _cloud_scan(rng, rows, cols, radius=6.0):
    az = np.tile(np.linspace(0, 2 * np.pi, cols, endpoint=False), rows)
    el = np.repeat(np.linspace(-0.4, 0.4, rows), cols)
    r = radius + 0.5 * np.sin(5 * az) + 0.01 * rng.standard_normal(rows * cols)
    scan = np.zeros(rows * cols, dtype=_capi.POINT4F)
    scan["x"] = (r * np.cos(el) * np.cos(az)).astype(np.float32)
    scan["y"] = (r * np.cos(el) * np.sin(az)).astype(np.float32)
    scan["z"] = (r * np.sin(el)).astype(np.float32)
    return scan


def test_matching_vs_bruteforce_and_insert_rule():
    rng = np.random.default_rng(3)
    rows, cols = 8, 512
    params = _capi.default_params(rows, cols)
    o = oracle_lib.Oracle(params, threads=2)
    ident = np.zeros((), dtype=_capi.POSE)
    ident["R"] = np.eye(3).reshape(9)
    poses = {0: ident}
    world = {0: [], 1: []}   # type -> list of (xyz world, scan, k)
    for k in range(3):
        scan = _cloud_scan(rng, rows, cols)
        pl, pt = o.extract(scan, k)
        poses[k] = perturbed(ident, rng, 0.01, 0.05) if k else ident
        sp = scan_poses(list(poses), [poses[s] for s in poses])
        o.map_rebuild(sp)
        pose_k = perturbed(poses[k], rng, 0.003, 0.02)
        counts = o.associate(pose_k)
        for t, feats in ((0, pl), (1, pt)):
            m = o.matches(t)
            assert len(m) == len(feats)
            pts = np.array([w[0] for w in world[t]]).reshape(-1, 3)
            ids = [(w[1], w[2]) for w in world[t]]
            q = npr.transform(pose_k["R"], pose_k["t"], np.stack([feats["x"], feats["y"], feats["z"]], 1))
            for j in range(0, len(feats), max(1, len(feats) // 80)):
                best = npr.nn_bruteforce(pts, ids, q[j], 0.8) if len(pts) else None
                if best is None:
                    assert m["found"][j] == 0 and m["dist_sqrd"][j] == np.finfo(np.float64).max
                else:
                    assert m["found"][j] == 1
                    assert (m["dist_sqrd"][j], m["scan"][j], m["k"][j]) == (best[0], best[2], best[3])
            # constraint counts = matches with dist^2 < max_dist^2 grouped by matched scan
            sel = m["dist_sqrd"] < 0.8 * 0.8
            for c in counts:
                n = int(np.sum(sel & (m["scan"] == c["i"])))
                assert n == (c["n_planar"] if t == 0 else c["n_point"])
        added = o.commit_scan()
        for t, feats in ((0, pl), (1, pt)):
            m = o.matches(t)
            novel = m["dist_sqrd"] > 0.1 * 0.1     # strict '>' (map.tpp:161); unmatched inserts
            assert added[t] == int(novel.sum())
            stored = o.keypoints(t, k)
            assert stored.tobytes() == feats[novel].tobytes()
            w = npr.transform(poses[k]["R"], poses[k]["t"], np.stack([stored["x"], stored["y"], stored["z"]], 1))
            world[t] += [(w[i], k, i) for i in range(len(stored))]
    # scan 0 inserted everything
    assert len(o.keypoints(0, 0)) > 0
    o.remove_scans([1])
    assert len(o.keypoints(0, 1)) == 0 and len(o.keypoints(1, 1)) == 0


# ---------------------------------------------------------------- factors
X0 = lambda: pose_rzryrx(0.1, 0.2, 0.3, (1, 2, 3))  # noqa: E731  test_SeparateFactor.cpp:26
X1 = lambda: pose_rzryrx(0.4, 0.5, 0.6, (4, 5, 6))  # noqa: E731  test_SeparateFactor.cpp:27


def _plane_point(p_i, n_i, p_j, Ti, Tj):
    n = len(p_i)
    r, H1, H2 = np.zeros(n), np.zeros((n, 6)), np.zeros((n, 6))
    a, b = np.array([Ti], dtype=_capi.POSE), np.array([Tj], dtype=_capi.POSE)
    oracle_lib.lib().oracle_plane_point(_capi.ptr(p_i), _capi.ptr(n_i), _capi.ptr(p_j), n, _capi.ptr(a),
                                        _capi.ptr(b), _capi.ptr(r), _capi.ptr(H1), _capi.ptr(H2))
    return r, H1, H2


def _point_point(p_i, p_j, Ti, Tj):
    m = len(p_i)
    r, H1, H2 = np.zeros(3 * m), np.zeros((3 * m, 6)), np.zeros((3 * m, 6))
    a, b = np.array([Ti], dtype=_capi.POSE), np.array([Tj], dtype=_capi.POSE)
    oracle_lib.lib().oracle_point_point(_capi.ptr(p_i), _capi.ptr(p_j), m, _capi.ptr(a), _capi.ptr(b),
                                        _capi.ptr(r), _capi.ptr(H1), _capi.ptr(H2))
    return r, H1, H2


def _numeric_jacobian(f, T, eps=1e-6):
    """Central differences under the right perturbation T * Exp(xi), xi = [omega, v]."""
    cols = []
    for k in range(6):
        xi = np.zeros(6)
        xi[k] = eps
        plus = f(compose(T, expmap(xi)))
        minus = f(compose(T, expmap(-xi)))
        cols.append((plus - minus) / (2 * eps))
    return np.stack(cols, axis=1)


def test_plane_point_jacobians_at_reference_seeds():
    p_i = np.array([[1.0, 2.0, 3.0]])
    n_i = np.array([[0.0, 0.0, 1.0]])
    p_j = np.array([[4.0, 5.0, 6.0]])      # test_SeparateFactor.cpp:56-58
    Ti, Tj = X0(), X1()
    r, H1, H2 = _plane_point(p_i, n_i, p_j, Ti, Tj)
    # closed form of the residual
    Ri, Rj = Ti["R"].reshape(3, 3), Tj["R"].reshape(3, 3)
    expect = (Ri @ n_i[0]) @ (Rj @ p_j[0] + Tj["t"] - Ri @ p_i[0] - Ti["t"])
    assert abs(r[0] - expect) < 1e-12
    N1 = _numeric_jacobian(lambda T: _plane_point(p_i, n_i, p_j, T, Tj)[0], Ti)
    N2 = _numeric_jacobian(lambda T: _plane_point(p_i, n_i, p_j, Ti, T)[0], Tj)
    assert np.allclose(H1, N1, rtol=1e-5, atol=1e-7)     # isApprox(., 1e-5), :16-19
    assert np.allclose(H2, N2, rtol=1e-5, atol=1e-7)


def test_point_point_jacobians_at_reference_seeds():
    p_i = np.tile([[1.0, 2.0, 3.0]], (4, 1))
    p_j = np.tile([[4.0, 5.0, 6.0]], (4, 1))   # test_SeparateFactor.cpp:30-31, four copies (:34-37)
    Ti, Tj = X0(), X1()
    r, H1, H2 = _point_point(p_i, p_j, Ti, Tj)
    Ri, Rj = Ti["R"].reshape(3, 3), Tj["R"].reshape(3, 3)
    expect = Rj @ p_j[0] + Tj["t"] - Ri @ p_i[0] - Ti["t"]
    assert np.allclose(r.reshape(4, 3), expect[None, :], atol=1e-12)
    N1 = _numeric_jacobian(lambda T: _point_point(p_i, p_j, T, Tj)[0], Ti)
    N2 = _numeric_jacobian(lambda T: _point_point(p_i, p_j, Ti, T)[0], Tj)
    assert np.allclose(H1, N1, rtol=1e-5, atol=1e-7)
    assert np.allclose(H2, N2, rtol=1e-5, atol=1e-7)


def test_random_jacobians_finite_difference():
    rng = np.random.default_rng(0)
    n, m = 7, 5
    p_i, p_j = rng.normal(size=(n, 3)) * 5, rng.normal(size=(n, 3)) * 5
    n_i = rng.normal(size=(n, 3))
    n_i /= np.linalg.norm(n_i, axis=1, keepdims=True)
    q_i, q_j = rng.normal(size=(m, 3)) * 5, rng.normal(size=(m, 3)) * 5
    Ti, Tj = expmap(rng.normal(size=6) * 0.5), expmap(rng.normal(size=6) * 0.5)
    _, H1, H2 = _plane_point(p_i, n_i, p_j, Ti, Tj)
    assert np.allclose(H1, _numeric_jacobian(lambda T: _plane_point(p_i, n_i, p_j, T, Tj)[0], Ti), rtol=1e-5, atol=1e-6)
    assert np.allclose(H2, _numeric_jacobian(lambda T: _plane_point(p_i, n_i, p_j, Ti, T)[0], Tj), rtol=1e-5, atol=1e-6)
    _, G1, G2 = _point_point(q_i, q_j, Ti, Tj)
    assert np.allclose(G1, _numeric_jacobian(lambda T: _point_point(q_i, q_j, T, Tj)[0], Ti), rtol=1e-5, atol=1e-6)
    assert np.allclose(G2, _numeric_jacobian(lambda T: _point_point(q_i, q_j, Ti, T)[0], Tj), rtol=1e-5, atol=1e-6)


def test_block_is_explicit_AtA():
    """13x13 block == [A b]^T [A b] with A = [J_i J_j]/sigma, b = -r/sigma (gtsam.hpp:67-86)."""
    rng = np.random.default_rng(1)
    n, m, sigma = 40, 25, 0.1
    p_i, p_j = rng.normal(size=(n, 3)) * 8, rng.normal(size=(n, 3)) * 8
    n_i = rng.normal(size=(n, 3))
    n_i /= np.linalg.norm(n_i, axis=1, keepdims=True)
    q_i, q_j = rng.normal(size=(m, 3)) * 8, rng.normal(size=(m, 3)) * 8
    Ti, Tj = expmap(rng.normal(size=6) * 0.3), expmap(rng.normal(size=6) * 0.3)
    out, err = np.zeros(91), C.c_double()
    a, b = np.array([Ti], dtype=_capi.POSE), np.array([Tj], dtype=_capi.POSE)
    oracle_lib.lib().oracle_linearize_raw(_capi.ptr(p_i), _capi.ptr(n_i), _capi.ptr(p_j), n, _capi.ptr(q_i),
                                          _capi.ptr(q_j), m, _capi.ptr(a), _capi.ptr(b), sigma, _capi.ptr(out),
                                          C.byref(err))
    r1, H1, H2 = _plane_point(p_i, n_i, p_j, Ti, Tj)
    r2, G1, G2 = _point_point(q_i, q_j, Ti, Tj)
    A = np.vstack([np.hstack([H1, H2]), np.hstack([G1, G2])]).astype(np.longdouble) / sigma
    bb = -np.concatenate([r1, r2]).astype(np.longdouble) / sigma
    Ab = np.hstack([A, bb[:, None]])
    M = (Ab.T @ Ab).astype(np.float64)
    got = unpack91(out)
    assert np.max(np.abs(got - M) / np.maximum(np.abs(M), 1e-300)) < 1e-12
    assert abs(err.value - 0.5 * M[12, 12]) < 1e-12 * M[12, 12]
    # planar rows come first, then point rows (factor.cpp:156-183): the block of a pair with
    # only planar rows equals the planar part alone
    out_p = np.zeros(91)
    oracle_lib.lib().oracle_linearize_raw(_capi.ptr(p_i), _capi.ptr(n_i), _capi.ptr(p_j), n, None, None, 0,
                                          _capi.ptr(a), _capi.ptr(b), sigma, _capi.ptr(out_p), None)
    Ap = np.hstack([H1, H2, -r1[:, None]]) / sigma
    assert np.allclose(unpack91(out_p), Ap.T @ Ap, rtol=1e-12, atol=1e-9)


def test_empty_pair_is_zero_block():
    ident = np.zeros(1, dtype=_capi.POSE)
    ident["R"] = np.eye(3).reshape(9)
    out = np.ones(91)
    oracle_lib.lib().oracle_linearize_raw(None, None, None, 0, None, None, 0, _capi.ptr(ident), _capi.ptr(ident),
                                          0.1, _capi.ptr(out), None)
    assert not out.any()
