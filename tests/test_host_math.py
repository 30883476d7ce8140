"""Host-side SE(3) math of the smoother (form/pose3.hpp), CPU only."""
import numpy as np
from scipy.linalg import expm, logm

from form_b200 import _capi

L = _capi.synth_lib


def P(R, t):
    out = np.zeros(1, dtype=_capi.POSE)
    out["R"] = np.asarray(R).reshape(9)
    out["t"] = t
    return out


def expmap(xi):
    out = np.zeros(1, dtype=_capi.POSE)
    x = np.asarray(xi, dtype=np.float64)
    L().formhost_pose_expmap(_capi.ptr(x), _capi.ptr(out))
    return out


def logmap(p):
    xi = np.zeros(6)
    L().formhost_pose_logmap(_capi.ptr(p), _capi.ptr(xi))
    return xi


def hat(xi):
    w, v = xi[:3], xi[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


def test_expmap_matches_matrix_exponential():
    rng = np.random.default_rng(0)
    for scale in (1e-9, 1e-4, 0.3, 2.5):
        for _ in range(5):
            xi = rng.normal(size=6) * scale
            T = expmap(xi)[0]
            E = expm(hat(xi))
            assert np.allclose(T["R"].reshape(3, 3), E[:3, :3], atol=1e-12)
            assert np.allclose(T["t"], E[:3, 3], atol=1e-12)
            if np.linalg.norm(xi[:3]) < 3.0:  # the log is only unique below pi
                assert np.allclose(logmap(expmap(xi)), xi, atol=1e-9)


def test_logmap_near_pi():
    for axis in np.eye(3):
        xi = np.concatenate([axis * (np.pi - 1e-5), [0.3, -0.2, 0.1]])
        assert np.allclose(np.abs(logmap(expmap(xi))[:3]), np.abs(xi[:3]), atol=1e-6)


def test_compose_inverse():
    rng = np.random.default_rng(1)
    a, b = expmap(rng.normal(size=6)), expmap(rng.normal(size=6))
    out, inv = np.zeros(1, dtype=_capi.POSE), np.zeros(1, dtype=_capi.POSE)
    L().formhost_pose_inverse(_capi.ptr(a), _capi.ptr(inv))
    L().formhost_pose_compose(_capi.ptr(a), _capi.ptr(inv), _capi.ptr(out))
    assert np.allclose(out["R"].reshape(3, 3), np.eye(3), atol=1e-14) and np.allclose(out["t"], 0, atol=1e-14)
    L().formhost_pose_compose(_capi.ptr(a), _capi.ptr(b), _capi.ptr(out))
    A, B = np.eye(4), np.eye(4)
    A[:3, :3], A[:3, 3] = a["R"].reshape(3, 3), a["t"]
    B[:3, :3], B[:3, 3] = b["R"].reshape(3, 3), b["t"]
    Cm = A @ B
    assert np.allclose(out["R"].reshape(3, 3), Cm[:3, :3]) and np.allclose(out["t"][0], Cm[:3, 3])


def test_logmap_derivative_finite_difference():
    """d Logmap(T Exp(d)) / d d at 0 - the Jacobian the pose prior uses."""
    rng = np.random.default_rng(2)
    for scale in (1e-7, 0.05, 0.8):
        xi0 = rng.normal(size=6) * scale
        T = expmap(xi0)
        J = np.zeros(36)
        L().formhost_pose_logmap_derivative(_capi.ptr(T), _capi.ptr(J))
        J = J.reshape(6, 6)
        num = np.zeros((6, 6))
        eps = 1e-6
        for k in range(6):
            d = np.zeros(6)
            d[k] = eps
            out_p, out_m = np.zeros(1, dtype=_capi.POSE), np.zeros(1, dtype=_capi.POSE)
            L().formhost_pose_compose(_capi.ptr(T), _capi.ptr(expmap(d)), _capi.ptr(out_p))
            L().formhost_pose_compose(_capi.ptr(T), _capi.ptr(expmap(-d)), _capi.ptr(out_m))
            num[:, k] = (logmap(out_p) - logmap(out_m)) / (2 * eps)
        assert np.allclose(J, num, atol=1e-6), (scale, np.abs(J - num).max())


def test_normalized_restores_orthonormality():
    rng = np.random.default_rng(3)
    T = expmap(rng.normal(size=6))
    T["R"] += rng.normal(size=9) * 1e-6
    out = np.zeros(1, dtype=_capi.POSE)
    L().formhost_pose_normalized(_capi.ptr(T), _capi.ptr(out))
    R = out["R"].reshape(3, 3)
    assert np.max(np.abs(R.T @ R - np.eye(3))) < 1e-10
    assert abs(np.linalg.det(R) - 1) < 1e-10
