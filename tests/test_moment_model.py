"""CPU model of the pair-moment cache of stage 3 (form_b200/csrc/moments.cu), checked against
the oracle's DenseFactor::linearize restatement (oracle_linearize_raw).

The kernels accumulate, ONCE per association, pose-independent second moments of every scan
pair's correspondences at a reference relative pose rel0 = (R0, t0) (q0 = R0 p_j + t0 in the
frame of scan i):
    plane-point:  phi  = [ vec(n q0^T) (9, index 3a+b),  n (3),  r0 = n.(q0 - p_i) ]   -> 13x13
    point-point:  zeta = [ q0 (3),  d0 = q0 - p_i (3),  1 ]                             -> 7x7
Every later linearisation / error evaluation at ANY pair of poses is a congruence of those
moments: with the requested relative pose (R, t), dR = R R0^T, dt = t - dR t0 one has
q = dR q0 + dt, the 7-vectors of the streaming kernel (linearize.cu) are LINEAR in phi / zeta,
    s = [n x q, n, n.(q - p_i)] = S(dR, dt) phi,      z = [p_i, q - p_i, 1] = Z(dR, dt) zeta,
so W_p = S M_p S^T and W_q = Z M_q Z^T are the 7x7 moment matrices that kernel would have
accumulated, and the 13x13 block follows from the same basis expansion.  The functions below
use the kernels' index conventions one to one."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from form_b200 import _capi
from helpers import block_rel_err
from test_oracle_map_factor import expmap

SYM3 = [[0, 1, 2], [1, 3, 4], [2, 4, 5]]  # index of the symmetric pair (a, c) among the 6 products


def accumulate_planar73(p_i, n, p_j, rel0):
    """The 73 distinct sums a thread keeps per plane-point correspondence."""
    R0, t0 = rel0
    acc = np.zeros(73)
    for k in range(len(p_i)):
        q = R0 @ p_j[k] + t0
        r0 = n[k] @ (q - p_i[k])
        nn = [n[k][a] * n[k][c] for a in range(3) for c in range(a, 3)]
        qq = [q[b] * q[d] for b in range(3) for d in range(b, 3)]
        e = 0
        for ac in range(6):          # [0, 36): (n_a n_c)(q_b q_d)
            for bd in range(6):
                acc[e] += nn[ac] * qq[bd]
                e += 1
        for ac in range(6):          # [36, 54): (n_a n_c) q_b
            for b in range(3):
                acc[e] += nn[ac] * q[b]
                e += 1
        for a in range(3):           # [54, 63): n_a r0 q_b
            for b in range(3):
                acc[e] += n[k][a] * r0 * q[b]
                e += 1
        for ac in range(6):          # [63, 69): n_a n_c
            acc[e] += nn[ac]
            e += 1
        for a in range(3):           # [69, 72): n_a r0
            acc[e] += n[k][a] * r0
            e += 1
        acc[72] += r0 * r0
    return acc


def expand73(acc):
    """73 sums -> the symmetric 13x13 matrix sum phi phi^T."""
    M = np.zeros((13, 13))
    for a in range(3):
        for b in range(3):
            for c in range(3):
                for d in range(3):
                    M[3 * a + b, 3 * c + d] = acc[6 * SYM3[a][c] + SYM3[b][d]]
                M[3 * a + b, 9 + c] = M[9 + c, 3 * a + b] = acc[36 + 3 * SYM3[a][c] + b]
            M[3 * a + b, 12] = M[12, 3 * a + b] = acc[54 + 3 * a + b]
        for c in range(3):
            M[9 + a, 9 + c] = acc[63 + SYM3[a][c]]
        M[9 + a, 12] = M[12, 9 + a] = acc[69 + a]
    M[12, 12] = acc[72]
    return M


def accumulate_point(p_i, p_j, rel0):
    R0, t0 = rel0
    M = np.zeros((7, 7))
    for k in range(len(p_i)):
        q = R0 @ p_j[k] + t0
        z = np.concatenate([q, q - p_i[k], [1.0]])
        M += np.outer(z, z)
    return M


def delta_pose(rel, rel0):
    (R, t), (R0, t0) = rel, rel0
    dR = R @ R0.T
    return dR, t - dR @ t0


EPS = np.zeros((3, 3, 3))
EPS[0, 1, 2] = EPS[1, 2, 0] = EPS[2, 0, 1] = 1.0
EPS[0, 2, 1] = EPS[2, 1, 0] = EPS[1, 0, 2] = -1.0


def coeff_S(dR, dt):
    S = np.zeros((7, 13))
    for k in range(3):
        for a in range(3):
            for c in range(3):
                S[k, 3 * a + c] = sum(EPS[k, a, b] * dR[b, c] for b in range(3))  # (n x dR q0)_k
            S[k, 9 + a] = sum(EPS[k, a, b] * dt[b] for b in range(3))             # (n x dt)_k
        S[3 + k, 9 + k] = 1.0
    for a in range(3):
        for c in range(3):
            S[6, 3 * a + c] = dR[a, c] - (1.0 if a == c else 0.0)
        S[6, 9 + a] = dt[a]
    S[6, 12] = 1.0
    return S


def coeff_Z(dR, dt):
    Z = np.zeros((7, 7))
    for k in range(3):
        Z[k, k] = 1.0
        Z[k, 3 + k] = -1.0
        for c in range(3):
            Z[3 + k, c] = dR[k, c] - (1.0 if k == c else 0.0)
        Z[3 + k, 3 + k] = 1.0
        Z[3 + k, 6] = dt[k]
    Z[6, 6] = 1.0
    return Z


def basis(rel):
    """build_basis of linearize.cu: Bp[7][13], Bq[7][3][13] at the relative pose (R, t)."""
    R, t = rel
    K = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Bp = np.zeros((7, 13))
    Bq = np.zeros((7, 3, 13))
    for k in range(3):
        Bp[k, k] = 1.0
        Bp[3 + k, 3 + k] = -1.0
        for c in range(3):
            Bp[k, 6 + c] = -R[k, c]
            Bp[3 + k, 6 + c] = -(R.T @ K)[c, k]
            Bp[3 + k, 9 + c] = R[k, c]
        E = np.zeros((3, 3))
        k1, k2 = (k + 1) % 3, (k + 2) % 3
        E[k2, k1], E[k1, k2] = 1.0, -1.0
        ER = E @ R
        for r in range(3):
            for c in range(3):
                Bq[k, r, c] = E[r, c]
                Bq[k, r, 6 + c] = -ER[r, c]
                Bq[3 + k, r, 6 + c] = -ER[r, c]
        Bq[3 + k, k, 12] = -1.0
        Bq[6, k, 3 + k] = -1.0
        for c in range(3):
            Bq[6, k, 9 + c] = R[k, c]
    Bp[6, 12] = -1.0
    KR = K @ R
    for r in range(3):
        for c in range(3):
            Bq[6, r, 6 + c] = KR[r, c]
    return Bp, Bq


def evaluate(Mp, Mq, rel0, rel, sigma):
    dR, dt = delta_pose(rel, rel0)
    S, Z = coeff_S(dR, dt), coeff_Z(dR, dt)
    Wp, Wq = S @ Mp @ S.T, Z @ Mq @ Z.T
    Bp, Bq = basis(rel)
    H = Bp.T @ Wp @ Bp
    for r in range(3):
        H += Bq[:, r, :].T @ Wq @ Bq[:, r, :]
    err = 0.5 * (Wp[6, 6] + Wq[3, 3] + Wq[4, 4] + Wq[5, 5]) / sigma**2
    return H / sigma**2, err


def relative(Ti, Tj):
    Ri, ti = Ti["R"].reshape(3, 3), Ti["t"]
    Rj, tj = Tj["R"].reshape(3, 3), Tj["t"]
    return Ri.T @ Rj, Ri.T @ (tj - ti)


def _pose(xi):
    out = np.zeros((), dtype=_capi.POSE)
    T = expmap(xi)
    out["R"], out["t"] = T["R"], T["t"]
    return out


def _oracle_block(p_i, n_i, p_j, q_i, q_j, Ti, Tj, sigma):
    out, err = np.zeros(91), C.c_double()
    a, b = np.array([Ti], dtype=_capi.POSE), np.array([Tj], dtype=_capi.POSE)
    oracle_lib.lib().oracle_linearize_raw(_capi.ptr(p_i), _capi.ptr(n_i), _capi.ptr(p_j), len(p_i),
                                          _capi.ptr(q_i), _capi.ptr(q_j), len(q_i), _capi.ptr(a), _capi.ptr(b),
                                          sigma, _capi.ptr(out), C.byref(err))
    return out, err.value


IU = np.triu_indices(13)


@pytest.mark.parametrize("spread,drift", [(8.0, 0.0), (8.0, 1e-3), (60.0, 3e-3), (60.0, 0.3), (100.0, 1e-2)])
def test_moment_cache_reproduces_the_streamed_block(spread, drift):
    """Moments accumulated at one relative pose, evaluated at another (drift = size of the
    tangent-space step between the two) == the oracle's block at the second pose."""
    rng = np.random.default_rng(int(spread * 1000 + drift * 1e6))
    n, m, sigma = 300, 120, 0.1
    Ti0, Tj0 = _pose(rng.normal(size=6) * 0.4), _pose(rng.normal(size=6) * 0.4)
    Ri0, Rj0 = Ti0["R"].reshape(3, 3), Tj0["R"].reshape(3, 3)
    # correspondences that agree to a few centimetres at (Ti0, Tj0), as after an association
    world = rng.normal(size=(n, 3)) * spread
    p_i = (world - Ti0["t"]) @ Ri0
    p_j = (world + rng.normal(size=(n, 3)) * 0.03 - Tj0["t"]) @ Rj0
    n_i = rng.normal(size=(n, 3))
    n_i /= np.linalg.norm(n_i, axis=1, keepdims=True)
    worldq = rng.normal(size=(m, 3)) * spread
    q_i = (worldq - Ti0["t"]) @ Ri0
    q_j = (worldq + rng.normal(size=(m, 3)) * 0.03 - Tj0["t"]) @ Rj0
    # the kernels keep keypoints as float-exact values
    p_i, p_j, n_i, q_i, q_j = [np.ascontiguousarray(v.astype(np.float32).astype(np.float64))
                               for v in (p_i, p_j, n_i, q_i, q_j)]
    rel0 = relative(Ti0, Tj0)
    Mp = expand73(accumulate_planar73(p_i, n_i, p_j, rel0))
    Mq = accumulate_point(q_i, q_j, rel0)
    # evaluate at the association poses and at drifted poses
    for step in (0.0, drift):
        Ti = _pose(np.zeros(6))
        Tj = _pose(np.zeros(6))
        di, dj = expmap(rng.normal(size=6) * step), expmap(rng.normal(size=6) * step)
        Ti["R"] = (Ri0 @ di["R"].reshape(3, 3)).reshape(9)
        Ti["t"] = Ti0["t"] + Ri0 @ di["t"]
        Tj["R"] = (Rj0 @ dj["R"].reshape(3, 3)).reshape(9)
        Tj["t"] = Tj0["t"] + Rj0 @ dj["t"]
        H, err = evaluate(Mp, Mq, rel0, relative(Ti, Tj), sigma)
        want, want_err = _oracle_block(p_i, n_i, p_j, q_i, q_j, Ti, Tj, sigma)
        assert block_rel_err(H[IU], want) < 1e-9
        assert abs(err - want_err) <= 1e-9 * want_err


def test_moment_cache_planar_only_and_point_only():
    rng = np.random.default_rng(5)
    sigma = 0.1
    Ti, Tj = _pose(rng.normal(size=6) * 0.2), _pose(rng.normal(size=6) * 0.2)
    rel = relative(Ti, Tj)
    p_i, p_j = rng.normal(size=(50, 3)) * 10, rng.normal(size=(50, 3)) * 10
    n_i = rng.normal(size=(50, 3))
    n_i /= np.linalg.norm(n_i, axis=1, keepdims=True)
    empty = np.zeros((0, 3))
    H, err = evaluate(expand73(accumulate_planar73(p_i, n_i, p_j, rel)), np.zeros((7, 7)), rel, rel, sigma)
    want, want_err = _oracle_block(p_i, n_i, p_j, empty, empty, Ti, Tj, sigma)
    assert block_rel_err(H[IU], want) < 1e-12 and abs(err - want_err) <= 1e-12 * want_err
    H, err = evaluate(np.zeros((13, 13)), accumulate_point(p_i, p_j, rel), rel, rel, sigma)
    want, want_err = _oracle_block(empty, empty, empty, p_i, p_j, Ti, Tj, sigma)
    assert block_rel_err(H[IU], want) < 1e-12 and abs(err - want_err) <= 1e-12 * want_err
