"""The host-side fixed-lag smoother (form_b200/host/form/constraints.hpp, the restatement of
/root/reference/form/optimization/constraints.cpp:39-336 over GTSAM's published semantics)
against an INDEPENDENT formulation of the same problem: the residuals written from their
definitions in numpy (factor.cpp:30-128: plane r = (R_i n_i).(T_j p_j - T_i p_i), point
r = T_j p_j - T_i p_i, prior r = Logmap(prior^-1 T)), minimised by scipy's least_squares, and the
marginal factors against a numpy Schur complement of finite-difference Jacobians.  GTSAM itself
does not exist in this image; what these tests pin is that the smoother finds the
optimum of the reference's cost function and that its marginals are the Gaussian marginals of
the dropped factors.  CPU only; the smoother is driven through oracle/oracle_smoother_capi.cpp
over hand-made correspondences."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.optimize import least_squares

import oracle_lib
from form_b200 import _capi

SIGMA, POSE_SIGMA = 0.1, 1e-3
_vp, _sz, _u64, _i, _d = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_double
_SYMS = {
    "oracle_smoother_create": (_vp, [_d, _d, _i, _i, _d, _d]),
    "oracle_smoother_destroy": (None, [_vp]),
    "oracle_smoother_error": (C.c_char_p, [_vp]),
    "oracle_smoother_set_pair": (None, [_vp, _u64, _u64, _vp, _vp, _vp, _sz, _vp, _vp, _sz]),
    "oracle_smoother_step": (_u64, [_vp, _vp]),
    "oracle_smoother_optimize": (_i, [_vp, _i]),
    "oracle_smoother_marginalize": (_i, [_vp, _vp, _sz]),
    "oracle_smoother_set_pose": (None, [_vp, _u64, _vp]),
    "oracle_smoother_values": (_sz, [_vp, _vp, _sz]),
    "oracle_smoother_stats": (None, [_vp, _vp]),
    "oracle_smoother_num_marginals": (_sz, [_vp]),
    "oracle_smoother_marginal_keys": (_sz, [_vp, _sz, _vp, _vp, _sz]),
    "oracle_smoother_marginal_quadratic": (None, [_vp, _sz, _vp, _vp, _vp]),
}


def L():
    lib = oracle_lib.lib()
    if not getattr(lib, "_smoother_ready", False):
        for name, (res, args) in _SYMS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib._smoother_ready = True
    return lib


# ---------------------------------------------------------------- SE(3), GTSAM conventions
def hat(xi):
    w, v = xi[:3], xi[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


def Exp(xi):
    return expm(hat(np.asarray(xi, dtype=np.float64)))


def Log(T):
    M = np.real(logm(T))
    return np.array([M[2, 1], M[0, 2], M[1, 0], M[0, 3], M[1, 3], M[2, 3]])


def to_pose(T):
    out = np.zeros(1, dtype=_capi.POSE)
    out["R"] = T[:3, :3].reshape(9)
    out["t"] = T[:3, 3]
    return out


def from_pose(p):
    T = np.eye(4)
    T[:3, :3] = np.asarray(p["R"]).reshape(3, 3)
    T[:3, 3] = p["t"]
    return T


# ---------------------------------------------------------------- a small world
class Problem:
    """K poses looking at random planes and points; pair (i, j) carries point-to-plane rows (a
    point of scan j against a DIFFERENT point of the same plane seen from scan i) and
    point-to-point rows, both with measurement noise."""

    def __init__(self, K, pairs, seed, n_plane=60, n_point=12, noise=0.01):
        rng = np.random.default_rng(seed)
        self.K = K
        self.truth = [np.eye(4)]
        for _ in range(1, K):
            step = np.concatenate([rng.normal(size=3) * 0.03, [0.5, 0.0, 0.0] + rng.normal(size=3) * 0.1])
            self.truth.append(self.truth[-1] @ Exp(step))
        self.pairs = {}
        for (i, j) in pairs:
            Ti, Tj = self.truth[i], self.truth[j]
            normals = rng.normal(size=(n_plane, 3))
            normals /= np.linalg.norm(normals, axis=1, keepdims=True)
            anchor = rng.uniform(-15, 15, size=(n_plane, 3))
            # two different points of each plane
            def on_plane():
                d = rng.uniform(-3, 3, size=(n_plane, 3))
                return anchor + d - (np.sum(d * normals, axis=1, keepdims=True)) * normals
            xi_w, xj_w = on_plane(), on_plane()
            inv = np.linalg.inv
            pl_pi = (inv(Ti)[:3, :3] @ xi_w.T).T + inv(Ti)[:3, 3]
            pl_ni = (Ti[:3, :3].T @ normals.T).T
            pl_pj = (inv(Tj)[:3, :3] @ xj_w.T).T + inv(Tj)[:3, 3] + rng.normal(size=(n_plane, 3)) * noise
            xq = rng.uniform(-15, 15, size=(n_point, 3))
            pt_pi = (inv(Ti)[:3, :3] @ xq.T).T + inv(Ti)[:3, 3]
            pt_pj = (inv(Tj)[:3, :3] @ xq.T).T + inv(Tj)[:3, 3] + rng.normal(size=(n_point, 3)) * noise
            self.pairs[(i, j)] = tuple(np.ascontiguousarray(a) for a in (pl_pi, pl_ni, pl_pj, pt_pi, pt_pj))
        # what the estimator would start from: the truth, perturbed (pose 0 is where the prior sits)
        self.initial = [self.truth[0]]
        for k in range(1, K):
            self.initial.append(self.truth[k] @ Exp(np.concatenate([rng.normal(size=3) * 0.02, rng.normal(size=3) * 0.05])))

    # -- the reference's cost, written from the definitions --
    def pair_residuals(self, key, Ti, Tj):
        pl_pi, pl_ni, pl_pj, pt_pi, pt_pj = self.pairs[key]
        wi = (Ti[:3, :3] @ pl_pi.T).T + Ti[:3, 3]
        wj = (Tj[:3, :3] @ pl_pj.T).T + Tj[:3, 3]
        wn = (Ti[:3, :3] @ pl_ni.T).T
        r_plane = np.sum(wn * (wj - wi), axis=1)
        qi = (Ti[:3, :3] @ pt_pi.T).T + Ti[:3, 3]
        qj = (Tj[:3, :3] @ pt_pj.T).T + Tj[:3, 3]
        return np.concatenate([r_plane, (qj - qi).reshape(-1)]) / SIGMA

    def residuals(self, poses, keys=None, prior=None, fixed=None):
        """poses: dict scan -> 4x4.  keys: the pairs to include.  prior: (scan, 4x4) or None."""
        out = []
        if prior is not None:
            out.append(Log(np.linalg.inv(prior[1]) @ poses[prior[0]]) / POSE_SIGMA)
        for key in (self.pairs if keys is None else keys):
            out.append(self.pair_residuals(key, poses[key[0]], poses[key[1]]))
        return np.concatenate(out)


class Smoother:
    def __init__(self, problem, fused=1, disable_smoothing=0, rel_tol=1e-13, abs_tol=1e-13):
        self.h = L().oracle_smoother_create(SIGMA, POSE_SIGMA, fused, disable_smoothing, rel_tol, abs_tol)
        for (i, j), arrs in problem.pairs.items():
            pl_pi, pl_ni, pl_pj, pt_pi, pt_pj = arrs
            L().oracle_smoother_set_pair(self.h, i, j, _capi.ptr(pl_pi), _capi.ptr(pl_ni), _capi.ptr(pl_pj),
                                         len(pl_pi), _capi.ptr(pt_pi), _capi.ptr(pt_pj), len(pt_pi))

    def __del__(self):
        L().oracle_smoother_destroy(self.h)

    def step(self, T):
        return L().oracle_smoother_step(self.h, _capi.ptr(to_pose(T)))

    def optimize(self, fast=False):
        rc = L().oracle_smoother_optimize(self.h, int(fast))
        assert rc == 0, L().oracle_smoother_error(self.h)

    def marginalize(self, scans):
        a = np.asarray(scans, dtype=np.uint64)
        rc = L().oracle_smoother_marginalize(self.h, _capi.ptr(a), len(a))
        assert rc == 0, L().oracle_smoother_error(self.h)

    def values(self):
        out = np.zeros(64, dtype=_capi.SCAN_POSE)
        n = L().oracle_smoother_values(self.h, _capi.ptr(out), len(out))
        return {int(e["scan"]): from_pose(e) for e in out[:n]}

    def stats(self):
        out = np.zeros(6, dtype=np.uint64)
        L().oracle_smoother_stats(self.h, _capi.ptr(out))
        return dict(zip(("optimize", "lm_iterations", "linearize", "error", "lin_pairs", "err_pairs"), map(int, out)))

    def marginals(self):
        res = []
        for idx in range(L().oracle_smoother_num_marginals(self.h)):
            keys = np.zeros(64, dtype=np.uint64)
            lin = np.zeros(64, dtype=_capi.POSE)
            n = L().oracle_smoother_marginal_keys(self.h, idx, _capi.ptr(keys), _capi.ptr(lin), 64)
            G, g, f = np.zeros((6 * n, 6 * n)), np.zeros(6 * n), C.c_double()
            L().oracle_smoother_marginal_quadratic(self.h, idx, _capi.ptr(G), _capi.ptr(g), C.byref(f))
            res.append(([int(k) for k in keys[:n]], [from_pose(p) for p in lin[:n]], G, g, f.value))
        return res


def scipy_optimum(problem, base, variables, keys=None, prior=None):
    """argmin over local coordinates of `variables` around base (T = base . Exp(x))."""
    def fun(x):
        poses = dict(base)
        for k, s in enumerate(variables):
            poses[s] = base[s] @ Exp(x[6 * k:6 * k + 6])
        return problem.residuals(poses, keys, prior)
    sol = least_squares(fun, np.zeros(6 * len(variables)), method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15,
                        x_scale=1.0, max_nfev=20000)
    poses = dict(base)
    for k, s in enumerate(variables):
        poses[s] = base[s] @ Exp(sol.x[6 * k:6 * k + 6])
    return poses, 0.5 * float(np.sum(sol.fun ** 2))


def pose_diff(A, B):
    d = Log(np.linalg.inv(A) @ B)
    return float(np.linalg.norm(d[:3])), float(np.linalg.norm(d[3:]))


ALL_PAIRS_4 = [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3)]


def drive(problem, sm):
    for k in range(problem.K):
        assert sm.step(problem.initial[k]) == k


@pytest.mark.parametrize("fused", [1, 0])
def test_lm_finds_the_optimum_of_the_reference_cost(fused):
    pr = Problem(4, ALL_PAIRS_4, seed=3)
    sm = Smoother(pr, fused=fused)
    drive(pr, sm)
    sm.optimize(fast=False)
    got = sm.values()
    base = {k: pr.initial[k] for k in range(pr.K)}
    want, cost = scipy_optimum(pr, base, list(range(pr.K)), prior=(0, pr.initial[0]))
    for k in range(pr.K):
        drot, dtr = pose_diff(want[k], got[k])
        assert drot < 1e-8 and dtr < 1e-7, (k, drot, dtr)
    # and it is a sensible estimate: closer to the truth than the start was
    assert max(pose_diff(pr.truth[k], got[k])[1] for k in range(pr.K)) < 0.01
    st = sm.stats()
    assert st["lm_iterations"] >= 3
    if fused:  # every step is a linearisation, no separate error evaluations
        assert st["error"] == 0 and st["linearize"] >= st["lm_iterations"] + 1
    else:      # GTSAM's schedule: one linearisation per iteration + one error per trial (+ the initial one)
        assert st["error"] >= st["lm_iterations"] + 1 and st["linearize"] >= st["lm_iterations"]


def test_fused_and_gtsam_schedules_take_the_same_steps():
    pr = Problem(4, ALL_PAIRS_4, seed=5)
    out = []
    for fused in (1, 0):
        sm = Smoother(pr, fused=fused, rel_tol=0, abs_tol=0)  # GTSAM's default tolerances
        drive(pr, sm)
        sm.optimize(fast=False)
        out.append((sm.values(), sm.stats()["lm_iterations"]))
    assert out[0][1] == out[1][1]
    for k in range(pr.K):
        drot, dtr = pose_diff(out[0][0][k], out[1][0][k])
        assert drot < 1e-10 and dtr < 1e-10


def numeric_jacobian(fun, n, h=1e-6):
    r0 = fun(np.zeros(n))
    J = np.zeros((len(r0), n))
    for c in range(n):
        e = np.zeros(n)
        e[c] = h
        J[:, c] = (fun(e) - fun(-e)) / (2 * h)
    return r0, J


def test_marginal_is_the_schur_complement_of_the_dropped_factors():
    pr = Problem(4, ALL_PAIRS_4, seed=7)
    sm = Smoother(pr)
    drive(pr, sm)
    sm.optimize(fast=False)
    before = sm.values()
    sm.marginalize([0])
    assert sorted(sm.values()) == [1, 2, 3]
    (keys, lin, G, g, f), = sm.marginals()
    assert keys == [1, 2, 3]
    for k, T in zip(keys, lin):
        assert np.allclose(T, before[k], atol=0, rtol=0)

    # the factors that touch scan 0: its prior and the pairs (0, j), linearised at `before`
    dropped = [key for key in pr.pairs if key[0] == 0]
    order = [0, 1, 2, 3]

    def fun(x):
        poses = {s: before[s] @ Exp(x[6 * k:6 * k + 6]) for k, s in enumerate(order)}
        return pr.residuals(poses, dropped, prior=(0, pr.initial[0]))
    r0, J = numeric_jacobian(fun, 24)
    H, b, f0 = J.T @ J, -J.T @ r0, float(r0 @ r0)
    Haa, Hac, Hcc = H[:6, :6], H[:6, 6:], H[6:, 6:]
    X = np.linalg.solve(Haa, np.column_stack([Hac, b[:6]]))
    G_want = Hcc - Hac.T @ X[:, :-1]
    g_want = b[6:] - Hac.T @ X[:, -1]
    f_want = f0 - b[:6] @ X[:, -1]
    scale = np.sqrt(np.outer(np.diag(G_want), np.diag(G_want)))
    assert np.max(np.abs(G - G_want) / scale) < 1e-6
    assert np.max(np.abs(g - g_want) / np.sqrt(np.diag(G_want))) < 1e-4 * max(1.0, np.sqrt(f_want))
    assert abs(f - f_want) <= 1e-6 * max(1.0, abs(f_want))

    # marginalising at the optimum leaves the optimum of the remaining window where it was
    sm.optimize(fast=False)
    after = sm.values()
    for k in (1, 2, 3):
        drot, dtr = pose_diff(before[k], after[k])
        assert drot < 1e-7 and dtr < 1e-6, (k, drot, dtr)


def test_fixed_lag_window_tracks_the_batch_solution():
    """Six scans, lag 3: every scan is optimised with the older ones marginalised out.  With
    factors that are nearly linear around the estimates the fixed-lag poses must stay close to
    the batch optimum over all scans (the marginals carry what the dropped scans knew)."""
    K = 6
    pairs = [(i, j) for j in range(K) for i in range(max(0, j - 2), j)]
    pr = Problem(K, pairs, seed=11, noise=0.005)
    base = {k: pr.initial[k] for k in range(K)}
    batch, _ = scipy_optimum(pr, base, list(range(K)), prior=(0, pr.initial[0]))
    sm = Smoother(pr)
    for k in range(K):
        sm.step(pr.initial[k])
        sm.optimize(fast=False)
        if k >= 3:
            sm.marginalize([k - 3])
    got = sm.values()
    assert sorted(got) == [3, 4, 5]
    assert len(sm.marginals()) >= 1
    for k in got:
        drot, dtr = pose_diff(batch[k], got[k])
        assert drot < 1e-6 and dtr < 1e-5, (k, drot, dtr)


def test_fast_graph_matches_full_graph_when_old_pairs_are_already_converged():
    """optimize(true) replaces the older pairs by ONE linearisation taken at the current values
    (constraints.cpp:268-288); started at the optimum of those pairs it must land where the full
    nonlinear graph lands, up to the curvature the linearisation drops."""
    pr = Problem(4, ALL_PAIRS_4, seed=13)
    full = Smoother(pr)
    fast = Smoother(pr)
    for sm in (full, fast):
        for k in range(3):
            sm.step(pr.initial[k])
        sm.optimize(fast=False)  # scans 0..2 converged
        sm.step(pr.initial[3])
    full.optimize(fast=False)
    fast.optimize(fast=True)
    a, b = full.values(), fast.values()
    for k in range(4):
        drot, dtr = pose_diff(a[k], b[k])
        assert drot < 1e-4 and dtr < 1e-3, (k, drot, dtr)


def test_disable_smoothing_optimises_the_current_pose_only():
    pr = Problem(4, ALL_PAIRS_4, seed=17)
    sm = Smoother(pr, disable_smoothing=1)
    drive(pr, sm)
    sm.optimize(fast=True)
    got = sm.values()
    for k in range(3):  # older poses are constants of the single-pose graph (constraints.cpp:235-250)
        assert np.array_equal(got[k], pr.initial[k])
    base = {k: pr.initial[k] for k in range(4)}
    want, _ = scipy_optimum(pr, base, [3], keys=[(0, 3), (1, 3), (2, 3)])
    drot, dtr = pose_diff(want[3], got[3])
    assert drot < 1e-8 and dtr < 1e-7
