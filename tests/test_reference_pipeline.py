"""The WHOLE pipeline against FORM's OWN code.

oracle/_ref/libformref.so also holds form/form.cpp (Estimator::register_scan) and
form/optimization/constraints.cpp (ConstraintManager: pair policy of get_graph, marginalize,
predict_next) compiled UNMODIFIED from /root/reference, over FORM's own extraction / map / matcher
/ factor / key-scanner sources and stand-ins for the missing libraries (oracle/shim; GTSAM's
optimiser is restated from its published behaviour in oracle/shim/gtsam/shim_smoother.h).  The
control flow is therefore the reference's: the ICP loop and its exit test, which factors enter the
fast and the full graph, when the one-off linear container is rebuilt, what is marginalised and in
which order the maps and constraints are erased.

These tests run that Estimator next to form_b200's host logic (form_b200/host/form/form.hpp +
constraints.hpp + keyscanner.hpp, the code the CUDA pipeline runs) over the CPU oracle, on the same
scans: keypoints must be byte-identical, the fixed-lag window must hold the same scans, and the
poses must agree to 1e-9 at every scan (observed: 2e-15) - which they can only do if both take the
same number of ICP iterations and LM steps and marginalise the same scans.  Together with the GPU
tests (CUDA pipeline == host logic over the oracle: identical keypoints and ICP / LM counts) this
pins SURVEY rows a25 / a26 and the north star's trajectory criterion to the reference's own code.
CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
import test_reference_pins as pins
from form_b200 import _capi, synth

_vp, _sz, _i, _d, _i64 = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int64


def ref():
    lib = pins.ref()
    if not hasattr(lib, "_est_ready"):
        lib.formref_est_create.restype = _vp
        lib.formref_est_create.argtypes = [C.POINTER(_capi.Params), _d, _d, _i, _i, _i64, _sz, _i64]
        lib.formref_est_destroy.argtypes = [_vp]
        lib.formref_est_register_scan.restype = _i
        lib.formref_est_register_scan.argtypes = [_vp, _vp, _sz, _vp, _sz, C.POINTER(_sz), _vp, _sz, C.POINTER(_sz)]
        lib.formref_est_pose.argtypes = [_vp, _vp]
        lib.formref_est_window.restype = _sz
        lib.formref_est_window.argtypes = [_vp, _vp, _sz]
        lib._est_ready = True
    return lib


class FormEstimator:
    """FORM's own form::Estimator (reference code)."""

    def __init__(self, p: _capi.EstParams):
        self.n = p.hot.num_rows * p.hot.num_columns
        self.h = ref().formref_est_create(C.byref(p.hot), p.new_pose_threshold, p.keyscan_match_ratio,
                                          p.max_num_rematches, p.disable_smoothing, p.max_num_keyscans,
                                          p.max_num_recent_scans, p.max_steps_unused_keyscan)
        self.pl = np.zeros(self.n, dtype=_capi.PLANAR_FEAT)
        self.pt = np.zeros(self.n, dtype=_capi.POINT_FEAT)

    def __del__(self):
        ref().formref_est_destroy(self.h)

    def register_scan(self, scan):
        a, b = _sz(), _sz()
        rc = ref().formref_est_register_scan(self.h, _capi.ptr(scan), self.n, _capi.ptr(self.pl), self.n, C.byref(a),
                                             _capi.ptr(self.pt), self.n, C.byref(b))
        assert rc == 0
        return self.pl[: a.value].copy(), self.pt[: b.value].copy()

    def pose(self):
        out = np.zeros(1, dtype=_capi.POSE)
        ref().formref_est_pose(self.h, _capi.ptr(out))
        return out[0]

    def window(self):
        out = np.zeros(256, dtype=_capi.SCAN_POSE)
        n = ref().formref_est_window(self.h, _capi.ptr(out), len(out))
        return out[:n].copy()


def compare(sensor, n_scans, tol, seq=0, check_every_pose=True, **over):
    rows, cols = synth.shape(sensor)
    p = _capi.default_est_params(rows, cols, num_threads=1, **over)
    ours = oracle_lib.OracleEstimator(p)
    theirs = FormEstimator(p)
    worst = 0.0
    sizes = set()
    reorders = []  # scans whose keypoints came out in another order (equal-curvature ties)
    for k in range(n_scans):
        scan = synth.scan(sensor, seq, k)
        a, b = ours.register_scan(scan)
        c, d = theirs.register_scan(scan)
        for mine, ref_ in ((a, c), (b, d)):
            if mine.tobytes() != ref_.tobytes():
                # the only licensed difference: FORM's unstable std::sort may emit two picks of EQUAL
                # curvature in either order (extraction.tpp:57-58; rule R1 of SURVEY A.1 fixes index
                # order) - the keypoints must then be the same set
                assert len(mine) == len(ref_) and sorted(r.tobytes() for r in mine) == sorted(r.tobytes() for r in ref_), \
                    f"keypoints differ at scan {k}"
                reorders.append(k)
        wo, wt = ours.window(), theirs.window()
        assert list(wo["scan"]) == list(wt["scan"]), f"window differs at scan {k}: {wo['scan']} vs {wt['scan']}"
        sizes.add(len(wo))
        for x, y in zip(wo, wt) if check_every_pose else [(ours.pose(), theirs.pose())]:
            worst = max(worst, float(np.linalg.norm(x["t"] - y["t"])), float(np.abs(x["R"] - y["R"]).max()))
        assert worst < tol, f"poses differ by {worst} at scan {k}"
    ours.tie_reorders = reorders
    return ours, worst, sizes


def test_default_parameters_reference_schedule():
    """GTSAM's LM schedule on both sides: 36 VLP-16 scans, the window fills (11 scans), scans fall
    out of it and are marginalised or promoted to key scans."""
    ours, worst, sizes = compare("vlp-16", 36, 1e-9, gtsam_lm_schedule=1)
    assert max(sizes) >= 12 and worst < 1e-12
    st = ours.stats()
    assert st["icp_iterations"] > 36 and st["error_calls"] > 0


def test_fused_trial_linearisation_is_the_same_estimator():
    """The product's default schedule (trial steps linearised, blocks reused) against FORM's own
    code: same iterates up to the rounding of f/2 vs the error evaluation."""
    ours, worst, _ = compare("vlp-16", 24, 1e-9)
    assert ours.stats()["error_calls"] == 0


def test_key_scan_churn():
    """Small window, key scans that are promoted, age out and hit the cap: several scans are
    marginalised at once and marginal factors are re-marginalised."""
    ours, worst, sizes = compare("vlp-16", 40, 1e-9, gtsam_lm_schedule=1, max_num_recent_scans=3, max_num_keyscans=3,
                                 max_steps_unused_keyscan=2, keyscan_match_ratio=0.01)
    assert max(sizes) <= 8 and worst < 1e-11


def test_disable_smoothing_ablation():
    compare("vlp-16", 16, 1e-9, gtsam_lm_schedule=1, disable_smoothing=1)


def test_no_point_features_ablation():
    compare("vlp-16", 12, 1e-9, gtsam_lm_schedule=1, point_feats_per_sector=0)


def test_os1_64_sequence():
    """BASELINE configs[0]'s sensor shape (64 x 1024), another seeded sequence."""
    compare("os1-64", 8, 1e-9, seq=3, gtsam_lm_schedule=1)


def test_equal_curvature_tie_is_the_only_licensed_difference():
    """Scan 40 of this sequence holds two planar picks of bit-identical curvature in one sector:
    FORM's std::sort emits them in one order, rule R1 (ties by column) in the other.  Same set,
    same window, poses unaffected - and nothing else differs anywhere in the 42 scans."""
    ours, worst, _ = compare("vlp-16", 42, 1e-9, seq=5, new_pose_threshold=1e-6, max_num_rematches=5)
    assert ours.tie_reorders == [40] and worst < 1e-12
