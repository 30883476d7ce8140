"""The CUDA pipeline against FORM's OWN pipeline: form::Estimator over the CUDA hot path next to
the reference's Estimator::register_scan (form.cpp + constraints.cpp and all of FORM's stage code,
compiled unmodified into oracle/_ref over the stand-ins of oracle/shim) on the same scans -
byte-identical keypoints, the same scans in the fixed-lag window, trajectory difference far below
the north star's 1 mm ATE criterion.  (tests/test_reference_pipeline.py is the CPU half of the
chain: the host logic over the oracle against the same reference code, to 4e-15.)"""
import numpy as np
import pytest

from form_b200 import _capi, synth

pytestmark = pytest.mark.gpu

ATE_TOL_M = 1e-3


def _run(sensor, n_scans, seq=0, **overrides):
    from form_b200.pipeline import Estimator
    from test_reference_pipeline import FormEstimator

    rows, cols = synth.shape(sensor)
    p = _capi.default_est_params(rows, cols, **overrides)
    gpu, form = Estimator(p), FormEstimator(p)
    dt = []
    for k in range(n_scans):
        scan = synth.scan(sensor, seq, k)
        pl, pt = gpu.register_scan(scan)
        rpl, rpt = form.register_scan(scan)
        assert pl.tobytes() == rpl.tobytes(), f"scan {k}: planar keypoints"
        assert pt.tobytes() == rpt.tobytes(), f"scan {k}: point keypoints"
        w, rw = gpu.window(), form.window()
        assert np.array_equal(w["scan"], rw["scan"]), f"scan {k}: window"
        assert np.max(np.abs(w["t"] - rw["t"])) < ATE_TOL_M and np.max(np.abs(w["R"] - rw["R"])) < 1e-6, f"scan {k}"
        dt.append(float(np.linalg.norm(gpu.pose()["t"] - form.pose()["t"])))
    assert float(np.sqrt(np.mean(np.square(dt)))) < ATE_TOL_M


def test_cuda_pipeline_matches_forms_own_estimator_vlp16():
    _run("vlp-16", 20)


def test_cuda_pipeline_matches_forms_own_estimator_os1_64_and_ablation():
    _run("os1-64", 6, seq=2)
    _run("vlp-16", 8, disable_smoothing=1)
