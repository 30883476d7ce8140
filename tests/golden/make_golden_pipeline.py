"""Generates tests/golden/form_pipeline_vlp16.npz from the REFERENCE'S OWN pipeline.

oracle/_ref/libformref.so holds FORM's Estimator::register_scan compiled unmodified from
/root/reference (form.cpp, constraints.cpp and all stage sources, over the stand-ins of
oracle/shim).  This script runs it on a seeded synthetic VLP-16 sequence and freezes, per scan,
the number of planar / point keypoints, a CRC of their bytes, the scans of the fixed-lag window and
the estimated pose - golden vectors that travel to boxes where neither /root/reference nor the
prebuilt library exists.  Run from the repository root (in the container that has /root/reference):

    python tests/golden/make_golden_pipeline.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from form_b200 import _capi, synth  # noqa: E402
from test_reference_pipeline import FormEstimator  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SENSOR, SEQ, N_SCANS = "vlp-16", 7, 30
# a small window so that promotion to key scan, ageing out and marginalisation all happen within 30 scans
OVERRIDES = dict(max_num_recent_scans=4, max_num_keyscans=4, max_steps_unused_keyscan=3, keyscan_match_ratio=0.02)


def main():
    rows, cols = synth.shape(SENSOR)
    p = _capi.default_est_params(rows, cols, num_threads=1, gtsam_lm_schedule=1, **OVERRIDES)
    est = FormEstimator(p)
    n_planar, n_point, crc, poses, windows = [], [], [], [], []
    for k in range(N_SCANS):
        pl, pt = est.register_scan(synth.scan(SENSOR, SEQ, k))
        n_planar.append(len(pl))
        n_point.append(len(pt))
        crc.append(zlib.crc32(pt.tobytes(), zlib.crc32(pl.tobytes())))
        pose = est.pose()
        poses.append(np.concatenate([pose["R"], pose["t"]]))
        w = np.full(16, -1, dtype=np.int64)
        ws = est.window()["scan"]
        w[: len(ws)] = ws
        windows.append(w)
    np.savez_compressed(os.path.join(HERE, "form_pipeline_vlp16.npz"), sensor=SENSOR, sequence=SEQ,
                        overrides=np.array(sorted(OVERRIDES.items()), dtype=object).astype(str),
                        n_planar=np.array(n_planar), n_point=np.array(n_point), crc=np.array(crc, dtype=np.uint32),
                        poses=np.array(poses), windows=np.array(windows))
    print("wrote form_pipeline_vlp16.npz:", N_SCANS, "scans, window sizes", sorted({int((w >= 0).sum()) for w in windows}))


if __name__ == "__main__":
    main()
