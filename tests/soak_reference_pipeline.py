"""Soak of the pipeline pin (tests/test_reference_pipeline.py) beyond what the CPU suite has time
for: FORM's OWN Estimator::register_scan (oracle/_ref) next to this repository's host logic over
the oracle on longer sequences, more seeds and corner-case key-scan parameters.  CPU only; needs
oracle/_ref (i.e. the container that has /root/reference).  Log of the round: profiles/r05/reference_pipeline_soak.txt

    python tests/soak_reference_pipeline.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # test infrastructure: the oracle is only used from tests/

import test_reference_pipeline as t  # noqa: E402

CASES = [
    ("os1-64", 200, dict(seq=0, gtsam_lm_schedule=1)),  # BASELINE configs[0]: 200 scans
    ("vlp-16", 120, dict(seq=1)),
    ("vlp-16", 120, dict(seq=2, gtsam_lm_schedule=1, keyscan_match_ratio=0.02)),  # the window fills up to 60 scans
    ("vlp-16", 80, dict(seq=3, max_num_recent_scans=2, max_num_keyscans=2, max_steps_unused_keyscan=1, keyscan_match_ratio=0.0)),
    ("vlp-16", 80, dict(seq=4, max_num_recent_scans=6, max_num_keyscans=0, max_steps_unused_keyscan=5, keyscan_match_ratio=0.01)),
    ("vlp-16", 60, dict(seq=5, new_pose_threshold=1e-6, max_num_rematches=5)),  # holds an equal-curvature tie (scan 40)
    ("os0-128", 30, dict(seq=6)),
    ("vlp-16", 100, dict(seq=8, gtsam_lm_schedule=1, disable_smoothing=1)),
    ("os1-64", 60, dict(seq=9, point_feats_per_sector=0)),
]

# twelve more seeds with mixed parameters (LM schedule, key-scan limits, feature density, matching radius)
for _seq in range(20, 32):
    _kw = dict(seq=_seq)
    if _seq % 2:
        _kw["gtsam_lm_schedule"] = 1
    if _seq % 3 == 0:
        _kw.update(max_num_recent_scans=5, max_num_keyscans=6, max_steps_unused_keyscan=4, keyscan_match_ratio=0.03)
    if _seq % 5 == 0:
        _kw.update(planar_feats_per_sector=20, radius=0.5)
    if _seq % 7 == 0:
        _kw.update(max_dist_matching=0.5, min_dist_map=0.2)
    CASES.append(("vlp-16" if _seq % 4 else "os1-64", 70 if _seq % 4 else 30, _kw))

# the benchmarked shape (configs[1])
CASES += [("os0-128", 40, dict(seq=40)), ("os0-128", 40, dict(seq=41, gtsam_lm_schedule=1)),
          ("os0-128", 40, dict(seq=42, max_num_recent_scans=4, max_num_keyscans=4, max_steps_unused_keyscan=3,
                               keyscan_match_ratio=0.02))]

if __name__ == "__main__":
    for sensor, n, kw in CASES:
        t0 = time.time()
        try:
            ours, worst, sizes = t.compare(sensor, n, 1e-9, **kw)
            print(sensor, n, kw, "| worst pose difference", worst, "| window sizes", min(sizes), "-", max(sizes),
                  "| ICP iterations", ours.stats()["icp_iterations"], "| tie reorders at scans", ours.tie_reorders,
                  "|", round(time.time() - t0), "s", flush=True)
        except AssertionError as e:
            print("FAIL", sensor, n, kw, e, flush=True)
