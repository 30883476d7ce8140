"""Replays the frozen golden sequence (tests/golden/hotpath_8x384.npz) on any backend
exposing the Context / Oracle call set and compares with the frozen outputs."""
import os

import numpy as np

from form_b200 import _capi
from helpers import block_rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_8x384.npz")
ROWS, COLS = 8, 384


def check_backend(make_backend, exact_blocks: bool, debug: bool = True):
    g = np.load(GOLDEN)
    params = _capi.default_params(ROWS, COLS)
    b = make_backend(params)
    worst = 0.0
    for k in range(3):
        pl, pt = b.extract(np.ascontiguousarray(g[f"scan{k}"]), k)
        assert pl.tobytes() == g[f"planar{k}"].tobytes(), f"scan {k} planar"
        assert pt.tobytes() == g[f"point{k}"].tobytes(), f"scan {k} point"
        if debug:
            d = b.extract_debug()
            for name in ("valid", "point_valid", "planar_indices", "planar_keep", "closest_prev", "closest_next",
                         "point_indices"):
                assert np.array_equal(d[name], g[f"{name}{k}"]), (k, name)
            assert np.array_equal(d["curvature"].view(np.uint32), g[f"curvature{k}"].view(np.uint32))
        b.map_rebuild(np.ascontiguousarray(g[f"map_poses{k}"]))
        counts = b.associate(np.ascontiguousarray(g[f"pose_k{k}"])[0])
        assert counts.tobytes() == g[f"counts{k}"].tobytes(), f"scan {k} counts"
        for t, name in ((0, "matches_planar"), (1, "matches_point")):
            assert b.matches(t).tobytes() == g[f"{name}{k}"].tobytes(), f"scan {k} {name}"
        if f"blocks{k}" in g:
            H = b.linearize(np.ascontiguousarray(g[f"lin_pairs{k}"]), np.ascontiguousarray(g[f"lin_poses{k}"]))
            e = b.error(np.ascontiguousarray(g[f"lin_pairs{k}"]), np.ascontiguousarray(g[f"lin_poses{k}"]))
            if exact_blocks:
                assert np.array_equal(H, g[f"blocks{k}"]) and np.array_equal(e, g[f"errors{k}"])
            else:
                for a, r in zip(H, g[f"blocks{k}"]):
                    worst = max(worst, block_rel_err(a, r))
                assert np.allclose(e, g[f"errors{k}"], rtol=1e-9)
        assert tuple(b.commit_scan()) == tuple(g[f"added{k}"])
        assert b.keypoints(0, k).tobytes() == g[f"stored_planar{k}"].tobytes()
        assert b.keypoints(1, k).tobytes() == g[f"stored_point{k}"].tobytes()
    assert worst < 1e-9, worst
    return worst
