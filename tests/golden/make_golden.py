"""Generates tests/golden/*.npz from the CPU oracle on seeded synthetic input.

The reference ships no golden vectors for this path and cannot be built or imported here
(SURVEY 8c), so these files freeze the oracle's own output (after it has been pinned by the
numpy restatement / finite-difference / hand-computed tests) so later refactors of the
oracle or the CUDA path cannot drift silently.  Run from the repository root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from form_b200 import _capi  # noqa: E402
from helpers import perturbed, scan_poses  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROWS, COLS = 8, 384


def make_scan(rng, k):
    az = np.tile(np.linspace(0, 2 * np.pi, COLS, endpoint=False), ROWS)
    el = np.repeat(np.linspace(-0.35, 0.35, ROWS), COLS)
    # box-ish room seen from a slowly moving sensor, with dropouts
    r = 4.0 / np.maximum(np.abs(np.cos(az + 0.02 * k)), np.abs(np.sin(az + 0.02 * k))) / np.cos(el)
    r = np.minimum(r, 1.4 / np.maximum(np.abs(np.sin(el)), 1e-3))
    r = r + 0.01 * rng.standard_normal(ROWS * COLS)
    r[rng.uniform(size=ROWS * COLS) < 0.02] = 0.0
    scan = np.zeros(ROWS * COLS, dtype=_capi.POINT4F)
    scan["x"] = (r * np.cos(el) * np.cos(az)).astype(np.float32)
    scan["y"] = (r * np.cos(el) * np.sin(az)).astype(np.float32)
    scan["z"] = (r * np.sin(el)).astype(np.float32)
    return scan


def main():
    rng = np.random.default_rng(20251018)
    params = _capi.default_params(ROWS, COLS)
    o = oracle_lib.Oracle(params, threads=1)
    ident = np.zeros((), dtype=_capi.POSE)
    ident["R"] = np.eye(3).reshape(9)
    out = {}
    poses = {}
    for k in range(3):
        scan = make_scan(rng, k)
        out[f"scan{k}"] = scan
        pl, pt = o.extract(scan, k)
        d = o.extract_debug()
        out[f"planar{k}"], out[f"point{k}"] = pl, pt
        for name in ("valid", "point_valid", "curvature", "planar_indices", "planar_keep", "closest_prev",
                     "closest_next", "point_indices"):
            out[f"{name}{k}"] = d[name]
        poses[k] = perturbed(ident, rng, 0.01, 0.05) if k else ident
        sp = scan_poses(list(poses), [poses[s] for s in poses])
        out[f"map_poses{k}"] = sp
        o.map_rebuild(sp)
        pose_k = perturbed(poses[k], rng, 0.004, 0.03)
        out[f"pose_k{k}"] = np.array([pose_k], dtype=_capi.POSE)
        counts = o.associate(pose_k)
        out[f"counts{k}"] = counts
        out[f"matches_planar{k}"], out[f"matches_point{k}"] = o.matches(0), o.matches(1)
        poses[k] = pose_k
        if len(counts):
            pairs = np.zeros(len(counts), dtype=_capi.PAIR)
            pairs["i"], pairs["j"] = counts["i"], k
            allp = scan_poses(list(poses), [poses[s] for s in poses])
            out[f"lin_pairs{k}"], out[f"lin_poses{k}"] = pairs, allp
            out[f"blocks{k}"] = o.linearize(pairs, allp)
            out[f"errors{k}"] = o.error(pairs, allp)
        out[f"added{k}"] = np.array(o.commit_scan())
        out[f"stored_planar{k}"], out[f"stored_point{k}"] = o.keypoints(0, k), o.keypoints(1, k)
    np.savez_compressed(os.path.join(HERE, "hotpath_8x384.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_8x384.npz"), {k: len(v) for k, v in out.items() if k.startswith("planar")})


if __name__ == "__main__":
    main()
