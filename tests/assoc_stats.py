"""What the association kernel's cost depends on, measured on the CPU oracle (no GPU needed):
bucket sizes of the voxel map, neighbour voxels that survive the face-distance pruning, and the
candidates c-bar a query examines (B_assoc = ... + 32 c-bar K ...).

A synthetic sequence is driven through the oracle's stages with the ground-truth poses and a
sliding window of `--window` scans (the bench's key-scan logic keeps ~13); the statistics are taken
for the last scan's association.  usage: python tests/assoc_stats.py [--sensor os0-128]
[--scans 30] [--window 13]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from form_b200 import _capi, synth  # noqa: E402
import oracle_lib  # noqa: E402
from helpers import scan_poses  # noqa: E402

SHIFTS = np.array([
    (0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
    (1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (1, 0, 1), (1, 0, -1), (-1, 0, 1),
    (-1, 0, -1), (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1), (1, 1, 1), (1, 1, -1),
    (1, -1, 1), (1, -1, -1), (-1, 1, 1), (-1, 1, -1), (-1, -1, 1), (-1, -1, -1)])


def world(kp, pose):
    p = np.stack([kp["x"], kp["y"], kp["z"]], axis=1)
    return p @ pose["R"].reshape(3, 3).T + pose["t"]


def stats_for(name, map_pts, queries, w):
    keys = np.floor(map_pts / w).astype(np.int64)
    uniq, inv, cnt = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    table = {tuple(k): i for i, k in enumerate(uniq)}
    order = np.argsort(inv, kind="stable")
    start = np.concatenate([[0], np.cumsum(cnt)])
    sorted_pts = map_pts[order]
    print(f"[{name}] map points {len(map_pts)}, occupied voxels {len(uniq)}, points/voxel mean {cnt.mean():.1f} "
          f"p50 {np.percentile(cnt, 50):.0f} p90 {np.percentile(cnt, 90):.0f} max {cnt.max()}")
    qk = np.floor(queries / w).astype(np.int64)
    centre, surv, surv_nonempty, cand, cand_all27, empty_centre, best_d = [], [], [], [], [], 0, []
    for q, k in zip(queries, qk):
        i = table.get(tuple(k))
        best = np.inf
        c0 = 0
        if i is not None:
            pts = sorted_pts[start[i]: start[i + 1]]
            c0 = len(pts)
            best = ((pts - q) ** 2).sum(axis=1).min()
        else:
            empty_centre += 1
        lo = k * w
        dm = np.maximum(q - lo, 0.0) ** 2
        dp = np.maximum(lo + w - q, 0.0) ** 2
        lb = np.where(SHIFTS[1:] < 0, dm, np.where(SHIFTS[1:] > 0, dp, 0.0)).sum(axis=1)
        keep = lb <= best
        n_keep = int(keep.sum())
        c_surv, n_ne, c_all = 0, 0, c0
        for s, kp_ in zip(SHIFTS[1:], keep):
            j = table.get(tuple(k + s))
            if j is None:
                continue
            c_all += cnt[j]
            if kp_:
                n_ne += 1
                c_surv += cnt[j]
                pts = sorted_pts[start[j]: start[j + 1]]
                best = min(best, ((pts - q) ** 2).sum(axis=1).min())
        centre.append(c0)
        surv.append(n_keep)
        surv_nonempty.append(n_ne)
        cand.append(c0 + c_surv)
        cand_all27.append(c_all)
        best_d.append(best)
    centre, surv, cand = np.array(centre), np.array(surv), np.array(cand)
    best_d = np.sqrt(np.array(best_d)[np.isfinite(best_d)])
    print(f"[{name}] queries {len(queries)}: centre bucket mean {centre.mean():.1f} p90 {np.percentile(centre, 90):.0f}; "
          f"empty centre {empty_centre / len(queries):.1%}; neighbour voxels surviving the face test mean "
          f"{surv.mean():.2f} (non-empty {np.mean(surv_nonempty):.2f}) of 26")
    print(f"[{name}] candidates per query c-bar: pruned search {cand.mean():.1f} (p90 {np.percentile(cand, 90):.0f}), "
          f"all 27 voxels as the reference scans them {np.mean(cand_all27):.1f}; NN distance median "
          f"{np.median(best_d) * 100:.1f} cm, p90 {np.percentile(best_d, 90) * 100:.1f} cm")


def subcell_model(name, map_pts, queries, w, n_sub=4):
    """What a sub-voxel ordering of the buckets would buy: every voxel's points grouped by an
    n_sub^3 cell code; a query scans its own cell, then every cell (of the 27 voxels) whose box is
    at most as far as the best so far, nearest first.  Reports candidates and cells per query and
    checks that the result is the brute-force nearest neighbour of the 27 voxels."""
    cw = w / n_sub
    vox = np.floor(map_pts / w).astype(np.int64)
    cell = np.clip(np.floor((map_pts - vox * w) / cw).astype(np.int64), 0, n_sub - 1)
    fine = vox * n_sub + cell  # global fine-cell coordinates
    cells = {}
    for i, f in enumerate(map(tuple, fine)):
        cells.setdefault(f, []).append(i)
    cells = {f: map_pts[np.array(ix)] for f, ix in cells.items()}
    rng = np.arange(-n_sub, 2 * n_sub)  # fine cells of the 3x3x3 voxel block, relative to the voxel
    rel = np.stack(np.meshgrid(rng, rng, rng, indexing="ij"), axis=-1).reshape(-1, 3)
    cand, visited, wrong = [], [], 0
    for q in queries:
        k = np.floor(q / w).astype(np.int64)
        base = k * n_sub
        lo = (base + rel) * cw
        d = np.maximum(np.maximum(lo - q, q - (lo + cw)), 0.0)
        lb = (d * d).sum(axis=1)
        best, n_c, n_v = np.inf, 0, 0
        for j in np.argsort(lb, kind="stable"):
            if lb[j] > best:
                break
            pts = cells.get(tuple(base + rel[j]))
            if pts is None:
                continue
            n_v += 1
            n_c += len(pts)
            best = min(best, ((pts - q) ** 2).sum(axis=1).min())
        # brute force over the 27 voxels
        m = np.all(np.abs(vox - k) <= 1, axis=1)
        ref = ((map_pts[m] - q) ** 2).sum(axis=1).min() if m.any() else np.inf
        wrong += int(best != ref)
        cand.append(n_c)
        visited.append(n_v)
    print(f"[{name}] with {n_sub}^3 sub-cells of {cw * 100:.0f} cm: candidates per query {np.mean(cand):.1f} "
          f"(p90 {np.percentile(cand, 90):.0f}), non-empty cells visited {np.mean(visited):.2f} "
          f"(p90 {np.percentile(visited, 90):.0f}); results differing from brute force: {wrong}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sensor", default="os0-128")
    ap.add_argument("--scans", type=int, default=30)
    ap.add_argument("--window", type=int, default=13)
    ap.add_argument("--sample", type=int, default=4000, help="queries sampled per keypoint type")
    args = ap.parse_args()
    rows, cols = synth.shape(args.sensor)
    params = _capi.default_params(rows, cols)
    o = oracle_lib.Oracle(params)
    w = float(params.max_dist_matching)
    live = []
    for k in range(args.scans):
        scan = synth.scan(args.sensor, 0, k)
        planar, point = o.extract(scan, k)
        live.append(k)
        if len(live) > args.window:  # drop the oldest non-first scan, roughly what key-scan logic does
            o.remove_scans([live.pop(0)])
        poses = [synth.gt_pose(0, s) for s in live]
        sp = scan_poses(live, poses)
        o.map_rebuild(sp)
        o.associate(poses[-1])
        if k == args.scans - 1:
            rng = np.random.default_rng(0)
            for t, cur in ((0, planar), (1, point)):
                mp = [world(o.keypoints(t, s), p) for s, p in zip(live[:-1], poses[:-1])]
                mp = np.concatenate([m for m in mp if len(m)])
                q = world(cur, poses[-1])
                if len(q) > args.sample:
                    q = q[rng.choice(len(q), args.sample, replace=False)]
                stats_for(("planar", "point")[t], mp, q, w)
                for n_sub in (2, 4):
                    subcell_model(("planar", "point")[t], mp, q[: args.sample // 4], w, n_sub)
        o.commit_scan()


if __name__ == "__main__":
    main()
