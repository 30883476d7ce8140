"""CPU model of the cell-search association kernel's SIMT cost under different query orders
(no GPU needed): `assoc_cells_body` (form_b200/csrc/map_assoc.cu) gives one thread to a query, so a
warp pays, per phase, for its slowest lane.  The model replays the kernel's control flow per query
(own cell or small voxel, the face-selected adjacent units popped in shift-rank order with the
summed bound re-tested, the whole-voxel fallback for queries that cannot prove completeness) on the
oracle's map, and charges a warp  sum over steps of (step overhead + max over lanes of the
candidates scanned in that step).  Orders compared: the keypoint order of the extraction (row,
sector, curvature rank), range-image tiles (rows/8 x cols/32), and a Morton order of the world
cell - the upper bound of what spatial ordering can buy.

usage: python tests/assoc_order_model.py [--sensor os0-128] [--scans 30] [--window 13]"""
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
from assoc_stats import SHIFTS, world  # noqa: E402

N_SUB, CELL_MIN = 4, 16
C_STEP, C_PT, C_P3 = 40.0, 10.0, 150.0  # instructions: per pop / probe, per candidate, per phase-3 query


def per_query_work(map_pts, queries, w):
    """-> list of (n0, [n_k per popped unit], phase3_candidates or 0)."""
    cw = w / N_SUB
    vox = np.floor(map_pts / w).astype(np.int64)
    uniq, inv, cnt = np.unique(vox, axis=0, return_inverse=True, return_counts=True)
    table = {tuple(k): i for i, k in enumerate(uniq)}
    order = np.argsort(inv, kind="stable")
    start = np.concatenate([[0], np.cumsum(cnt)])
    pts_sorted = map_pts[order]
    cell_of = np.clip(np.floor((pts_sorted - np.floor(pts_sorted / w) * w) / cw).astype(np.int64), 0, N_SUB - 1)

    def voxel_pts(key):
        i = table.get(tuple(key))
        if i is None:
            return None, None
        return pts_sorted[start[i]:start[i + 1]], cell_of[start[i]:start[i + 1]]

    out = []
    for q in queries:
        k = np.floor(q / w).astype(np.int64)
        pts, cells = voxel_pts(k)
        by_cell = pts is not None and len(pts) >= CELL_MIN
        best = np.inf
        if by_cell:
            f0 = np.clip(np.floor((q - k * w) / cw).astype(np.int64), 0, N_SUB - 1)
            ulo, uw = k * w + f0 * cw, cw
            m = np.all(cells == f0, axis=1)
            n0 = int(m.sum())
            if n0:
                best = ((pts[m] - q) ** 2).sum(axis=1).min()
        else:
            ulo, uw = k * w, w
            n0 = 0 if pts is None else len(pts)
            if n0:
                best = ((pts - q) ** 2).sum(axis=1).min()
        dm = np.maximum(q - ulo, 0.0) ** 2
        dp = np.maximum(ulo + uw - q, 0.0) ** 2
        steps = []
        for s in SHIFTS[1:]:
            # six face comparisons select the candidate shifts (against the best after phase 1)
            if any((s[a] < 0 and dm[a] > best) or (s[a] > 0 and dp[a] > best) for a in range(3)):
                continue
            steps.append(s)
        best1 = best
        ns = []
        for s in steps:
            lb = sum(dm[a] if s[a] < 0 else dp[a] if s[a] > 0 else 0.0 for a in range(3))
            n = 0
            if lb <= best:
                if by_cell:
                    f = f0 + s
                    carry = np.where(f < 0, -1, np.where(f >= N_SUB, 1, 0))
                    ci = f - N_SUB * carry
                    p2, c2 = voxel_pts(k + carry)
                    if p2 is not None:
                        m = np.all(c2 == ci, axis=1)
                        n = int(m.sum())
                        if n:
                            best = min(best, ((p2[m] - q) ** 2).sum(axis=1).min())
                else:
                    p2, _ = voxel_pts(k + s)
                    if p2 is not None:
                        n = len(p2)
                        best = min(best, ((p2 - q) ** 2).sum(axis=1).min())
            ns.append(n)
        p3 = 0
        if by_cell and not best < cw * cw:
            # whole-voxel search by all 32 lanes: centre bucket + surviving neighbour voxels
            vlo = k * w
            dmv, dpv = np.maximum(q - vlo, 0.0) ** 2, np.maximum(vlo + w - q, 0.0) ** 2
            p3 = len(pts)
            for s in SHIFTS[1:]:
                lb = sum(dmv[a] if s[a] < 0 else dpv[a] if s[a] > 0 else 0.0 for a in range(3))
                if lb <= best:
                    p2, _ = voxel_pts(k + s)
                    if p2 is not None:
                        p3 += len(p2)
            p3 = max(p3, 1)
        del best1
        out.append((n0, ns, p3))
    return out


def warp_cost(work, order):
    total, thread_work = 0.0, 0.0
    for b in range(0, len(order), 32):
        lanes = [work[i] for i in order[b:b + 32]]
        cost = C_STEP + C_PT * max(l[0] for l in lanes)
        iters = max(len(l[1]) for l in lanes)
        for k in range(iters):
            cost += C_STEP + C_PT * max((l[1][k] if k < len(l[1]) else 0) for l in lanes)
        for l in lanes:
            if l[2]:
                cost += C_P3 + C_PT * ((l[2] + 31) // 32)
            thread_work += C_STEP + C_PT * l[0] + sum(C_STEP + C_PT * n for n in l[1])
            if l[2]:
                thread_work += C_P3 + C_PT * l[2] / 32.0
        total += cost
    return total, thread_work / (32.0 * total)


def morton(ix):
    ix = ix - ix.min(axis=0)
    code = np.zeros(len(ix), dtype=np.int64)
    for b in range(12):
        for a in range(3):
            code |= ((ix[:, a] >> b) & 1) << (3 * b + a)
    return code


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sensor", default="os0-128")
    ap.add_argument("--scans", type=int, default=30)
    ap.add_argument("--window", type=int, default=13)
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
        if len(live) > args.window:
            o.remove_scans([live.pop(0)])
        poses = [synth.gt_pose(0, s) for s in live]
        o.map_rebuild(scan_poses(live, poses))
        o.associate(poses[-1])
        if k == args.scans - 1:
            for t, cur in ((0, planar), (1, point)):
                mp = np.concatenate([m for m in (world(o.keypoints(t, s), p) for s, p in zip(live[:-1], poses[:-1])) if len(m)])
                local = np.stack([cur["x"], cur["y"], cur["z"]], axis=1).astype(np.float64)
                q = world(cur, poses[-1])
                work = per_query_work(mp, q, w)
                n = len(q)
                az = np.arctan2(local[:, 1], local[:, 0])
                el = np.arctan2(local[:, 2], np.hypot(local[:, 0], local[:, 1]))
                orders = {"extraction order (row, sector, curvature)": np.arange(n)}
                for (tr, tc) in ((8, 32), (4, 64), (16, 16)):
                    eb = np.clip(((el - el.min()) / (np.ptp(el) + 1e-9) * (rows // tr)).astype(int), 0, rows // tr - 1)
                    ab = np.clip(((az + np.pi) / (2 * np.pi) * (cols // tc)).astype(int), 0, cols // tc - 1)
                    orders[f"range-image tiles {tr} rows x {tc} cols"] = np.lexsort((eb, ab))
                orders["Morton order of the world 20 cm cell (bound)"] = np.argsort(morton(np.floor(q / (w / N_SUB)).astype(np.int64)), kind="stable")
                cand = np.array([l[0] + sum(l[1]) for l in work])
                print(f"[{('planar', 'point')[t]}] {n} queries, candidates/query {cand.mean():.1f}, units popped/query "
                      f"{np.mean([len(l[1]) for l in work]):.2f}, phase-3 queries {np.mean([l[2] > 0 for l in work]):.1%}")
                base = None
                for name, od in orders.items():
                    c, eff = warp_cost(work, od)
                    base = base or c
                    print(f"    {name:55s} warp cost {c / n:7.1f} instr/query  ({base / c:4.2f}x)  lanes active {32 * eff:4.1f}/32")
        o.commit_scan()


if __name__ == "__main__":
    main()
