"""Independent numpy / pure-Python restatement of FORM's extraction and voxel-map rules,
used to pin the C++ oracle (SURVEY 8c pins (ii)-(v)).  Written from the reference's
semantics (extraction.tpp / map.tpp line numbers in the comments), not from the oracle."""
from __future__ import annotations

import numpy as np

FLT_MAX = np.finfo(np.float32).max


def sqnorm4_f32(d):
    """Eigen 4-lane float packet reduction: (d0^2 + d2^2) + (d1^2 + d3^2)."""
    d = d.astype(np.float32)
    return (d[..., 0] * d[..., 0] + d[..., 2] * d[..., 2]) + (d[..., 1] * d[..., 1] + d[..., 3] * d[..., 3])


def masks(scan4, rows, cols, np_, min2, max2):
    """valid (dilated) and point_valid masks in closed form (extraction.tpp:136-222)."""
    P = scan4.reshape(rows, cols, 4).astype(np.float32)
    r2 = sqnorm4_f32(P).astype(np.float64)
    range_ok = ~((r2 < min2) | (r2 > max2))
    c = np.arange(cols)
    edge = (c < np_) | (c >= cols - np_)
    bad = (~range_ok) & (~edge)[None, :]          # only non-edge bad points dilate (:159-175)
    dil = bad.copy()
    for k in range(1, np_ + 1):
        dil[:, k:] |= bad[:, :-k]
        dil[:, :-k] |= bad[:, k:]
    valid = (~edge)[None, :] & ~dil
    pvalid = (~edge)[None, :] & range_ok
    return valid.reshape(-1), pvalid.reshape(-1)


def curvature(scan4, rows, cols, np_, valid):
    """extraction.tpp:226-261: f64 accumulation in the reference order, rounded to f32."""
    P = scan4.reshape(rows, cols, 4)[..., :3].astype(np.float64)
    out = np.full(rows * cols, FLT_MAX, dtype=np.float32)
    v = valid.reshape(rows, cols)
    for r in range(rows):
        for c in np.nonzero(v[r])[0]:
            d = -(2.0 * np_) * P[r, c]
            for n in range(1, np_ + 1):
                d = (d + P[r, c - n]) + P[r, c + n]
            out[r * cols + c] = np.float32((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    return out


def sectors(cols, num_sectors):
    pps = cols // num_sectors
    return [(s * pps, cols if s == num_sectors - 1 else (s + 1) * pps) for s in range(num_sectors)]


def planar_picks(curv, valid, rows, cols, np_, num_sectors, thr, per_sector):
    """extraction.tpp:44-68 + :332-358 with rule R1 (stable sort)."""
    used = valid.copy()
    picks = []
    for r in range(rows):
        for a, b in sectors(cols, num_sectors):
            idx = np.arange(r * cols + a, r * cols + b)
            order = idx[np.argsort(curv[idx], kind="stable")]
            count = 0
            for i in order:
                if used[i] and float(curv[i]) < thr:
                    picks.append(int(i))
                    used[i - (np_ - 1): i + np_] = False
                    count += 1
                if count > per_sector:
                    break
    return np.array(picks, dtype=np.uint32), used


def point_picks(used, valid, pvalid, rows, cols, np_, num_sectors, per_sector):
    """extraction.tpp:72-96 + :360-399 (break leaves only the inner loop)."""
    mask = (used == valid) & pvalid
    picks = []
    if per_sector == 0:
        return np.array(picks, dtype=np.uint32)
    for r in range(rows):
        for a, b in sectors(cols, num_sectors):
            unused = [i for i in range(r * cols + a, r * cols + b) if mask[i]]
            factor = 1 + len(unused) // per_sector
            count = 0
            for offset in range(factor):
                for u in range(offset, len(unused), factor):
                    i = unused[u]
                    if mask[i]:
                        picks.append(i)
                        mask[i - (np_ - 1): i + np_] = False
                        count += 1
                    if count > per_sector:
                        break
    return np.array(picks, dtype=np.uint32)


def closest_in_row(scan4, valid, p, start, end):
    """find_closest (extraction.tpp:402-420), rule R2."""
    idx = np.arange(start, end)[valid[start:end]]
    if len(idx) == 0:
        return -1
    d2 = sqnorm4_f32(scan4[idx] - p[None, :])
    return int(idx[np.argmin(d2)])  # argmin returns the first minimum = lowest index


def neighbor_list(scan4, idx, np_, radius2):
    """find_neighbors (extraction.tpp:422-448)."""
    out = []
    for sgn in (+1, -1):
        for i in range(1, np_ + 1):
            q = scan4[idx + sgn * i]
            if float(sqnorm4_f32((q - scan4[idx])[None, :])[0]) < radius2:
                out.append(q)
            else:
                break
    return out


def normal_f64(scan4, valid, idx, rows, cols, np_, radius, min_points):
    """compute_normal (extraction.tpp:263-329) in float64 with numpy's eigh: returns
    (ok, normal, eigenvalues) - a tolerance reference for the oracle's float solver."""
    row = idx // cols
    nbrs = neighbor_list(scan4, idx, np_, radius * radius)
    other = False
    for rr in (row - 1, row + 1):
        if 0 <= rr < rows:
            c = closest_in_row(scan4, valid, scan4[idx], rr * cols, (rr + 1) * cols)
            if c >= 0:
                other = True
                nbrs.append(scan4[c])
                nbrs += neighbor_list(scan4, c, np_, radius * radius)
    if not other or len(nbrs) < min_points:
        return False, None, None
    A = (np.array(nbrs, dtype=np.float64)[:, :3] - scan4[idx][:3].astype(np.float64)) / len(nbrs)
    w, v = np.linalg.eigh(A.T @ A)
    return True, v[:, 0], w


# ---- voxel map ----
SHIFTS = np.array([
    (0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
    (1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (1, 0, 1), (1, 0, -1), (-1, 0, 1),
    (-1, 0, -1), (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1), (1, 1, 1), (1, 1, -1),
    (1, -1, 1), (1, -1, -1), (-1, 1, 1), (-1, 1, -1), (-1, -1, 1), (-1, -1, -1)], dtype=np.int64)


def voxel_key(p, width):
    return np.floor(np.asarray(p, dtype=np.float64) / width).astype(np.int64)


def transform(R, t, p):
    """R p + t with the oracle's dot-product order ((r0 x + r1 y) + r2 z) + t."""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    p = np.asarray(p, dtype=np.float64)
    return ((R[:, 0] * p[..., 0:1] + R[:, 1] * p[..., 1:2]) + R[:, 2] * p[..., 2:3]) + np.asarray(t)


def nn_bruteforce(world_pts, ids, query, width):
    """VoxelMap::find_closest by brute force over the 27-voxel neighbourhood with rule R5:
    arg-min of (dist2, shift rank, scan, k).  `ids` rows are (scan, k)."""
    kq = voxel_key(query, width)
    keys = voxel_key(world_pts, width)
    best = None
    for rank, s in enumerate(SHIFTS):
        sel = np.nonzero(np.all(keys == kq + s, axis=1))[0]
        for j in sel:
            d = world_pts[j] - query
            d2 = (d[0] * d[0] + d[2] * d[2]) + (d[1] * d[1] + 0.0)
            cand = (d2, rank, int(ids[j][0]), int(ids[j][1]))
            if best is None or cand < best:
                best = cand
    return best
