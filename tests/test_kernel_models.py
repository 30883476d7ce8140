"""CPU models of transformations the CUDA kernels apply to the reference's algorithm,
checked against the reference's own sequential form on random and adversarial inputs
(the GPU parity tests check the kernels themselves; these pin the REASONING they rest on):

* extract.cu, point-feature walk: the reference visits u = offset + m * factor pass by pass
  (/root/reference/form/feature/extraction.tpp:379-398); the kernel resolves 32 consecutive
  elements of the whole (offset, m) visiting sequence per warp step, with a second resolution
  when the `count > point_feats_per_sector` cut-off falls inside the group.
* map_assoc.cu, voxel_coord: floor(v * fl(1 / w)) with a fall-back to the IEEE division when the
  product is within 8.9e-16 (|q| + 1) of an integer must equal floor(fl(v / w)) always
  (/root/reference/form/mapping/map.tpp:34-38).
* blueprint of the NEXT association kernel (not built yet): buckets ordered by a 4x4x4 cell
  code, own cell + adjacent cells by box bound, whole-voxel fall-back - must give the rule-R5
  arg-min of the 27-voxel search (/root/reference/form/mapping/map.tpp:54-91) exactly.
"""
import zlib

import numpy as np
import pytest


# --------------------------------------------------------------------------- point walk
def reference_point_walk(ulist, mask, np_, pfps):
    """extraction.tpp:379-398 verbatim (mask = the row's valid_mask, mutated)."""
    out = []
    if pfps == 0:
        return out
    U = len(ulist)
    factor = 1 + U // pfps
    count = 0
    for offset in range(factor):
        u = offset
        while u < U:
            idx = ulist[u]
            if mask[idx]:
                out.append(idx)
                for n in range(np_):
                    mask[idx + n] = False
                    mask[idx - n] = False
                count += 1
            if count > pfps:
                break
            u += factor
    return out


def resolve_group(cand, col, np_):
    """extract.cu resolve_group: lane l survives iff it is a candidate and no SURVIVING earlier
    lane lies within np-1 columns - computed, like the kernel, by a ballot fixed point."""
    n = len(cand)
    conf = [[k for k in range(l) if cand[k] and abs(col[l] - col[k]) < np_] for l in range(n)]
    undecided = {l for l in range(n) if cand[l]}
    alive = set()
    while undecided:
        kill = {l for l in undecided if any(k in alive for k in conf[l])}
        ok = {l for l in undecided if l not in kill and not any(k in undecided for k in conf[l])}
        assert kill or ok, "the lowest undecided lane is always decidable"
        alive |= ok
        undecided -= kill | ok
    return [l in alive for l in range(n)]


def kernel_point_walk(ulist, mask, np_, pfps, lanes=32):
    """The grouped walk of extract_select_body (extract.cu), lane by lane."""
    out = []
    if pfps == 0:
        return out
    U = len(ulist)
    factor = 1 + U // pfps
    n_full, rem = U // factor, U % factor
    split = rem * (n_full + 1)

    def pass_of(t):
        if t < split:
            return t // (n_full + 1), t % (n_full + 1)
        t2 = t - split
        return rem + t2 // n_full, t2 % n_full

    def clear(c):
        for n in range(np_):
            mask[c + n] = False
            mask[c - n] = False

    count, offset, base = 0, factor, 0
    while base < U and count <= pfps:
        ts = [base + l for l in range(lanes)]
        inr = [t < U for t in ts]
        om = [pass_of(t) if i else (0, 0) for t, i in zip(ts, inr)]
        col = [ulist[o + m * factor] if i else 0 for (o, m), i in zip(om, inr)]
        cand = [i and bool(mask[c]) for i, c in zip(inr, col)]
        alive = resolve_group(cand, col, np_)
        before = [count + sum(alive[:l]) for l in range(lanes)]
        over = [b > pfps for b in before]
        if any(over):
            first_over = over.index(True)
            dropped = [cand[l] and om[l][1] > 0 and l >= first_over for l in range(lanes)]
            if any(dropped):
                alive = resolve_group([c and not d for c, d in zip(cand, dropped)], col, np_)
        for l in range(lanes):
            if alive[l]:
                out.append(col[l])
        for l in range(lanes):  # the kernel clears after the group's decisions are taken
            if alive[l]:
                clear(col[l])
        count += sum(alive)
        if count > pfps:
            offset = pass_of(min(base + lanes - 1, U - 1))[0] + 1
        base += lanes
    base = offset
    while base < factor:  # phase B: only the first element of every remaining pass
        os_ = [base + l for l in range(lanes)]
        inr = [o < factor and o < U for o in os_]
        if not any(inr):
            break
        col = [ulist[o] if i else 0 for o, i in zip(os_, inr)]
        cand = [i and bool(mask[c]) for i, c in zip(inr, col)]
        alive = resolve_group(cand, col, np_)
        for l in range(lanes):
            if alive[l]:
                out.append(col[l])
        for l in range(lanes):
            if alive[l]:
                clear(col[l])
        base += lanes
    return out


def _sector_case(rng, sector_len, np_, density):
    pad = 2 * np_
    mask = np.zeros(sector_len + 2 * pad, dtype=bool)
    mask[pad: pad + sector_len] = rng.uniform(size=sector_len) < density
    ulist = [int(i) for i in np.nonzero(mask)[0]]
    return ulist, mask


@pytest.mark.parametrize("pfps", [1, 2, 3, 5, 10, 40])
@pytest.mark.parametrize("np_", [1, 3, 5, 8])
def test_grouped_point_walk_equals_the_reference_loop(pfps, np_):
    rng = np.random.default_rng(1000 * pfps + np_)
    for trial in range(60):
        sector_len = int(rng.integers(1, 360))
        density = float(rng.choice([0.02, 0.1, 0.3, 0.6, 0.9, 1.0]))
        ulist, mask = _sector_case(rng, sector_len, np_, density)
        m_ref, m_ker = mask.copy(), mask.copy()
        ref = reference_point_walk(ulist, m_ref, np_, pfps)
        ker = kernel_point_walk(ulist, m_ker, np_, pfps)
        assert ker == ref, (pfps, np_, trial, sector_len, density)
        assert np.array_equal(m_ref, m_ker)  # the mask feeds the next sector


def test_grouped_point_walk_small_groups_and_cutoff_inside_a_group():
    """Narrow 'warps' put many group boundaries and cut-offs inside groups."""
    rng = np.random.default_rng(7)
    for lanes in (2, 3, 4, 7):
        for trial in range(150):
            pfps, np_ = int(rng.integers(1, 6)), int(rng.integers(1, 4))
            ulist, mask = _sector_case(rng, int(rng.integers(1, 60)), np_, float(rng.uniform(0.2, 1.0)))
            m_ref, m_ker = mask.copy(), mask.copy()
            assert kernel_point_walk(ulist, m_ker, np_, pfps, lanes) == reference_point_walk(ulist, m_ref, np_, pfps)
            assert np.array_equal(m_ref, m_ker)


# --------------------------------------------------------------------------- voxel_coord
def voxel_coord_fast(v, w):
    inv = np.float64(1.0) / np.float64(w)
    q = v * inv
    f = np.floor(q)
    fr = q - f
    tol = 8.9e-16 * (np.abs(q) + 1.0)
    slow = (fr < tol) | (fr > 1.0 - tol)
    out = f.copy()
    out[slow] = np.floor(v[slow] / np.float64(w))
    return out, slow


@pytest.mark.parametrize("w", [0.8, 0.1, 0.3, 1.0, 0.7, 2.5, 1e-3])
def test_voxel_coord_fast_path_is_exact(w):
    rng = np.random.default_rng(int(w * 1e6))
    k = rng.integers(-(2 ** 20), 2 ** 20, size=400_000).astype(np.float64)
    faces = k * np.float64(w)
    near = [faces]
    for _ in range(6):  # a few ulps either side of every face, where floor() flips
        near.append(np.nextafter(near[-1], np.inf))
    lo = faces
    for _ in range(6):
        lo = np.nextafter(lo, -np.inf)
        near.append(lo)
    v = np.concatenate(near + [rng.uniform(-1e5, 1e5, 2_000_000), rng.uniform(-3, 3, 500_000),
                               np.array([0.0, -0.0, 5e-324, -5e-324, 1e-300, -1e-300])])
    got, slow = voxel_coord_fast(v, w)
    assert np.array_equal(got, np.floor(v / np.float64(w)))
    # the fall-back really is rare away from the faces
    rnd = rng.uniform(-1e5, 1e5, 1_000_000)
    assert voxel_coord_fast(rnd, w)[1].mean() < 1e-6


# --------------------------------------------------------------------------- cell-ordered buckets
# Blueprint of the next association kernel (tests/assoc_stats.py): every voxel's
# bucket is ordered by a 4x4x4 cell code; a query scans its own cell, then the adjacent fine cells
# whose box bound does not exceed the best so far, and falls back to the whole-voxel search (the
# current kernel) only when the best is not inside the adjacency radius.  The model carries the
# full rule-R5 key (dist^2, shift rank, tie) and must reproduce the brute-force search of the 27
# voxels exactly - including ties, points on voxel / cell faces and voxels too small for a table.
SHIFTS27 = [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
            (1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (1, 0, 1), (1, 0, -1), (-1, 0, 1),
            (-1, 0, -1), (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1), (1, 1, 1), (1, 1, -1),
            (1, -1, 1), (1, -1, -1), (-1, 1, 1), (-1, 1, -1), (-1, -1, 1), (-1, -1, -1)]
RANK = {s: r for r, s in enumerate(SHIFTS27)}
N_SUB, TABLE_MIN = 4, 16


def _dist2(p, q):  # (d0^2 + d2^2) + (d1^2 + 0), the reference's 4-lane order
    d = p - q
    return (d[0] * d[0] + d[2] * d[2]) + (d[1] * d[1] + 0.0)


def build_cell_map(points, w):
    """voxel -> {'pts': [(xyz, tie)] ordered by cell code, 'off': 65 offsets or None}."""
    cw = w / N_SUB
    vox = {}
    for tie, p in enumerate(points):
        v = tuple(int(np.floor(c / w)) for c in p)
        vox.setdefault(v, []).append((p, tie))
    out = {}
    for v, pts in vox.items():
        if len(pts) < TABLE_MIN:
            out[v] = dict(pts=pts, off=None)
            continue
        code = []
        for p, _ in pts:
            rel = p - np.array(v, dtype=np.float64) * w
            c = [min(max(int(np.floor(r / cw)), 0), N_SUB - 1) for r in rel]
            code.append((c[0] * N_SUB + c[1]) * N_SUB + c[2])
        order = sorted(range(len(pts)), key=lambda i: code[i])  # stable: any order inside a cell is fine
        cnt = np.bincount([code[i] for i in order], minlength=N_SUB ** 3)
        out[v] = dict(pts=[pts[i] for i in order], off=np.concatenate([[0], np.cumsum(cnt)]))
    return out


def brute_force_r5(cmap, q, w):
    c = tuple(int(np.floor(x / w)) for x in q)
    best = (np.inf, 99, 1 << 62)
    for s in SHIFTS27:
        b = cmap.get(tuple(c[a] + s[a] for a in range(3)))
        if b:
            for p, tie in b["pts"]:
                best = min(best, (_dist2(p, q), RANK[s], tie))
    return best


def cell_search_r5(cmap, q, w, stats=None):
    cw = w / N_SUB
    c = tuple(int(np.floor(x / w)) for x in q)
    margin = [1e-9 * (1.0 + abs(x)) for x in q]
    best = (np.inf, 99, 1 << 62)
    seen_whole = set()  # voxels without a table that were scanned entirely
    n_cand = 0

    def scan(v, lo, hi):
        nonlocal best, n_cand
        s = tuple(v[a] - c[a] for a in range(3))
        for p, tie in cmap[v]["pts"][lo:hi]:
            n_cand += 1
            best = min(best, (_dist2(p, q), RANK[s], tie))

    def visit_cell(f):  # f = global fine-cell coordinates
        v = tuple(x // N_SUB for x in f)  # floor division: the voxel that holds the fine cell
        b = cmap.get(v)
        if b is None:
            return
        if b["off"] is None:
            if v not in seen_whole:
                seen_whole.add(v)
                scan(v, 0, len(b["pts"]))
            return
        i = [f[a] - v[a] * N_SUB for a in range(3)]
        code = (i[0] * N_SUB + i[1]) * N_SUB + i[2]
        scan(v, b["off"][code], b["off"][code + 1])

    def cell_bound(f):
        lb = 0.0
        for a in range(3):
            lo = f[a] * cw
            gap = max(lo - q[a] - margin[a], q[a] - (lo + cw) - margin[a], 0.0)
            lb += gap * gap
        return lb

    rel = [q[a] - c[a] * w for a in range(3)]
    f0 = tuple(c[a] * N_SUB + min(max(int(np.floor(rel[a] / cw)), 0), N_SUB - 1) for a in range(3))
    visit_cell(f0)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if (dx, dy, dz) == (0, 0, 0):
                    continue
                f = (f0[0] + dx, f0[1] + dy, f0[2] + dz)
                if cell_bound(f) <= best[0]:
                    visit_cell(f)
    # every fine cell that is not adjacent to f0 is at least one cell width away along some axis
    reach = cw - max(margin)
    if not best[0] < reach * reach:  # fall back to the whole-voxel search of today's kernel
        if stats is not None:
            stats["fallback"] = stats.get("fallback", 0) + 1
        for s in SHIFTS27:
            v = tuple(c[a] + s[a] for a in range(3))
            if v in cmap:
                scan(v, 0, len(cmap[v]["pts"]))
    if stats is not None:
        stats["cand"] = stats.get("cand", 0) + n_cand
        stats["n"] = stats.get("n", 0) + 1
    return best


def _cloud(rng, kind, w):
    if kind == "surface":  # a wall and a floor through a block of voxels, 3-8 cm spacing
        n = 3000
        a = rng.uniform(-1.5 * w, 2.5 * w, (n, 2))
        wall = np.stack([a[:, 0], np.full(n, 0.31) + 0.01 * rng.standard_normal(n), a[:, 1]], axis=1)
        floor = np.stack([a[:, 0], a[:, 1], np.full(n, -0.52) + 0.01 * rng.standard_normal(n)], axis=1)
        return np.concatenate([wall, floor])
    if kind == "sparse":
        return rng.uniform(-2 * w, 3 * w, (150, 3))
    if kind == "lattice":  # points ON voxel and cell faces, many exact distance ties
        g = np.arange(-8, 13) * (w / 4)
        x, y, z = np.meshgrid(g, g, g, indexing="ij")
        pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
        return np.concatenate([pts, pts[::7]])  # duplicates: the tie id decides
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["surface", "sparse", "lattice"])
@pytest.mark.parametrize("w", [0.8, 0.3])
def test_cell_ordered_search_equals_brute_force_r5(kind, w):
    rng = np.random.default_rng(zlib.crc32(f"{kind}-{w}".encode()))
    pts = _cloud(rng, kind, w)
    cmap = build_cell_map(pts, w)
    stats = {}
    queries = list(rng.uniform(-1.2 * w, 2.2 * w, (250, 3)))
    queries += [p + rng.normal(scale=0.03, size=3) for p in pts[rng.choice(len(pts), 250)]]
    queries += [pts[i].copy() for i in rng.choice(len(pts), 60)]  # exactly on map points / faces
    for q in queries:
        assert cell_search_r5(cmap, q, w, stats) == brute_force_r5(cmap, q, w), (kind, w, q)
    if kind == "surface":  # the point of the exercise: far fewer candidates, rare fall-backs
        assert stats["fallback"] < 0.45 * stats["n"]


# --------------------------------------------------------------------------- face pruning (today's kernel)
def face_pruned_search_r5(cmap, q, w):
    """map_assoc.cu assoc_nn_body as built today: the centre voxel's bucket first, then only the
    neighbour voxels whose box (bounded from the six face distances of the centre voxel, shrunk by
    the rounding margin) can still hold a point at least as close."""
    c = tuple(int(np.floor(x / w)) for x in q)
    best = (np.inf, 99, 1 << 62)

    def scan(s):
        nonlocal best
        b = cmap.get(tuple(c[a] + s[a] for a in range(3)))
        for p, tie in (b["pts"] if b else []):
            best = min(best, (_dist2(p, q), RANK[s], tie))

    scan((0, 0, 0))
    bound = best[0]
    face2 = []
    for a in range(3):
        lo = c[a] * w
        margin = 1e-9 * (1.0 + abs(q[a])) + 4e-16 * abs(lo)
        dm, dp = max(q[a] - lo - margin, 0.0), max(lo + w - q[a] - margin, 0.0)
        face2.append((dm * dm, dp * dp))
    n_scanned = 0
    for s in SHIFTS27[1:]:
        lb = sum(0.0 if s[a] == 0 else face2[a][0] if s[a] < 0 else face2[a][1] for a in range(3))
        if lb <= bound:  # the bound is the centre's best: it is not tightened between voxels
            n_scanned += 1
            scan(s)
    return best, n_scanned


@pytest.mark.parametrize("kind", ["surface", "sparse", "lattice"])
def test_face_pruned_search_equals_brute_force_r5(kind):
    w = 0.8
    rng = np.random.default_rng(zlib.crc32(f"face-{kind}".encode()))
    pts = _cloud(rng, kind, w)
    cmap = build_cell_map(pts, w)
    queries = list(rng.uniform(-1.2 * w, 2.2 * w, (300, 3)))
    queries += [p + rng.normal(scale=0.03, size=3) for p in pts[rng.choice(len(pts), 300)]]
    queries += [pts[i].copy() for i in rng.choice(len(pts), 80)]
    scanned = []
    for q in queries:
        got, n = face_pruned_search_r5(cmap, q, w)
        assert got == brute_force_r5(cmap, q, w), (kind, q)
        scanned.append(n)
    if kind == "surface":  # queries near the surfaces (the second group) touch few of the 26 neighbours
        assert np.mean(scanned[300:600]) < 4


# --------------------------------------------------------------------------- sector-parallel walks
# extract.cu runs the greedy walks of a row's sectors in parallel on private copies of the mask
# and validates them in sector order (the reference walks the sectors one after the other on ONE
# mask, extraction.tpp:44-68 / :332-399, so a pick near the end of sector s suppresses the first
# np-1 columns of sector s+1).  These models pin the two validation rules the kernel rests on.
def _planar_sector(order, curv, thr, mask, np_, cap):
    """extract_planar (extraction.tpp:332-358) on one sector: `order` = its columns sorted by
    (curvature, column); returns the picks, mutates mask."""
    out = []
    for c in order:
        if mask[c] and curv[c] < thr:
            out.append(c)
            for n in range(np_):
                mask[c + n] = False
                mask[c - n] = False
        if len(out) > cap:
            break
    return out


def _point_sector(start, end, mask, np_, pfps):
    ulist = [c for c in range(start, end) if mask[c]]
    return reference_point_walk(ulist, mask, np_, pfps)


def _row_case(rng, cols, S, np_, density):
    valid = np.zeros(cols + 2 * np_, dtype=bool)  # columns 0..cols-1 live at [np_, np_ + cols)
    valid[np_: np_ + cols] = rng.random(cols) < density
    valid[np_: 2 * np_] = False
    valid[cols: np_ + cols] = False
    pps = cols // S
    bounds = [(np_ + s * pps, np_ + (cols if s == S - 1 else (s + 1) * pps)) for s in range(S)]
    curv = rng.random(cols + 2 * np_)
    return valid, bounds, curv


@pytest.mark.parametrize("np_,cap,density", [(5, 50, 0.9), (5, 3, 0.9), (3, 50, 0.5), (2, 8, 1.0), (5, 50, 0.2)])
def test_speculative_sector_parallel_planar_walk(np_, cap, density):
    rng = np.random.default_rng(zlib.crc32(f"pl{np_}{cap}{density}".encode()))
    redone = 0
    for _ in range(300):
        cols, S = int(rng.integers(40, 200)), int(rng.integers(1, 7))
        valid, bounds, curv = _row_case(rng, cols, S, np_, density)
        thr = float(rng.choice([0.3, 0.8, 2.0]))
        orders = [sorted(range(a, b), key=lambda c: (curv[c], c)) for a, b in bounds]
        # reference: one mask, sectors in order
        m_ref = valid.copy()
        ref = [_planar_sector(orders[s], curv, thr, m_ref, np_, cap) for s in range(S)]
        # kernel: every sector on a private copy of the INITIAL mask ...
        spec = [_planar_sector(orders[s], curv, thr, valid.copy(), np_, cap) for s in range(S)]
        # ... then, in sector order on the shared mask: a speculative result stands iff none of its
        # picks has been suppressed by the sectors before it; otherwise the sector is walked again
        m = valid.copy()
        got = []
        for s in range(S):
            if all(m[c] for c in spec[s]):
                got.append(spec[s])
                for c in spec[s]:
                    for n in range(np_):
                        m[c + n] = False
                        m[c - n] = False
            else:
                redone += 1
                got.append(_planar_sector(orders[s], curv, thr, m, np_, cap))
        assert got == ref
        assert np.array_equal(m, m_ref)
    assert redone > 0 or density < 0.5  # the redo path is exercised


@pytest.mark.parametrize("np_,pfps,density", [(5, 10, 0.9), (5, 3, 0.6), (3, 20, 0.5), (2, 4, 1.0), (5, 1, 0.3), (4, 0, 0.9)])
def test_speculative_sector_parallel_point_walk(np_, pfps, density):
    rng = np.random.default_rng(zlib.crc32(f"pt{np_}{pfps}{density}".encode()))
    redone = 0
    for _ in range(300):
        cols, S = int(rng.integers(40, 200)), int(rng.integers(1, 7))
        valid, bounds, _ = _row_case(rng, cols, S, np_, density)
        m_ref = valid.copy()
        ref = [_point_sector(a, b, m_ref, np_, pfps) for a, b in bounds]
        spec = [_point_sector(a, b, valid.copy(), np_, pfps) for a, b in bounds]
        # a speculative result stands iff the shared mask still equals the initial one on the
        # sector's own columns (the list of unused points, and with it the pass structure, is
        # built from exactly those bits)
        m = valid.copy()
        got = []
        for s, (a, b) in enumerate(bounds):
            if np.array_equal(m[a:b], valid[a:b]):
                got.append(spec[s])
                for c in spec[s]:
                    for n in range(np_):
                        m[c + n] = False
                        m[c - n] = False
            else:
                redone += 1
                got.append(_point_sector(a, b, m, np_, pfps))
        assert got == ref
        assert np.array_equal(m, m_ref)
    assert redone > 0 or pfps == 0
