"""CPU models of two transformations the CUDA kernels apply to the reference's algorithm,
checked against the reference's own sequential form on random and adversarial inputs
(the GPU parity tests check the kernels themselves; these pin the REASONING they rest on):

* extract.cu, point-feature walk: the reference visits u = offset + m * factor pass by pass
  (/root/reference/form/feature/extraction.tpp:379-398); the kernel resolves 32 consecutive
  elements of the whole (offset, m) visiting sequence per warp step, with a second resolution
  when the `count > point_feats_per_sector` cut-off falls inside the group.
* map_assoc.cu, voxel_coord: floor(v * fl(1 / w)) with a fall-back to the IEEE division when the
  product is within 8.9e-16 (|q| + 1) of an integer must equal floor(fl(v / w)) always
  (/root/reference/form/mapping/map.tpp:34-38).
"""
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
