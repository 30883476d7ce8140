"""Stage 2 + 3 parity through the C-ABI: map rebuild, association (bit-exact ids and
squared distances), commit (bit-exact stored keypoints), linearisation and error
(relative tolerance), driven with perturbed ground-truth poses instead of a smoother."""
import numpy as np
import pytest

from form_b200 import _capi, synth
from helpers import block_rel_err, gt, perturbed, scan_poses

pytestmark = pytest.mark.gpu

H_TOL = 1e-5   # north-star tolerance on H/b
H_TIGHT = 1e-9 # what we actually expect (f64 accumulation both sides)


def _run_sequence(sensor, n_scans, seq=0, remove_at=None, overrides=None, icp_iters=2, max_window=None,
                  min_final_pairs=1):
    """max_window: once the window holds more scans than this, one of the older ones (never one
    of the newest ten) is removed after every scan - the steady-state churn of a full fixed-lag
    window (slot reuse, pairs erased on both sides)."""
    import oracle_lib
    from form_b200.context import Context

    rows, cols = synth.shape(sensor)
    params = _capi.default_params(rows, cols, **(overrides or {}))
    ref = oracle_lib.Oracle(params)
    rng = np.random.default_rng(42)
    window = []
    est = {}
    worst = 0.0
    with Context(params) as ctx:
        for k in range(n_scans):
            scan = synth.scan(sensor, seq, k)
            pl, pt = ctx.extract(scan, k)
            rpl, rpt = ref.extract(scan, k)
            assert pl.tobytes() == rpl.tobytes() and pt.tobytes() == rpt.tobytes()
            est[k] = perturbed(gt(seq, k), rng, 0.0005, 0.005)
            poses = scan_poses(window + [k], [est[s] for s in window + [k]])
            ctx.map_rebuild(poses)
            ref.map_rebuild(poses)
            for it in range(icp_iters):
                pose_k = perturbed(gt(seq, k), rng, 0.002 / (it + 1), 0.03 / (it + 1))
                counts = ctx.associate(pose_k)
                rcounts = ref.associate(pose_k)
                for t in (0, 1):
                    m, rm = ctx.matches(t), ref.matches(t)
                    assert len(m) == len(rm)
                    assert np.array_equal(m["found"], rm["found"]), f"scan {k} type {t} found flags"
                    assert np.array_equal(m["scan"], rm["scan"]), f"scan {k} type {t} neighbour scan"
                    assert np.array_equal(m["k"], rm["k"]), f"scan {k} type {t} neighbour index"
                    assert np.array_equal(m["dist_sqrd"].view(np.uint64), rm["dist_sqrd"].view(np.uint64)), \
                        f"scan {k} type {t} dist^2 bits"
                assert counts.tobytes() == rcounts.tobytes(), f"scan {k} pair counts"
                est[k] = pose_k
                all_poses = scan_poses(window + [k], [est[s] for s in window + [k]])
                if len(counts):
                    pairs = np.zeros(len(counts), dtype=_capi.PAIR)
                    pairs["i"] = counts["i"]
                    pairs["j"] = k
                    H = ctx.linearize(pairs, all_poses)
                    Hr = ref.linearize(pairs, all_poses)
                    for a, b in zip(H, Hr):
                        worst = max(worst, block_rel_err(a, b))
                    e = ctx.error(pairs, all_poses)
                    er = ref.error(pairs, all_poses)
                    assert np.allclose(e, er, rtol=1e-9, atol=1e-12)
                    assert np.allclose(e, 0.5 * H[:, 90], rtol=1e-9, atol=1e-12)  # f = b^T b
            # the fused entry point (association + linearisation of the current scan's pairs in
            # one device round trip) against the two separate oracle calls
            pose_k = perturbed(gt(seq, k), rng, 0.001, 0.01)
            est[k] = pose_k
            all_poses = scan_poses(window + [k], [est[s] for s in window + [k]])
            fcounts, fH = ctx.associate_linearize(all_poses)
            rcounts = ref.associate(pose_k)
            assert fcounts.tobytes() == rcounts.tobytes(), f"scan {k} fused pair counts"
            for t in (0, 1):
                assert ctx.matches(t).tobytes() == ref.matches(t).tobytes(), f"scan {k} fused matches"
            if len(rcounts):
                pairs = np.zeros(len(rcounts), dtype=_capi.PAIR)
                pairs["i"] = rcounts["i"]
                pairs["j"] = k
                Hr = ref.linearize(pairs, all_poses)
                for a, b in zip(fH, Hr):
                    worst = max(worst, block_rel_err(a, b))
                assert np.array_equal(fH, ctx.linearize(pairs, all_poses))  # same kernel, same order
            added = ctx.commit_scan()
            radded = ref.commit_scan()
            assert added == radded, f"scan {k} novel keypoints"
            for t in (0, 1):
                assert ctx.keypoints(t, k).tobytes() == ref.keypoints(t, k).tobytes()
            window.append(k)
            drop = list(remove_at[k]) if remove_at and k in remove_at else []
            if max_window and len(window) - len(drop) > max_window:
                older = [s for s in window[:-10] if s not in drop]
                drop.append(older[(7 * k) % len(older)])
            if drop:
                ctx.remove_scans(drop)
                ref.remove_scans(drop)
                window = [s for s in window if s not in drop]
        # full-window relinearisation at freshly perturbed poses (HOT LOOP C)
        for s in window:
            est[s] = perturbed(est[s], rng, 0.001, 0.01)
        all_poses = scan_poses(window, [est[s] for s in window])
        pairs = np.array([(i, j) for j in window for i in window if i < j], dtype=_capi.PAIR)
        H = ctx.linearize(pairs, all_poses)
        Hr = ref.linearize(pairs, all_poses)
        nonzero = 0
        for a, b in zip(H, Hr):
            worst = max(worst, block_rel_err(a, b))
            nonzero += bool(np.any(b))
        assert nonzero >= min_final_pairs, nonzero
        e = ctx.error(pairs, all_poses)
        assert np.allclose(e, ref.error(pairs, all_poses), rtol=1e-9, atol=1e-12)
        wpl, wpt = ctx.world_keypoints(all_poses)
        n_pl = sum(len(ref.keypoints(0, s)) for s in window)
        n_pt = sum(len(ref.keypoints(1, s)) for s in window)
        assert len(wpl) == n_pl and len(wpt) == n_pt
    assert worst < H_TOL, worst
    assert worst < H_TIGHT, worst
    return worst


def test_stages_os1_64_sequence():
    _run_sequence("os1-64", 8, remove_at={4: [1], 6: [0, 3]})


def test_stages_vlp16_sparse():
    _run_sequence("vlp-16", 10, seq=2, remove_at={5: [2]})


def test_stages_os0_128():
    _run_sequence("os0-128", 4, seq=1)


def test_stages_no_point_features():
    _run_sequence("os1-64", 4, overrides=dict(point_feats_per_sector=0))


def test_stages_small_matching_distance():
    _run_sequence("vlp-16", 5, overrides=dict(max_dist_matching=0.3, min_dist_map=0.05))
