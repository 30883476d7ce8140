"""GPU parity on the BASELINE.json configurations the short tests do not reach: a FULL fixed-lag
window (61 scans, 1830 pairs, slot reuse under steady-state churn), the 200-scan OS1-64 sequence
of configs[0] end to end (ATE difference <= 1 mm), and the 128x2048 stress scans of configs[4]
against a map of about a million occupied voxels.  Every comparison is against the CPU oracle on
the same inputs, through the C-ABI; both stage-3 implementations (pair-moment cache and the
streaming kernels, FORMGPU_STREAM_LINEARIZE=1) are covered."""
import numpy as np
import pytest

import test_gpu_pipeline
import test_gpu_stages
from form_b200 import _capi, synth
from helpers import block_rel_err, perturbed, scan_poses

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ full window, stage level
@pytest.mark.parametrize("stream", [False, True], ids=["moment-cache", "streaming"])
def test_full_window_stages_os1_64(monkeypatch, stream):
    """70 OS1-64 scans, window capped at 61 by removing older scans (slot reuse): neighbour ids,
    dist^2 bits and pair counts bit-exact at every scan; the final relinearisation covers the
    whole window (> 1000 non-empty pairs of the 1830, far beyond the 48 tasks that travel in
    kernel parameters)."""
    if stream:
        monkeypatch.setenv("FORMGPU_STREAM_LINEARIZE", "1")
    worst = test_gpu_stages._run_sequence("os1-64", 70, seq=3, icp_iters=1, max_window=61,
                                          min_final_pairs=1000)
    assert worst < 1e-9


# ------------------------------------------------------------------ full window, end to end
def test_full_window_pipeline_os1_64():
    """form::Estimator with key-scan parameters that keep every scan (ratio 0, no ageing): the
    window fills to 1 + 10 + 50 scans and then churns through the hard cap; full LM over up to
    1830 pairs every scan.  Identical keypoints and control flow, poses within 1 mm."""
    gpu, ref, scans, p, ate = test_gpu_pipeline._run_both(
        "os1-64", 85, seq=1, keyscan_match_ratio=0.0, max_steps_unused_keyscan=100000)
    gs, rs = gpu.stats(), ref.stats()
    assert gs["window_size"] == rs["window_size"] == 60, (gs["window_size"], rs["window_size"])
    assert gs == rs, (gs, rs)
    w, rw = gpu.window(), ref.window()
    assert np.array_equal(w["scan"], rw["scan"])
    assert np.max(np.abs(w["t"] - rw["t"])) < test_gpu_pipeline.ATE_TOL_M


def test_config0_200_scans_os1_64_ate():
    """BASELINE.json configs[0]: the 200-scan OS1-64 sequence, default parameters."""
    gpu, ref, scans, p, ate = test_gpu_pipeline._run_both("os1-64", 200)
    assert gpu.stats() == ref.stats()
    assert ate < test_gpu_pipeline.ATE_TOL_M
    # the estimate follows the synthetic ground truth
    g0, gk = synth.gt_pose(0, 0), synth.gt_pose(0, 199)
    rel = g0["R"].reshape(3, 3).T @ (gk["t"] - g0["t"])
    assert np.linalg.norm(gpu.pose()["t"] - rel) < 0.25


# ------------------------------------------------------------------ configs[4]: 1 M-voxel map
N_TILES = 40  # ~25 k planar + ~5 k point voxels per 128x2048 scan of one hall tile


def test_config4_stress_scans_against_million_voxel_map():
    """128x2048 scans matched against a map seeded with one scan per tile of the tiled hall
    (form/synth.hpp): about a million occupied 0.8 m voxels, 1.5 M map points.  Association
    (ids, dist^2 bits, counts), the blocks of the matched pairs and the commit are compared with
    the oracle for two consecutive query scans in tile 0."""
    import oracle_lib
    from form_b200.context import Context

    rows, cols = synth.shape("stress-128x2048")
    params = _capi.default_params(rows, cols)
    ref = oracle_lib.Oracle(params)
    rng = np.random.default_rng(7)
    poses = {}
    with Context(params) as ctx:
        empty = scan_poses([], [])
        ctx.map_rebuild(empty)
        ref.map_rebuild(empty)
        # seed: one scan per far tile; nothing is near it in the (empty) map, so every keypoint
        # is novel and is committed
        for tile in range(1, N_TILES + 1):
            scan = synth.stress_scan(tile, 0)
            pl, pt = ctx.extract(scan, tile)
            rpl, rpt = ref.extract(scan, tile)
            assert pl.tobytes() == rpl.tobytes() and pt.tobytes() == rpt.tobytes(), f"tile {tile}"
            poses[tile] = synth.stress_pose(tile, 0)
            assert len(ctx.associate(poses[tile])) == 0 and len(ref.associate(poses[tile])) == 0
            assert ctx.commit_scan() == ref.commit_scan() == (len(pl), len(pt))
        worst = 0.0
        for q in range(3):  # tile 0: scan 0 joins the map, scans 1 and 2 are matched against it
            idx = 1000 + q
            scan = synth.stress_scan(0, q)
            pl, pt = ctx.extract(scan, idx)
            rpl, rpt = ref.extract(scan, idx)
            assert pl.tobytes() == rpl.tobytes() and pt.tobytes() == rpt.tobytes()
            window = sorted(poses)
            sp = scan_poses(window, [poses[s] for s in window])
            ctx.map_rebuild(sp)
            ref.map_rebuild(sp)
            if q == 0:
                total = ref.num_voxels(0) + ref.num_voxels(1)
                assert total > 1_000_000, total
            pose_k = perturbed(synth.stress_pose(0, q), rng, 0.001, 0.02)
            counts, rcounts = ctx.associate(pose_k), ref.associate(pose_k)
            assert counts.tobytes() == rcounts.tobytes()
            for t in (0, 1):
                m, rm = ctx.matches(t), ref.matches(t)
                assert np.array_equal(m["found"], rm["found"])
                assert np.array_equal(m["scan"], rm["scan"]) and np.array_equal(m["k"], rm["k"])
                assert np.array_equal(m["dist_sqrd"].view(np.uint64), rm["dist_sqrd"].view(np.uint64))
            poses[idx] = pose_k
            if q > 0:
                assert len(counts) >= 1 and counts["n_planar"].sum() > 10000
                all_poses = scan_poses(sorted(poses), [poses[s] for s in sorted(poses)])
                pairs = np.zeros(len(counts), dtype=_capi.PAIR)
                pairs["i"], pairs["j"] = counts["i"], idx
                for a, b in zip(ctx.linearize(pairs, all_poses), ref.linearize(pairs, all_poses)):
                    worst = max(worst, block_rel_err(a, b))
                assert np.allclose(ctx.error(pairs, all_poses), ref.error(pairs, all_poses), rtol=1e-9)
            assert ctx.commit_scan() == ref.commit_scan()
            for t in (0, 1):
                assert ctx.keypoints(t, idx).tobytes() == ref.keypoints(t, idx).tobytes()
        assert worst < 1e-9, worst
