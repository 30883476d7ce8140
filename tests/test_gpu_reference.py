"""CUDA path vs the REFERENCE'S OWN CODE (oracle/_ref/libformref.so: FORM's hot-path
sources compiled unmodified against API stand-ins, see tests/test_reference_pins.py),
through the C-ABI: stages 1-2 bit for bit, stage-3 blocks to 1e-9 (tolerance 1e-5).  The library is prebuilt where /root/reference exists and
travels to the GPU box with the snapshot."""
import ctypes as C

import numpy as np
import pytest

from form_b200 import _capi, synth
from form_b200.context import Context
from helpers import block_rel_err, perturbed, scan_poses
import test_reference_pins as refpins

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sensor,k", [("vlp-16", 2), ("os1-64", 5), ("os0-128", 1)])
def test_cuda_extraction_is_bit_identical_to_reference_code(sensor, k):
    rows, cols = synth.shape(sensor)
    params = _capi.default_params(rows, cols)
    scan = synth.scan(sensor, 1, k)
    rc, rpl, rpt = refpins.ref_extract(params, scan, k)
    assert rc == 0
    with Context(params) as ctx:
        pl, pt = ctx.extract(scan, k)
    assert len(pl) == len(rpl) and len(pt) == len(rpt)
    assert pl.tobytes() == rpl.tobytes(), "planar keypoints / normals differ from FORM's extract()"
    assert pt.tobytes() == rpt.tobytes(), "point keypoints differ from FORM's extract()"


def test_cuda_extraction_matches_reference_on_random_scans_and_variants():
    rng = np.random.default_rng(23)
    rows, cols = 8, 384
    for overrides in ({}, {"point_feats_per_sector": 0}, {"neighbor_points": 3, "num_sectors": 5},
                      {"radius": 0.3, "min_norm_squared": 0.04}, {"min_points": 20}):
        params = _capi.default_params(rows, cols, **overrides)
        with Context(params) as ctx:
            for trial in range(3):
                scan = refpins.random_scan(rng, rows, cols)
                rc, rpl, rpt = refpins.ref_extract(params, scan, trial)
                pl, pt = ctx.extract(scan, trial)
                assert pl.tobytes() == rpl.tobytes() and pt.tobytes() == rpt.tobytes(), (overrides, trial)


def test_cuda_association_and_commit_match_reference_code():
    """Matcher::match + insert_matches of the reference vs formgpu_associate / _commit_scan:
    matched scan, dist^2 bit pattern, per-pair counts, stored novel keypoints; and the
    reference's DenseFactor::linearize over ITS correspondences vs formgpu_linearize."""
    worst, checked = 0.0, 0
    rng = np.random.default_rng(29)
    sensor, n_scans = "vlp-16", 5
    rows, cols = synth.shape(sensor)
    params = _capi.default_params(rows, cols)
    world = refpins.ref().formref_world_create(C.byref(params))
    try:
        with Context(params) as ctx:
            poses = {}
            for k in range(n_scans):
                scan = synth.scan(sensor, 2, k)
                pl, pt = ctx.extract(scan, k)
                gt = synth.gt_pose(2, k)
                poses[k] = gt if k == 0 else perturbed(gt, rng, 0.002, 0.02)
                sp = scan_poses(list(poses), [poses[s] for s in poses])
                ctx.map_rebuild(sp)
                counts = ctx.associate(poses[k])
                r = refpins._ref_associate(world, sp, k, pl, pt)
                for t, (rq, rp, rd, rf) in ((0, r[0:4]), (1, r[4:8])):
                    m = ctx.matches(t)
                    assert np.array_equal(rf, m["found"].astype(np.uint8))
                    assert rd.tobytes() == m["dist_sqrd"].tobytes()
                    found = m["found"] == 1
                    assert np.array_equal(rp["scan"][found], m["scan"][found])
                by_scan = {int(c["i"]): (int(c["n_planar"]), int(c["n_point"])) for c in counts}
                for s in range(k):
                    a, b = C.c_size_t(), C.c_size_t()
                    refpins.ref().formref_world_constraint_counts(world, s, C.byref(a), C.byref(b))
                    assert (a.value, b.value) == by_scan.get(s, (0, 0)), (k, s)
                if len(counts):
                    pairs = np.zeros(len(counts), dtype=_capi.PAIR)
                    pairs["i"] = counts["i"]
                    pairs["j"] = k
                    H = ctx.linearize(pairs, sp)
                    for c, h in zip(counts, H):
                        got = np.zeros(91)
                        Ti = np.array([poses[int(c["i"])]], dtype=_capi.POSE)
                        Tj = np.array([poses[k]], dtype=_capi.POSE)
                        rc = refpins.ref().formref_world_linearize(world, int(c["i"]), k, _capi.ptr(Ti), _capi.ptr(Tj),
                                                                   params.sigma, _capi.ptr(got))
                        assert rc == 0
                        worst = max(worst, block_rel_err(h, got))
                        checked += 1
                ctx.commit_scan()
                refpins.ref().formref_world_commit(world)
                for t in (0, 1):
                    stored = ctx.keypoints(t, k)
                    buf = np.zeros(len(stored) + 8, dtype=stored.dtype)
                    n = refpins.ref().formref_world_keypoints(world, t, k, _capi.ptr(buf), len(buf))
                    assert n == len(stored) and buf[:n].tobytes() == stored.tobytes()
    finally:
        refpins.ref().formref_world_destroy(world)
    assert checked >= 4 and worst < 1e-9, (checked, worst)  # north-star tolerance: 1e-5
