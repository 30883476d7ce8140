"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np

from form_b200 import _capi, synth

IU = np.triu_indices(13)


def unpack91(v):
    """Packed upper triangle (91) -> symmetric 13x13."""
    M = np.zeros((13, 13))
    M[IU] = v
    return M + np.triu(M, 1).T


def block_rel_err(a91, b91):
    """Scale-aware relative error between two 13x13 information blocks: each entry is
    compared against sqrt(G_aa * G_bb) (the Cauchy-Schwarz bound of that entry), so
    entries that vanish by cancellation do not blow the ratio up."""
    A, B = unpack91(a91), unpack91(b91)
    d = np.sqrt(np.maximum(np.diag(B), 0.0))
    scale = np.outer(d, d)
    scale[scale == 0] = 1.0
    return float(np.max(np.abs(A - B) / scale))


def scan_poses(scans, poses):
    out = np.zeros(len(scans), dtype=_capi.SCAN_POSE)
    for n, (s, p) in enumerate(zip(scans, poses)):
        out[n]["scan"] = s
        out[n]["R"] = p["R"]
        out[n]["t"] = p["t"]
    return out


def perturbed(pose, rng, rot=0.002, trans=0.02):
    """Pose composed with a small random rotation / translation."""
    w = rng.normal(scale=rot, size=3)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    dR = np.eye(3) + (np.sin(th) / th if th > 0 else 1.0) * K + ((1 - np.cos(th)) / th**2 if th > 0 else 0.5) * K @ K
    out = np.zeros((), dtype=_capi.POSE)
    out["R"] = (pose["R"].reshape(3, 3) @ dR).reshape(9)
    out["t"] = pose["t"] + rng.normal(scale=trans, size=3)
    return out


def gt(seq, k):
    return synth.gt_pose(seq, k)
