"""Point-sharded stage 3 (formgpu_set_shard, formgpu_linearize_device; SURVEY 8e mode 2,
BASELINE.json configs[4]): the blocks of the shards add up to the block of the whole pair.
One GPU plays every rank in turn; with two or more GPUs a real two-rank NCCL run sums the
shard blocks with one all-reduce and compares them with the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest

from form_b200 import _capi, synth
from helpers import block_rel_err, gt, perturbed, scan_poses

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_world(ctx, sensor, n_scans, seq=0):
    """Replicated stages 1-2: the same inputs give every rank a bit-identical context."""
    rng = np.random.default_rng(7)
    est, window = {}, []
    for k in range(n_scans):
        ctx.extract(synth.scan(sensor, seq, k), k)
        est[k] = perturbed(gt(seq, k), rng, 0.001, 0.01)
        poses = scan_poses(window + [k], [est[s] for s in window + [k]])
        ctx.map_rebuild(poses)
        ctx.associate(est[k])
        ctx.commit_scan()
        window.append(k)
    poses = scan_poses(window, [est[s] for s in window])
    pairs = np.array([(i, j) for j in window for i in window if i < j], dtype=_capi.PAIR)
    return pairs, poses


def test_shard_blocks_sum_to_the_full_block_on_one_gpu():
    import torch

    from form_b200.context import Context

    rows, cols = synth.shape("vlp-16")
    params = _capi.default_params(rows, cols)
    # a real (non-default) stream shared by torch and the context: the legacy default stream's
    # handle is NULL, which formgpu_create reads as "create a private stream"
    side = torch.cuda.Stream()
    with torch.cuda.stream(side), Context(params, stream=side.cuda_stream) as ctx:
        pairs, poses = _build_world(ctx, "vlp-16", 4)
        full = ctx.linearize(pairs, poses)
        full_err = ctx.error(pairs, poses)
        assert np.abs(full).sum() > 0
        for world in (2, 3, 8):
            acc = torch.zeros(len(pairs) * 91, dtype=torch.float64, device="cuda")
            acc_err = np.zeros(len(pairs))
            acc_host = np.zeros_like(full)
            for rank in range(world):
                ctx.set_shard(rank, world)
                part = torch.empty_like(acc)
                ctx.linearize_device(pairs, poses, part.data_ptr())  # same stream as torch
                acc += part
                acc_host += ctx.linearize(pairs, poses)              # host-output path shards too
                acc_err += ctx.error(pairs, poses)
            ctx.set_shard(0, 1)
            got = acc.cpu().numpy().reshape(len(pairs), 91)
            for a, b, c in zip(got, full, acc_host):
                assert block_rel_err(a, b) < 1e-12
                assert block_rel_err(c, b) < 1e-12
            assert np.allclose(acc_err, full_err, rtol=1e-12, atol=0)
        assert ctx.linearize(pairs, poses).tobytes() == full.tobytes()  # mode off again: identical
        with pytest.raises(Exception):
            ctx.set_shard(2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from form_b200 import multi
    from form_b200.context import Context

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rows, cols = synth.shape("vlp-16")
    params = _capi.default_params(rows, cols)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side), Context(params, device=rank, stream=side.cuda_stream) as ctx:
        pairs, poses = _build_world(ctx, "vlp-16", 4)
        full = ctx.linearize(pairs, poses)
        ctx.set_shard(rank, world)
        summed = multi.sharded_linearize(ctx, pairs, poses)
        errs = multi.sharded_linearize(ctx, pairs, poses, error_only=True)
        ctx.set_shard(0, 1)
        full_err = ctx.error(pairs, poses)
        worst = max(block_rel_err(a, b) for a, b in zip(summed, full))
        ok = worst < 1e-12 and np.allclose(errs, full_err, rtol=1e-12, atol=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        with open(out_path, "w") as f:
            f.write("ok" if ok else f"mismatch {worst}")


def test_two_rank_nccl_allreduce_of_shard_blocks():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import tempfile

    out = tempfile.mktemp()
    mp.spawn(_nccl_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


# ---------------------------------------------------------------------------------------------
# point-sharded mode over NCCL in the C++ library (formgpu_comm_init): keypoint-sharded association
# with an all-gather of the matches, correspondence-sharded pair moments, blocks all-reduced
# ---------------------------------------------------------------------------------------------
def _comm_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle_lib
    from form_b200.context import Context

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # only carries the 128-byte id
    ident = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    sensor, n_scans, seq = "os1-64", 5, 2
    rows, cols = synth.shape(sensor)
    params = _capi.default_params(rows, cols)
    ref = oracle_lib.Oracle(params)
    rng = np.random.default_rng(11)  # same stream on every rank: identical calls
    worst, ok, why = 0.0, True, ""
    with Context(params, device=rank) as ctx:
        ctx.comm_init(ident[0], rank, world)
        est, window = {}, []
        for k in range(n_scans):
            scan = synth.scan(sensor, seq, k)
            ctx.extract(scan, k)
            ref.extract(scan, k)
            est[k] = perturbed(gt(seq, k), rng, 0.001, 0.01)
            poses = scan_poses(window + [k], [est[s] for s in window + [k]])
            ctx.map_rebuild(poses)
            ref.map_rebuild(poses)
            for it in range(2):
                est[k] = perturbed(gt(seq, k), rng, 0.001, 0.01)
                all_poses = scan_poses(window + [k], [est[s] for s in window + [k]])
                if it == 0:
                    counts = ctx.associate(est[k])
                    blocks = None
                else:
                    counts, blocks = ctx.associate_linearize(all_poses)
                rcounts = ref.associate(est[k])
                # every rank holds ALL matches (all-gather), bit-identical to one GPU / the oracle
                if counts.tobytes() != rcounts.tobytes():
                    ok, why = False, f"scan {k}: pair counts"
                for t in (0, 1):
                    if ctx.matches(t).tobytes() != ref.matches(t).tobytes():
                        ok, why = False, f"scan {k}: matches of type {t}"
                if len(rcounts):
                    pairs = np.zeros(len(rcounts), dtype=_capi.PAIR)
                    pairs["i"], pairs["j"] = rcounts["i"], k
                    want = ref.linearize(pairs, all_poses)
                    got = ctx.linearize(pairs, all_poses)  # partial moments -> blocks -> ncclAllReduce
                    for a, b in zip(got, want):
                        worst = max(worst, block_rel_err(a, b))
                    if blocks is not None:
                        for a, b in zip(blocks, want):
                            worst = max(worst, block_rel_err(a, b))
                    e, er = ctx.error(pairs, all_poses), ref.error(pairs, all_poses)
                    if not np.allclose(e, er, rtol=1e-9, atol=1e-12):
                        ok, why = False, f"scan {k}: errors"
            if ctx.commit_scan() != ref.commit_scan():
                ok, why = False, f"scan {k}: novel keypoints"
            for t in (0, 1):
                if ctx.keypoints(t, k).tobytes() != ref.keypoints(t, k).tobytes():
                    ok, why = False, f"scan {k}: stored keypoints"
            window.append(k)
        # full-window relinearisation (more than 48 pairs would take the uploaded-task path; here 10)
        pairs = np.array([(i, j) for j in window for i in window if i < j], dtype=_capi.PAIR)
        all_poses = scan_poses(window, [est[s] for s in window])
        for a, b in zip(ctx.linearize(pairs, all_poses), ref.linearize(pairs, all_poses)):
            worst = max(worst, block_rel_err(a, b))
        ctx.comm_destroy()
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, why, worst))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        bad = [g for g in gathered if not g[0] or not g[2] < 1e-9]
        with open(out_path, "w") as f:
            f.write("ok" if not bad else f"mismatch {bad}")


def test_comm_sharded_sequence_matches_oracle_on_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import tempfile

    out = tempfile.mktemp()
    mp.spawn(_comm_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_comm_world_of_one_is_the_plain_context():
    """formgpu_comm_init with world = 1: a communicator of one rank; results unchanged."""
    from form_b200.context import Context

    rows, cols = synth.shape("vlp-16")
    params = _capi.default_params(rows, cols)
    with Context(params) as plain, Context(params) as solo:
        solo.comm_init(Context.comm_unique_id(), 0, 1)
        pa, posa = _build_world(plain, "vlp-16", 3)
        pb, posb = _build_world(solo, "vlp-16", 3)
        assert plain.linearize(pa, posa).tobytes() == solo.linearize(pb, posb).tobytes()
        solo.comm_destroy()
