"""World-size-2 gloo test of the N>1 plumbing (form_b200/multi.py): sequence partition,
max-over-ranks timing, and the point-sharded block all-reduce checked against the
unsharded oracle block.  CPU only."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib
    from form_b200 import _capi, multi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert multi.sequence_of_rank(rank, world) == rank
    # slowest rank decides
    assert multi.max_over_ranks(1.0 + rank) == float(world)
    # same correspondences on every rank (seeded), each linearises its shard
    rng = np.random.default_rng(99)
    n, m, sigma = 1001, 333, 0.1
    p_i, p_j = rng.normal(size=(n, 3)) * 6, rng.normal(size=(n, 3)) * 6
    n_i = rng.normal(size=(n, 3))
    n_i /= np.linalg.norm(n_i, axis=1, keepdims=True)
    q_i, q_j = rng.normal(size=(m, 3)) * 6, rng.normal(size=(m, 3)) * 6
    Ti = np.zeros(1, dtype=_capi.POSE)
    Tj = np.zeros(1, dtype=_capi.POSE)
    Ti["R"] = np.eye(3).reshape(9)
    c, s = np.cos(0.3), np.sin(0.3)
    Tj["R"] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]).reshape(9)
    Tj["t"] = [0.4, -0.2, 0.1]

    def lin(a0, a1, b0, b1):
        out, err = np.zeros(91), C.c_double()
        pi, ni, pj = (np.ascontiguousarray(x[a0:a1]) for x in (p_i, n_i, p_j))
        qi, qj = (np.ascontiguousarray(x[b0:b1]) for x in (q_i, q_j))
        oracle_lib.lib().oracle_linearize_raw(_capi.ptr(pi), _capi.ptr(ni), _capi.ptr(pj), a1 - a0, _capi.ptr(qi),
                                              _capi.ptr(qj), b1 - b0, _capi.ptr(Ti), _capi.ptr(Tj), sigma,
                                              _capi.ptr(out), C.byref(err))
        return out, err.value

    a0, a1 = multi.shard_bounds(n, rank, world)
    b0, b1 = multi.shard_bounds(m, rank, world)
    part, perr = lin(a0, a1, b0, b1)
    total = multi.allreduce_blocks(part[None, :])[0]
    terr = multi.allreduce_blocks(np.array([perr]))[0]
    full, ferr = lin(0, n, 0, m)
    scale = np.abs(full).max()
    assert np.max(np.abs(total - full)) < 1e-12 * scale
    assert abs(terr - ferr) < 1e-12 * ferr
    # shards tile the range exactly
    bounds = [multi.shard_bounds(n, r, world) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_two_rank_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
