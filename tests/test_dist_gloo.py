"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: BN-group sharding of inference / evaluation and the
data-parallel gradient-bucket exchange.  The kernels themselves need a GPU; what is checked here is
that the partition covers every pose exactly once at batch boundaries and that the collectives reproduce the
single-process sums."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lcn_pose_b200.dist import (all_reduce_eval_sums, average_gradient_bucket, broadcast_parameters, shard_groups)
from oracle import lcn_oracle as O


def test_shard_groups_partition_properties():
    for n, bs in ((1000, 256), (256, 256), (257, 256), (5, 200), (64 << 20, 256), (1000, 200)):
        for world in (1, 2, 4, 8):
            spans = [shard_groups(n, bs, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            for a0, a1 in spans:
                assert a0 % bs == 0 or a0 == a1 == n      # shards start at BN-group boundaries (or are empty)
            sizes = [(a1 - a0 + bs - 1) // bs for a0, a1 in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, bs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        gt = rng.normal(0, 300, (n, 17, 3)) + np.array([0, 0, 4500.0])
        pred = gt + rng.normal(0, 25, gt.shape)
        # (1) sharded evaluation: per-rank oracle errors on the rank's BN groups, summed with one all-reduce
        r0, r1 = shard_groups(n, bs, rank, world)
        err = np.linalg.norm(pred[r0:r1] - gt[r0:r1], axis=2)            # Protocol-1 error in the camera frame
        sums = torch.zeros((1, 19), dtype=torch.float64)
        sums[0, :17] = torch.from_numpy(err.sum(0))
        sums[0, 17] = r1 - r0
        all_reduce_eval_sums(sums)
        # (2) data-parallel exchange: mean of the per-rank gradient buckets
        g = torch.from_numpy(np.random.default_rng(100 + rank).normal(0, 1, 1000).astype(np.float32))
        average_gradient_bucket(g)
        # (3) replicas start from rank 0's parameters
        p = torch.full((16,), float(rank + 1))
        broadcast_parameters(p)
        if rank == 0:
            out.put((sums.numpy().copy(), g.numpy().copy(), p.numpy().copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_eval_and_gradient_exchange_match_single_process():
    n, bs, world = 1000, 256, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bs, out)) for r in range(world)]
    for p in procs:
        p.start()
    sums, g, par = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    gt = rng.normal(0, 300, (n, 17, 3)) + np.array([0, 0, 4500.0])
    pred = gt + rng.normal(0, 25, gt.shape)
    ref = O.eval_errors(pred, gt, None, None, None, False, camera_frame=True) if "camera_frame" in O.eval_errors.__code__.co_varnames \
        else np.linalg.norm(pred - gt, axis=2)
    assert np.allclose(sums[0, :17], ref.sum(0), rtol=1e-12)
    assert sums[0, 17] == n
    g_ref = np.mean([np.random.default_rng(100 + r).normal(0, 1, 1000).astype(np.float32) for r in range(world)], axis=0)
    assert np.allclose(g, g_ref, atol=1e-6)
    assert np.all(par == 1.0)
