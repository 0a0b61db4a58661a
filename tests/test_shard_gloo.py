"""N > 1 host logic on CPU: two gloo ranks shard a batch, each evaluates its slice (with the
oracle standing in for the device, since there is no GPU here), rank 0 reassembles and checks the
result equals the unsharded evaluation; timings are MAX-reduced like bench.py does."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from leela_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leela_b200 import synth
    from oracle import oracle
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))
    n = 11                                   # not divisible by the world size
    lo, hi = shard.shard_range(n, rank, world)
    vn = oracle.OracleNet(synth.value_weights())
    local = oracle.value_forward(vn, g["value_planes"][lo:hi], g["rotation"][lo:hi])
    full = shard.gather_sharded(local.reshape(-1, 1), n, rank, world, dist)
    t = shard.max_over_ranks([10.0 + rank, 5.0 - rank], dist)
    dist.barrier()
    if rank == 0:
        q.put((full.reshape(-1), t, g["value"][:n]))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29533 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    full, t, want = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert np.abs(full - want).max() < 2e-5          # sharded == the reference's unsharded answers
    assert t == [11.0, 5.0]                           # MAX over ranks


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 4, 8):
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    assert shard.aggregate_throughput(256, 10, 8, 5.0) == pytest.approx(8 * 256 * 10 / 5e-3)
    assert shard.batch_order(4, 1) == [1, 2, 3, 0]
