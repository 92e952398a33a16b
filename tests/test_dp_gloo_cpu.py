"""Data-parallel host logic with world_size 2 on the gloo backend (CPU): sharding, bucketing, gradient averaging."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nn_conformer_for_speech_recognition_b200 import dp


def test_shard_range_partitions_in_rank_order():
    for n in [0, 1, 7, 10000, 10001]:
        for world in [1, 2, 3, 8]:
            parts = [dp.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert dp.shard_range(10000, 3, 8) == (3750, 5000)          # C5: 1250 utterances per GPU


def test_make_buckets_preserves_order_and_limits():
    sizes = [10, 20, 100, 5, 5, 5, 200, 1]
    b = dp.make_buckets(sizes, 100)
    assert [i for bucket in b for i in bucket] == list(range(len(sizes)))
    for bucket in b:
        assert sum(sizes[i] for i in bucket) <= 100 or len(bucket) == 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                        # same init on both ranks ...
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    with torch.no_grad():
        for p in model.parameters():
            p.add_(rank)                                        # ... then perturbed, so the broadcast has work to do
    sync = dp.GradSync(model, world=world, bucket_mb=1e-4)      # tiny buckets: several all-reduces
    sync.broadcast_parameters()
    assert len(sync.buckets) > 1
    x = torch.arange(24, dtype=torch.float32).reshape(4, 6) * (rank + 1)
    model(x).square().sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    sync.finish()
    gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
    for g, out in zip(local, gathered):
        dist.all_gather(out, g)
    ok = all(torch.allclose(p.grad, sum(out) / world, atol=1e-6) for p, out in zip(model.parameters(), gathered))
    same_params = True
    for p in model.parameters():
        ref = p.data.clone()
        dist.broadcast(ref, 0)
        same_params &= bool(torch.equal(ref, p.data))
    q.put((rank, ok, same_params))
    dist.destroy_process_group()


def test_grad_sync_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and same for _, ok, same in res), res
