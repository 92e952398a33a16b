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
    # gradients living in ONE arena (what the Conformer stack's backward returns): reduced in place, span by span
    deep = torch.nn.Sequential(*[torch.nn.Linear(4, 4) for _ in range(6)])
    with torch.no_grad():
        for p in deep.parameters():
            p.fill_(0.1 * (rank + 1))
    deep(torch.ones(2, 4) * (rank + 1)).sum().backward()
    n = sum(p.numel() for p in deep.parameters())
    arena, off = torch.zeros(n + 12), 4
    local2 = []
    for p in deep.parameters():
        view = arena[off:off + p.numel()].view_as(p)
        view.copy_(p.grad)
        p.grad = view
        local2.append(view.clone())
        off += p.numel()
    # the protocol autograd._ConformerFn.backward follows: attach() -> reduce_span() per finished group of layers (async, in
    # place, no packing) -> join() before the gradients are handed out; finish() then only handles the OTHER parameters
    extra = torch.nn.Linear(3, 2)
    extra(torch.ones(1, 3) * (rank + 1)).sum().backward()
    local3 = [p.grad.clone() for p in extra.parameters()]
    deep.num_layers = 6
    sync2 = dp.GradSync(deep, extra, world=world).attach(deep, groups=2)
    assert [id(p) for p in sync2.loose] == [id(p) for p in extra.parameters()]
    calls = []
    orig = dist.all_reduce
    dist.all_reduce = lambda t, **kw: (calls.append(t.numel()), orig(t, **kw))[1]
    half = n // 2
    sync2.reduce_span(arena[4 + half:4 + n])                    # the later layers finish first
    sync2.reduce_span(arena[4:4 + half])
    sync2.join()
    sync2.finish()
    dist.all_reduce = orig
    g2 = [[torch.zeros_like(g) for _ in range(world)] for g in local2]
    for g, out in zip(local2, g2):
        dist.all_gather(out, g)
    ok &= all(torch.allclose(p.grad, sum(out) / world, atol=1e-6) for p, out in zip(deep.parameters(), g2))
    g3 = [[torch.zeros_like(g) for _ in range(world)] for g in local3]
    for g, out in zip(local3, g3):
        dist.all_gather(out, g)
    ok &= all(torch.allclose(p.grad, sum(out) / world, atol=1e-6) for p, out in zip(extra.parameters(), g3))
    ok &= calls == [n - half, half, sum(p.numel() for p in extra.parameters())]   # two spans in place + one packed bucket
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


# ---------------------------------------------------------------------------------------------- NST sweep (config C5)
def test_plan_sweep_covers_shard_once_and_buckets_by_length():
    import random
    from nn_conformer_for_speech_recognition_b200 import nst
    rng = random.Random(0)
    lengths = [rng.randint(500, 1500) for _ in range(10000)]          # C5: U{500..1500} frames, seed 0
    seen = []
    for rank in range(8):
        plan = nst.plan_sweep(lengths, 64, rank, 8)
        lo, hi = dp.shard_range(10000, rank, 8)
        idx = [i for b in plan for i in b.indices]
        assert sorted(idx) == list(range(lo, hi))                      # the shard, exactly once
        assert len(plan) == -(-(hi - lo) // 64) and all(len(b.indices) <= 64 for b in plan)
        for b in plan:
            assert b.max_len == max(lengths[i] for i in b.indices)
        firsts = [lengths[b.indices[0]] for b in plan]
        assert firsts == sorted(firsts)                                # buckets ascend in length
        padded = sum(b.max_len * len(b.indices) for b in plan)
        unbucketed = sum(b.max_len * len(b.indices) for b in nst.plan_sweep(lengths, 64, rank, 8, bucket_by_length=False))
        true = sum(lengths[lo:hi])
        assert padded < 1.03 * true < 0.8 * unbucketed                 # bucketing removes the padding the in-order batches carry
        seen += idx
    assert sorted(seen) == list(range(10000))
    ref = nst.plan_sweep(lengths, 64, 0, 1, bucket_by_length=False)    # reference order (runner.py:270-275)
    assert [i for b in ref for i in b.indices] == list(range(10000))
    with pytest.raises(ValueError):
        nst.plan_sweep(lengths, 0)


def _sweep_worker(rank, world, port, q):
    from nn_conformer_for_speech_recognition_b200 import nst
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [7 + (i * 13) % 23 for i in range(37)]
    utt = lambda i: torch.full((4, lengths[i]), float(i))              # noqa: E731
    enc = lambda x, lens: (x[:, 0].transpose(1, 2) * 2.0, lens)        # noqa: E731  stand-in encoder: (B,T,F), lengths kept
    part = nst.run_sweep(enc, utt, lengths, 5, rank, world, device="cpu", reduce=lambda y, n: (float(y.sum()), n))
    gathered = [None] * world
    dist.all_gather_object(gathered, part)                             # test-side gather only: the sweep itself has no collective
    merged = nst.merge_shards(gathered, len(lengths))
    ok = all(abs(m[0] - 2.0 * i * 4 * lengths[i]) < 1e-3 and m[1] == lengths[i] for i, m in enumerate(merged))
    q.put((rank, ok, len(part)))
    dist.destroy_process_group()


def test_nst_sweep_world2_gloo_restores_dataset_order():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and sorted(n for _, _, n in res) == [18, 19], res


def test_plan_sweep_edge_cases():
    from nn_conformer_for_speech_recognition_b200 import nst
    assert nst.plan_sweep([], 8) == []
    assert nst.plan_sweep([5, 7], 8, rank=3, world=4) == []                      # more ranks than utterances
    one = nst.plan_sweep([5, 7, 6], 2, rank=0, world=1)
    assert [b.indices for b in one] == [[0, 2], [1]] and [b.max_len for b in one] == [6, 7]
    assert nst.merge_shards([{0: "a", 2: "c"}, {1: "b"}], 3) == ["a", "b", "c"]
    with pytest.raises(ValueError):
        nst.merge_shards([{0: "a"}], 2)
