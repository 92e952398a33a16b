"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch on the GPU box,
gloo in the CPU tests).  The path shards by utterance: inference and NST pseudo-labelling use
NO collective; a training step has exactly one exchange -- the gradient all-reduce (mean) -- issued in buckets.

`GradSync.finish()` is called between backward and the optimizer step.  Gradients are packed into flat buckets of
about `bucket_mb` MB, each bucket is all-reduced asynchronously (NCCL runs them on its own stream, overlapping the
packing of the next bucket), then unpacked.  BatchNorm statistics stay local to each rank, as in the single-device
reference; parameters and buffers are broadcast from rank 0 once.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `n_items` utterances for `rank` (sizes differ by at most one; concatenating the
    shards in rank order restores dataset order, as `Runner.generate_labels` requires, runner.py:269-277)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_buckets(sizes: List[int], bucket_elems: int) -> List[List[int]]:
    """Greedy, order-preserving partition of parameter indices into buckets of at most `bucket_elems` elements
    (a parameter larger than the limit gets a bucket of its own)."""
    buckets, cur, cur_n = [], [], 0
    for i, n in enumerate(sizes):
        if cur and cur_n + n > bucket_elems:
            buckets.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += n
    if cur:
        buckets.append(cur)
    return buckets


class GradSync:
    def __init__(self, *modules, world: int = None, bucket_mb: float = 32.0):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad]
        self.buffers = [b for m in modules for b in m.buffers()]
        self.world = world if world is not None else dist.get_world_size()
        self.buckets = make_buckets([p.numel() for p in self.params], int(bucket_mb * (1 << 20) / 4))
        self.avg = dist.get_backend() == "nccl"

    def broadcast_parameters(self, src: int = 0):
        for t in self.params + self.buffers:
            dist.broadcast(t.data, src)

    def finish(self):
        """All-reduce (mean) every gradient; returns when the results are visible to the current stream."""
        if self.world == 1:
            return
        pending = []
        for idx in self.buckets:
            grads = [self.params[i].grad for i in idx if self.params[i].grad is not None]
            if not grads:
                continue
            flat = torch.cat([g.reshape(-1) for g in grads])
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, async_op=True)
            pending.append((work, flat, grads))
        for work, flat, grads in pending:
            work.wait()
            if not self.avg:
                flat.div_(self.world)
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
