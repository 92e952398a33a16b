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
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.buckets = make_buckets([p.numel() for p in self.params], self.bucket_elems)
        self.avg = dist.get_backend() == "nccl"

    def broadcast_parameters(self, src: int = 0):
        for t in self.params + self.buffers:
            dist.broadcast(t.data, src)

    def finish(self):
        """All-reduce (mean) every gradient; returns when the results are visible to the current stream.

        Gradients that are views of ONE buffer (the Conformer stack returns all of its ~500 parameter gradients as slices
        of a single fp32 arena, autograd.py) are reduced IN PLACE as one flat span: no packing, no per-parameter
        unpacking copies (those were ~500 launches = a tenth of the step at world size 2).  Everything else goes
        through the bucketed pack / all-reduce / multi-tensor unpack path."""
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        by_storage = {}
        for p in self.params:
            g = p.grad
            if g is None:
                continue
            by_storage.setdefault(g.untyped_storage().data_ptr(), []).append(g)
        pending, loose = [], []
        for grads in by_storage.values():
            g0 = grads[0]
            if len(grads) >= 8 and all(g.is_contiguous() and g.dtype == g0.dtype for g in grads):
                lo = min(g.storage_offset() for g in grads)
                hi = max(g.storage_offset() + g.numel() for g in grads)
                if hi - lo <= 1.25 * sum(g.numel() for g in grads):         # dense enough: reduce the span itself
                    flat = torch.empty(0, device=g0.device, dtype=g0.dtype).set_(g0.untyped_storage(), lo, (hi - lo,))
                    pending.append((dist.all_reduce(flat, op=op, async_op=True), flat, None))
                    continue
            loose += grads
        for idx in make_buckets([g.numel() for g in loose], self.bucket_elems):
            grads = [loose[i] for i in idx]
            flat = torch.cat([g.reshape(-1) for g in grads])
            pending.append((dist.all_reduce(flat, op=op, async_op=True), flat, grads))
        for work, flat, grads in pending:
            work.wait()
            if not self.avg:
                flat.div_(self.world)
            if grads is not None:
                outs, off = [], 0
                for g in grads:
                    outs.append(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
                torch._foreach_copy_(grads, outs)
