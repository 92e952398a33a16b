"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch on the GPU box,
gloo in the CPU tests).  The path shards by utterance: inference and NST pseudo-labelling use
NO collective; a training step has exactly one exchange -- the gradient all-reduce (mean) -- issued in buckets.

See `GradSync`: the Conformer stack's gradient arena is all-reduced in layer-group spans from INSIDE the backward pass
(overlapping the rest of it); `GradSync.finish()`, called between backward and the optimizer step, handles the few
remaining gradients in packed buckets.
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
    """Gradient exchange of a training step: exactly one all-reduce (mean) per gradient element, issued as early as possible.

    * The Conformer stack computes all of its ~500 parameter gradients into ONE fp32 arena (autograd.py), last layer first.
      `attach(conformer, groups)` makes its backward pass hand the arena span of every finished group of layers to
      `reduce_span`, which all-reduces it IN PLACE and asynchronously -- NCCL works on the spans of layers 12-15 while the
      backward kernels of layers 0-11 still run (and, when the backward pass is captured into a CUDA graph, the collectives
      are captured with it as a parallel branch).  The backward pass joins them before it returns.
    * Everything else (conv-subsampling, bridge, head: their gradients are the LAST the backward pass produces) goes through
      `finish()`: packed into flat buckets, all-reduced, unpacked with one multi-tensor copy.  The parameter list of that
      path is fixed at construction time (no per-step grouping by storage).
    BatchNorm statistics stay local to each rank, as in the single-device reference; parameters and buffers are broadcast
    from rank 0 once."""

    def __init__(self, *modules, world: int = None, bucket_mb: float = 32.0):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad]
        self.buffers = [b for m in modules for b in m.buffers()]
        self.world = world if world is not None else dist.get_world_size()
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.avg = dist.get_backend() == "nccl"
        self.op = dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM
        self.pending = []
        self.in_backward_ids = set()           # parameters whose gradients are reduced inside the Conformer backward
        self.loose = list(self.params)
        self.buckets = make_buckets([p.numel() for p in self.loose], self.bucket_elems)

    def attach(self, conformer, groups: int = 4):
        """Reduce the Conformer stack's gradient arena from inside its backward pass, in `groups` spans of layers."""
        conformer._grad_sync = self
        conformer._grad_sync_groups = max(1, min(groups, conformer.num_layers))
        self.in_backward_ids = {id(p) for p in conformer.parameters()}
        self.loose = [p for p in self.params if id(p) not in self.in_backward_ids]
        self.buckets = make_buckets([p.numel() for p in self.loose], self.bucket_elems)
        return self

    def broadcast_parameters(self, src: int = 0):
        for t in self.params + self.buffers:
            dist.broadcast(t.data, src)

    # ---- called by autograd._ConformerFn.backward --------------------------------------------------------------------
    def reduce_span(self, flat: torch.Tensor):
        if self.world == 1:
            return
        self.pending.append((dist.all_reduce(flat, op=self.op, async_op=True), flat))

    def join(self):
        for work, flat in self.pending:
            work.wait()
            if not self.avg:
                flat.div_(self.world)
        self.pending = []

    # ---- called between backward and the optimizer step ---------------------------------------------------------------
    def finish(self):
        """All-reduce (mean) the gradients the backward pass has not already reduced; returns when the results are visible
        to the current stream."""
        if self.world == 1:
            return
        self.join()
        pending = []
        for idx in self.buckets:
            grads = [self.loose[i].grad for i in idx if self.loose[i].grad is not None]
            if not grads:
                continue
            flat = torch.cat([g.reshape(-1) for g in grads])
            pending.append((dist.all_reduce(flat, op=self.op, async_op=True), flat, grads))
        for work, flat, grads in pending:
            work.wait()
            if not self.avg:
                flat.div_(self.world)
            outs, off = [], 0
            for g in grads:
                outs.append(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            torch._foreach_copy_(grads, outs)
