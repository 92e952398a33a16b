"""Noisy-student pseudo-labelling sweep: the teacher encoder's eval forward over a
pool of unlabeled utterances, sharded data-parallel with NO collective.

Reference loop: `Runner.generate_labels` (/root/reference/lib/standard/runner.py:255-281) walks the dataset in
order, `batch_size` utterances at a time, `model.eval()` + `torch.no_grad()`, and concatenates the per-batch
results, so the output order IS dataset order.  Here

  * rank r owns the contiguous utterance range `dp.shard_range(N, r, world)` (concatenating the ranks' results in
    rank order restores dataset order),
  * inside a shard utterances are bucketed by length (sorted, `batch_size` per bucket) so a batch is padded only to
    ITS longest utterance -- the kernels skip key blocks past `lengths[b]` and the tail tiles vanish,
  * results are scattered back to the utterances' original positions, so bucketing is invisible to the caller.

`plan_sweep` is pure host logic (tested on CPU, incl. world_size-2 gloo); `run_sweep` drives any callable
`encode(x (B,1,F,T) cuda, lengths (B,) cuda) -> (y (B,T',d), out_lengths (B,))`, e.g. `ConformerEncoder`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Sequence

import torch

from .dp import shard_range  # noqa: F401  (re-exported)


@dataclass
class SweepBatch:
    indices: List[int]      # dataset positions of the utterances in this batch (ascending length)
    max_len: int            # frames the batch is padded to


def plan_sweep(lengths: Sequence[int], batch_size: int, rank: int = 0, world: int = 1,
               bucket_by_length: bool = True) -> List[SweepBatch]:
    """Batches for `rank`: its contiguous shard, sorted by (length, position) when bucketing, cut every `batch_size`.
    bucket_by_length=False reproduces the reference's in-order batches (runner.py:270-275)."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    lo, hi = shard_range(len(lengths), rank, world)
    idx = list(range(lo, hi))
    if bucket_by_length:
        idx.sort(key=lambda i: (int(lengths[i]), i))
    out = []
    for s in range(0, len(idx), batch_size):
        chunk = idx[s:s + batch_size]
        out.append(SweepBatch(chunk, max(int(lengths[i]) for i in chunk)))
    return out


def run_sweep(encode: Callable, get_utterance: Callable[[int], torch.Tensor], lengths: Sequence[int], batch_size: int,
              rank: int = 0, world: int = 1, bucket_by_length: bool = True, device: str = "cuda",
              reduce: Callable = None):
    """Runs `encode` over this rank's shard.  `get_utterance(i)` returns the fp32 log-mel (F, len_i) of utterance i
    (host or device).  Returns {dataset position: reduce(y_i (T'_i, d), out_len_i)} for the shard (default reduce keeps
    the valid frames of the encoder output on the device).  No collective is used; the caller gathers per-rank
    dictionaries in rank order if it needs the global list (`merge_shards`)."""
    results = {}
    with torch.no_grad():
        for sb in plan_sweep(lengths, batch_size, rank, world, bucket_by_length):
            first = get_utterance(sb.indices[0])
            F = first.shape[0]
            x = torch.zeros((len(sb.indices), 1, F, sb.max_len), device=device, dtype=torch.float32)
            for j, i in enumerate(sb.indices):
                u = first if j == 0 else get_utterance(i)
                x[j, 0, :, :u.shape[1]] = u.to(device, non_blocking=True)
            lens = torch.tensor([int(lengths[i]) for i in sb.indices], device=device)
            y, out_lens = encode(x, lens)
            ol = out_lens.tolist()
            for j, i in enumerate(sb.indices):
                yi = y[j, :ol[j]]
                results[i] = reduce(yi, ol[j]) if reduce is not None else yi.clone()
    return results


def merge_shards(per_rank: Sequence[dict], n_items: int) -> list:
    """Per-rank result dictionaries -> one list in dataset order (what `generate_labels` returns)."""
    merged = {}
    for d in per_rank:
        merged.update(d)
    if sorted(merged) != list(range(n_items)):
        raise ValueError("shards do not cover the dataset exactly once")
    return [merged[i] for i in range(n_items)]
