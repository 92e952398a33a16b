"""Host -> device input pipeline: the next batch crosses PCIe on a copy stream while the current step computes.

The reference feeds its model straight from the DataLoader (`/root/reference/lib/standard/runner.py:140-146`:
`x = x.to(device)` in front of every step, so the copy and the step serialise).  A 64 x 15 s batch of 80-bin features is
30.7 MB = ~0.55 ms of PCIe time against a 16.7 ms training step (3 %) or a 3.9 ms inference batch (14 %): worth hiding.

    pre = HostPrefetcher(x_host.shape, device)
    pre.submit(x_host)                 # pinned host tensor
    for ...:
        x = pre.get()                  # device tensor; the current stream waits for its copy
        pre.submit(next_host_batch)    # overlaps the step launched below
        step(x)

Two device buffers alternate.  A buffer is refilled only after the work that was enqueued on the consumer stream up to
the NEXT `get()` has finished with it (event recorded there), so `x` stays valid until the second `get()` after it.

`ResultFetcher` is the same idea in the other direction: the reference reads `loss.item()` right behind every step
(`runner.py:147-149`), which parks the host until the GPU has drained and leaves the GPU idle while the host launches the next
step.  `fetch(t)` enqueues the device -> pinned-host copy of a step's result in stream order and returns a handle; its
`value()` is read one step later, while the next step computes:

    fut = None
    for ...:
        loss = step(x)
        nxt = fetcher.fetch(loss)      # D2H of THIS step's loss, asynchronous
        if fut is not None:
            log(fut.value())           # the previous step's loss (its copy finished long ago)
        fut = nxt
    log(fut.value())
"""
from __future__ import annotations

import torch


class HostPrefetcher:
    def __init__(self, shape, device, dtype=torch.float32, depth: int = 2):
        if torch.device(device).type != "cuda":
            raise ValueError("HostPrefetcher needs a CUDA device (there is no CPU fallback)")
        if depth < 2:
            raise ValueError("depth must be >= 2 (one buffer in use, one in flight)")
        self.bufs = [torch.empty(shape, device=device, dtype=dtype) for _ in range(depth)]
        self.stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]        # copy into buffer i finished
        self.free = [None] * depth                                     # consumer finished with buffer i
        self.head = 0                                                  # next buffer to fill
        self.tail = 0                                                  # next buffer to hand out
        self.in_flight = 0
        self.last = None                                               # buffer handed out by the previous get()

    def submit(self, host: torch.Tensor) -> None:
        """Start copying `host` (pinned memory for a truly asynchronous copy) into the next free device buffer."""
        if self.in_flight >= len(self.bufs) - (1 if self.last is not None else 0):
            raise RuntimeError("HostPrefetcher: every buffer is in use; call get() before submitting more")
        if host.shape != self.bufs[0].shape or host.dtype != self.bufs[0].dtype:
            raise ValueError(f"HostPrefetcher: expected {tuple(self.bufs[0].shape)} {self.bufs[0].dtype}, "
                             f"got {tuple(host.shape)} {host.dtype}")
        i = self.head
        if self.free[i] is not None:
            self.stream.wait_event(self.free[i])
        with torch.cuda.stream(self.stream):
            self.bufs[i].copy_(host, non_blocking=True)
            self.ready[i].record(self.stream)
        self.head = (i + 1) % len(self.bufs)
        self.in_flight += 1

    def get(self) -> torch.Tensor:
        """Oldest submitted batch as a device tensor; the current stream waits for its copy."""
        if self.in_flight == 0:
            raise RuntimeError("HostPrefetcher: get() without a submitted batch")
        cur = torch.cuda.current_stream(self.bufs[0].device)
        if self.last is not None:                   # everything enqueued so far may have read the previous buffer
            ev = torch.cuda.Event()
            ev.record(cur)
            self.free[self.last] = ev
        i = self.tail
        cur.wait_event(self.ready[i])
        self.tail = (i + 1) % len(self.bufs)
        self.in_flight -= 1
        self.last = i
        return self.bufs[i]


class _Fetched:
    def __init__(self, host: torch.Tensor, event: torch.cuda.Event):
        self._host, self._event = host, event

    def value(self) -> float:
        """Blocks until the copy has landed (normally it has), then reads the pinned word."""
        self._event.synchronize()
        return float(self._host[0])

    def tensor(self) -> torch.Tensor:
        self._event.synchronize()
        return self._host.clone()


class ResultFetcher:
    """Asynchronous device -> host read-back of small per-step results (loss, metric) through `depth` pinned slots."""

    def __init__(self, device, numel: int = 1, dtype=torch.float32, depth: int = 4):
        if torch.device(device).type != "cuda":
            raise ValueError("ResultFetcher needs a CUDA device (there is no CPU fallback)")
        self.hosts = [torch.zeros(numel, dtype=dtype).pin_memory() for _ in range(depth)]
        self.slot = 0
        self.bytes_per_fetch = numel * self.hosts[0].element_size()

    def fetch(self, t: torch.Tensor) -> _Fetched:
        """Enqueue the copy of `t` (numel elements) on the current stream; the slot is reused `depth` fetches later."""
        h = self.hosts[self.slot]
        self.slot = (self.slot + 1) % len(self.hosts)
        h.copy_(t.detach().reshape(-1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return _Fetched(h, ev)
