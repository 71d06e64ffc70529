"""Host-side input pipeline for callers whose embeddings arrive from pinned host memory (evaluation caches, the
benchmark's end-to-end leg): a double-buffered prefetcher that issues the host->device copies of batch k+1 on a
side stream while batch k is being processed on the compute stream. Pure stream/event plumbing, no math."""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Tuple

import torch


class HostBatchPrefetcher:
    """Iterates device copies of pinned host batches, overlapping each H2D copy with the previous step.

    ``batches`` yields tuples of pinned CPU tensors. Each yielded tuple lives in one of two device buffer sets; it is
    valid until the iterator is advanced twice. Every copy is issued inside ``__next__`` of the step before it is
    consumed, so per-step H2D traffic stays inside any timed region that brackets the loop."""

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device: torch.device):
        self.it = iter(batches)
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot = 0
        self.pending = None
        self._issue()

    def _issue(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            self.pending = None
            return
        s = self.slot
        if self.bufs[s] is None:
            self.bufs[s] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
        else:
            self.copy_stream.wait_event(self.consumed[s])      # the step that used this buffer set has finished
        with torch.cuda.stream(self.copy_stream):
            for d, h in zip(self.bufs[s], host):
                d.copy_(h, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self.pending = s
        self.slot ^= 1

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        return self

    def __next__(self) -> Tuple[torch.Tensor, ...]:
        if self.pending is None:
            raise StopIteration
        s = self.pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        out = self.bufs[s]
        self._issue()                                          # next batch's copies overlap this step's kernels
        return out

    def release(self, batch: Tuple[torch.Tensor, ...]) -> None:
        """Marks ``batch`` as consumed on the current stream (call after the last kernel that reads it is queued)."""
        for s in (0, 1):
            if self.bufs[s] is batch:
                self.consumed[s].record(torch.cuda.current_stream(self.device))
