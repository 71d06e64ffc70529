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


class GraphedLossStep:
    """One loss forward + backward captured in a CUDA graph (kernels AND the NCCL collectives of the sharded path) and
    replayed: removes the per-step host work (≈25 kernel launches, four collective enqueues, a dozen allocations) that
    makes the eager step host-bound at 8 ranks. Under torch.distributed drop the object (``del step; gc.collect()``)
    BEFORE ``destroy_process_group()`` — a live graph holding NCCL kernels keeps the teardown waiting. Inputs live in static
    buffers: ``step(video, text)`` copies the new batch in (device or pinned-host tensors), replays, and returns
    ``(loss, dvideo, dtext, dlog_temp)`` — views of static tensors, valid until the next call. Shapes, dtypes and the loss
    configuration are fixed at construction; the parity of a replayed step with the eager module is covered by
    tests/test_gpu_clip_loss.py::test_graphed_step_matches_eager.

    Parameters: the temperature is held in a PRIVATE static tensor (a clone of the ``log_temp`` given at construction), so
    a learnable temperature must be handed to every step — ``step(video, text, log_temp=param)`` copies its current value
    in (device copy, no sync) before the replay; ``dlog_temp`` is the gradient for the caller's parameter. The loss
    module's own parameters (SigLIP ``bias``) are captured in place: update them in place (every torch optimizer does) and
    read their ``.grad`` after the step."""

    def __init__(self, loss_module, video: torch.Tensor, text: torch.Tensor, log_temp: torch.Tensor, warmup: int = 3,
                 **forward_kwargs):
        dev = video.device
        self.loss_module = loss_module
        self.kwargs = forward_kwargs
        self.video = video.detach().clone().requires_grad_(True)
        self.text = text.detach().clone().requires_grad_(True)
        self.log_temp = log_temp.detach().clone().requires_grad_(True)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # library attribute calls, NCCL communicators, allocator warm-up
                self._zero()
                self._eager().backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
            self.loss.backward()

    def _zero(self) -> None:
        self.video.grad = None
        self.text.grad = None
        self.log_temp.grad = None
        for p in self.loss_module.parameters():
            p.grad = None

    def _eager(self) -> torch.Tensor:
        return self.loss_module(video_features=self.video, text_features=self.text, log_temp=self.log_temp,
                                **self.kwargs)

    def step(self, video: torch.Tensor = None, text: torch.Tensor = None, log_temp: torch.Tensor = None):
        with torch.no_grad():
            if video is not None:
                self.video.copy_(video, non_blocking=True)
            if text is not None:
                self.text.copy_(text, non_blocking=True)
            if log_temp is not None:
                self.log_temp.copy_(log_temp.detach().reshape(self.log_temp.shape), non_blocking=True)
        self.graph.replay()
        return self.loss, self.video.grad, self.text.grad, self.log_temp.grad

    __call__ = step
