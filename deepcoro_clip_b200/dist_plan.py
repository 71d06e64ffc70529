"""Host-side sharding plan of the contrastive head (pure PyTorch tensor plumbing + torch.distributed; no kernels),
kept separate so it can be exercised on CPU with the gloo backend (tests/test_dist_plan_gloo.py).

CLIP / InfoNCE over W ranks, global batch N = W * B:
  * every rank L2-normalises its B rows and all-gathers the bf16 operands -> [N, ld];
  * rank r owns the ROW SLAB  logits[r*B:(r+1)*B, :]: row sums complete locally; column sums are partial over the
    slab rows -> all_reduce(SUM) [N]; row sums all_gather -> [N] (the text-side backward needs every row's scale);
  * loss = 0.5/N * (sum_i r_i + sum_j c_j) - sum_tgt / N is assembled identically on every rank;
  * backward: dVhat for the slab rows needs nothing else; dThat for the rank's OWN text rows is computed against all
    video rows (roles swapped), so no [N, D] reduce-scatter is needed; the log_temp gradient is one scalar all-reduce.
SigLIP: row-parallel, text replicated; loss / dbias / dlog_temp scalars and dThat [T, D] are all-reduced.
Retrieval: text database sharded by rows; rank counts all-reduced, per-shard top-k lists all-gathered and merged."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world(use_ddp: bool = True, group=None) -> Tuple[int, int]:
    if use_ddp and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def text_shard(n_text: int, world_size: int, rank: int) -> Tuple[int, int]:
    per = (n_text + world_size - 1) // world_size
    lo = min(rank * per, n_text)
    return lo, min(lo + per, n_text)


def gather_rows(x: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """[B, ...] on every rank -> [W*B, ...] (rank-major), one contiguous all_gather_into_tensor."""
    if world_size == 1:
        return x
    out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def gather_rows_async(x: torch.Tensor, world_size: int, group=None):
    """Like gather_rows but returns (out, work) with the all-gather running on the communicator's stream; call
    ``work.wait()`` (a stream dependency, not a host sync) before the first kernel that reads ``out``."""
    if world_size == 1:
        return x, None
    out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    work = dist.all_gather_into_tensor(out, x.contiguous(), group=group, async_op=True)
    return out, work
