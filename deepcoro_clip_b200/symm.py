"""Symmetric-memory plan of the multi-GPU row-slab CLIP path (one 8 x B200 box, NVLink 5 / NVSwitch).

The NCCL version of the step had three collectives on its critical path: the all-gather of the text operands in front of
the forward, the all-gather of the video operands, and the latency-bound all-reduce of the per-row / per-column statistics
in front of the finalize kernel (~100 us of a 0.75 ms step at 8 ranks). Here the exchange is done by
the producing / consuming kernels themselves over peer memory:

  * operands : the normalise kernel stores every bf16 operand row into the [N, ld] operand buffer of EVERY rank
               (b200clip_l2norm_fwd_multi: st.global to the peers' symmetric-memory blocks) - normalise + all-gather in one
               launch, followed by ONE cross-rank barrier for both operands;
  * statistics: every rank accumulates into its own block; after a second barrier the finalize kernel of every rank reads
               the W blocks directly (b200clip_clip_finalize_peers) - a one-shot exchange instead of an all-reduce.

Buffers are allocated once per (group, N, ld) with torch.distributed._symmetric_memory (PyTorch supplies the allocation,
the peer pointers and the barrier kernel: plumbing) and DOUBLE-BUFFERED by step parity: step k + 2 may overwrite the slot
of step k only after the barriers of step k + 1, which every rank reaches after finishing step k in stream order. A third
forward before the backward of the first (three live autograd graphs on the same plan) would break that; the plan counts
live slots and falls back to the NCCL path instead.

Opt-out: B200CLIP_SYMM=0 keeps the NCCL collectives (the A/B baseline; also used automatically when symmetric memory
cannot be set up - e.g. gloo / CPU tests, bf16x3 operands, odd embedding widths)."""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import call as _call, stream_ptr as _stream_ptr

_PLANS: Dict[Tuple, "SymmPlan"] = {}
_FAILED = set()


def enabled() -> bool:
    return os.environ.get("B200CLIP_SYMM", "1") != "0"


class SymmPlan:
    """Symmetric operand / statistics buffers of one (process group, N, ld) CLIP problem."""

    NVEC = 7

    def __init__(self, group, N: int, ld: int, dev: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        pg = group if group is not None else dist.group.WORLD
        self.W, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
        self.N, self.ld = N, ld
        self.dev_index = dev.index
        # [slot][video | text][N, ld] bf16 operands; [slot][NVEC * N] fp32 statistics
        self.ops = symm_mem.empty((2, 2, N, ld), dtype=torch.bfloat16, device=dev)
        # per slot: NVEC * N fp32 statistics + 16 floats (8 fp64: the backward's scalar sums, sum G L first)
        self.stats = symm_mem.empty((2, self.NVEC * N + 16), dtype=torch.float32, device=dev)
        self.flags = symm_mem.empty((64,), dtype=torch.int32, device=dev)      # barrier epochs (b200clip_symm_barrier)
        self.flags.zero_()
        self.h_ops = symm_mem.rendezvous(self.ops, pg)
        self.h_stats = symm_mem.rendezvous(self.stats, pg)
        self.h_flags = symm_mem.rendezvous(self.flags, pg)
        self.h_flags.barrier(channel=0)       # every rank's flags are zero before the first epoch is written
        self.own_barrier = os.environ.get("B200CLIP_SYMM_BARRIER", "own") != "torch"
        self.step = 0
        self.live = [None, None]              # weak references to the autograd contexts that still need a slot's operands
        op_bytes = N * ld * 2
        st_bytes = (self.NVEC * N + 16) * 4

        def arr(ptrs):
            return (ctypes.c_void_p * len(ptrs))(*ptrs)
        # peer pointer tables (host arrays handed to the C ABI): [slot][video | text] and [slot]
        self.op_ptrs = [[arr([int(p) + (2 * s + side) * op_bytes for p in self.h_ops.buffer_ptrs]) for side in (0, 1)]
                        for s in (0, 1)]
        self.stat_ptrs = [arr([int(p) + s * st_bytes for p in self.h_stats.buffer_ptrs]) for s in (0, 1)]
        # multicast mapping of the operand buffer (NVLS), when the platform provides one: rows leave the GPU once
        mc = int(getattr(self.h_ops, "multicast_ptr", 0) or 0) if os.environ.get("B200CLIP_SYMM_MC", "1") != "0" else 0
        self.mc_op = [[mc + (2 * s + side) * op_bytes for side in (0, 1)] for s in (0, 1)] if mc else None
        self.scal_ptrs = [arr([int(p) + s * st_bytes + self.NVEC * N * 4 for p in self.h_stats.buffer_ptrs]) for s in (0, 1)]
        self.flag_ptrs = arr([int(p) for p in self.h_flags.buffer_ptrs])

    def acquire(self, owner) -> Optional[int]:
        """Next slot, or None when the slot is still referenced by a live autograd graph (caller uses the NCCL path)."""
        s = self.step & 1
        ref = self.live[s]
        if ref is not None and ref() is not None:
            return None
        self.live[s] = weakref.ref(owner) if owner is not None else None
        self.step += 1
        return s

    def release(self, slot: int, owner) -> None:
        ref = self.live[slot]
        if ref is not None and ref() is owner:
            self.live[slot] = None

    def _barrier(self, channel: int, handle) -> None:
        if self.own_barrier:
            _call("symm_barrier", self.flag_ptrs, self.W, self.rank, channel, _stream_ptr(self.dev_index))
        else:
            handle.barrier(channel=0)

    def barrier_ops(self) -> None:
        self._barrier(0, self.h_ops)

    def barrier_stats(self) -> None:
        self._barrier(1, self.h_stats)

    def barrier_scal(self) -> None:
        self._barrier(2, self.h_stats)


def get_plan(group, N: int, ld: int, dev: torch.device) -> Optional[SymmPlan]:
    """The cached plan, created collectively on first use (every rank of the group reaches this call with the same
    arguments: the loss is called by all ranks in lock step). None when symmetric memory is unavailable."""
    if not enabled() or not dist.is_available() or not dist.is_initialized():
        return None
    pg = group if group is not None else dist.group.WORLD
    if dist.get_backend(pg) != "nccl" or dist.get_world_size(pg) > 8:
        return None
    key = (id(pg), N, ld, dev.index)
    if key in _PLANS:
        return _PLANS[key]
    if (id(pg), dev.index) in _FAILED:
        return None
    ok = torch.ones(1, device=dev)
    plan = None
    try:
        plan = SymmPlan(pg, N, ld, dev)
    except Exception as e:       # noqa: BLE001 - any failure of the optional fast path selects the NCCL path on ALL ranks
        ok.zero_()
        if dist.get_rank(pg) == 0:
            print(f"[deepcoro_clip_b200] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL collectives",
                  flush=True)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=pg)      # one-time agreement (set-up only, never per step)
    if ok.item() == 0:
        _FAILED.add((id(pg), dev.index))
        return None
    _PLANS[key] = plan
    return plan


class SymmReducePlan:
    """Symmetric buffers for an in-place sum-all-reduce of one fp32 tensor + a few fp64 scalars per step (the replicated-text
    SigLIP path: the [T, D] text gradient and the loss / bias / temperature sums), without any NCCL call in the step: every
    rank accumulates into its own copy, `b200clip_symm_allreduce_f32` reduces slice r on rank r and writes it back into every
    copy, `b200clip_symm_sum_f64` adds the scalars, both bracketed by `b200clip_symm_barrier`. Double-buffered by step parity
    like SymmPlan (the reduced tensor is a saved tensor of the autograd graph until its backward has run)."""

    NSCAL = 32          # floats reserved behind the tensor (16 fp64 scalars)

    def __init__(self, group, nfloat: int, dev: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        pg = group if group is not None else dist.group.WORLD
        self.W, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
        self.n = nfloat
        self.dev_index = dev.index
        self.per = (nfloat + self.NSCAL + 63) // 64 * 64
        self.buf = symm_mem.empty((2, self.per), dtype=torch.float32, device=dev)
        self.flags = symm_mem.empty((64,), dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.h_buf = symm_mem.rendezvous(self.buf, pg)
        self.h_flags = symm_mem.rendezvous(self.flags, pg)
        self.h_flags.barrier(channel=0)
        self.step = 0
        self.live = [None, None]

        def arr(ptrs):
            return (ctypes.c_void_p * len(ptrs))(*ptrs)
        self.buf_ptrs = [arr([int(p) + s * self.per * 4 for p in self.h_buf.buffer_ptrs]) for s in (0, 1)]
        self.scal_ptrs = [arr([int(p) + (s * self.per + nfloat) * 4 for p in self.h_buf.buffer_ptrs]) for s in (0, 1)]
        self.flag_ptrs = arr([int(p) for p in self.h_flags.buffer_ptrs])

    acquire = SymmPlan.acquire
    release = SymmPlan.release

    def barrier(self, channel: int) -> None:
        _call("symm_barrier", self.flag_ptrs, self.W, self.rank, channel, _stream_ptr(self.dev_index))

    def allreduce(self, slot: int, nscal: int, scal_out: torch.Tensor) -> None:
        """Tensor of `slot` summed in place on every rank, its first `nscal` fp64 scalars summed into `scal_out`."""
        st = _stream_ptr(self.dev_index)
        self.barrier(0)                                   # every rank's partial sums are complete
        _call("symm_sum_f64", self.scal_ptrs[slot], nscal, self.W, scal_out, st)
        _call("symm_allreduce_f32", self.buf_ptrs[slot], self.n, self.W, self.rank, st)
        self.barrier(1)                                   # every slice has been written back everywhere


_RPLANS: Dict[Tuple, "SymmReducePlan"] = {}


def get_reduce_plan(group, nfloat: int, dev: torch.device) -> Optional[SymmReducePlan]:
    """Cached SymmReducePlan, created collectively on first use; None when symmetric memory is unavailable / switched off."""
    if not enabled() or not dist.is_available() or not dist.is_initialized() or nfloat % 4:
        return None
    pg = group if group is not None else dist.group.WORLD
    if dist.get_backend(pg) != "nccl" or dist.get_world_size(pg) > 8:
        return None
    key = (id(pg), nfloat, dev.index)
    if key in _RPLANS:
        return _RPLANS[key]
    if (id(pg), dev.index) in _FAILED:
        return None
    ok = torch.ones(1, device=dev)
    plan = None
    try:
        plan = SymmReducePlan(pg, nfloat, dev)
    except Exception as e:       # noqa: BLE001
        ok.zero_()
        if dist.get_rank(pg) == 0:
            print(f"[deepcoro_clip_b200] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL collectives",
                  flush=True)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=pg)
    if ok.item() == 0:
        _FAILED.add((id(pg), dev.index))
        return None
    _RPLANS[key] = plan
    return plan
