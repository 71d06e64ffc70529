"""Drop-in for utils/retrieval_metrics_streaming.py (reference :10-101, :104-197) on sm_100a.

Same function names, positional/keyword arguments and result keys. One pass of tcgen05 similarity tiles with a
rank-count epilogue replaces the chunked matmul + topk + cat + topk + gather loops and the per-row argsort loop;
recall@k and MRR come from exact integer rank counts (rank_i = 1 + #{columns ranked before the ground truth}).

Tie rule: lowest index first (BASELINE.json north_star) — the reference inherits torch.topk's unspecified order;
on tie-free inputs the results are identical.

Precision: similarities are fp32-accumulated products of bf16 operands. ``precision="auto"`` uses plain bf16 when
every input value is exactly representable in bf16 (e.g. the exact-grid evaluation embeddings: results are then
bit-exact in any summation order) and the error-compensated bf16x3 operands (≈fp32) otherwise.

Multi-GPU: the reference calls these functions on RANK 0 ONLY (``if self.config.is_ref_device:`` in
runners/multitask_runner.py:642 -> :1276), so the reference-signature entry points default to ``use_ddp=False``: one
process sweeps the whole text set and no collective is issued (a collective here would hang rank 0 forever). Callers
that invoke them on EVERY rank of a process group (``embedding_store.epoch_end_retrieval_metrics``, ``bench.py``) pass
``use_ddp=True``: the text database is then sharded by rows across ranks (every rank passes the same full
``text_features``), rank counts are all-reduced and the per-shard top-k lists merged.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import dist_plan, ops
from ._lib import call, stream_ptr


def _exact_in_bf16(*xs: torch.Tensor) -> bool:
    """True when every element of every tensor is exactly representable in bf16 (one streaming kernel pass per tensor and
    ONE host read for all of them; bf16 inputs are exact by construction)."""
    from ._lib import DTYPE_CODE, i64
    flag = None
    for x in xs:
        if x.dtype == torch.bfloat16:
            continue
        if x.dim() != 2 or x.dtype not in DTYPE_CODE or x.stride(1) != 1:
            x = x.float().reshape(x.shape[0], -1).contiguous() if x.dim() >= 1 else x.float().reshape(1, 1)
        if flag is None:
            flag = torch.zeros(1, dtype=torch.int32, device=x.device)
        call("inexact_bf16", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), x.shape[0], x.shape[1], flag, stream_ptr(x.device))
    return True if flag is None else flag.item() == 0


def _operands(video, text, normalize: bool, precision: str):
    if precision == "auto":
        x3 = normalize or not _exact_in_bf16(video, text)
    elif precision in ("bf16", "bf16x3"):
        x3 = precision == "bf16x3"
    else:
        raise ValueError(f"precision must be 'auto', 'bf16' or 'bf16x3', got {precision!r}")
    vop, vn, Kp = ops.l2norm_operand(video, 0 if x3 else -1, normalize)
    top, tn, _ = ops.l2norm_operand(text, 1 if x3 else -1, normalize)
    return vop, top, vn, tn, Kp


def _shard_rows(M: int, use_ddp: bool, group=None) -> Tuple[int, int, int, int]:
    W, r = dist_plan.world(use_ddp, group)
    lo, hi = dist_plan.text_shard(M, W, r)
    return W, r, lo, hi


def _topk2_enabled() -> bool:
    """Two-sweep top-k (threshold, then collect) instead of the register lists. Opt-in: written after the GPU budget of
    round 1 was spent, so it has not been measured yet (tests/test_gpu_z_topk_two_sweeps.py)."""
    return os.environ.get("B200CLIP_TOPK2", "0") == "1"


def _topk_two_sweeps(vop, top, lo: int, hi: int, k: int):
    """Exact top-k of the text rows [lo, hi) for every video row, or (None, None) when a candidate buffer overflowed (the
    caller then runs the register-list sweep). Sweep 1: maxima of 64 * segs disjoint column subsets per row -> their k-th
    largest bounds the row's k-th best score from below; sweep 2: every s_ij >= that bound becomes a candidate (all ties
    with the k-th score included); ``topk_merge`` orders them by (score desc, index asc)."""
    dev = vop.device
    N, K, Ms = vop.shape[0], vop.shape[1], hi - lo
    st = stream_ptr(dev)
    segs = ops._lib.lib().b200clip_retrieval_segments(N, Ms)
    # -inf: a (row, slot) the sweep never visits (empty segment) must not contribute a maximum
    pm = torch.full((N, 2 * segs, 32), float("-inf"), dtype=torch.float32, device=dev)
    call("retrieval_colmax", vop, top[lo:hi], N, Ms, K, vop.stride(0), top.stride(0), segs, pm, st)
    thr = torch.empty(N, dtype=torch.float32, device=dev)
    call("kth_largest", pm, N, 2 * segs * 32, k, thr, st)
    cap = max(64, 8 * k)
    cnt = torch.zeros(N + 1, dtype=torch.int32, device=dev)          # [N] candidate counts | overflow flag
    bs = torch.empty((N, cap), dtype=torch.float32, device=dev)
    bi = torch.full((N, cap), 0x7FFFFFFF, dtype=torch.int32, device=dev)
    call("retrieval_collect", vop, top[lo:hi], N, Ms, K, vop.stride(0), top.stride(0), thr, lo, segs, cnt, bs, bi, cap,
         cnt[N:], st)
    if int(cnt[N].item()) != 0:
        return None, None
    out_s = torch.empty((N, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((N, k), dtype=torch.int64, device=dev)
    call("topk_merge", bs, bi, N, cap, k, out_s, out_i, st)
    return out_s, out_i


def _sweep(vop, top, gt, k: int, use_ddp: bool, group=None, _shard=None):
    """Returns (counts [N] int32 or None, top-k scores [N,k] / indices [N,k] int64 or None).
    ``_shard=(lo, hi)`` restricts the sweep to text rows [lo, hi) of a single process (tests: shard additivity)."""
    dev = vop.device
    N, M, K = vop.shape[0], top.shape[0], vop.shape[1]
    W, rank, lo, hi = _shard_rows(M, use_ddp, group)
    if _shard is not None:
        W, rank, (lo, hi) = 1, 0, _shard
    st = stream_ptr(dev)
    counts = sgt = gt64 = None
    if gt is not None:
        gt64 = gt.to(device=dev, dtype=torch.int64).contiguous()
        # ground-truth similarity with the tensor core's own rounding (ties with duplicate texts stay exact ties)
        if W > 1 and N >= 1024 * W:
            # every rank holds all video rows and the whole text database: the ground-truth dots are sharded by VIDEO rows
            # (each rank gathers / multiplies N / W rows) and all-gathered, instead of W identical copies of the work
            ch = (N + W - 1) // W
            a, b = min(rank * ch, N), min((rank + 1) * ch, N)
            sg_all = torch.zeros(W * ch, dtype=torch.float32, device=dev)
            if b > a:
                tg = torch.empty((b - a, K), dtype=torch.bfloat16, device=dev)
                call("gather_rows_bf16", top, top.stride(0), gt64[a:b], b - a, M, K, tg, K, st)
                call("rowdot_tc", vop[a:b], vop.stride(0), tg, K, b - a, K, sg_all[rank * ch:rank * ch + (b - a)], st)
            dist.all_gather_into_tensor(sg_all, sg_all[rank * ch:(rank + 1) * ch].clone(), group=group)
            sgt = sg_all[:N]
        else:
            sgt = torch.empty(N, dtype=torch.float32, device=dev)
            tg = torch.empty((N, K), dtype=torch.bfloat16, device=dev)
            call("gather_rows_bf16", top, top.stride(0), gt64, N, M, K, tg, K, st)
            call("rowdot_tc", vop, vop.stride(0), tg, K, N, K, sgt, st)
        counts = torch.zeros(N, dtype=torch.int32, device=dev)
    k = min(k, M)
    Ms = hi - lo
    out_s = out_i = None
    if Ms > 0 and gt is None and 0 < k <= 16 and _topk2_enabled():
        out_s, out_i = _topk_two_sweeps(vop, top, lo, hi, k)
    if out_s is not None:
        pass                                     # exact lists from the two-sweep path
    elif Ms > 0:
        segs = ops._lib.lib().b200clip_retrieval_segments(N, Ms)
        ps = pi = None
        if k > 0:
            ps = torch.empty((N, 2 * segs, k), dtype=torch.float32, device=dev)
            pi = torch.empty((N, 2 * segs, k), dtype=torch.int32, device=dev)
        call("retrieval_sweep", vop, top[lo:hi], N, Ms, K, vop.stride(0), top.stride(0), sgt, gt64, lo, counts, k, segs,
             ps, pi, st)
        if k > 0:
            out_s = torch.empty((N, k), dtype=torch.float32, device=dev)
            out_i = torch.empty((N, k), dtype=torch.int64, device=dev)
            call("topk_merge", ps, pi, N, 2 * segs * k, k, out_s, out_i, st)
    elif k > 0:
        out_s = torch.full((N, k), float("-inf"), dtype=torch.float32, device=dev)
        out_i = torch.full((N, k), -1, dtype=torch.int64, device=dev)
    if W > 1:
        if counts is not None:
            dist.all_reduce(counts, group=group)
        if k > 0:
            # concatenated [W * N, k] outputs: the form every backend accepts (gloo rejects the stacked [W, N, k] shape)
            gs = torch.empty((W * N, k), dtype=torch.float32, device=dev)
            gi = torch.empty((W * N, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gs, out_s, group=group)
            dist.all_gather_into_tensor(gi, out_i, group=group)
            cs = gs.view(W, N, k).permute(1, 0, 2).reshape(N, W * k).contiguous()
            ci = gi.view(W, N, k).permute(1, 0, 2).reshape(N, W * k)
            ci = torch.where(ci < 0, torch.full_like(ci, 0x7FFFFFFF), ci).to(torch.int32).contiguous()
            call("topk_merge", cs, ci, N, W * k, k, out_s, out_i, st)
    return counts, out_s, out_i


@torch.no_grad()
def streaming_topk(video_features, text_features, k: int, *, normalize: bool = False, precision: str = "auto",
                   use_ddp: bool = False, group=None):
    """Top-k text indices per video by (similarity desc, index asc): (scores [N,k] fp32, indices [N,k] int64)."""
    ops.require_cuda(video_features, text_features)
    if k < 1 or k > 64:
        raise ValueError("k must be in [1, 64]")
    vop, top, _, _, _ = _operands(video_features, text_features, normalize, precision)
    _, s, i = _sweep(vop, top, None, k, use_ddp, group)
    return s, i


@torch.no_grad()
def inference_topk_indices(video_embeddings: torch.Tensor, text_embeddings: torch.Tensor, topk: int, **kw) -> torch.Tensor:
    """``inference()``'s retrieval step (runners/video_constrative_learning_runner.py:1757-1758):
    ``torch.topk(video_embeddings @ text_embeddings.t(), k=topk, dim=1)[1]`` — [N, topk] int64 indices, without the
    [N, M] similarity matrix. Ties go to the lowest index (torch.topk leaves them unspecified)."""
    return streaming_topk(video_embeddings, text_embeddings, min(int(topk), text_embeddings.shape[0]), **kw)[1]


@torch.no_grad()
def top5_predictions(video_embeddings: torch.Tensor, text_embeddings: torch.Tensor, k: int = 5, **kw):
    """The per-row ``torch.topk(similarity_matrix[i], k=min(5, M))`` of ``save_retrieval_results``
    (utils/wandb_logger.py:951-957) for ALL rows at once from the embeddings: (scores [N, k] fp32, indices [N, k] int64);
    one sweep instead of N host-driven top-k calls with two ``.item()`` per entry."""
    return streaming_topk(video_embeddings, text_embeddings, min(int(k), text_embeddings.shape[0]), **kw)


@torch.no_grad()
def compute_recall_at_k_streaming(video_features: torch.Tensor, text_features: torch.Tensor,
                                  ground_truth_indices: torch.Tensor, k_values: List[int] = [1, 5, 10, 50],
                                  video_chunk_size: int = 2048, text_chunk_size: int = 8192, device: str = "cuda",
                                  *, precision: str = "auto", use_ddp: bool = False, group=None,
                                  _counts_out: Optional[list] = None) -> Dict[str, float]:
    """Reference :10-101 — Recall@k in PERCENT; inputs are used as given (not normalised)."""
    ops.require_cuda(video_features, text_features)
    vop, top, _, _, _ = _operands(video_features, text_features, False, precision)
    return _recall_from_operands(vop, top, ground_truth_indices, k_values, use_ddp, group, _counts_out)


def mrr_sum_from_counts(counts: torch.Tensor, n_texts: int) -> torch.Tensor:
    """sum_i 1 / (counts[i] + 1) as a 0-d float64 device tensor (deterministic histogram-order fp64 sum; no host pass over
    the rows — the host loop of the reference costs 2.6 ms for 203,808 rows even vectorised with numpy)."""
    dev = counts.device
    hist = torch.zeros(max(1, int(n_texts)), dtype=torch.int32, device=dev)
    out = torch.empty(1, dtype=torch.float64, device=dev)
    call("mrr_from_counts", counts, counts.numel(), hist.numel(), hist, out, stream_ptr(dev))
    return out[0]


def _recall_from_operands(vop, top, gt, k_values, use_ddp, group, counts_out=None) -> Dict[str, float]:
    dev = vop.device
    N, M = vop.shape[0], top.shape[0]
    counts, _, _ = _sweep(vop, top, gt, 0, use_ddp, group)
    kv = torch.tensor(list(k_values), dtype=torch.int32, device=dev)
    hits = torch.zeros(len(k_values), dtype=torch.int64, device=dev)
    call("recall_hits", counts, N, kv, len(k_values), hits, stream_ptr(dev))
    hits_h = hits.cpu().tolist()
    if counts_out is not None:
        counts_out.append(counts)
    width = min(max(k_values), M)        # the reference's best_indices has min(k_max, M) columns (:64, :80, :90)
    return {f"Recall@{k}": ((h / N) * 100 if k <= width else 0.0) for k, h in zip(k_values, hits_h)}


@torch.no_grad()
def compute_metrics_streaming(video_features: torch.Tensor, text_features: torch.Tensor,
                              ground_truth_indices: torch.Tensor, k_values: List[int] = [1, 5, 10, 50],
                              video_chunk_size: int = 2048, text_chunk_size: int = 8192, device: str = "cuda",
                              *, precision: str = "auto", use_ddp: bool = False, group=None) -> Dict[str, float]:
    """Reference :104-197 — normalises both sides, then Recall@k, MRR_V2T, alignment_score, norms, median_rank."""
    ops.require_cuda(video_features, text_features)
    dev = video_features.device
    vop, top, vinv, tinv, Kp = _operands(video_features, text_features, True, precision)
    N, M, K = vop.shape[0], top.shape[0], vop.shape[1]
    keep: list = []
    metrics = _recall_from_operands(vop, top, ground_truth_indices, k_values, use_ddp, group, keep)
    counts = keep[0]
    # MRR = mean_i 1 / rank_i (reference :162-172), summed on the device over the rank histogram in fp64
    metrics["MRR_V2T"] = float(mrr_sum_from_counts(counts, M).item() / N) if N else 0.0
    # alignment = mean_i vhat_i . that_gt(i)  (:175-190)
    gt64 = ground_truth_indices.to(device=dev, dtype=torch.int64).contiguous()
    sg = torch.empty(N, dtype=torch.float32, device=dev)
    call("rowdot_bf16", vop, vop.stride(0), top, top.stride(0), gt64, N, M, K, sg, stream_ptr(dev))
    metrics["alignment_score"] = float(sg.double().sum().item() / N)
    # norms of the NORMALISED features (:193-194): 1 for every non-degenerate row, ||x||/1e-12 below the eps clamp
    metrics["video_norm"] = float(_normalised_norm_mean(video_features, vinv))
    metrics["text_norm"] = float(_normalised_norm_mean(text_features, tinv))
    metrics["median_rank"] = 1           # placeholder in the reference (:195)
    return metrics


def _normalised_norm_mean(x: torch.Tensor, inv: torch.Tensor) -> float:
    # ||x * inv|| = ||x|| * inv ; inv = 1/max(||x||, 1e-12)  =>  1 unless ||x|| < 1e-12
    degenerate = inv >= 1e12
    if not bool(degenerate.any().item()):
        return 1.0
    nrm = torch.linalg.vector_norm(x.float(), dim=1) * inv
    return nrm.mean().item()
