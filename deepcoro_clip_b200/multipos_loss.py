"""Multi-positive softmax losses over materialised logits, same class names / arguments as the reference:

  * ``MultiPositiveInfoNCELoss``  utils/loss/multi_positive_infonce.py:8-100   (registry key "multi_positive_infonce")
  * ``WeightedSigLIPLoss``        utils/loss/weighted_siglip.py:6-51           (the runner's inline multi-positive branch,
                                   runners/video_constrative_learning_runner.py:121, 1275-1283, 1604-1612)

Their input is the [N, M] logits matrix the caller already computed, so these are HBM-bound kernels (csrc/multipos.cu):
one read of logits and weights per direction for the row / column statistics, one elementwise pass for the gradient —
instead of the reference's two log_softmax matrices, two transposes and the weighted products."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import call, i64, lib, stream_ptr


def _rowmajor_f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.stride(-1) == 1 else t.contiguous()


class _MultiPosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, pos_weights, pos_mask, mode, eps, reduce_sum):
        dev = ops.require_cuda(logits, pos_weights, pos_mask)
        L = _rowmajor_f32(logits)
        pw, mk = _rowmajor_f32(pos_weights), _rowmajor_f32(pos_mask)
        if pw is not None and mk is not None and pw.stride(0) != mk.stride(0):
            pw, mk = pw.contiguous(), mk.contiguous()
        N, M = L.shape
        ldw = (pw if pw is not None else mk).stride(0)
        st = stream_ptr(dev)
        stats = torch.empty((N + M, 4), dtype=torch.float32, device=dev)
        coef = torch.empty(N + M, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        ws = torch.empty(max(4, lib().b200clip_multipos_workspace_bytes(N, M)), dtype=torch.uint8, device=dev)
        call("multipos_fwd", L, i64(L.stride(0)), pw, mk, i64(ldw), N, M, int(mode), float(eps), int(reduce_sum),
             stats[:N], stats[N:], coef, loss, ws, st)
        ctx.save_for_backward(L, pw if pw is not None else torch.empty(0, device=dev),
                              mk if mk is not None else torch.empty(0, device=dev), stats, coef)
        ctx.has = (pw is not None, mk is not None, ldw, logits.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        L, pw, mk, stats, coef = ctx.saved_tensors
        has_pw, has_mk, ldw, dtype = ctx.has
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None, None
        N, M = L.shape
        g = grad_out.detach().reshape(1).float().contiguous()
        dL = torch.empty((N, M), dtype=torch.float32, device=L.device)
        call("multipos_bwd", L, i64(L.stride(0)), pw if has_pw else None, mk if has_mk else None, i64(ldw), N, M,
             stats[:N], stats[N:], coef, g, dL, i64(M), stream_ptr(L.device))
        return (dL if dtype == torch.float32 else dL.to(dtype)), None, None, None, None, None


class WeightedSigLIPLoss(nn.Module):
    """utils/loss/weighted_siglip.py:6-51 — 0.5 * (mean_i l_i + mean_j l_j) with
    l = -sum(pos * log_softmax(logits)) / clamp_min(sum(pos), eps), pos = clamp(positive_weights, min=0)."""

    def __init__(self, eps: float = 1e-6) -> None:
        super().__init__()
        self.eps = eps

    def forward(self, logits: torch.Tensor, positive_weights: torch.Tensor) -> torch.Tensor:
        if positive_weights.shape != logits.shape:
            raise ValueError(
                f"positive_weights shape {positive_weights.shape} must match logits shape {logits.shape}.")
        return _MultiPosFn.apply(logits, positive_weights, None, 0, self.eps, 0)


class MultiPositiveInfoNCELoss(nn.Module):
    """utils/loss/multi_positive_infonce.py:8-100 — symmetric multi-positive InfoNCE: rows / columns that have at least one
    positive contribute -sum(w * log_softmax) with w = clamp_min(pos_mask [* pos_weights], 0) / clamp_min(sum w, 1)."""

    def __init__(self, reduction: str = "mean", use_importance_weighting: bool = False):
        super().__init__()
        if reduction not in {"mean", "sum"}:
            raise ValueError(f"Unsupported reduction '{reduction}'. Expected 'mean' or 'sum'.")
        self.reduction = reduction
        self.use_importance_weighting = use_importance_weighting

    def forward(self, logits: torch.Tensor, pos_mask: torch.Tensor,
                pos_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        if logits.dim() != 2:
            raise ValueError(f"logits must be 2D, got shape {tuple(logits.shape)}")
        if pos_mask.shape != logits.shape:
            raise ValueError("pos_mask must match logits shape.")
        if pos_weights is not None and pos_weights.shape != logits.shape:
            raise ValueError("pos_weights must match logits shape.")
        # mode 2: rows / columns weighted by their summed raw importance (pos_weights, else pos_mask), :57-93
        return _MultiPosFn.apply(logits, pos_weights, pos_mask, 2 if self.use_importance_weighting else 1, 0.0,
                                 int(self.reduction == "sum"))
