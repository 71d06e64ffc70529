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


# --------------------------------------------------------------------------------------------------------------------
# the runner's inline multi-positive branch, from features (no [B, M] matrix is materialised)
# --------------------------------------------------------------------------------------------------------------------
class _InlineMultiPosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, video, text, log_temp, targets, pos_weights, abnormal, margin, mode, eps, neg_weight):
        dev = ops.require_cuda(video, text, log_temp, targets)
        v = _rowmajor_f32(video)
        t = _rowmajor_f32(text)
        tg = _rowmajor_f32(targets)
        pw = _rowmajor_f32(pos_weights)
        if pw is not None and pw.stride(0) != tg.stride(0):
            pw, tg = pw.contiguous(), tg.contiguous()
        ab = None if abnormal is None else abnormal.detach().to(device=dev, dtype=torch.float32).contiguous()
        lt = log_temp.detach().reshape(-1)[:1].float().contiguous()
        B, D = v.shape
        M = t.shape[0]
        if t.shape[1] != D or tuple(tg.shape) != (B, M) or (pw is not None and tuple(pw.shape) != (B, M)):
            raise ValueError(f"inline multi-positive loss: video {tuple(v.shape)}, text {tuple(t.shape)}, positive_mask "
                             f"{tuple(tg.shape)} do not fit together")
        if D > 1024:
            raise ValueError("inline multi-positive loss: embedding width above 1024 is not supported")
        stats = torch.empty((B + M, 8), dtype=torch.float32, device=dev)
        scal = torch.empty(10, dtype=torch.float32, device=dev)          # [0:8] scalars, [8:10] one int flag + padding
        flag = scal[8:9].view(torch.int32)
        call("inline_mp_fwd", v, i64(v.stride(0)), t, i64(t.stride(0)), tg, pw, i64(tg.stride(0)), ab, float(margin), lt, B, M,
             D, int(mode), float(eps), float(neg_weight), stats[:B], stats[B:], scal, flag, stream_ptr(dev))
        ctx.save_for_backward(v, t, tg, pw, ab, lt, stats, scal)
        ctx.cfg = (B, M, D, int(mode), float(eps), float(neg_weight), float(margin), video.dtype, text.dtype, log_temp.shape,
                   log_temp.dtype)
        ctx.mark_non_differentiable(scal)
        return scal[0:1].clone().reshape(()), scal

    @staticmethod
    def backward(ctx, grad_loss, _grad_scal):
        v, t, tg, pw, ab, lt, stats, scal = ctx.saved_tensors
        B, M, D, mode, eps, neg_weight, margin, vdt, tdt, lt_shape, lt_dtype = ctx.cfg
        dev = v.device
        g = grad_loss.detach().reshape(1).float().contiguous()
        dv = torch.empty((B, D), dtype=torch.float32, device=dev)
        dt = torch.empty((M, D), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        acc = torch.empty(1, dtype=torch.float64, device=dev) if ctx.needs_input_grad[2] else None
        call("inline_mp_bwd", v, i64(v.stride(0)), t, i64(t.stride(0)), tg, pw, i64(tg.stride(0)), ab, margin, lt, B, M, D,
             mode, eps, neg_weight, stats[:B], stats[B:], scal, scal[8:9].view(torch.int32), g, dv, dt, acc, stream_ptr(dev))
        dlt = acc.to(lt_dtype).reshape(lt_shape) if acc is not None else None
        return (dv.to(vdt) if ctx.needs_input_grad[0] else None, dt.to(tdt) if dt is not None else None, dlt,
                None, None, None, None, None, None, None)


def inline_multipositive_loss(video_emb: torch.Tensor, text_emb: torch.Tensor, log_temp: torch.Tensor,
                              positive_mask: torch.Tensor, positive_weights: Optional[torch.Tensor] = None, *,
                              abnormal_vector: Optional[torch.Tensor] = None, abnormal_margin: float = 0.0,
                              use_weighted_siglip: bool = True, eps: float = 1e-6, negative_weight: float = 1.0):
    """The runner's inline branch for batches that carry ``positive_mask`` (reference
    runners/video_constrative_learning_runner.py:1256-1322; validation :1585-1641), computed from the FEATURES:

        similarity = normalize(video) @ normalize(text).T;  logits = similarity * sigmoid(similarity) / exp(log_temp)
        logits += abnormal_vector[None, :] * abnormal_margin                      (siglip_abnormal_margin > 0, :1269-1278)
        use_weighted_siglip: WeightedSigLIPLoss(logits, targets [* positive_weights])          (:1280-1288)
        else:                BCEWithLogits(logits, targets, weight, 'sum') / max(1, targets.sum())   (:1289-1303)

    Returns ``(loss, diagnostics)`` where ``diagnostics`` holds the three 0-d tensors the runner logs from the same
    logits (:1305-1322): ``alignment_logprob``, ``alignment_prob``, ``alignment_cosine`` (NaN where the reference leaves
    them ``None``: no row with a positive / no positive pair). The loss is differentiable w.r.t. ``video_emb``, ``text_emb``
    and ``log_temp``; nothing of size [B, M] is allocated (csrc/inline_mp.cu). INTEGRATION.md §2c shows the runner edit."""
    if abnormal_vector is not None and abnormal_margin == 0.0:
        abnormal_vector = None
    loss, scal = _InlineMultiPosFn.apply(video_emb, text_emb, log_temp, positive_mask, positive_weights, abnormal_vector,
                                         abnormal_margin, 0 if use_weighted_siglip else 1, eps, negative_weight)
    return loss, {"alignment_logprob": scal[2], "alignment_prob": scal[3], "alignment_cosine": scal[4]}
