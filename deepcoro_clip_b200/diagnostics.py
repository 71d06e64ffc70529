"""Per-step alignment diagnostics without the dense [B, B] pass (the logging half).

After every training step the reference runner rebuilds the batch's similarity matrix and its row log-softmax only to
log three scalars (runners/video_constrative_learning_runner.py:1323-1335):

    similarity            = normalize(video) @ normalize(text).T
    alignment_cosine      = diag(similarity).mean()
    logits                = (similarity | similarity * sigmoid(similarity)) / exp(log_temp)     # gated for "siglip" losses
    alignment_logprob     = diag(log_softmax(logits, dim=1)).mean()
    alignment_prob        = alignment_logprob.exp()

Here the same three numbers come from ONE sweep of the fused logits forward (row sums + the tensor core's own diagonal,
``b200clip_logits_lse_fwd``) and a one-CTA scalar kernel (``b200clip_alignment_diag``); the [B, B] matrix never exists.
Like the runner's block, this is computed on the LOCAL batch of a rank (no gather) and carries no gradient.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .loss import _pick_precision


@torch.no_grad()
def alignment_diagnostics(video_features: torch.Tensor, text_features: torch.Tensor, log_temp, *,
                          use_siglip: bool = False, precision: str = "auto") -> Dict[str, torch.Tensor]:
    """{"alignment_cosine", "alignment_logprob", "alignment_prob"} as 0-d fp32 device tensors (no host sync).

    ``use_siglip`` selects the gated logits the runner uses when ``"siglip" in config.loss_name`` (:1328-1331); the
    temperature is ``exp(log_temp)`` without a clamp (:1332)."""
    dev = ops.require_cuda(video_features, text_features)
    if video_features.dim() != 2 or video_features.shape != text_features.shape:
        raise ValueError(f"video_features {tuple(video_features.shape)} and text_features "
                         f"{tuple(text_features.shape)} must both be [B, D]")
    if not isinstance(log_temp, torch.Tensor):
        log_temp = torch.tensor(float(log_temp), device=dev)
    elif log_temp.device != dev:
        log_temp = log_temp.to(dev)
    B = video_features.shape[0]
    x3 = _pick_precision(precision, B, B)
    gated = int(bool(use_siglip))
    top, _, _ = ops.l2norm_operand(text_features.detach(), 1 if x3 else -1)
    vop, _, _ = ops.l2norm_operand(video_features.detach(), 0 if x3 else -1)
    K = vop.shape[1]
    dyn = ops.dyn_prep(log_temp, None, 0.0, ops.GATED_BOUND if gated else 1.0)
    # [colsum (B) | rowsum (B) | S_ii (B)] zeroed (the forward accumulates into the sums) | out (3, padded to 4) | row
    # tickets of the stable sweep (B int32, zero)
    ws = torch.zeros(4 * B + 4, dtype=torch.float32, device=dev)
    st = ops.stream_ptr(dev)
    # both variants are enqueued, dyn[11] (set from tau on the device) lets exactly one run: the fixed-shift sweep, or the
    # running-maximum row sweep that leaves the log2-domain row log-sum-exp in the rowsum slot (tau below ~0.013; the
    # runner applies no temperature floor here, :1332)
    ops.call("logits_lse_fwd", vop, top, B, B, K, vop.stride(0), top.stride(0), 0.0, 0.0, gated, dyn, 1, ws[B:2 * B],
             ws[:B], ws[2 * B:3 * B], 0, st)
    slots = ops._lib.lib().b200clip_rowlse_slots(B, B, K)
    part = torch.empty(2 * B * slots, dtype=torch.float32, device=dev)
    ops.call("logits_rowlse", vop, top, B, B, K, vop.stride(0), top.stride(0), gated, dyn, 1, part, slots,
             ws[3 * B + 4:].view(torch.int32), ws[B:2 * B], ws[2 * B:3 * B], 0, None, st)
    out = ws[3 * B:3 * B + 3]
    ops.call("alignment_diag", ws, B, dyn, gated, out, st)
    return {"alignment_cosine": out[0], "alignment_logprob": out[1], "alignment_prob": out[2]}
