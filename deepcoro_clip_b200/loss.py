"""Drop-in contrastive losses backed by the sm_100a kernels (no PyTorch math on the hot path).

Class names, constructor arguments, forward keyword arguments, registry keys and error behaviour mirror the
reference:
  * ``CLIPLoss``                       utils/loss/contrastive.py:108-164
  * ``SigLIPLoss``                     utils/loss/contrastive.py:171-319
  * ``ContrastiveLoss`` / ``...DDP``   utils/loss/losses.py:12-64, 104-158
  * ``SiglipLoss`` / ``SiglipLossDDP`` utils/loss/losses.py:160-276   (legacy *gated* softmax loss)
  * ``InfoNCELoss``                    utils/loss/losses.py:278-331
``install()`` registers them into the reference's ``LossRegistry`` under the same ``LossType`` keys.

Distributed semantics (pinned against the reference on 2-rank gloo): every rank returns the
FULL global loss; the gradient of the local rows equals the corresponding rows of the full-batch gradient;
``log_temp.grad`` is the identical full value on every rank. Unlike the reference, each rank only computes its
own row slab of the logits (1/W of the work): embeddings are all-gathered as bf16 operands, the per-column
sums are all-reduced and the per-row sums all-gathered (three [N]-sized collectives).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from ._lib import B200ClipError

BW_CLIP, BW_GATED, BW_SIGLIP = 0, 1, 2


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def _world(use_ddp: bool = True, group=None):
    if use_ddp and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather_rows(x: torch.Tensor, world: int, group=None) -> torch.Tensor:
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _pick_precision(precision: str, n_rows: int, n_cols: int) -> bool:
    """True -> bf16x3 error-compensated operands (≈fp32 logits), False -> plain bf16 operands."""
    if precision == "bf16x3":
        return True
    if precision == "bf16":
        return False
    if precision != "auto":
        raise ValueError(f"precision must be 'auto', 'bf16' or 'bf16x3', got {precision!r}")
    # small problems are latency bound: spend 3x MMA work to meet the 1e-5 loss tolerance of the fp32 reference
    return n_rows * n_cols <= 4096 * 4096


class _ClipLossFn(torch.autograd.Function):
    """Softmax-CE both directions, diagonal targets."""

    @staticmethod
    def forward(ctx, video, text, log_temp, label_smoothing, gated, clamp_min, precision, use_ddp, group):
        dev = ops.require_cuda(video, text, log_temp)
        if video.dim() != 2 or text.dim() != 2 or video.shape != text.shape:
            raise ValueError(
                f"video_features {tuple(video.shape)} and text_features {tuple(text.shape)} must both be [B, D]")
        B, D = video.shape
        W, rank = _world(use_ddp, group)
        N = B * W
        x3 = _pick_precision(precision, N, N)
        eps = float(label_smoothing)
        mode = BW_GATED if gated else BW_CLIP

        vop, vinv, Kp = ops.l2norm_operand(video, 0 if x3 else -1)
        top, tinv, _ = ops.l2norm_operand(text, 1 if x3 else -1)
        K = vop.shape[1]
        vall = _all_gather_rows(vop, W, group)
        tall = _all_gather_rows(top, W, group)
        dyn = ops.dyn_prep(log_temp, None, clamp_min, ops.GATED_BOUND if gated else 1.0)

        rowsum = torch.zeros(B, dtype=torch.float32, device=dev)
        colsum = torch.zeros(N, dtype=torch.float32, device=dev)
        dots = torch.empty(B, dtype=torch.float32, device=dev)      # S_ii as the tensor core rounded it
        ops.lse_fwd(vop, tall, B, N, K, dyn, gated, rowsum, colsum, dots, rank * B)
        if W > 1:
            dist.all_reduce(colsum, group=group)
            rowsum_all = _all_gather_rows(rowsum, W, group)
        else:
            rowsum_all = rowsum

        c = 0.5 / N
        acc = torch.zeros(4, dtype=torch.float64, device=dev)      # [sum r_i, sum c_j, sum f(S_ii), -]
        rowscale_all = torch.empty(N, dtype=torch.float32, device=dev)
        colscale_all = torch.empty(N, dtype=torch.float32, device=dev)
        ops.lse_finalize(rowsum_all, dyn, c, rowscale_all, acc[0:1])
        ops.lse_finalize(colsum, dyn, c, colscale_all, acc[1:2])
        ops.vec_fsum(dots, gated, acc[2:3])
        if W > 1:
            dist.all_reduce(acc[2:3], group=group)
        inv_tau = dyn[2].double()
        sum_tgt = (1.0 - eps) * acc[2] * inv_tau
        unif_tgt = torch.zeros((), dtype=torch.float64, device=dev)
        vsum = tsum = None
        if eps != 0.0:
            if gated:
                raise B200ClipError("label_smoothing is not defined for the gated legacy loss")
            vsum = ops.colsum_bf16(vall[:, K - Kp:], N, D)      # hi panel (last in bf16x3 mode)
            tsum = ops.colsum_bf16(tall[:, K - Kp:], N, D)
            unif_tgt = (eps / N) * torch.dot(vsum.double(), tsum.double()) * inv_tau
            sum_tgt = sum_tgt + unif_tgt
        loss = (0.5 / N) * (acc[0] + acc[1]) - sum_tgt / N

        ctx.save_for_backward(video, text, vop, top, vall, tall, vinv, tinv, dyn, rowscale_all, colscale_all, dots,
                              vsum, tsum, unif_tgt)
        ctx.cfg = (B, D, Kp, K, W, rank, N, eps, mode, group, log_temp.shape, log_temp.dtype)
        return loss.float()

    @staticmethod
    def backward(ctx, grad_out):
        (video, text, vop, top, vall, tall, vinv, tinv, dyn, rowscale_all, colscale_all, dots, vsum, tsum,
         unif_tgt) = ctx.saved_tensors
        B, D, Kp, K, W, rank, N, eps, mode, group, lt_shape, lt_dtype = ctx.cfg
        dev = video.device
        gmul = grad_out.detach().reshape(1).float().contiguous()
        lo, hi = rank * B, (rank + 1) * B
        dV = dT = dLT = None
        scal = torch.zeros(4, dtype=torch.float32, device=dev)
        need_lt = ctx.needs_input_grad[2]
        if ctx.needs_input_grad[0] or need_lt:
            dVh = torch.zeros((B, D), dtype=torch.float32, device=dev)
            dcv = torch.zeros((B, 2), dtype=torch.float32, device=dev)
            ops.logits_bwd(mode, vop, tall, B, N, K, Kp, D, dyn, rowscale_all[lo:hi], colscale_all, dVh, scal,
                           ydiag=(1.0 - eps) / N, diag_off=lo, diag_corr=dcv)
            if ctx.needs_input_grad[0]:
                dV = ops.l2norm_backward(dVh, video, vinv, other_x=text, other_inv=tinv, other_hi=top[:, K - Kp:],
                                         diag_corr=dcv, usum=tsum, ucoef=-eps / (N * N), dev_omul=dyn[2:3],
                                         dev_gmul=gmul).to(video.dtype)
        if ctx.needs_input_grad[1]:
            dTh = torch.zeros((B, D), dtype=torch.float32, device=dev)
            dct = torch.zeros((B, 2), dtype=torch.float32, device=dev)
            ops.logits_bwd(mode, top, vall, B, N, K, Kp, D, dyn, colscale_all[lo:hi], rowscale_all, dTh, None,
                           ydiag=(1.0 - eps) / N, diag_off=lo, diag_corr=dct)
            dT = ops.l2norm_backward(dTh, text, tinv, other_x=video, other_inv=vinv, other_hi=vop[:, K - Kp:],
                                     diag_corr=dct, usum=vsum, ucoef=-eps / (N * N), dev_omul=dyn[2:3],
                                     dev_gmul=gmul).to(text.dtype)
        if need_lt:
            s0 = scal[0:1].double()
            if W > 1:
                dist.all_reduce(s0, group=group)
            # d loss / d log_temp = -sum_ij G_ij L_ij  (zero while the tau clamp is active)
            # (the kernel's sum already contains the diagonal target; only the uniform label-smoothing part is added)
            dlt = (unif_tgt / N - s0 * dyn[2].double()) * dyn[7].double() * gmul.double()
            dLT = dlt.to(lt_dtype).reshape(lt_shape)
        return dV, dT, dLT, None, None, None, None, None, None


def clip_loss(video_features, text_features, log_temp, *, label_smoothing: float = 0.0, gated: bool = False,
              clamp_min: float = 1e-4, precision: str = "auto", use_ddp: bool = True, group=None) -> torch.Tensor:
    """Functional form of the fused softmax contrastive loss (forward + custom backward)."""
    if not isinstance(log_temp, torch.Tensor):
        log_temp = torch.tensor(float(log_temp), device=video_features.device)
    if log_temp.device != video_features.device:
        log_temp = log_temp.to(video_features.device)
    return _ClipLossFn.apply(video_features, text_features, log_temp, label_smoothing, gated, clamp_min, precision,
                             use_ddp, group)


# --------------------------------------------------------------------------------------------------
# nn.Module front-ends (reference signatures)
# --------------------------------------------------------------------------------------------------
class CLIPLoss(nn.Module):
    """utils/loss/contrastive.py:108-164 — 0.5*(CE(v->t) + CE(t->v)), tau = exp(log_temp).clamp(min=1e-4)."""

    def __init__(self, label_smoothing: float = 0.0, precision: str = "auto"):
        super().__init__()
        self.label_smoothing = label_smoothing
        self.precision = precision

    def forward(self, video_features: torch.Tensor, text_features: torch.Tensor, log_temp: torch.Tensor):
        return clip_loss(video_features, text_features, log_temp, label_smoothing=self.label_smoothing,
                         clamp_min=1e-4, precision=self.precision, use_ddp=True)


class ContrastiveLoss(nn.Module):
    """utils/loss/losses.py:12-64 — single-process CLIP loss without the tau clamp (never gathers)."""

    def __init__(self, precision: str = "auto"):
        super().__init__()
        self.precision = precision

    def forward(self, video_features, text_features, log_temp=None):
        if log_temp is None:
            log_temp = torch.log(torch.tensor(0.1, device=video_features.device))
        return clip_loss(video_features, text_features, log_temp, clamp_min=0.0, precision=self.precision,
                         use_ddp=False)


class ContrastiveLossDDP(nn.Module):
    """utils/loss/losses.py:104-158 — gathers across ranks when a process group exists; no tau clamp."""

    def __init__(self, precision: str = "auto"):
        super().__init__()
        self.precision = precision

    def forward(self, video_features, text_features, log_temp=None):
        if log_temp is None:
            log_temp = torch.log(torch.tensor(0.1, device=video_features.device))
        return clip_loss(video_features, text_features, log_temp, clamp_min=0.0, precision=self.precision,
                         use_ddp=True)


class SiglipLoss(nn.Module):
    """utils/loss/losses.py:160-211 — legacy *gated* loss: logits = S*sigmoid(S)/tau, symmetric softmax CE."""

    def __init__(self, precision: str = "auto"):
        super().__init__()
        self.precision = precision

    def forward(self, video_features, text_features, log_temp=None):
        if log_temp is None:
            log_temp = torch.log(torch.tensor(0.1, device=video_features.device))
        return clip_loss(video_features, text_features, log_temp, gated=True, clamp_min=0.0,
                         precision=self.precision, use_ddp=False)


class SiglipLossDDP(nn.Module):
    """utils/loss/losses.py:213-276 — gated loss, features rounded to fp16 before the gather (:243-244)."""

    def __init__(self, precision: str = "auto"):
        super().__init__()
        self.precision = precision

    def forward(self, video_features, text_features, log_temp=None):
        if log_temp is None:
            log_temp = torch.log(torch.tensor(0.1, device=video_features.device))
        return clip_loss(video_features.half(), text_features.half(), log_temp, gated=True, clamp_min=0.0,
                         precision=self.precision, use_ddp=True).float()


class InfoNCELoss(nn.Module):
    """utils/loss/losses.py:278-331 — dispatcher that owns a ``log_temp`` Parameter."""

    def __init__(self, temperature: float = 0.07, use_ddp: bool = False, loss_type: str = "contrastive"):
        super().__init__()
        self.temperature = temperature
        self.use_ddp = use_ddp
        self.loss_type = loss_type
        self.log_temp = nn.Parameter(torch.log(torch.tensor(temperature)))

    def forward(self, video_features, text_features, log_temp: Optional[torch.Tensor] = None):
        temp = log_temp if log_temp is not None else self.log_temp
        ddp = self.use_ddp and dist.is_available() and dist.is_initialized()
        if self.loss_type == "siglip":
            return (SiglipLossDDP() if ddp else SiglipLoss())(video_features, text_features, temp)
        if self.loss_type == "contrastive":
            return (ContrastiveLossDDP() if ddp else ContrastiveLoss())(video_features, text_features, temp)
        raise ValueError(f"Invalid loss type: {self.loss_type}")
