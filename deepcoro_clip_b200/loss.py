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

import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import dist_plan, ops, symm
from ._lib import B200ClipError

BW_CLIP, BW_GATED, BW_SIGLIP, BW_SIGLIP_ENT = 0, 1, 2, 3
_GSTORE_MAX_BYTES = 4 << 30   # largest stored gradient-tile matrix (bf16 [N, N]: 2 GiB at N = 32,768) of the one-recompute backward


_LOCAL = "local"   # cfg['group'] marker: never gather, even inside an initialised process group


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
_world = dist_plan.world
_all_gather_rows = dist_plan.gather_rows


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


class _SymmToken:
    """Lives as long as the autograd graph of one forward: the symmetric operand slot it used stays reserved."""
    __slots__ = ("__weakref__",)


def _symm_eligible(x: torch.Tensor) -> bool:
    return x.stride(1) == 1 and x.data_ptr() % 16 == 0 and (x.stride(0) * x.element_size()) % 16 == 0


class _ClipLossFn(torch.autograd.Function):
    """Softmax-CE both directions, diagonal targets.

    Host side kept lean on purpose (at 8 GPUs a step is ~1 ms and Python dispatch is the limiter): one zeroed fp32 arena
    per pass, ONE all-reduce of [colsum | rowsum | target dots] (3N floats), one finalize kernel for the scalar tail."""

    @staticmethod
    def forward(ctx, video, text, log_temp, label_smoothing, gated, clamp_min, precision, use_ddp, group, stable=None):
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
        st = ops.stream_ptr(dev)

        lo = rank * B
        vraw, traw = ops._rowmajor(video.detach()), ops._rowmajor(text.detach())
        # Multi-GPU exchange over symmetric memory (symm.py): the normalise kernel stores the operand rows into every rank's
        # [N, ld] buffer (normalise + all-gather in one launch), the statistics are read from the peers by the finalize
        # kernel (one-shot exchange instead of an all-reduce). Falls back to the NCCL collectives when unavailable.
        plan = slot = token = None
        if W > 1 and not x3 and D % 8 == 0 and ops.round_up(D, 64) <= 1024 and _symm_eligible(vraw) and _symm_eligible(traw):
            plan = symm.get_plan(group, N, ops.round_up(D, 64), dev)
            if plan is not None:
                token = _SymmToken()
                slot = plan.acquire(token)
                if slot is None:
                    plan = None
        if plan is not None:
            Kp = K = ops.round_up(D, 64)
            if torch.cuda.is_current_stream_capturing():
                plan.barrier_ops()        # a replayed graph reuses THIS slot every step: wait until every rank left the last one
            vall, tall = plan.ops[slot, 0], plan.ops[slot, 1]
            tinv = torch.empty(B, dtype=torch.float32, device=dev)
            vinv = torch.empty(B, dtype=torch.float32, device=dev)
            if plan.mc_op is not None:        # one multicast store per row (NVSwitch replicates it to every rank)
                ops.call("l2norm_fwd_mc", traw, ops.DTYPE_CODE[traw.dtype], ops.i64(traw.stride(0)), B, D,
                         plan.mc_op[slot][1], ops.i64(lo), K, Kp, tinv, 1, st)
                ops.call("l2norm_fwd_mc", vraw, ops.DTYPE_CODE[vraw.dtype], ops.i64(vraw.stride(0)), B, D,
                         plan.mc_op[slot][0], ops.i64(lo), K, Kp, vinv, 1, st)
            else:
                ops.call("l2norm_fwd_multi", traw, ops.DTYPE_CODE[traw.dtype], ops.i64(traw.stride(0)), B, D,
                         plan.op_ptrs[slot][1], W, ops.i64(lo), K, Kp, tinv, 1, st)
                ops.call("l2norm_fwd_multi", vraw, ops.DTYPE_CODE[vraw.dtype], ops.i64(vraw.stride(0)), B, D,
                         plan.op_ptrs[slot][0], W, ops.i64(lo), K, Kp, vinv, 1, st)
            plan.barrier_ops()            # every rank's rows have landed in every buffer
            vop, top = vall[lo:lo + B], tall[lo:lo + B]
            t_work = v_work = None
        else:
            # text operands first: their all-gather overlaps the video normalise; the video all-gather (needed only by
            # the backward's text-side pass) overlaps the forward tile kernel
            top, tinv, Kp = ops.l2norm_operand(text, 1 if x3 else -1)
            tall, t_work = dist_plan.gather_rows_async(top, W, group)
            vop, vinv, _ = ops.l2norm_operand(video, 0 if x3 else -1)
            vall, v_work = dist_plan.gather_rows_async(vop, W, group)
            K = vop.shape[1]
        dyn = ops.dyn_prep(log_temp, None, clamp_min, ops.GATED_BOUND if gated else 1.0)
        if stable is not None:              # A/B override of the device-side choice (tests, tools)
            ops.call("dyn_set_stable", dyn, int(bool(stable)), st)

        # statistics: [colsum (N) | rowsum (N) | dots (N) | lse2 rows (N) | lse2 columns (N) | fp32 dots (N) | column-sweep
        # dots (N)] zeroed (other ranks' slices stay 0) — this rank's symmetric block, or a local arena that is all-reduced —
        # then scales (2N) | tickets of the stable sweeps (2B int32)
        if plan is not None:
            plan.stats[slot].zero_()          # statistics + the backward's fp64 scalar sums of this slot
            sums = plan.stats[slot][:7 * N]
            ws = torch.zeros(2 * N + 2 * B + 2, dtype=torch.float32, device=dev)
        else:
            ws_all = torch.zeros(9 * N + 2 * B + 2, dtype=torch.float32, device=dev)
            sums = ws_all[:7 * N]
            ws = ws_all[7 * N:]
        # target logits S_ii from the raw features in fp32 (local pairs: video row r <-> text row r)
        ops.call("rowdot_raw", vraw, ops.DTYPE_CODE[vraw.dtype], ops.i64(vraw.stride(0)), vinv, traw,
                 ops.DTYPE_CODE[traw.dtype], ops.i64(traw.stride(0)), tinv, B, D, sums[5 * N + lo:5 * N + lo + B], st)
        if t_work is not None:
            t_work.wait()
        # Fixed-shift sweep (row + column sums in one pass) and the stable pair of row-LSE sweeps (running maxima; the
        # column statistics are the row statistics of the role-swapped problem) are BOTH enqueued: dyn[11], written by
        # dyn_prep from tau on the device, lets exactly one variant run (the other grid returns at once), so a learnable
        # temperature is never read by the host. tau >= ~0.013 (every shipped config): fixed shift.
        ops.call("logits_lse_fwd", vop, tall, B, N, K, vop.stride(0), tall.stride(0), 0.0, 0.0, int(gated), dyn, 1,
                 sums[N + lo:N + lo + B], sums[:N], sums[2 * N + lo:2 * N + lo + B], lo, st)
        slots = ops._lib.lib().b200clip_rowlse_slots(B, N, K)
        part = torch.empty(2 * B * slots * 2, dtype=torch.float32, device=dev)
        tick = ws[2 * N:2 * N + 2 * B].view(torch.int32)
        ops.call("logits_rowlse", vop, tall, B, N, K, vop.stride(0), tall.stride(0), int(gated), dyn, 1, part[:2 * B * slots],
                 slots, tick[:B], sums[3 * N + lo:3 * N + lo + B], sums[2 * N + lo:2 * N + lo + B], lo,
                 sums[N + lo:N + lo + B], st)
        if v_work is not None:
            v_work.wait()
            v_work = None
        ops.call("logits_rowlse", top, vall, B, N, K, top.stride(0), vall.stride(0), int(gated), dyn, 1, part[2 * B * slots:],
                 slots, tick[B:], sums[4 * N + lo:4 * N + lo + B], sums[6 * N + lo:6 * N + lo + B], lo, sums[lo:lo + B], st)
        if W > 1 and plan is None:
            dist.all_reduce(sums, group=group)
        rowscale_all = ws[:N]
        colscale_all = ws[N:2 * N]
        unif_tgt = vsum = tsum = None
        if eps != 0.0:
            if gated:
                raise B200ClipError("label_smoothing is not defined for the gated legacy loss")
            vsum = ops.colsum_bf16(vall[:, K - Kp:], N, D)      # hi panel (last in bf16x3 mode)
            tsum = ops.colsum_bf16(tall[:, K - Kp:], N, D)
            if x3:
                # bf16x3 operands: add the lo panels (video [lo|hi|hi], text [hi|lo|hi]) so the uniform-target term
                # sum_ij S_ij keeps the fp32-level accuracy of the logits (it dominates the error at small N otherwise)
                ops.colsum_bf16(vall[:, :Kp], N, D, out=vsum)
                ops.colsum_bf16(tall[:, Kp:2 * Kp], N, D, out=tsum)
            unif_tgt = ((eps / N) * torch.dot(vsum.double(), tsum.double()) * dyn[2].double()).reshape(1)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        if plan is not None:
            plan.barrier_stats()          # every rank's block is complete: read the peers' blocks directly
            ops.call("clip_finalize_peers", plan.stat_ptrs[slot], W, N, 7, dyn, eps, int(gated), unif_tgt, rowscale_all,
                     colscale_all, loss, None, st)
        else:
            ops.call("clip_finalize", sums, N, 7, dyn, eps, int(gated), unif_tgt, rowscale_all, colscale_all, loss, None, st)
        ctx.symm = (plan, slot, token)

        ctx.save_for_backward(video, text, vop, top, vall, tall, vinv, tinv, dyn, rowscale_all, colscale_all, vsum, tsum,
                              unif_tgt)
        ctx.cfg = (B, D, Kp, K, W, rank, N, eps, mode, group, log_temp.shape, log_temp.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (video, text, vop, top, vall, tall, vinv, tinv, dyn, rowscale_all, colscale_all, vsum, tsum,
         unif_tgt) = ctx.saved_tensors
        B, D, Kp, K, W, rank, N, eps, mode, group, lt_shape, lt_dtype = ctx.cfg
        dev = video.device
        gmul = grad_out.detach().reshape(1)
        if gmul.dtype != torch.float32:
            gmul = gmul.float()
        lo, hi = rank * B, (rank + 1) * B
        need_v, need_t, need_lt = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dV = dT = dLT = None
        # arena: dVhat [B, D] | dThat [B, D] | diag corrections 2 x [B, 2] | fp64 scalars (8 floats)
        nbd = B * D
        plan, slot, token = ctx.symm
        ws = torch.zeros(2 * nbd + 4 * B + 8, dtype=torch.float32, device=dev)
        # fp64 scalar sums: in the rank's symmetric block when the exchange runs over peer memory (zeroed by the forward),
        # so that d log_temp needs no all-reduce: every rank reads the W partials after one more barrier
        peers_lt = plan is not None and W > 1 and need_lt
        scal = plan.stats[slot][7 * N:7 * N + 8].view(torch.float64) if peers_lt else ws[2 * nbd + 4 * B:].view(torch.float64)
        if peers_lt:
            if getattr(ctx, "scal_used", False):
                scal.zero_()                   # a second backward through the same graph (retain_graph)
            ctx.scal_used = True
        ydiag = (1.0 - eps) / N
        lt_work = None
        # One recompute of the logits for both gradients (single GPU, plain bf16 operands, padded D in {256, 512, 768}): the video-side pass stores
        # its G tiles (bf16, 2 N^2 bytes, blocked layout) and the text-side gradient is the plain product G^T V̂ (csrc/gt_gemm.cu) instead of a
        # second pass with a second recompute. B200CLIP_GSTORE=0 keeps the two passes (A/B measurements).
        both = False
        if (W == 1 and need_v and need_t and K == Kp and Kp in (256, 512, 768) and 2 * ops.gstore_elems(B, N) <= _GSTORE_MAX_BYTES
                and os.environ.get("B200CLIP_GSTORE", "1") != "0"):
            dVh, dTh = ws[:nbd].view(B, D), ws[nbd:2 * nbd].view(B, D)
            dcv = ws[2 * nbd:2 * nbd + 2 * B]
            G = torch.empty(ops.gstore_elems(B, N), dtype=torch.bfloat16, device=dev)
            both = ops.logits_bwd_both(mode, vop, tall, B, N, K, D, dyn, rowscale_all, colscale_all, dVh, dTh, scal, G,
                                       ydiag=ydiag, diag_off=0, diag_corr=dcv, gnorm=2.0 * N)
            del G
        if both:
            dV = ops.l2norm_backward(dVh, video, vinv, other_x=text, other_inv=tinv, other_hi=top[:, K - Kp:], diag_corr=dcv,
                                     usum=tsum, ucoef=-eps / (N * N), dev_omul=dyn[2:3], dev_gmul=gmul)
            # G_jj is the same number on both sides: the text rows take the same diagonal corrections
            dT = ops.l2norm_backward(dTh, text, tinv, other_x=video, other_inv=vinv, other_hi=vop[:, K - Kp:], diag_corr=dcv,
                                     usum=vsum, ucoef=-eps / (N * N), dev_omul=dyn[2:3], dev_gmul=gmul)
            if dV.dtype != video.dtype:
                dV = dV.to(video.dtype)
            if dT.dtype != text.dtype:
                dT = dT.to(text.dtype)
            need_v = need_t = False        # done; d log_temp below
        if need_v or (need_lt and not both):
            dVh = ws[:nbd].view(B, D)
            dcv = ws[2 * nbd:2 * nbd + 2 * B]
            ops.logits_bwd(mode, vop, tall, B, N, K, Kp, D, dyn, rowscale_all[lo:hi], colscale_all, dVh, scal,
                           ydiag=ydiag, diag_off=lo, diag_corr=dcv, gnorm=2.0 * N, hp=(K != Kp))
            if need_lt and W > 1 and not peers_lt:
                # sum_ij G_ij L_ij is complete after the video-side pass: its all-reduce overlaps the text-side pass
                lt_work = dist.all_reduce(scal[0:1], group=group, async_op=True)
            if need_v:
                dV = ops.l2norm_backward(dVh, video, vinv, other_x=text, other_inv=tinv, other_hi=top[:, K - Kp:],
                                         diag_corr=dcv, usum=tsum, ucoef=-eps / (N * N), dev_omul=dyn[2:3],
                                         dev_gmul=gmul)
                if dV.dtype != video.dtype:
                    dV = dV.to(video.dtype)
        if need_t:
            dTh = ws[nbd:2 * nbd].view(B, D)
            dct = ws[2 * nbd + 2 * B:2 * nbd + 4 * B]
            ops.logits_bwd(mode, top, vall, B, N, K, Kp, D, dyn, colscale_all[lo:hi], rowscale_all, dTh, None,
                           ydiag=ydiag, diag_off=lo, diag_corr=dct, gnorm=2.0 * N, hp=(K != Kp))
            dT = ops.l2norm_backward(dTh, text, tinv, other_x=video, other_inv=vinv, other_hi=vop[:, K - Kp:],
                                     diag_corr=dct, usum=vsum, ucoef=-eps / (N * N), dev_omul=dyn[2:3],
                                     dev_gmul=gmul)
            if dT.dtype != text.dtype:
                dT = dT.to(text.dtype)
        if need_lt:
            if lt_work is not None:
                lt_work.wait()
            # d loss / d log_temp = -sum_ij G_ij L_ij  (zero while the tau clamp is active); the kernel's sum already
            # contains the diagonal target, only the uniform label-smoothing part is added
            dlt = torch.empty(1, dtype=torch.float32, device=dev)
            if peers_lt:
                plan.barrier_scal()            # every rank's video-side pass (enqueued before its text-side pass) has finished
                ops.call("clip_dlogtemp_peers", plan.scal_ptrs[slot], W, dyn, gmul, unif_tgt, N, dlt, ops.stream_ptr(dev))
            else:
                ops.call("clip_dlogtemp", scal, dyn, gmul, unif_tgt, N, dlt, ops.stream_ptr(dev))
            dLT = (dlt if lt_dtype == torch.float32 else dlt.to(lt_dtype)).reshape(lt_shape)
        if plan is not None:
            plan.release(slot, token)      # the operand slot may be overwritten by the step after next
        return dV, dT, dLT, None, None, None, None, None, None, None


def clip_loss(video_features, text_features, log_temp, *, label_smoothing: float = 0.0, gated: bool = False,
              clamp_min: float = 1e-4, precision: str = "auto", use_ddp: bool = True, group=None,
              stable=None) -> torch.Tensor:
    """Functional form of the fused softmax contrastive loss (forward + custom backward).

    ``stable``: None (default) lets the device choose between the fixed-shift sweep and the running-maximum (stable) sweeps
    from tau; True / False force one of them (False is only valid inside the fixed-shift window, tau >= ~0.013)."""
    if not isinstance(log_temp, torch.Tensor):
        log_temp = torch.tensor(float(log_temp), device=video_features.device)
    if log_temp.device != video_features.device:
        log_temp = log_temp.to(video_features.device)
    return _ClipLossFn.apply(video_features, text_features, log_temp, label_smoothing, gated, clamp_min, precision,
                             use_ddp, group, stable)


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


# --------------------------------------------------------------------------------------------------
# SigLIP multi-positive sigmoid loss
# --------------------------------------------------------------------------------------------------
_NO_CLAMP = 3.0e38   # logit clamp of the variants that do not clamp (utils/loss/siglip2_bce.py:88-90)


def _siglip_cfg(**kw) -> dict:
    cfg = dict(positive_weight=1.0, negative_weight=1.0, use_severity_weights=True, auto_balance=False,
               precision="auto", max_positives=64, group=None, tau_clamp=1e-4, logit_clamp=30.0, label_smoothing=0.0,
               pos_rule_mask=False, entropy=False, entropy_weight=0.1, entropy_threshold=2.0, ent_holder=None,
               text_replicated=False)
    cfg.update(kw)
    return cfg


class _SigLIPFn(torch.autograd.Function):
    """utils/loss/contrastive.py:250-315; SURVEY Appendix A.2. Backward is a pure recompute, so when any input
    needs a gradient the two gradient passes run eagerly in forward (no separate forward GEMM) and ``backward``
    only applies the normalise-backward with the upstream scale.

    ``cfg`` selects the reference variant: ``tau_clamp`` (1e-4 | 0), ``logit_clamp`` (30 | none), ``label_smoothing``
    (y (1 - eps) + eps / 2 on every pair, siglip2_bce.py:98-99), ``pos_rule_mask`` (positive weight where
    pos_mask > 0, siglip_pairwise.py:352), ``entropy`` (contrastive.py:19-68, 306-313).

    Process groups. The reference gathers video, pos_mask and pos_weights and multiplies against THE RANK'S OWN text
    (contrastive.py:252-256, 263): every rank holds the loss over [B_global, T_local], computed W times over when the ranks
    happen to hold the same texts. Default (``text_replicated=False``): exactly that — the raw video rows and the dense
    masks are all-gathered, the rank evaluates its [B_global, T_local] problem, keeps its own rows of the video gradient
    (the gather's backward, :87-91) and the full gradient of its own text; no result is reduced across ranks, texts may
    differ per rank. ``text_replicated=True`` is the sharded fast path for callers that pass the SAME text matrix on every
    rank (BASELINE config 2: the global text batch all-gathered by the caller): each rank computes only its
    [B_local, T] row slab — 1/W of the reference's per-rank work — and the loss scalars and the text gradient are
    all-reduced; a checksum of the text operand rides in the same all-reduce and the loss is NaN when the ranks' texts
    differ (loud, no host sync)."""

    @staticmethod
    def forward(ctx, video, text, log_temp, bias, pos_mask, pos_weights, cfg):
        dev = ops.require_cuda(video, text, log_temp, bias)
        if video.dim() != 2 or text.dim() != 2 or video.shape[1] != text.shape[1]:
            raise ValueError(f"video_features {tuple(video.shape)} / text_features {tuple(text.shape)} must be "
                             "[B, D] and [T, D]")
        B, D = video.shape
        T = text.shape[0]
        W, rank = (1, 0) if cfg["group"] is _LOCAL else _world(True, cfg["group"])
        Bg = B * W
        if pos_mask is not None and tuple(pos_mask.shape) != (B, T):
            raise ValueError(f"pos_mask must be [B, T] = {(B, T)}, got {tuple(pos_mask.shape)}")
        if pos_weights is not None and tuple(pos_weights.shape) != (B, T):
            raise ValueError(f"pos_weights must be [B, T] = {(B, T)}, got {tuple(pos_weights.shape)}")
        own = None
        video_in = video
        if W > 1 and not cfg["text_replicated"]:
            # reference semantics for per-rank texts (:252-256): gather video / mask / weights, evaluate the
            # [B_global, T_local] problem locally, keep the own rows of the video gradient
            grp = cfg["group"]
            video = dist_plan.gather_rows(ops._rowmajor(video.detach()), W, grp)
            if pos_mask is not None:
                pos_mask = dist_plan.gather_rows(pos_mask.detach().float().contiguous(), W, grp)
            if pos_weights is not None:
                pos_weights = dist_plan.gather_rows(pos_weights.detach().float().contiguous(), W, grp)
            own = (rank * B, (rank + 1) * B)
            B, W, rank = Bg, 1, 0
        x3 = _pick_precision(cfg["precision"], Bg, T)
        vop, vinv, Kp = ops.l2norm_operand(video, 0 if x3 else -1)
        top, tinv, _ = ops.l2norm_operand(text, 1 if x3 else -1)
        K = vop.shape[1]
        st = ops.stream_ptr(dev)
        dyn = ops.dyn_prep(log_temp, bias, cfg["tau_clamp"], 1.0)
        eps = float(cfg["label_smoothing"])
        if cfg["logit_clamp"] != 30.0 or eps != 0.0:
            ops.call("dyn_set_siglip", dyn, float(cfg["logit_clamp"]), 0.5 * eps, st)
        c = 1.0 / (Bg * T)
        wp, wn = cfg["positive_weight"], cfg["negative_weight"]
        gn = 1.0 / ((wn if wn > 0.0 else 1.0) * c)     # the dense G is fed to the tensor core as G * gn = O(1)

        need_grad = any(ctx.needs_input_grad[:4])
        # one zeroed arena (128-byte header keeps the gradient buffers 16-byte aligned for the vectorised reductions):
        # fp64 sums [8] | overflow flag | pad | dVhat [B, D] | dThat [T, D]
        # replicated text on several ranks: the [T, D] text gradient and the scalar sums are reduced over symmetric memory
        # (symm.SymmReducePlan) instead of two NCCL all-reduces, whose HOST cost (150-350 us each) bounds the eager step
        rplan = rslot = rtoken = None
        if W > 1 and need_grad and video.is_cuda and not torch.cuda.is_current_stream_capturing():
            rplan = symm.get_reduce_plan(cfg["group"], T * D, dev)
            if rplan is not None:
                rtoken = _SymmToken()
                rslot = rplan.acquire(rtoken)
                if rslot is None:
                    rplan = None
        nz = ((B * D) if rplan is not None else (B * D + T * D)) if need_grad else 0
        arena = torch.zeros(32 + nz, dtype=torch.float32, device=dev)
        acc = arena[:16].view(torch.float64)     # [0] sum g*s [1] sum softplus [2] sum g | [4..6] positives
        overflow = arena[16:17].view(torch.int32)
        dVh = dTh = None

        # ---- positives: one streaming pass over the dense mask / weights ----
        cap = cfg["max_positives"]
        # one allocation for the compacted positive lists: col | y | w  [B, cap] each, then cnt [B], ysum [B]
        lists = torch.empty(3 * B * cap + 2 * B, dtype=torch.int32, device=dev)
        col = lists[:B * cap].view(B, cap)
        yv = lists[B * cap:2 * B * cap].view(torch.float32).view(B, cap)
        wv = lists[2 * B * cap:3 * B * cap].view(torch.float32).view(B, cap)
        cnt = lists[3 * B * cap:3 * B * cap + B]
        ysum = lists[3 * B * cap + B:].view(torch.float32)
        pm = pw = None
        if pos_mask is not None:
            pm = pos_mask.detach().float()
            pm = pm if pm.stride(1) == 1 else pm.contiguous()
            if pos_weights is not None and cfg["use_severity_weights"]:
                pw = pos_weights.detach().float()
                pw = pw if pw.stride(1) == 1 else pw.contiguous()
        if pos_mask is None and W > 1:
            # diagonal targets on the GLOBAL [B_global, T] matrix (:274-278): local row r is global row rank*B + r
            rows = torch.arange(B, device=dev, dtype=torch.int64) + rank * B
            valid = rows < min(Bg, T)
            col[:, 0] = rows.clamp(max=T - 1).to(torch.int32)
            yv[:, 0] = 1.0
            wv[:, 0] = 1.0
            cnt.copy_(valid.to(torch.int32))
            ysum.copy_(valid.float())
        else:
            ops.call("siglip_compact", pm, ops.i64(pm.stride(0) if pm is not None else 0), pw,
                     ops.i64(pw.stride(0) if pw is not None else 0), B, T, cap, col, yv, wv, cnt, ysum, overflow, st)

        # ---- entropy regulariser: row statistics of softmax_j(L_ij) before the gradient passes ----
        mode, rowvec, ent = BW_SIGLIP, None, None
        if cfg["entropy"]:
            zhq = torch.zeros((3, B), dtype=torch.float32, device=dev)
            ops.call("siglip_entropy_rowsum", vop, top, B, T, K, vop.stride(0), top.stride(0), dyn, zhq[0], st)
            ops.call("siglip_entropy_stats", vop, top, B, T, K, vop.stride(0), top.stride(0), dyn, zhq[0], zhq[1], zhq[2],
                     st)
            rowvec = torch.empty((B, 2), dtype=torch.float32, device=dev)
            stats = torch.empty(3, dtype=torch.float64, device=dev)
            ops.call("siglip_entropy_rows", zhq[0], zhq[1], zhq[2], B, rowvec, stats, st)
            stats_all = stats
            if W > 1:
                stats_all = torch.empty(W * 3, dtype=torch.float64, device=dev)      # [W, 3], flat for every backend
                dist.all_gather_into_tensor(stats_all, stats, group=cfg["group"])
            ent = torch.empty(8, dtype=torch.float32, device=dev)
            ops.call("siglip_entropy_coef", stats_all, W, Bg, T, float(cfg["entropy_weight"]),
                     float(cfg["entropy_threshold"]), dyn, ent, st)
            mode = BW_SIGLIP_ENT

        if need_grad:
            dVh = arena[32:32 + B * D].view(B, D)
            if rplan is not None:
                rplan.buf[rslot].zero_()                   # the text gradient and the scalar block of this slot
                dTh = rplan.buf[rslot][:T * D].view(T, D)
            else:
                dTh = arena[32 + B * D:].view(T, D)
            # one recompute of the logits for both gradients (CLIPLoss above): the [B, T] slab of G is stored and the text-side
            # gradient is G^T V̂ — valid for any world size here, the text gradient is summed over the ranks anyway
            both = False
            if (mode == BW_SIGLIP and not x3 and K == Kp and Kp in (256, 512, 768)
                    and 2 * ops.gstore_elems(B, T) <= _GSTORE_MAX_BYTES and os.environ.get("B200CLIP_GSTORE", "1") != "0"):
                G = torch.empty(ops.gstore_elems(B, T), dtype=torch.bfloat16, device=dev)
                both = ops.logits_bwd_both(mode, vop, top, B, T, K, D, dyn, None, None, dVh, dTh, acc[0:4], G, gnorm=gn,
                                           wneg_c=wn * c)
                del G
            if not both:
                ops.logits_bwd(mode, vop, top, B, T, K, Kp, D, dyn, rowvec, None, dVh, acc[0:4], wneg_c=wn * c, gnorm=gn,
                               hp=x3)
                ops.logits_bwd(mode, top, vop, T, B, K, Kp, D, dyn, None, rowvec, dTh, None, wneg_c=wn * c, gnorm=gn, hp=x3)
        else:
            ops.call("siglip_dense_fwd", vop, top, B, T, K, vop.stride(0), top.stride(0), dyn, acc[1:2], st)
        flags = int(pw is not None) | (2 if cfg["pos_rule_mask"] else 0)
        # positives: gradient formed with the fp32 normalised partner rows (raw features x 1/norm), not the bf16 operands
        vraw, traw = ops._rowmajor(video.detach()), ops._rowmajor(text.detach())
        ops.call("siglip_pos", vop, vop.stride(0), top, top.stride(0), K, Kp, D, K - Kp, B, T, cap, col, yv, wv, cnt,
                 ysum, dyn, float(wp), float(wn), float(c), float(gn), int(x3), flags, int(cfg["auto_balance"]), dVh,
                 D if dVh is not None else 0, dTh, D if dTh is not None else 0, acc[4:7],
                 vraw, ops.DTYPE_CODE[vraw.dtype], ops.i64(vraw.stride(0)), vinv,
                 traw, ops.DTYPE_CODE[traw.dtype], ops.i64(traw.stride(0)), tinv, st)
        # local sums -> global (every rank returns the full loss, reference DDP semantics); scalar tails on the device
        # loss, dbias, sum G*s | checksum of the text operand and its square (replication check, see the class docstring)
        red = torch.empty(6, dtype=torch.float64, device=dev)      # [5] carries the overflow flag through the same all-reduce
        if rplan is not None:
            red_sym = rplan.buf[rslot][T * D:T * D + 12].view(torch.float64)
            ops.call("siglip_combine", acc, float(wn * c), tinv, T, red_sym, st)
            red_sym[5:6].copy_(overflow)
            rplan.allreduce(rslot, 6, red)
            overflow.copy_(red[5:6])
        else:
            ops.call("siglip_combine", acc, float(wn * c), tinv, T, red, st)
        if rplan is not None:
            pass
        elif W > 1:
            red[5:6].copy_(overflow)
            dist.all_reduce(red, group=cfg["group"])
            if dTh is not None:
                dist.all_reduce(dTh, group=cfg["group"])      # text is replicated: every rank gets the full text grad
            overflow.copy_(red[5:6])
        elif own is not None:
            dist.all_reduce(overflow, op=dist.ReduceOp.MAX, group=cfg["group"])
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        diag = None
        if ent is not None and cfg["ent_holder"] is not None:
            diag = torch.empty(7, dtype=torch.float32, device=dev)
        # NaN on overflow (never silently drop positives) and when the ranks of a text_replicated job hold different texts
        ops.call("siglip_loss_out", red, overflow, ent, W, loss, diag, st)
        if diag is not None:
            cfg["ent_holder"]["raw"] = diag
        if own is not None:             # the gather's backward keeps this rank's rows (contrastive.py:87-91)
            if dVh is not None:
                dVh = dVh[own[0]:own[1]]
            vinv = vinv[own[0]:own[1]]
        ctx.save_for_backward(video_in, text, vinv, tinv, dyn, dVh, dTh, red)
        ctx.rsymm = (rplan, rslot, rtoken)
        ctx.meta = (log_temp.shape, log_temp.dtype, None if bias is None else (bias.shape, bias.dtype))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        video, text, vinv, tinv, dyn, dVh, dTh, red = ctx.saved_tensors
        lt_shape, lt_dtype, bmeta = ctx.meta
        gmul = grad_out.detach().reshape(1).float().contiguous()
        dV = dT = dLT = dB = None
        if ctx.needs_input_grad[0]:
            dV = ops.l2norm_backward(dVh, video, vinv, dev_gmul=gmul).to(video.dtype)
        if ctx.needs_input_grad[1]:
            dT = ops.l2norm_backward(dTh, text, tinv, dev_gmul=gmul).to(text.dtype)
        rplan, rslot, rtoken = ctx.rsymm
        if rplan is not None:
            rplan.release(rslot, rtoken)       # the symmetric text-gradient slot may be overwritten by the step after next
        need_lt = ctx.needs_input_grad[2]
        need_b = bmeta is not None and ctx.needs_input_grad[3]
        if need_lt or need_b:
            # d loss / d log_temp = -sum G (R - b) = -(sum G*s)/tau (zero while the tau clamp is active); dbias = sum G
            sg = torch.empty(2, dtype=torch.float32, device=video.device)
            ops.call("siglip_scalar_grads", red, dyn, gmul, sg[0:1] if need_lt else None, sg[1:2] if need_b else None,
                     ops.stream_ptr(video.device))
            if need_lt:
                dLT = (sg[0:1] if lt_dtype == torch.float32 else sg[0:1].to(lt_dtype)).reshape(lt_shape)
            if need_b:
                dB = (sg[1:2] if bmeta[1] == torch.float32 else sg[1:2].to(bmeta[1])).reshape(bmeta[0])
        return dV, dT, dLT, dB, None, None, None


def _as_log_temp(log_temp, dev) -> torch.Tensor:
    if not isinstance(log_temp, torch.Tensor):
        log_temp = torch.tensor(float(log_temp), device=dev)
    return log_temp.to(dev)


_ENT_KEYS = ("entropy_mean", "entropy_min", "entropy_max", "entropy_normalized", "entropy_deficit", "entropy_loss",
             "bce_loss")


def _entropy_diagnostics(holder: dict) -> dict:
    """The reference fills this dict with seven .item() calls per step (contrastive.py:58-64, 311-312); here it is ONE
    device->host copy of seven floats."""
    raw = holder.pop("raw", None)
    return dict(zip(_ENT_KEYS, raw.tolist())) if raw is not None else {}


class SigLIPLoss(nn.Module):
    """utils/loss/contrastive.py:171-319 — sigmoid BCE over every (video, text) pair with multi-positive masks,
    severity weights, auto-balance and the optional entropy regulariser. Same constructor / forward signature /
    ``bias`` attribute / ``get_entropy_diagnostics()``."""

    def __init__(self, bias_init: float = -10.0, learnable_bias: bool = True, positive_weight: float = 1.0,
                 negative_weight: float = 1.0, use_severity_weights: bool = True, auto_balance: bool = False,
                 entropy_regularization: bool = False, entropy_weight: float = 0.1,
                 min_entropy_threshold: float = 2.0, precision: str = "auto", max_positives_per_row: int = 64,
                 text_replicated: bool = False):
        super().__init__()
        self.text_replicated = bool(text_replicated)     # see _SigLIPFn: sharded fast path for identical texts on all ranks
        self.positive_weight = max(float(positive_weight), 1e-6)
        self.negative_weight = max(float(negative_weight), 1e-6)
        self.use_severity_weights = use_severity_weights
        self.auto_balance = auto_balance
        self.entropy_regularization = entropy_regularization
        self.entropy_weight = entropy_weight
        self.min_entropy_threshold = min_entropy_threshold
        self.precision = precision
        self.max_positives_per_row = int(max_positives_per_row)
        self._last_entropy_diagnostics: dict = {}
        if learnable_bias:
            self.bias = nn.Parameter(torch.tensor(bias_init))
        else:
            self.register_buffer("bias", torch.tensor(bias_init))

    def forward(self, video_features: torch.Tensor, text_features: torch.Tensor, log_temp: torch.Tensor,
                pos_mask: Optional[torch.Tensor] = None, pos_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = video_features.device
        log_temp = _as_log_temp(log_temp, dev)
        bias = self.bias if self.bias.device == dev else self.bias.to(dev)
        holder: dict = {}
        cfg = _siglip_cfg(positive_weight=self.positive_weight, negative_weight=self.negative_weight,
                          use_severity_weights=self.use_severity_weights, auto_balance=self.auto_balance,
                          precision=self.precision, max_positives=self.max_positives_per_row,
                          entropy=bool(self.entropy_regularization), entropy_weight=self.entropy_weight,
                          entropy_threshold=self.min_entropy_threshold, ent_holder=holder,
                          text_replicated=self.text_replicated)
        loss = _SigLIPFn.apply(video_features, text_features, log_temp, bias, pos_mask, pos_weights, cfg)
        if self.entropy_regularization:
            self._last_entropy_diagnostics = _entropy_diagnostics(holder)
        return loss

    def get_entropy_diagnostics(self) -> dict:
        return self._last_entropy_diagnostics


class SiglipPairwiseFeatureLoss(nn.Module):
    """utils/loss/siglip_pairwise.py:261-376 — the feature-level pairwise loss kept by the reference for backwards
    compatibility: no bias, tau NOT clamped (:333), logits clamped to +-30 (:337), positive weight wherever
    ``pos_mask > 0`` (:352), optional entropy regulariser; gathers video / mask / weights under DDP (:322-326)."""

    def __init__(self, *, positive_weight: float = 1.0, negative_weight: float = 1.0, use_positive_weights: bool = True,
                 auto_positive_weight: bool = False, entropy_regularization: bool = False, entropy_weight: float = 0.1,
                 min_entropy_threshold: float = 2.0, precision: str = "auto", max_positives_per_row: int = 64,
                 text_replicated: bool = False) -> None:
        super().__init__()
        self.text_replicated = bool(text_replicated)
        self.positive_weight = max(float(positive_weight), 1e-6)
        self.negative_weight = max(float(negative_weight), 0.0)
        self.use_positive_weights = use_positive_weights
        self.auto_positive_weight = auto_positive_weight
        self.entropy_regularization = entropy_regularization
        self.entropy_weight = entropy_weight
        self.min_entropy_threshold = min_entropy_threshold
        self.precision = precision
        self.max_positives_per_row = int(max_positives_per_row)
        self._last_entropy_diagnostics: dict = {}

    def forward(self, video_features: torch.Tensor, text_features: torch.Tensor, log_temp: torch.Tensor,
                pos_mask: torch.Tensor, pos_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = video_features.device
        holder: dict = {}
        cfg = _siglip_cfg(positive_weight=self.positive_weight, negative_weight=self.negative_weight,
                          use_severity_weights=self.use_positive_weights, auto_balance=self.auto_positive_weight,
                          precision=self.precision, max_positives=self.max_positives_per_row, tau_clamp=0.0,
                          pos_rule_mask=True, entropy=bool(self.entropy_regularization),
                          entropy_weight=self.entropy_weight, entropy_threshold=self.min_entropy_threshold,
                          ent_holder=holder, text_replicated=self.text_replicated)
        loss = _SigLIPFn.apply(video_features, text_features, _as_log_temp(log_temp, dev), None, pos_mask, pos_weights,
                               cfg)
        if self.entropy_regularization:
            self._last_entropy_diagnostics = _entropy_diagnostics(holder)
        return loss

    def get_entropy_diagnostics(self) -> dict:
        return self._last_entropy_diagnostics


class _GatherTextRows(torch.autograd.Function):
    """all_gather of the local text rows for the SigLIP2 DDP loss (siglip2_bce.py:194-224). ``_SigLIPFn`` already
    all-reduces the text gradient, so the backward is the reference's: keep this rank's chunk, no reduce."""

    @staticmethod
    def forward(ctx, x, group):
        W, rank = _world(True, group)
        ctx.meta = (rank, x.shape[0])
        out = torch.empty((W * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        rank, n = ctx.meta
        return g[rank * n:(rank + 1) * n], None


class SigLIP2BCELoss(nn.Module):
    """utils/loss/siglip2_bce.py:22-108 — one-to-one sigmoid BCE: identity labels, learnable bias, tau and logits NOT
    clamped, label smoothing ``y (1 - eps) + eps / 2`` (:98-99); never gathers."""

    _gather = False

    def __init__(self, bias_init: float = -10.0, learnable_bias: bool = True, label_smoothing: float = 0.0,
                 precision: str = "auto"):
        super().__init__()
        self.label_smoothing = label_smoothing
        self.precision = precision
        if learnable_bias:
            self.bias = nn.Parameter(torch.tensor(bias_init))
        else:
            self.register_buffer("bias", torch.tensor(bias_init))

    def forward(self, video_features: torch.Tensor, text_features: torch.Tensor,
                log_temp: torch.Tensor) -> torch.Tensor:
        dev = video_features.device
        if video_features.shape != text_features.shape:
            raise ValueError(f"video_features {tuple(video_features.shape)} and text_features "
                             f"{tuple(text_features.shape)} must both be [B, D]")
        bias = self.bias if self.bias.device == dev else self.bias.to(dev)
        ddp = self._gather and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if ddp:
            text_features = _GatherTextRows.apply(text_features, None)
        # the DDP variant gathers the text itself (one line above): replicated by construction
        cfg = _siglip_cfg(precision=self.precision, tau_clamp=0.0, logit_clamp=_NO_CLAMP,
                          label_smoothing=float(self.label_smoothing), group=None if ddp else _LOCAL,
                          text_replicated=True)
        return _SigLIPFn.apply(video_features, text_features, _as_log_temp(log_temp, dev), bias, None, None, cfg)


class SigLIP2BCELossDDP(SigLIP2BCELoss):
    """utils/loss/siglip2_bce.py:115-186 — the same loss over the global batch: video AND text are gathered."""

    _gather = True


class SigLIP2MultiPositiveBCELoss(nn.Module):
    """utils/loss/siglip2_bce.py:227-335 — multi-positive sigmoid BCE with label smoothing: logits clamped to +-30
    (:296), tau not clamped, weights ``labels > 0.5 ? positive_weight [* pos_weights] : negative_weight`` (:313-322);
    single process only (the reference never gathers here)."""

    def __init__(self, bias_init: float = -10.0, learnable_bias: bool = True, positive_weight: float = 1.0,
                 negative_weight: float = 1.0, label_smoothing: float = 0.0, precision: str = "auto",
                 max_positives_per_row: int = 64):
        super().__init__()
        self.positive_weight = positive_weight
        self.negative_weight = negative_weight
        self.label_smoothing = label_smoothing
        self.precision = precision
        self.max_positives_per_row = int(max_positives_per_row)
        if learnable_bias:
            self.bias = nn.Parameter(torch.tensor(bias_init))
        else:
            self.register_buffer("bias", torch.tensor(bias_init))

    def forward(self, video_features: torch.Tensor, text_features: torch.Tensor, log_temp: torch.Tensor,
                pos_mask: Optional[torch.Tensor] = None, pos_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = video_features.device
        bias = self.bias if self.bias.device == dev else self.bias.to(dev)
        cfg = _siglip_cfg(positive_weight=float(self.positive_weight), negative_weight=float(self.negative_weight),
                          precision=self.precision, max_positives=self.max_positives_per_row, tau_clamp=0.0,
                          label_smoothing=float(self.label_smoothing), group=_LOCAL)
        return _SigLIPFn.apply(video_features, text_features, _as_log_temp(log_temp, dev), bias, pos_mask, pos_weights,
                               cfg)
