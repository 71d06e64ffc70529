"""Gated-attention pooling of the multi-instance probing head: drop-in for
models/multi_instance_linear_probing.py::MultiInstanceLinearProbing._attention_pooling (reference :493-507) and
._hierarchical_attention_pooling (:509-536), SURVEY §8f #4. Same parameters (attention_V / attention_U / attention_w,
attn_dropout) and argument meaning; the gate products, the masked softmax and the weighted sum run in csrc/milpool.cu
(fp32 FMA tiles, or tcgen05 products on split-precision operands for >= 1024 rows and D in {256, 512, 768}). ``install()`` rebinds the two methods on the reference class."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import call, i64, lib, stream_ptr, try_call


def _plan(S: int, L: int, D: int, Hd: int):
    buf = (ctypes.c_int * 4)()
    call("milpool_plan", S, L, D, Hd, buf)
    return tuple(buf)


def _tc_plan(S: int, L: int, D: int, Hd: int):
    """(P, chunks, slots, Hp, g_elems) of the tensor-core variant, or None when the shape does not qualify
    (R = S L >= 1024 rows, D in {256, 512, 768}; B200CLIP_MIL_TC=0 keeps the fp32 FMA tiles)."""
    if os.environ.get("B200CLIP_MIL_TC", "1") == "0":
        return None
    buf = (ctypes.c_int64 * 5)()
    return tuple(buf) if try_call("milpool_tc_plan", S, L, D, Hd, buf) else None


_ONE3: dict = {}


def _one3(dev):
    t = _ONE3.get(dev)
    if t is None:
        t = _ONE3[dev] = torch.ones(4, dtype=torch.float32, device=dev)
    return t


class _GatedPool(torch.autograd.Function):
    """x [S, L, D] fp32 (row-contiguous view), valid [S, L] uint8 or None -> out [S, D]."""

    @staticmethod
    def forward(ctx, x, valid, V, bV, U, bU, w, bw, drop_p, seed):
        S, L, D = x.shape
        Hd = V.shape[0]
        dev = x.device
        R = S * L
        Vc, Uc, wc = V.detach().contiguous(), U.detach().contiguous(), w.detach().reshape(Hd).contiguous()
        bVc, bUc, bwc = bV.detach().contiguous(), bU.detach().contiguous(), bw.detach().reshape(1).contiguous()
        tcp = _tc_plan(S, L, D, Hd)
        tg = torch.empty((R, 2 * Hd), dtype=torch.float32, device=dev)
        attn = torch.empty(R, dtype=torch.float32, device=dev)
        out = torch.empty((S, D), dtype=torch.float32, device=dev)
        vs = i64(valid.stride(0) if valid is not None else 0)
        if tcp is not None:
            P, chunks, slots, Hp, g_elems = tcp
            x3 = torch.empty((R, 3 * D), dtype=torch.bfloat16, device=dev)
            w3 = torch.empty((2 * Hd, 3 * D), dtype=torch.bfloat16, device=dev)
            wt3 = torch.empty((D, 3 * Hp), dtype=torch.bfloat16, device=dev)
            spart = torch.empty((slots, R), dtype=torch.float32, device=dev)
            opart = torch.empty((S, P, D), dtype=torch.float32, device=dev) if P > 1 else None
            call("milpool_tc_fwd", x, i64(x.stride(0)), i64(x.stride(1)), valid, vs, Vc, bVc, Uc, bUc, wc, bwc, S, L, D, Hd,
                 float(drop_p), i64(seed), x3, w3, wt3, tg, spart, attn, opart, out, stream_ptr(dev))
            ctx.save_for_backward(x, wc, tg, attn, x3, wt3)
            ctx.cfg = (float(drop_p), int(seed), tcp, tuple(w.shape), tuple(bw.shape), Hd)
            return out
        P, Z, chunks, nut = _plan(S, L, D, Hd)
        spart = torch.empty((nut, R), dtype=torch.float32, device=dev)
        opart = torch.empty((S, P, D), dtype=torch.float32, device=dev) if P > 1 else None
        call("milpool_fwd", x, i64(x.stride(0)), i64(x.stride(1)), valid, vs, Vc, bVc, Uc, bUc, wc, bwc, S, L, D, Hd,
             float(drop_p), i64(seed), tg, spart, attn, opart, out, stream_ptr(dev))
        ctx.save_for_backward(x, Vc, Uc, wc, tg, attn)
        ctx.cfg = (float(drop_p), int(seed), None, Z, chunks, tuple(w.shape), tuple(bw.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        dev = dout.device
        dout = dout.float().contiguous()
        if ctx.cfg[2] is not None:
            x, wc, tg, attn, x3, wt3 = ctx.saved_tensors
            drop_p, seed, (P, chunks, slots, Hp, g_elems), w_shape, bw_shape, Hd = ctx.cfg
            S, L, D = x.shape
            R = S * L
            ds = torch.empty(R, dtype=torch.float32, device=dev)
            ad = torch.empty(R, dtype=torch.float32, device=dev)
            dx = torch.empty((S, L, D), dtype=torch.float32, device=dev)
            dpre3 = torch.empty((R, 3 * Hp), dtype=torch.bfloat16, device=dev)
            g2 = torch.empty((2, g_elems), dtype=torch.bfloat16, device=dev)
            fpart = torch.empty((chunks, 3 * Hd + 4), dtype=torch.float32, device=dev)
            dW = torch.empty((Hd, 2, D), dtype=torch.float32, device=dev)             # rows interleaved: [u, (V, U), :]
            dsm = torch.empty(3 * Hd + 1, dtype=torch.float32, device=dev)
            call("milpool_tc_bwd", x, i64(x.stride(0)), i64(x.stride(1)), wc, S, L, D, Hd, drop_p, i64(seed), x3, wt3, tg, attn,
                 dout, ds, dx, dpre3, g2[0], g2[1], ad, fpart, dW, dsm, _one3(dev), stream_ptr(dev))
            return (dx, None, dW[:, 0], dsm[:Hd], dW[:, 1], dsm[Hd:2 * Hd], dsm[2 * Hd:3 * Hd].view(w_shape),
                    dsm[3 * Hd:].view(bw_shape), None, None)
        x, Vc, Uc, wc, tg, attn = ctx.saved_tensors
        drop_p, seed, _, Z, chunks, w_shape, bw_shape = ctx.cfg
        S, L, D = x.shape
        Hd = Vc.shape[0]
        R = S * L
        ds = torch.empty(R, dtype=torch.float32, device=dev)
        dx = torch.empty((S, L, D), dtype=torch.float32, device=dev)
        wpart = torch.empty((Z, 2 * Hd, D), dtype=torch.float32, device=dev)
        dpre = torch.empty((R, 2 * Hd), dtype=torch.float32, device=dev)
        fpart = torch.empty((chunks, 3 * Hd + 4), dtype=torch.float32, device=dev)
        dW = torch.empty((2 * Hd, D), dtype=torch.float32, device=dev)
        dsm = torch.empty(3 * Hd + 1, dtype=torch.float32, device=dev)
        call("milpool_bwd", x, i64(x.stride(0)), i64(x.stride(1)), Vc, Uc, wc, S, L, D, Hd, drop_p, i64(seed), tg, attn,
             dout, ds, dx, dpre, wpart, fpart, dW, dsm, stream_ptr(dev))
        return (dx, None, dW[:Hd], dsm[:Hd], dW[Hd:], dsm[Hd:2 * Hd], dsm[2 * Hd:3 * Hd].view(w_shape),
                dsm[3 * Hd:].view(bw_shape), None, None)


def _rows(x: torch.Tensor) -> torch.Tensor:
    """[S, L, D] fp32 view the kernels can read: unit stride along D, 16-byte aligned rows."""
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(2) != 1 or x.stride(0) % 4 or x.stride(1) % 4 or x.data_ptr() % 16:
        x = x.contiguous()
    return x


def gated_attention_pool(x: torch.Tensor, mask: Optional[torch.Tensor], V_weight, V_bias, U_weight, U_bias, w_weight, w_bias,
                         dropout_p: float = 0.0, training: bool = False) -> torch.Tensor:
    """One level of gated-attention pooling (reference :499-507): x [S, L, D], mask [S, L] with True = instance present
    (or None) -> [S, D]. Weights as in nn.Linear(D, Hd) x 2 and nn.Linear(Hd, 1)."""
    ops.require_cuda(x)
    if x.dim() != 3:
        raise ValueError(f"expected [S, L, D], got {tuple(x.shape)}")
    S, L, D = x.shape
    Hd = V_weight.shape[0]
    if not lib().b200clip_milpool_ok(L, D, Hd):
        raise ValueError(f"gated attention pooling needs D % 16 == 0, hidden % 8 == 0, L <= 49152 (got L={L}, D={D}, hidden={Hd})")
    for t in (V_weight, V_bias, U_weight, U_bias, w_weight, w_bias):
        if t.dtype != torch.float32:
            raise ValueError("gated attention pooling weights must be float32")
    valid = None
    if mask is not None:
        if tuple(mask.shape) != (S, L):
            raise ValueError(f"Mask shape {tuple(mask.shape)} does not match input shape {(S, L)}")
        valid = mask.to(torch.bool).contiguous().view(torch.uint8)
    p = float(dropout_p) if training else 0.0
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0.0 else 0
    out = _GatedPool.apply(_rows(x), valid, V_weight, V_bias, U_weight, U_bias, w_weight, w_bias, p, seed)
    return out if x.dtype == torch.float32 else out.to(x.dtype)


def _params(mod):
    drop = getattr(mod, "attn_dropout", None)
    p = float(drop.p) if isinstance(drop, nn.Dropout) else 0.0
    return (mod.attention_V.weight, mod.attention_V.bias, mod.attention_U.weight, mod.attention_U.bias,
            mod.attention_w.weight, mod.attention_w.bias, p, mod.training)


def attention_pooling(self, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Replacement for MultiInstanceLinearProbing._attention_pooling (reference :493-507)."""
    if x.ndim == 4:
        return hierarchical_attention_pooling(self, x, mask)
    return gated_attention_pool(x, mask, *_params(self))


def hierarchical_attention_pooling(self, x: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Replacement for ._hierarchical_attention_pooling (reference :509-536): patch level without a mask over the L tokens of
    every video, then the video level with the mask, same weights."""
    B, N, L, D = x.shape
    prm = _params(self)
    video_emb = gated_attention_pool(x.reshape(B * N, L, D), None, *prm).view(B, N, D)
    return gated_attention_pool(video_emb, mask, *prm)


class GatedAttentionPooling(nn.Module):
    """The attention-pooling parameters of MultiInstanceLinearProbing (reference :185-189) as a module of their own:
    forward(x [B, N, D] or [B, N, L, D], mask [B, N] True = valid)."""

    def __init__(self, embedding_dim: int, attention_hidden: int = 128, dropout: float = 0.0):
        super().__init__()
        self.attention_V = nn.Linear(embedding_dim, attention_hidden)
        self.attention_U = nn.Linear(embedding_dim, attention_hidden)
        self.attention_w = nn.Linear(attention_hidden, 1)
        self.attn_dropout = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        for m in (self.attention_V, self.attention_U, self.attention_w):     # reference _reset_parameters :538-544
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if mask is None:
            mask = torch.ones(x.shape[0], x.shape[1], dtype=torch.bool, device=x.device)
        return attention_pooling(self, x, mask)
