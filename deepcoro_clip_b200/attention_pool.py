"""Drop-in for models/attention_pool.py::AttentionPool (reference :10-101): same constructor, parameter names
(``query``, ``attn.in_proj_weight``, ``attn.in_proj_bias``, ``attn.out_proj.weight``, ``attn.out_proj.bias``,
``norm.*``, ``proj.*``) and forward signature, so checkpoints and the ``video_attention_pool`` optimizer group keep
working. The pass over the tokens is the streaming sm_100a kernel (csrc/attnpool.cu); the O(B*D^2) projections and
the LayerNorm around it act on [B, D] vectors and stay ordinary dense ops under autograd."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import DTYPE_CODE, call, i64, lib, stream_ptr


class _StreamPool(torch.autograd.Function):
    """xbar[b, h, :] = sum_n softmax_n(x[b, n] . qt[h]) x[b, n]   (fp32 [B, H, D])"""

    @staticmethod
    def forward(ctx, x, qt, mask):
        ops.require_cuda(x, qt)
        if x.dtype not in DTYPE_CODE:
            x = x.float()
        x = x.contiguous()
        B, N, D = x.shape
        H = qt.shape[0]
        dev = x.device
        if (D * x.element_size()) % 512 != 0 or H > 16:
            raise ValueError(f"AttentionPool kernel needs embed_dim*itemsize % 512 == 0 and heads <= 16 (got D={D}, "
                             f"{x.dtype}, heads={H})")
        qt32 = qt.detach().float().contiguous()
        mk = None
        if mask is not None:
            mk = mask.to(torch.bool).contiguous().view(torch.uint8)
        S = lib().b200clip_attnpool_splits(B, N)
        pm = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        pl = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        pa = torch.empty((B, S, H, D), dtype=torch.float32, device=dev)
        st = stream_ptr(dev)
        call("attnpool_fwd", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)), mk,
             i64(mk.stride(0) if mk is not None else 0), qt32, None, i64(0), i64(0), B, N, D, H, S, pm, pl, pa, st)
        xbar = torch.empty((B, H, D), dtype=torch.float32, device=dev)
        m = torch.empty((B, H), dtype=torch.float32, device=dev)
        l = torch.empty((B, H), dtype=torch.float32, device=dev)
        call("attnpool_merge", pm, pl, pa, B, S, H, D, xbar, m, l, 0, st)
        ctx.save_for_backward(x, qt32, mk if mk is not None else torch.empty(0, device=dev), xbar, m, l)
        ctx.has_mask = mk is not None
        ctx.S = S
        return xbar

    @staticmethod
    def backward(ctx, dxbar):
        x, qt32, mk, xbar, m, l = ctx.saved_tensors
        mk = mk if ctx.has_mask else None
        B, N, D = x.shape
        H = qt32.shape[0]
        dev = x.device
        st = stream_ptr(dev)
        dxbar = dxbar.float().contiguous()
        dx = torch.empty((B, N, D), dtype=x.dtype, device=dev)
        ds = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        mb = i64(mk.stride(0) if mk is not None else 0)
        call("attnpool_bwd_dx", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)), mk, mb, qt32, dxbar, xbar, m,
             l, B, N, D, H, dx, ds, st)
        dqt = None
        if ctx.needs_input_grad[1]:
            S = ctx.S
            pa = torch.empty((B, S, H, D), dtype=torch.float32, device=dev)
            call("attnpool_fwd", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)), None, i64(0), None, ds,
                 i64(ds.stride(0)), i64(ds.stride(1)), B, N, D, H, S, None, None, pa, st)
            dqt = torch.zeros((H, D), dtype=torch.float32, device=dev)
            call("attnpool_merge", None, None, pa, B, S, H, D, dqt, None, None, 1, st)
        return (dx if ctx.needs_input_grad[0] else None), dqt, None


class AttentionPool(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int = 8, output_dim: int = None, dropout: float = 0.0):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.output_dim = output_dim or embed_dim
        self.query = nn.Parameter(torch.randn(1, 1, embed_dim))
        nn.init.trunc_normal_(self.query, std=0.02)
        # parameter container only (keeps the reference state-dict keys); its forward is never called
        self.attn = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads, dropout=dropout, batch_first=True)
        self.norm = nn.LayerNorm(embed_dim)
        self.proj = nn.Linear(embed_dim, self.output_dim) if self.output_dim != embed_dim else nn.Identity()
        self.dropout = dropout

    def forward(self, x: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
        B, N, D = x.shape
        assert D == self.embed_dim, f"Input dim {D} != expected {self.embed_dim}"
        if self.training and self.dropout > 0.0:
            raise NotImplementedError("AttentionPool kernel: attention dropout > 0 in training mode is not supported "
                                      "(the folded single-query algebra needs sum(a) = 1); use dropout=0.0")
        H, Dh = self.num_heads, D // self.num_heads
        with torch.autocast("cuda", enabled=False):
            W = self.attn.in_proj_weight.float()
            bias = self.attn.in_proj_bias.float()
            Wq, Wk, Wv = W[:D], W[D:2 * D], W[2 * D:]
            bq, bv = bias[:D], bias[2 * D:]
            q0 = F.linear(self.query.float().view(1, D), Wq, bq).view(H, Dh)
            qt = torch.einsum("hkd,hk->hd", Wk.view(H, Dh, D), q0) * (1.0 / math.sqrt(Dh))     # [H, D]
            xbar = _StreamPool.apply(x, qt, mask)                                              # [B, H, D] fp32
            o = torch.einsum("bhd,hkd->bhk", xbar, Wv.view(H, Dh, D)).reshape(B, D) + bv
            y = F.linear(o, self.attn.out_proj.weight.float(), self.attn.out_proj.bias.float())
            y = F.layer_norm(y, (D,), self.norm.weight.float(), self.norm.bias.float(), self.norm.eps)
            if isinstance(self.proj, nn.Linear):
                y = F.linear(y, self.proj.weight.float(), self.proj.bias.float())
        return y.to(x.dtype)
