"""Drop-in for models/video_aggregator.py::EnhancedVideoAggregator (reference :57-159). Same constructor, parameter
names and forward; the transformer blocks over the N <= ~15 views stay PyTorch ("next"), the tail
(positional add when there are no blocks, final LayerNorm, masked query pooling :128-158) is one fused kernel."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import call, i64, stream_ptr


class _QueryPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, ln_w, ln_b, query, mask, eps):
        ops.require_cuda(x)
        x = x.float()
        if x.stride(2) != 1:
            x = x.contiguous()
        B, N, D = x.shape
        dev = x.device
        posc = None if pos is None else pos.detach().float().reshape(-1, D)[:N].contiguous()
        lw, lb, q = ln_w.detach().float().contiguous(), ln_b.detach().float().contiguous(), \
            query.detach().float().reshape(D).contiguous()
        mk = None if mask is None else mask.to(torch.bool).contiguous().view(torch.uint8)
        out = torch.empty((B, D), dtype=torch.float32, device=dev)
        call("querypool", 0, x, i64(x.stride(0)), i64(x.stride(1)), posc, lw, lb, q, mk,
             i64(mk.stride(0) if mk is not None else 0), B, N, D, float(eps), out, None, None, None, None, None, None,
             stream_ptr(dev))
        ctx.save_for_backward(x, posc if posc is not None else torch.empty(0, device=dev), lw, lb, q,
                              mk if mk is not None else torch.empty(0, device=dev))
        ctx.flags = (posc is not None, mk is not None, float(eps), None if pos is None else tuple(pos.shape),
                     tuple(query.shape))
        return out

    @staticmethod
    def backward(ctx, dout):
        x, posc, lw, lb, q, mk = ctx.saved_tensors
        has_pos, has_mask, eps, pos_shape, q_shape = ctx.flags
        B, N, D = x.shape
        dev = x.device
        dx = torch.empty((B, N, D), dtype=torch.float32, device=dev)
        dpos = torch.zeros((N, D), dtype=torch.float32, device=dev) if has_pos else None
        dlw = torch.zeros(D, dtype=torch.float32, device=dev)
        dlb = torch.zeros(D, dtype=torch.float32, device=dev)
        dq = torch.zeros(D, dtype=torch.float32, device=dev)
        mkp = mk if has_mask else None
        call("querypool", 1, x, i64(x.stride(0)), i64(x.stride(1)), posc if has_pos else None, lw, lb, q, mkp,
             i64(mkp.stride(0) if mkp is not None else 0), B, N, D, eps, None, dout.float().contiguous(), dx, dpos, dlw,
             dlb, dq, stream_ptr(dev))
        gpos = None
        if has_pos:
            gpos = torch.zeros(pos_shape, dtype=torch.float32, device=dev)
            gpos.view(-1, D)[:N] = dpos
        return dx, gpos, dlw, dlb, dq.view(q_shape), None, None


def query_pool(x, pos_encoding, ln_weight, ln_bias, attn_query, mask=None, eps: float = 1e-5):
    """Fused tail of EnhancedVideoAggregator.forward: [B, N, D] -> [B, D]."""
    return _QueryPool.apply(x, pos_encoding, ln_weight, ln_bias, attn_query, mask, eps)


class TransformerBlock(nn.Module):
    """Pre-LN block (reference :7-54): x + drop(MHA(LN(x))) ; x + drop(MLP(LN(x))). Plain PyTorch."""

    def __init__(self, embedding_dim, num_heads, dropout):
        super().__init__()
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.attn = nn.MultiheadAttention(embed_dim=embedding_dim, num_heads=num_heads, dropout=dropout,
                                          batch_first=True)
        self.dropout1 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = nn.Sequential(nn.Linear(embedding_dim, embedding_dim * 4), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(embedding_dim * 4, embedding_dim))
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, x, key_padding_mask: Optional[torch.Tensor] = None):
        h = self.norm1(x)
        a, _ = self.attn(h, h, h, key_padding_mask=key_padding_mask)
        x = x + self.dropout1(a)
        return x + self.dropout2(self.mlp(self.norm2(x)))


class EnhancedVideoAggregator(nn.Module):
    def __init__(self, embedding_dim: int, num_heads: int = 4, dropout: float = 0.1,
                 use_positional_encoding: bool = True, aggregator_depth: int = 2, max_segments: int = 1024):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.use_positional_encoding = use_positional_encoding
        self.aggregator_depth = aggregator_depth
        if use_positional_encoding:
            self.pos_encoding = nn.Parameter(torch.zeros(1, max_segments, embedding_dim))
            nn.init.trunc_normal_(self.pos_encoding, std=0.02)
        else:
            self.pos_encoding = None
        self.blocks = nn.ModuleList([TransformerBlock(embedding_dim, num_heads, dropout)
                                     for _ in range(aggregator_depth)])
        self.final_ln = nn.LayerNorm(embedding_dim)
        self.attn_query = nn.Parameter(torch.randn(1, 1, embedding_dim))
        nn.init.normal_(self.attn_query, std=0.02)

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, N, D = x.shape
        if mask is not None:
            mask = mask.to(torch.bool)
        pos = self.pos_encoding
        if len(self.blocks) > 0:
            if pos is not None:
                x = x + pos[:, :N, :]
                pos = None
            for blk in self.blocks:
                x = blk(x, key_padding_mask=mask)
        return query_pool(x, pos, self.final_ln.weight, self.final_ln.bias, self.attn_query, mask, self.final_ln.eps)
