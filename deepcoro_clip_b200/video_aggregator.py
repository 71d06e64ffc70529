"""Drop-in for models/video_aggregator.py::EnhancedVideoAggregator (reference :57-159). Same constructor, parameter
names and forward; each transformer block over the N <= 16 views is one cluster kernel per direction (csrc/xfblock.cu,
SURVEY §8f #4), the tail (positional add when there are no blocks, final LayerNorm, masked query pooling :128-158) is one
fused kernel."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

import os

from . import ops
from ._lib import call, i64, lib, stream_ptr


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


def _al(n: int) -> int:
    return (n + 63) // 64 * 64


class _XfBlockFn(torch.autograd.Function):
    """One transformer block of the aggregator in 1 + 5 library launches (csrc/xfblock.cu): the forward cluster kernel,
    the backward cluster kernel and four rank-(B N) weight-gradient updates."""

    @staticmethod
    def forward(ctx, x, mask, ln1w, ln1b, w_in, b_in, w_o, b_o, ln2w, ln2b, w1, b1, w2, b2, H, eps1, eps2, drop_p, seed):
        B, N, D = x.shape
        F_ = w1.shape[0]
        dev = x.device
        R = B * N
        xc = x.detach().contiguous()
        mk = mask.to(torch.bool).contiguous().view(torch.uint8) if mask is not None else None
        seg = (("xhat1", R * D), ("rstd1", R), ("h1", R * D), ("qkv", 3 * R * D), ("attn", B * H * N * N), ("o", R * D),
               ("x1", R * D), ("xhat2", R * D), ("rstd2", R), ("h2", R * D), ("z", R * F_), ("u", R * F_))
        off, cur = {}, 0
        for name, n in seg:
            off[name] = 4 * cur
            cur += _al(n)
        ws = torch.empty(cur, dtype=torch.float32, device=dev)
        out = torch.empty((B, N, D), dtype=torch.float32, device=dev)
        w = ws.data_ptr()
        ptrs = [xc.data_ptr(), out.data_ptr(), mk.data_ptr() if mk is not None else None] + \
               [t.data_ptr() for t in (ln1w, ln1b, w_in, b_in, w_o, b_o, ln2w, ln2b, w1, b1, w2, b2)] + \
               [w + off[k] for k, _ in seg] + [None] * 8
        import ctypes
        table = (ctypes.c_void_p * 35)(*ptrs)
        call("xfblock", 0, table, B, N, D, H, F_, float(eps1), float(eps2), float(drop_p), int(seed),
             i64(mk.stride(0) if mk is not None else 0), stream_ptr(dev))
        ctx.save_for_backward(xc, mk, ws, ln1w, ln1b, w_in, b_in, w_o, b_o, ln2w, ln2b, w1, b1, w2, b2)
        ctx.cfg = (B, N, D, H, F_, float(eps1), float(eps2), float(drop_p), int(seed), off, [k for k, _ in seg])
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes
        xc, mk, ws, ln1w, ln1b, w_in, b_in, w_o, b_o, ln2w, ln2b, w1, b1, w2, b2 = ctx.saved_tensors
        B, N, D, H, F_, eps1, eps2, drop_p, seed, off, names = ctx.cfg
        dev = xc.device
        R = B * N
        st = stream_ptr(dev)
        dout = dout.float().contiguous()
        seg2 = (("d_f2", R * D), ("d_z", R * F_), ("d_ao", R * D), ("d_qkv", 3 * R * D), ("d_h2", R * D), ("d_h1", R * D))
        o2, cur = {}, 0
        for name, n in seg2:
            o2[name] = 4 * cur
            cur += _al(n)
        ws2 = torch.empty(cur, dtype=torch.float32, device=dev)
        dx = torch.empty((B, N, D), dtype=torch.float32, device=dev)
        w, q = ws.data_ptr(), ws2.data_ptr()
        ptrs = [xc.data_ptr(), None, mk.data_ptr() if mk is not None else None] + \
               [t.data_ptr() for t in (ln1w, ln1b, w_in, b_in, w_o, b_o, ln2w, ln2b, w1, b1, w2, b2)] + \
               [w + off[k] for k in names] + [dout.data_ptr(), dx.data_ptr()] + [q + o2[k] for k, _ in seg2]
        table = (ctypes.c_void_p * 35)(*ptrs)
        call("xfblock", 1, table, B, N, D, H, F_, eps1, eps2, drop_p, seed, i64(mk.stride(0) if mk is not None else 0), st)
        g = {k: torch.empty_like(t) for k, t in (("ln1w", ln1w), ("ln1b", ln1b), ("w_in", w_in), ("b_in", b_in), ("w_o", w_o),
                                                 ("b_o", b_o), ("ln2w", ln2w), ("ln2b", ln2b), ("w1", w1), ("b1", b1),
                                                 ("w2", w2), ("b2", b2))}
        call("xfblock_wgrad", q + o2["d_f2"], i64(D), w + off["u"], i64(F_), g["w2"], g["b2"], D, F_, R, None, None, None,
             None, 0, st)
        call("xfblock_wgrad", q + o2["d_z"], i64(F_), w + off["h2"], i64(D), g["w1"], g["b1"], F_, D, R, q + o2["d_h2"],
             w + off["xhat2"], g["ln2w"], g["ln2b"], D, st)
        call("xfblock_wgrad", q + o2["d_ao"], i64(D), w + off["o"], i64(D), g["w_o"], g["b_o"], D, D, R, None, None, None,
             None, 0, st)
        call("xfblock_wgrad", q + o2["d_qkv"], i64(3 * D), w + off["h1"], i64(D), g["w_in"], g["b_in"], 3 * D, D, R,
             q + o2["d_h1"], w + off["xhat1"], g["ln1w"], g["ln1b"], D, st)
        return (dx, None, g["ln1w"], g["ln1b"], g["w_in"], g["b_in"], g["w_o"], g["b_o"], g["ln2w"], g["ln2b"], g["w1"],
                g["b1"], g["w2"], g["b2"], None, None, None, None, None)


class TransformerBlock(nn.Module):
    """Pre-LN block (reference :7-54): x + drop(MHA(LN(x))) ; x + drop(MLP(LN(x))). fp32 CUDA inputs with N <= 16 views,
    D <= 512 run as one cluster kernel per direction (csrc/xfblock.cu; B200CLIP_XFBLOCK=0 keeps the PyTorch ops below, which
    are also the path for every other shape / dtype). Dropout in training mode uses the library's counter-based mask seeded
    from torch's CPU generator (like AttentionPool), not PyTorch's Philox stream."""

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

    def _fused_params(self):
        return (self.norm1.weight, self.norm1.bias, self.attn.in_proj_weight, self.attn.in_proj_bias,
                self.attn.out_proj.weight, self.attn.out_proj.bias, self.norm2.weight, self.norm2.bias,
                self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight, self.mlp[3].bias)

    def forward(self, x, key_padding_mask: Optional[torch.Tensor] = None):
        if x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and os.environ.get("B200CLIP_XFBLOCK", "1") != "0":
            B, N, D = x.shape
            params = self._fused_params()
            if (lib().b200clip_xfblock_ok(N, D, self.attn.num_heads, self.mlp[0].out_features)
                    and all(p is not None and p.dtype == torch.float32 and p.is_contiguous() for p in params)):
                drop_p = float(self.dropout1.p) if self.training else 0.0
                seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0.0 else 0
                return _XfBlockFn.apply(x, key_padding_mask, *params, self.attn.num_heads, self.norm1.eps, self.norm2.eps,
                                        drop_p, seed)
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
