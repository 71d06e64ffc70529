"""Drop-in for models/video_aggregator.py::EnhancedVideoAggregator (reference :57-159). Same constructor, parameter
names and forward; each transformer block over the N <= 16 views is one cluster kernel per direction (csrc/xfblock.cu,
SURVEY §8f #4), the tail (positional add when there are no blocks, final LayerNorm, masked query pooling :128-158) is one
fused kernel."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

import ctypes
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


_SAVED = ("xhat1", "rstd1", "h1", "qkv", "attn", "o", "x1", "xhat2", "rstd2", "h2", "z", "u")
_WORK = ("d_f2", "d_z", "d_ao", "d_qkv", "d_h2", "d_h1")
_LAYOUTS: dict = {}


def _xf_layout(B: int, N: int, D: int, H: int, F_: int):
    """Byte offsets of the saved tensors / backward work buffers of one block inside their flat fp32 workspaces."""
    key = (B, N, D, H, F_)
    lay = _LAYOUTS.get(key)
    if lay is None:
        R = B * N
        sizes = dict(xhat1=R * D, rstd1=R, h1=R * D, qkv=3 * R * D, attn=B * H * N * N, o=R * D, x1=R * D, xhat2=R * D, rstd2=R,
                     h2=R * D, z=R * F_, u=R * F_, d_f2=R * D, d_z=R * F_, d_ao=R * D, d_qkv=3 * R * D, d_h2=R * D, d_h1=R * D)
        off, cur = {}, 0
        for k in _SAVED:
            off[k] = 4 * cur
            cur += _al(sizes[k])
        n_saved, cur = cur, 0
        for k in _WORK:
            off[k] = 4 * cur
            cur += _al(sizes[k])
        lay = _LAYOUTS[key] = (off, n_saved, cur)
    return lay


def _xf_forward(x_ptr, out_ptr, mk, params, ws_ptr, off, dims, eps1, eps2, drop_p, seed, st):
    B, N, D, H, F_ = dims
    table = (ctypes.c_void_p * 35)(x_ptr, out_ptr, mk.data_ptr() if mk is not None else None, *[t.data_ptr() for t in params],
                                   *[ws_ptr + off[k] for k in _SAVED], *([None] * 8))
    call("xfblock", 0, table, B, N, D, H, F_, eps1, eps2, drop_p, seed, i64(mk.stride(0) if mk is not None else 0), st)


# element counts of the twelve parameter gradients of a block, in _fused_params order
def _xf_grad_sizes(D: int, F_: int):
    return (D, D, 3 * D * D, 3 * D, D * D, D, D, D, F_ * D, F_, D * F_, D)


def _xf_backward(x_ptr, dout_ptr, dx_ptr, mk, params, ws_ptr, ws2_ptr, off, dims, eps1, eps2, drop_p, seed, g_ptr, st):
    """Backward cluster kernel + the four weight-gradient launches; the twelve gradients are written at g_ptr in
    _fused_params order (fp32, packed)."""
    B, N, D, H, F_ = dims
    R = B * N
    table = (ctypes.c_void_p * 35)(x_ptr, None, mk.data_ptr() if mk is not None else None, *[t.data_ptr() for t in params],
                                   *[ws_ptr + off[k] for k in _SAVED], dout_ptr, dx_ptr, *[ws2_ptr + off[k] for k in _WORK])
    call("xfblock", 1, table, B, N, D, H, F_, eps1, eps2, drop_p, seed, i64(mk.stride(0) if mk is not None else 0), st)
    g, cur = [], g_ptr
    for n in _xf_grad_sizes(D, F_):
        g.append(cur)
        cur += 4 * n
    ln1w, ln1b, w_in, b_in, w_o, b_o, ln2w, ln2b, w1, b1, w2, b2 = g
    w, q = ws_ptr, ws2_ptr
    call("xfblock_wgrad", q + off["d_f2"], i64(D), w + off["u"], i64(F_), w2, b2, D, F_, R, None, None, None, None, 0, st)
    call("xfblock_wgrad", q + off["d_z"], i64(F_), w + off["h2"], i64(D), w1, b1, F_, D, R, q + off["d_h2"], w + off["xhat2"],
         ln2w, ln2b, D, st)
    call("xfblock_wgrad", q + off["d_ao"], i64(D), w + off["o"], i64(D), w_o, b_o, D, D, R, None, None, None, None, 0, st)
    call("xfblock_wgrad", q + off["d_qkv"], i64(3 * D), w + off["h1"], i64(D), w_in, b_in, 3 * D, D, R, q + off["d_h1"],
         w + off["xhat1"], ln1w, ln1b, D, st)


def _grad_views(flat, params):
    """Views of the packed gradient buffer with the shapes of ``params`` (one split + reshapes)."""
    return [v.view(p.shape) for v, p in zip(torch.split(flat, [p.numel() for p in params]), params)]


class _XfBlockFn(torch.autograd.Function):
    """One transformer block of the aggregator in 1 + 5 library launches (csrc/xfblock.cu): the forward cluster kernel,
    the backward cluster kernel and four rank-(B N) weight-gradient updates."""

    @staticmethod
    def forward(ctx, x, mask, H, eps1, eps2, drop_p, seed, *params):
        B, N, D = x.shape
        dims = (B, N, D, H, params[8].shape[0])
        dev = x.device
        xc = x.detach().contiguous()
        mk = mask.to(torch.bool).contiguous().view(torch.uint8) if mask is not None else None
        off, n_saved, _ = _xf_layout(*dims)
        ws = torch.empty(n_saved, dtype=torch.float32, device=dev)
        out = torch.empty((B, N, D), dtype=torch.float32, device=dev)
        _xf_forward(xc.data_ptr(), out.data_ptr(), mk, params, ws.data_ptr(), off, dims, float(eps1), float(eps2), float(drop_p),
                    int(seed), stream_ptr(dev))
        ctx.save_for_backward(xc, mk, ws, *params)
        ctx.cfg = (dims, float(eps1), float(eps2), float(drop_p), int(seed))
        return out

    @staticmethod
    def backward(ctx, dout):
        xc, mk, ws, *params = ctx.saved_tensors
        dims, eps1, eps2, drop_p, seed = ctx.cfg
        B, N, D, H, F_ = dims
        dev = xc.device
        off, _, n_work = _xf_layout(*dims)
        dout = dout.float().contiguous()
        ws2 = torch.empty(n_work, dtype=torch.float32, device=dev)
        dx = torch.empty((B, N, D), dtype=torch.float32, device=dev)
        flat = torch.empty(sum(_xf_grad_sizes(D, F_)), dtype=torch.float32, device=dev)
        _xf_backward(xc.data_ptr(), dout.data_ptr(), dx.data_ptr(), mk, params, ws.data_ptr(), ws2.data_ptr(), off, dims, eps1,
                     eps2, drop_p, seed, flat.data_ptr(), stream_ptr(dev))
        return (dx, None, None, None, None, None, None, *_grad_views(flat, params))


_AGG_SIZES: dict = {}


def _agg_sizes(B: int, N: int, D: int, H: int, F_: int):
    key = (B, N, D, H, F_)
    sz = _AGG_SIZES.get(key)
    if sz is None:
        buf = (ctypes.c_int64 * 3)()
        call("aggregator_sizes", B, N, D, H, F_, buf)
        sz = _AGG_SIZES[key] = tuple(buf)
    return sz


class _AggregatorFn(torch.autograd.Function):
    """The whole EnhancedVideoAggregator.forward with depth >= 1 as ONE autograd node and ONE library call per direction
    (b200clip_aggregator: positional add, the blocks of csrc/xfblock.cu, LayerNorm + query-pool tail of csrc/querypool.cu;
    2 + depth kernels forward, 2 + 5 depth + 2 backward). Same kernels as the per-module path; what it removes is the Python /
    autograd / allocator work between the launches, which bounded the eager step (840-950 us eager against 556 us
    graph-replayed at 8 studies x 4 views, depth 2)."""

    @staticmethod
    def forward(ctx, x, mask, pos, ln_w, ln_b, query, H, eps, drop_p, seeds, *params):
        B, N, D = x.shape
        depth = len(params) // 12
        dims = (B, N, D, H, params[8].shape[0])
        dev = x.device
        if x.stride(2) != 1:
            x = x.contiguous()
        mk = mask.to(torch.bool).contiguous().view(torch.uint8) if mask is not None else None
        n_saved, _, _ = _agg_sizes(*dims)
        acts = torch.empty((depth + 1, B, N, D), dtype=torch.float32, device=dev)
        ws = torch.empty(depth * n_saved, dtype=torch.float32, device=dev)
        out = torch.empty((B, D), dtype=torch.float32, device=dev)
        table = (ctypes.c_void_p * (15 + 12 * depth))(
            x.data_ptr(), None if pos is None else pos.data_ptr(), None if mk is None else mk.data_ptr(), acts.data_ptr(),
            ws.data_ptr(), out.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), query.data_ptr(), None, None, None, None, None, None,
            *[t.data_ptr() for t in params])
        eps_c = (ctypes.c_float * len(eps))(*eps)
        seeds_c = (ctypes.c_int64 * depth)(*seeds)
        mask_sb = mk.stride(0) if mk is not None else 0
        pos_rows = pos.shape[1] if pos is not None else 0
        call("aggregator", 0, table, depth, *dims, eps_c, drop_p, seeds_c, mask_sb, x.stride(0), x.stride(1), pos_rows,
             stream_ptr(dev))
        ctx.save_for_backward(acts, mk, ws, pos, ln_w, ln_b, query, *params)
        ctx.cfg = (dims, depth, eps_c, drop_p, seeds_c, mask_sb, pos_rows)
        return out

    @staticmethod
    def backward(ctx, dout):
        acts, mk, ws, pos, ln_w, ln_b, query, *params = ctx.saved_tensors
        dims, depth, eps_c, drop_p, seeds_c, mask_sb, pos_rows = ctx.cfg
        B, N, D, H, F_ = dims
        dev = acts.device
        _, n_work, per_block = _agg_sizes(*dims)
        n_pos = pos_rows * D if pos is not None else 0
        # returned gradients: d(final_ln.weight | final_ln.bias | attn_query) | d pos_encoding | block gradients
        grads = torch.empty(3 * D + n_pos + depth * per_block, dtype=torch.float32, device=dev)
        scratch = torch.empty(2 * B * N * D + n_work, dtype=torch.float32, device=dev)        # dact [2, B, N, D] | work
        if dout.dtype != torch.float32 or not dout.is_contiguous():
            dout = dout.float().contiguous()
        gp, sp = grads.data_ptr(), scratch.data_ptr()
        table = (ctypes.c_void_p * (15 + 12 * depth))(
            None, None if pos is None else pos.data_ptr(), None if mk is None else mk.data_ptr(), acts.data_ptr(), ws.data_ptr(),
            None, ln_w.data_ptr(), ln_b.data_ptr(), query.data_ptr(), dout.data_ptr(), sp, sp + 8 * B * N * D, gp,
            gp + 12 * D if pos is not None else None, gp + 4 * (3 * D + n_pos), *[t.data_ptr() for t in params])
        call("aggregator", 1, table, depth, *dims, eps_c, drop_p, seeds_c, mask_sb, 0, 0, pos_rows, stream_ptr(dev))
        dx = scratch[:B * N * D].view(B, N, D)
        sizes = [D, D, D] + ([n_pos] if pos is not None else []) + [t.numel() for t in params]
        parts = torch.split(grads, sizes)
        k = 4 if pos is not None else 3
        gpos = parts[3].view(pos.shape) if pos is not None else None
        gparams = [v if t.dim() == 1 else v.view(t.shape) for v, t in zip(parts[k:], params)]
        return (dx, None, gpos, parts[0], parts[1], parts[2].view(query.shape), None, None, None, None, *gparams)


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

    def _fused_eligible(self, x) -> bool:
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and os.environ.get("B200CLIP_XFBLOCK", "1") != "0"):
            return False
        params = self._fused_params()
        return bool(lib().b200clip_xfblock_ok(x.shape[1], x.shape[2], self.attn.num_heads, self.mlp[0].out_features)
                    and all(p is not None and p.dtype == torch.float32 and p.is_contiguous() for p in params))

    def _drop_seed(self):
        drop_p = float(self.dropout1.p) if self.training else 0.0
        return drop_p, (int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0.0 else 0)

    def forward(self, x, key_padding_mask: Optional[torch.Tensor] = None):
        if self._fused_eligible(x):
            drop_p, seed = self._drop_seed()
            return _XfBlockFn.apply(x, key_padding_mask, self.attn.num_heads, self.norm1.eps, self.norm2.eps, drop_p, seed,
                                    *self._fused_params())
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
            if x.is_cuda and x.dtype == torch.float32 and os.environ.get("B200CLIP_AGG_FUSED", "1") != "0" \
                    and os.environ.get("B200CLIP_XFBLOCK", "1") != "0":
                plan = self.__dict__.get("_plan")
                if plan is None:       # (module, parameter name) of everything the fused step reads; heads, widths, eps
                    blk0 = self.blocks[0]
                    owners = []
                    for b in self.blocks:
                        owners += [(b.norm1, "weight"), (b.norm1, "bias"), (b.attn, "in_proj_weight"), (b.attn, "in_proj_bias"),
                                   (b.attn.out_proj, "weight"), (b.attn.out_proj, "bias"), (b.norm2, "weight"), (b.norm2, "bias"),
                                   (b.mlp[0], "weight"), (b.mlp[0], "bias"), (b.mlp[3], "weight"), (b.mlp[3], "bias")]
                    uniform = all(b.attn.num_heads == blk0.attn.num_heads and b.mlp[0].out_features == blk0.mlp[0].out_features
                                  and b.dropout1.p == blk0.dropout1.p for b in self.blocks)
                    eps = tuple(float(e) for b in self.blocks for e in (b.norm1.eps, b.norm2.eps)) + (float(self.final_ln.eps),)
                    plan = self.__dict__["_plan"] = (owners, uniform, blk0.attn.num_heads, blk0.mlp[0].out_features, eps,
                                                     float(blk0.dropout1.p), self.final_ln, list(self.blocks))
                owners, uniform, heads, width, eps, p_drop, fln, blocks = plan
                params = [m._parameters[n] for m, n in owners]
                tail = (fln._parameters["weight"], fln._parameters["bias"], self.attn_query)
                if (uniform and lib().b200clip_xfblock_ok(N, D, heads, width)
                        and all(t is not None and t.dtype == torch.float32 and t.is_contiguous() for t in params)
                        and all(t is not None and t.dtype == torch.float32 and t.is_contiguous() for t in tail)
                        and (pos is None or (pos.dtype == torch.float32 and pos.is_contiguous()))):
                    drop_p = p_drop if self.training else 0.0
                    # one draw per block, like TransformerBlock._drop_seed: the same seeds as the per-module path
                    seeds = [int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0.0 else 0 for _ in blocks]
                    return _AggregatorFn.apply(x, mask, pos, *tail, heads, eps, drop_p, seeds, *params)
            if pos is not None:
                x = x + pos[:, :N, :]
                pos = None
            for blk in self.blocks:
                x = blk(x, key_padding_mask=mask)
        return query_pool(x, pos, self.final_ln.weight, self.final_ln.bias, self.attn_query, mask, self.final_ln.eps)
