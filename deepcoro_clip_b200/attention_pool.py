"""Drop-ins for models/attention_pool.py::AttentionPool (reference :10-101) and ::AttentionPoolWithCLS (:104-197):
same constructors, parameter names (``query``, ``attn.in_proj_weight``, ``attn.in_proj_bias``,
``attn.out_proj.weight``, ``attn.out_proj.bias``, ``norm.*``, ``proj.*``; ``cls_token``, ``transformer.layers.0.*``)
and forward signatures, so checkpoints and the ``video_attention_pool`` optimizer group keep working. The pass over
the tokens is the streaming sm_100a kernel (csrc/attnpool.cu, csrc/attnpool_mma.cu); the O(B*D^2) projections, the
feed-forward block and the LayerNorms around it act on [B, D] vectors and stay ordinary dense ops under autograd."""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import DTYPE_CODE, call, i64, lib, stream_ptr


def _fused_dq_enabled() -> bool:
    """B200CLIP_POOL_FUSED_DQ=1: the backward accumulates the query gradient in the same pass over x instead of launching
    the weighted-sum kernel over all of x again. Opt-in: written after the round-1 GPU budget was spent — executed under
    the CPU emulation of the MMA kernels (tests/test_emulated_pool_kernels.py), not yet on hardware."""
    return os.environ.get("B200CLIP_POOL_FUSED_DQ", "0") == "1"


def _tc_splits(x: torch.Tensor, B: int, N: int, D: int, H: int) -> int:
    """Token splits of the tcgen05 pool kernels for this problem, 0 when they do not apply (fp32 x, wide rows, > 8 heads)
    or are switched off (B200CLIP_POOL_TC=0: the mma.sync / CUDA-core kernels, kept as the A/B baseline)."""
    if x.dtype not in (torch.bfloat16, torch.float16) or not x.is_cuda or os.environ.get("B200CLIP_POOL_TC", "1") == "0":
        return 0
    return int(lib().b200clip_attnpool_tc_splits(x.data_ptr(), DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)),
                                                 B, N, D, H))


class _StreamPool(torch.autograd.Function):
    """xbar[b, h, :] = sum_n a'[b, h, n] x[b, n],  a = softmax_n(x[b, n] . qt[h]),  a' = dropout(a)  (fp32 [B, H, D]);
    second output sa[b, h] = sum_n a'[b, h, n] (== 1 without dropout; it multiplies the value bias); third output (only
    with want_lse) lse[b, h] = log sum_n exp(x[b, n] . qt[h]) over the unmasked tokens, -inf when all are masked -- the
    statistic that lets a caller merge further keys into the same softmax (AttentionPoolWithCLS)."""

    @staticmethod
    def forward(ctx, x, qt, mask, drop_p, drop_seed, want_lse=False):
        ops.require_cuda(x, qt)
        if x.dtype not in DTYPE_CODE:
            x = x.float()
        x = x.contiguous()
        B, N, D = x.shape
        H = qt.shape[0]
        dev = x.device
        qt32 = qt.detach().float().contiguous()
        mk = None
        if mask is not None:
            mk = mask.to(torch.bool).contiguous().view(torch.uint8)
        # tcgen05 path (csrc/attnpool_tc.cu) when the shape allows it: 0 = not applicable
        S_tc = _tc_splits(x, B, N, D, H)
        if S_tc == 0 and ((D * x.element_size()) % 512 != 0 or H > 16):
            raise ValueError(f"AttentionPool kernel needs embed_dim*itemsize % 512 == 0 and heads <= 16 (got D={D}, "
                             f"{x.dtype}, heads={H})")
        S = S_tc if S_tc > 0 else lib().b200clip_attnpool_splits(B, N)
        pm = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        pl = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        pl2 = torch.empty((B, S, H), dtype=torch.float32, device=dev) if drop_p > 0.0 else None
        pa = torch.empty((B, S, H, D), dtype=torch.float32, device=dev)
        st = stream_ptr(dev)
        if S_tc > 0:
            call("attnpool_tc_fwd", x, DTYPE_CODE[x.dtype], mk, i64(mk.stride(0) if mk is not None else 0), qt32, None, B, N,
                 D, H, S, pm, pl, pa, float(drop_p), int(drop_seed), pl2, st)
        else:
            call("attnpool_fwd", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)), mk,
                 i64(mk.stride(0) if mk is not None else 0), qt32, None, i64(0), i64(0), B, N, D, H, S, pm, pl, pa,
                 float(drop_p), int(drop_seed), pl2, st)
        xbar = torch.empty((B, H, D), dtype=torch.float32, device=dev)
        m = torch.empty((B, H), dtype=torch.float32, device=dev)
        l = torch.empty((B, H), dtype=torch.float32, device=dev)
        sa = torch.empty((B, H), dtype=torch.float32, device=dev)
        call("attnpool_merge", pm, pl, pa, B, S, H, D, xbar, m, l, 0, pl2, sa, st)
        lse = None
        if want_lse:
            lse = m + torch.log(l)
            # fully masked rows: xbar is 0/0 there and the caller weighs it with exp(lse - LSE) = 0; keep 0 * NaN out of
            # both passes (AttentionPool itself keeps nn.MultiheadAttention's NaN for such rows)
            torch.nan_to_num_(xbar, nan=0.0)
            if pl2 is not None:
                torch.nan_to_num_(sa, nan=0.0)             # l2 / l of such a row
        ctx.save_for_backward(x, qt32, mk if mk is not None else torch.empty(0, device=dev), xbar, m, l, sa)
        ctx.has_mask = mk is not None
        ctx.S = S
        ctx.S_tc = S_tc
        ctx.drop = (float(drop_p), int(drop_seed))
        return xbar, sa, lse

    @staticmethod
    def backward(ctx, dxbar, dsa, dlse=None):
        x, qt32, mk, xbar, m, l, sa = ctx.saved_tensors
        mk = mk if ctx.has_mask else None
        drop_p, drop_seed = ctx.drop
        B, N, D = x.shape
        H = qt32.shape[0]
        dev = x.device
        st = stream_ptr(dev)
        dxbar = dxbar.float().contiguous()
        dsa = dsa.float().contiguous() if (dsa is not None and drop_p > 0.0) else None
        dlse = dlse.float().contiguous() if dlse is not None else None
        dx = torch.empty((B, N, D), dtype=x.dtype, device=dev)
        mb = i64(mk.stride(0) if mk is not None else 0)
        dqt = None
        if ctx.S_tc > 0:
            # one pass over x: dx and the per-split partials of the query gradient (no ds tensor, no second read of x)
            S = ctx.S_tc
            pdq = torch.empty((B, S, H, D), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
            call("attnpool_tc_bwd", x, DTYPE_CODE[x.dtype], mk, mb, qt32, dxbar, xbar, None, None, m, l, B, N, D, H, S, dx,
                 sa if dsa is not None else None, dsa, drop_p if dsa is not None else 0.0, drop_seed, dlse, pdq, st)
            if pdq is not None:
                dqt = torch.zeros((H, D), dtype=torch.float32, device=dev)
                call("attnpool_merge", None, None, pdq, B, S, H, D, dqt, None, None, 1, None, None, st)
            return (dx if ctx.needs_input_grad[0] else None), dqt, None, None, None, None
        ds = torch.empty((B, H, N), dtype=torch.float32, device=dev)
        if (ctx.needs_input_grad[1] and _fused_dq_enabled() and x.dtype in (torch.bfloat16, torch.float16) and H <= 8
                and D % 128 == 0 and D <= 1024 and x.data_ptr() % 16 == 0):
            # query gradient from the SAME pass over x (kDq kernels): per-(b, split) partials, summed by the merge kernel
            Sb = lib().b200clip_attnpool_bwd_splits(B, N)
            pdq = torch.zeros((B, Sb, H, D), dtype=torch.float32, device=dev)
            call("attnpool_bwd_dx_dq", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)), mk, mb, qt32, dxbar,
                 xbar, m, l, B, N, D, H, dx, ds, sa if dsa is not None else None, dsa,
                 drop_p if dsa is not None else 0.0, drop_seed, dlse, pdq, st)
            dqt = torch.zeros((H, D), dtype=torch.float32, device=dev)
            call("attnpool_merge", None, None, pdq, B, Sb, H, D, dqt, None, None, 1, None, None, st)
            return (dx if ctx.needs_input_grad[0] else None), dqt, None, None, None, None
        call("attnpool_bwd_dx", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)), mk, mb, qt32, dxbar, xbar, m,
             l, B, N, D, H, dx, ds, sa if dsa is not None else None, dsa, drop_p if dsa is not None else 0.0, drop_seed,
             dlse, st)
        if ctx.needs_input_grad[1]:
            S = ctx.S
            pa = torch.empty((B, S, H, D), dtype=torch.float32, device=dev)
            call("attnpool_fwd", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), i64(x.stride(1)), None, i64(0), None, ds,
                 i64(ds.stride(0)), i64(ds.stride(1)), B, N, D, H, S, None, None, pa, 0.0, 0, None, st)
            dqt = torch.zeros((H, D), dtype=torch.float32, device=dev)
            call("attnpool_merge", None, None, pa, B, S, H, D, dqt, None, None, 1, None, None, st)
        return (dx if ctx.needs_input_grad[0] else None), dqt, None, None, None, None


def _al(n: int) -> int:
    return (n + 63) // 64 * 64          # workspace segments start on 256-byte boundaries


_LAYOUTS: dict = {}


def _layout(key, segments):
    """Byte offsets of the named fp32 segments of one workspace (cached per problem configuration) + total floats."""
    lay = _LAYOUTS.get(key)
    if lay is None:
        off, cur = {}, 0
        for name, n in segments():
            off[name] = 4 * cur
            cur += _al(n)
        lay = _LAYOUTS[key] = (off, cur)
    return lay


class _FusedPool(torch.autograd.Function):
    """The whole AttentionPool forward / backward in 3 + 4 launches of the library (csrc/pooltail.cu + csrc/attnpool_tc.cu):
    pool_prep -> attnpool_tc_fwd -> pool_tail_fwd, and pool_tail_bwd -> pool_param_grads -> attnpool_tc_bwd -> pool_qgrads.
    Every intermediate lives in one fp32 workspace per pass."""

    @staticmethod
    def forward(ctx, x, mask, query, in_w, in_b, out_w, out_b, ln_w, ln_b, proj_w, proj_b, H, eps, drop_p, drop_seed, S):
        B, N, D = x.shape
        dev = x.device
        st = stream_ptr(dev)
        has_proj = proj_w is not None
        Do = proj_w.shape[0] if has_proj else D
        drop = drop_p > 0.0
        mk = mask.to(torch.bool).contiguous().view(torch.uint8) if mask is not None else None
        mb = mk.stride(0) if mk is not None else 0
        code = DTYPE_CODE[x.dtype]
        fp16 = 1 if x.dtype == torch.float16 else 0
        off, total = _layout(("f", B, D, H, S, drop, has_proj), lambda: (
            ("q0", D), ("qt", H * D), ("img", D * 8), ("pm", B * S * H), ("pl", B * S * H), ("pl2", B * S * H if drop else 0),
            ("pa", B * S * H * D), ("xbar", B * H * D), ("m", B * H), ("l", B * H), ("sa", B * H), ("o", B * D),
            ("yhat", B * D), ("rstd", B), ("yln", B * D if has_proj else 0)))
        ws = torch.empty(total, dtype=torch.float32, device=dev)
        w = ws.data_ptr()
        out = torch.empty((B, Do), dtype=x.dtype, device=dev)
        wv, bv = in_w.data_ptr() + 8 * D * D, in_b.data_ptr() + 8 * D
        pl2 = w + off["pl2"] if drop else None
        call("pool_prep", query, in_w, in_b, D, H, w + off["q0"], w + off["qt"], w + off["img"], fp16, st)
        call("attnpool_tc_fwd", x, code, mk, mb, None, w + off["img"], B, N, D, H, S, w + off["pm"], w + off["pl"],
             w + off["pa"], drop_p, drop_seed, pl2, st)
        call("pool_tail_fwd", w + off["pm"], w + off["pl"], pl2, w + off["pa"], B, S, H, D, wv, bv, out_w, out_b, ln_w, ln_b,
             eps, proj_w, proj_b, Do if has_proj else 0, w + off["xbar"], w + off["m"], w + off["l"], w + off["sa"],
             w + off["o"], w + off["yhat"], w + off["rstd"], w + off["yln"] if has_proj else None, out, code, st)
        ctx.save_for_backward(x, mk, ws, query, in_w, in_b, out_w, ln_w, proj_w)
        ctx.cfg = (B, N, D, H, S, Do, has_proj, drop_p, drop_seed, off)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mk, ws, query, in_w, in_b, out_w, ln_w, proj_w = ctx.saved_tensors
        B, N, D, H, S, Do, has_proj, drop_p, drop_seed, off = ctx.cfg
        dev = x.device
        st = stream_ptr(dev)
        mb = mk.stride(0) if mk is not None else 0
        code = DTYPE_CODE[x.dtype]
        fp16 = 1 if x.dtype == torch.float16 else 0
        if dout.dtype not in DTYPE_CODE:
            dout = dout.float()
        if not dout.is_contiguous():
            dout = dout.contiguous()
        dcode = DTYPE_CODE[dout.dtype]
        w = ws.data_ptr()
        o2, total2 = _layout(("b", B, D, H, S), lambda: (
            ("dyln", B * D), ("dy", B * D), ("do", B * D), ("dxbar", B * H * D), ("dsa", B * H), ("cdot", B * H),
            ("wimg", B * D * 16), ("pdq", B * S * H * D), ("dqt", H * D)))
        ws2 = torch.empty(total2, dtype=torch.float32, device=dev)
        q = ws2.data_ptr()
        # parameter gradients: fresh tensors (AccumulateGrad takes them over without a copy)
        gq, gw, gb = torch.empty_like(query), torch.empty_like(in_w), torch.empty_like(in_b)
        gow, gob = torch.empty_like(out_w), torch.empty((D,), dtype=torch.float32, device=dev)
        glw, glb = torch.empty_like(ln_w), torch.empty_like(ln_w)
        gpw = torch.empty_like(proj_w) if has_proj else None
        gpb = torch.empty((Do,), dtype=torch.float32, device=dev) if has_proj else None
        dx = torch.empty((B, N, D), dtype=x.dtype, device=dev)
        use_sa = 1 if drop_p > 0.0 else 0
        wv, bv = in_w.data_ptr() + 8 * D * D, in_b.data_ptr() + 8 * D
        call("pool_tail_bwd", dout, dcode, w + off["yhat"], w + off["rstd"], w + off["xbar"], w + off["sa"], wv, bv, out_w,
             ln_w, proj_w, Do if has_proj else 0, w + off["qt"], B, H, D, q + o2["dyln"], q + o2["dy"], q + o2["do"],
             q + o2["dxbar"], q + o2["dsa"] if use_sa else None, q + o2["cdot"], q + o2["wimg"], fp16, st)
        call("pool_param_grads", q + o2["dy"], w + off["o"], q + o2["do"], w + off["xbar"], w + off["sa"], use_sa,
             q + o2["dyln"], w + off["yhat"], dout, dcode, w + off["yln"] if has_proj else None, Do if has_proj else 0, B, H,
             D, gow, gob, gw.data_ptr() + 8 * D * D, gb.data_ptr() + 8 * D, glw, glb, gpw, gpb, st)
        call("attnpool_tc_bwd", x, code, mk, mb, None, None, None, q + o2["wimg"], q + o2["cdot"], w + off["m"], w + off["l"],
             B, N, D, H, S, dx, w + off["sa"] if use_sa else None, q + o2["dsa"] if use_sa else None, drop_p, drop_seed, None,
             q + o2["pdq"], st)
        call("pool_qgrads", q + o2["pdq"], B * S, w + off["q0"], query, in_w, H, D, q + o2["dqt"], gw, gb, gq, st)
        return dx, None, gq, gw, gb, gow, gob, glw, glb, gpw, gpb, None, None, None, None, None


def _fused_ok(x, params, D, H, Do) -> bool:
    if os.environ.get("B200CLIP_POOL_FUSED", "1") == "0":
        return False
    for p in params:
        if p is not None and (p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda):
            return False
    return bool(lib().b200clip_pooltail_ok(D, H, Do))


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
        drop_p, drop_seed = 0.0, 0
        if self.training and self.dropout > 0.0:
            # nn.MultiheadAttention drops attention WEIGHTS (after the softmax). Same here, with a counter-based mask
            # seeded from torch's CPU generator (reproducible under torch.manual_seed; not PyTorch's Philox stream).
            drop_p = float(self.dropout)
            drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        H, Dh = self.num_heads, D // self.num_heads
        has_proj = isinstance(self.proj, nn.Linear)
        params = (self.query, self.attn.in_proj_weight, self.attn.in_proj_bias, self.attn.out_proj.weight,
                  self.attn.out_proj.bias, self.norm.weight, self.norm.bias,
                  self.proj.weight if has_proj else None, self.proj.bias if has_proj else None)
        if x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and _fused_ok(x, params, D, H, self.output_dim if has_proj else 0):
            xc = x.contiguous()
            S = _tc_splits(xc, B, N, D, H)
            if S > 0:
                return _FusedPool.apply(xc, mask, *params, H, self.norm.eps, drop_p, drop_seed, S)
        with torch.autocast("cuda", enabled=False):
            W = self.attn.in_proj_weight.float()
            bias = self.attn.in_proj_bias.float()
            Wq, Wk, Wv = W[:D], W[D:2 * D], W[2 * D:]
            bq, bv = bias[:D], bias[2 * D:]
            q0 = F.linear(self.query.float().view(1, D), Wq, bq).view(H, Dh)
            qt = torch.einsum("hkd,hk->hd", Wk.view(H, Dh, D), q0) * (1.0 / math.sqrt(Dh))     # [H, D]
            xbar, sa, _ = _StreamPool.apply(x, qt, mask, drop_p, drop_seed, False)             # [B, H, D], [B, H] fp32
            o = torch.einsum("bhd,hkd->bhk", xbar, Wv.view(H, Dh, D))
            o = (o + bv.view(1, H, Dh) * sa.unsqueeze(-1) if drop_p > 0.0 else o + bv.view(1, H, Dh)).reshape(B, D)
            y = F.linear(o, self.attn.out_proj.weight.float(), self.attn.out_proj.bias.float())
            y = F.layer_norm(y, (D,), self.norm.weight.float(), self.norm.bias.float(), self.norm.eps)
            if isinstance(self.proj, nn.Linear):
                y = F.linear(y, self.proj.weight.float(), self.proj.bias.float())
        return y.to(x.dtype)


def token_mean_pool(token_feats: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
    """The mean branch of ``VideoEncoder._pool_video_tokens`` (models/video_encoder.py:603): ``token_feats.mean(dim=2)``,
    [B, N_views, L, D] -> [B, N_views, D] (also accepts [B, L, D] -> [B, D]).

    It is the streaming pool kernel in its uniform-weights mode: one head with a zero query makes every score 0, the
    softmax is exactly 1 / L, so x is read once, accumulated in fp32 and never re-materialised; the backward writes
    dx = dout / L in the same kernel family. With a key-padding ``mask`` ([.., L] bool, True = ignore) it is the mean over the
    unmasked tokens (the reference's branch has no mask)."""
    shape = token_feats.shape
    if token_feats.dim() not in (3, 4):
        raise ValueError(f"token_feats must be [B, N, L, D] or [B, L, D], got {tuple(shape)}")
    L, D = shape[-2], shape[-1]
    x = token_feats.reshape(-1, L, D)
    mk = mask.reshape(-1, L) if mask is not None else None
    qt = torch.zeros((1, D), dtype=torch.float32, device=x.device)
    xbar, _, _ = _StreamPool.apply(x, qt, mk, 0.0, 0, False)          # [B * N, 1, D] fp32
    return xbar[:, 0, :].reshape(*shape[:-2], D).to(token_feats.dtype)


def pool_video_tokens(encoder, token_feats: torch.Tensor) -> torch.Tensor:
    """Drop-in body of ``VideoEncoder._pool_video_tokens`` (models/video_encoder.py:589-603): [B, N, L, D] -> [B, N, D].
    The reference loops over the N views in Python and concatenates; here the views are folded into the batch — ONE pass
    of the pool kernel over all B * N views (same parameters for every view, so the result is identical) — and the
    ``attention_pool is None`` case is the uniform-weights mode (``token_mean_pool``). ``install(modules=True)`` binds it."""
    B, N, L, D = token_feats.shape
    pool = getattr(encoder, "attention_pool", None)
    if pool is not None:
        return pool(token_feats.reshape(B * N, L, D)).reshape(B, N, -1)
    return token_mean_pool(token_feats)


class AttentionPoolWithCLS(nn.Module):
    """models/attention_pool.py:104-197 (built by VideoEncoder for ``token_pooling_mode == "cls_token"``,
    models/video_encoder.py:214-219): a learnable CLS token is prepended, one post-LN ``nn.TransformerEncoderLayer``
    runs over [CLS; x] and only the CLS row is kept (:187-195). Only that row is computed here: its attention is the
    one-query pool over the N tokens (streaming kernel, x read once, never concatenated or projected to K / V) plus
    the CLS key itself, merged through the log-sum-exp of the streamed scores:

        w_c = exp(s_c - logaddexp(lse_x, s_c)),   attn = W_v ((1 - w_c) xbar + w_c cls) + b_v

    followed by out_proj, residual + norm1, the feed-forward block, norm2, the final norm and proj on [B, D] vectors.
    The reference computes all N + 1 rows of the layer (QKV projections 6 (N + 1) D^2 FLOP per sample and a dense
    [N + 1, N + 1] attention) and throws N of them away. ``num_layers`` must be 1, the only value the reference ever
    constructs (a deeper stack needs every row of the earlier layers)."""

    def __init__(self, embed_dim: int, num_heads: int = 8, num_layers: int = 1, output_dim: int = None,
                 dropout: float = 0.0):
        super().__init__()
        if num_layers != 1:
            raise ValueError("AttentionPoolWithCLS: only num_layers=1 is implemented (the reference never builds "
                             "another value: models/video_encoder.py:215-219)")
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.num_layers = num_layers
        self.output_dim = output_dim or embed_dim
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        # parameter container only (keeps the reference state-dict keys transformer.layers.0.*); never called
        self.transformer = nn.TransformerEncoder(
            nn.TransformerEncoderLayer(d_model=embed_dim, nhead=num_heads, dropout=dropout, batch_first=True),
            num_layers=num_layers)
        self.norm = nn.LayerNorm(embed_dim)
        self.proj = nn.Linear(embed_dim, self.output_dim) if self.output_dim != embed_dim else nn.Identity()
        self.dropout = dropout

    def forward(self, x: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
        B, N, D = x.shape
        assert D == self.embed_dim, f"Input dim {D} != expected {self.embed_dim}"
        layer = self.transformer.layers[0]
        train_drop = self.training and self.dropout > 0.0
        drop_p, drop_seed = 0.0, 0
        if train_drop:
            drop_p = float(self.dropout)
            drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        H, Dh = self.num_heads, D // self.num_heads
        with torch.autocast("cuda", enabled=False):
            W = layer.self_attn.in_proj_weight.float()
            bias = layer.self_attn.in_proj_bias.float()
            Wq, Wk, Wv = W[:D], W[D:2 * D], W[2 * D:]
            bq, bv = bias[:D], bias[2 * D:]
            c = self.cls_token.float().view(1, D)
            q0 = F.linear(c, Wq, bq).view(H, Dh)
            qt = torch.einsum("hkd,hk->hd", Wk.view(H, Dh, D), q0) * (1.0 / math.sqrt(Dh))      # [H, D]
            # the key bias adds the same q0 . b_k to every score, CLS included: it cancels in the softmax
            xbar, sa, lse = _StreamPool.apply(x, qt, mask, drop_p, drop_seed, True)             # [B,H,D], [B,H], [B,H]
            s_c = (qt * c).sum(dim=1)                                                           # [H] score of the CLS key
            tot = torch.logaddexp(lse, s_c.unsqueeze(0))                                        # [B, H]
            w_x = torch.exp(lse - tot)                                                          # weight of the N tokens
            w_c = torch.exp(s_c.unsqueeze(0) - tot)                                             # weight of the CLS key
            if train_drop:                                                                      # attention dropout:
                keep = (torch.rand(B, H, device=x.device) >= drop_p).to(w_c.dtype) / (1.0 - drop_p)  # the CLS key
                w_c = w_c * keep
                w_sum = w_x * sa + w_c
            mix = w_x.unsqueeze(-1) * xbar + w_c.unsqueeze(-1) * c.view(1, 1, D)                # [B, H, D]
            o = torch.einsum("bhd,hkd->bhk", mix, Wv.view(H, Dh, D))
            o = (o + bv.view(1, H, Dh) * w_sum.unsqueeze(-1) if train_drop else o + bv.view(1, H, Dh)).reshape(B, D)
            y = F.linear(o, layer.self_attn.out_proj.weight.float(), layer.self_attn.out_proj.bias.float())
            y = F.dropout(y, drop_p, train_drop)
            y = F.layer_norm(c + y, (D,), layer.norm1.weight.float(), layer.norm1.bias.float(), layer.norm1.eps)
            f = F.relu(F.linear(y, layer.linear1.weight.float(), layer.linear1.bias.float()))
            f = F.linear(F.dropout(f, drop_p, train_drop), layer.linear2.weight.float(), layer.linear2.bias.float())
            y = F.layer_norm(y + F.dropout(f, drop_p, train_drop), (D,), layer.norm2.weight.float(),
                             layer.norm2.bias.float(), layer.norm2.eps)
            y = F.layer_norm(y, (D,), self.norm.weight.float(), self.norm.bias.float(), self.norm.eps)
            if isinstance(self.proj, nn.Linear):
                y = F.linear(y, self.proj.weight.float(), self.proj.bias.float())
        return y.to(x.dtype)
