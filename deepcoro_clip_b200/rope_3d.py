"""Drop-in for models/rope_3d.py (reference :47-252, :255-282): same constructor, attributes, forward signature,
CLS auto-detection and "return the inputs unchanged on a shape mismatch" behaviour — the rotation itself (forward and
backward, q and k together) is one sm_100a kernel launch instead of ~30 elementwise kernels.

Tables: built per (T, H, W, device, dtype, n_special) IN THE TENSOR DTYPE with the same sequence of roundings as the
reference (angles = position * base^(-2i/d) rounded to dtype, then cos / sin), cached when not training."""
from __future__ import annotations

from typing import Literal, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import DTYPE_CODE, call, i64, stream_ptr


def _axis_angles(dim: int, length: int, base: float, device, dtype) -> torch.Tensor:
    """[length, dim] angles, each frequency repeated for the (even, odd) channel pair."""
    exponent = torch.arange(0, dim, 2, device=device, dtype=dtype) / dim
    inv_freq = 1.0 / (base ** exponent)
    pos = torch.arange(length, device=device, dtype=dtype)
    ang = torch.outer(pos, inv_freq)
    return ang.repeat_interleave(2, dim=1)


class _RopeApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, sin, cos):
        ctx.save_for_backward(sin, cos)
        return _launch(q, k, sin, cos, backward=False)

    @staticmethod
    def backward(ctx, gq, gk):
        sin, cos = ctx.saved_tensors
        dq, dk = _launch(gq, gk, sin, cos, backward=True)
        return dq, dk, None, None


def _launch(q: torch.Tensor, k: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor, backward: bool):
    ops.require_cuda(q, k)
    if q.dtype not in DTYPE_CODE or k.dtype != q.dtype:
        raise TypeError(f"rope: unsupported dtypes {q.dtype} / {k.dtype}")
    B, Hh, N, Dh = q.shape

    def prep(t):
        if t.stride(3) != 1:
            t = t.contiguous()
        return t

    q, k = prep(q), prep(k)
    if k.shape != q.shape:
        # different token counts (pooled K): two independent launches
        qo = _launch_one(q, sin, cos, backward)
        ko = _launch_one(k, sin, cos, backward)
        return qo, ko
    sin = sin.reshape(-1, Dh).to(q.dtype).contiguous()
    cos = cos.reshape(-1, Dh).to(q.dtype).contiguous()
    qo = torch.empty((B, Hh, N, Dh), dtype=q.dtype, device=q.device)
    ko = torch.empty((B, Hh, N, Dh), dtype=q.dtype, device=q.device)
    call("rope3d_apply", q, i64(q.stride(0)), i64(q.stride(1)), i64(q.stride(2)), qo, k, i64(k.stride(0)),
         i64(k.stride(1)), i64(k.stride(2)), ko, sin, cos, DTYPE_CODE[q.dtype], B, Hh, N, Dh, int(backward),
         stream_ptr(q.device))
    return qo, ko


def _launch_one(x, sin, cos, backward):
    B, Hh, N, Dh = x.shape
    sin = sin.reshape(-1, Dh).to(x.dtype).contiguous()
    cos = cos.reshape(-1, Dh).to(x.dtype).contiguous()
    out = torch.empty((B, Hh, N, Dh), dtype=x.dtype, device=x.device)
    call("rope3d_apply", x, i64(x.stride(0)), i64(x.stride(1)), i64(x.stride(2)), out, None, i64(0), i64(0), i64(0),
         None, sin, cos, DTYPE_CODE[x.dtype], B, Hh, N, Dh, int(backward), stream_ptr(x.device))
    return out


class Rope3D(nn.Module):
    """3D axial RoPE: head_dim split into T / H / W thirds, rotary embedding per axis (reference :47-252)."""

    def __init__(self, embed_dim: int, num_heads: int, *, temporal_base: float = 10000.0,
                 spatial_base: float = 10000.0, temporal_scale: float = 1.0,
                 normalize_mode: Literal["separate", "max", "min"] = "separate",
                 device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None):
        super().__init__()
        assert embed_dim % num_heads == 0, f"embed_dim ({embed_dim}) must be divisible by num_heads ({num_heads})"
        self.head_dim = embed_dim // num_heads
        if self.head_dim % 6 != 0:
            raise ValueError(f"For 3D RoPE, head_dim ({self.head_dim}) must be divisible by 6. "
                             f"Got embed_dim={embed_dim}, num_heads={num_heads}")
        self.num_heads = num_heads
        self.temporal_base = temporal_base
        self.spatial_base = spatial_base
        self.temporal_scale = temporal_scale
        self.normalize_mode = normalize_mode          # stored, never used by the reference math (:92)
        self.t_dim = self.head_dim // 3
        self.h_dim = self.head_dim // 3
        self.w_dim = self.head_dim - self.t_dim - self.h_dim
        assert self.t_dim % 2 == 0 and self.h_dim % 2 == 0 and self.w_dim % 2 == 0
        self._cache = {}

    @torch.no_grad()
    def _get_cached_freqs(self, T: int, H: int, W: int, device, dtype, n_special: int = 0):
        key = (T, H, W, device, dtype, n_special)
        if not self.training and key in self._cache:
            return self._cache[key]
        ta = _axis_angles(self.t_dim, T, self.temporal_base * self.temporal_scale, device, dtype)
        ha = _axis_angles(self.h_dim, H, self.spatial_base, device, dtype)
        wa = _axis_angles(self.w_dim, W, self.spatial_base, device, dtype)
        cos = torch.cat([ta.cos().view(T, 1, 1, -1).expand(T, H, W, -1), ha.cos().view(1, H, 1, -1).expand(T, H, W, -1),
                         wa.cos().view(1, 1, W, -1).expand(T, H, W, -1)], dim=-1).reshape(T * H * W, self.head_dim)
        sin = torch.cat([ta.sin().view(T, 1, 1, -1).expand(T, H, W, -1), ha.sin().view(1, H, 1, -1).expand(T, H, W, -1),
                         wa.sin().view(1, 1, W, -1).expand(T, H, W, -1)], dim=-1).reshape(T * H * W, self.head_dim)
        if n_special > 0:
            cos = torch.cat([torch.ones((n_special, self.head_dim), device=device, dtype=dtype), cos], dim=0)
            sin = torch.cat([torch.zeros((n_special, self.head_dim), device=device, dtype=dtype), sin], dim=0)
        sin, cos = sin.contiguous(), cos.contiguous()
        if not self.training:
            self._cache[key] = (sin, cos)
        return sin, cos

    def forward(self, q: torch.Tensor, k: torch.Tensor, T: int, H: int, W: int,
                n_special: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        B, Hh, N, Dh = q.shape
        assert Dh == self.head_dim, f"Expected head_dim={self.head_dim}, got {Dh}"
        expected = T * H * W
        if n_special == 0 and N == expected + 1:
            n_special = 1
        if N != n_special + expected:
            return q, k                                   # pooling stages: leave untouched (reference :218-221)
        sin, cos = self._get_cached_freqs(T, H, W, q.device, q.dtype, n_special)
        return _RopeApply.apply(q, k, sin, cos)


def apply_rope_qk(q: torch.Tensor, k: torch.Tensor, sin: torch.Tensor, cos: torch.Tensor):
    """Reference :255-282 — apply precomputed sin/cos ([N, Dh] or [1, 1, N, Dh]) to q and k."""
    return _RopeApply.apply(q, k, sin, cos)
