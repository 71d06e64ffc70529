"""Thin tensor-level wrappers over the C ABI (include/b200clip.h). No math happens in Python/PyTorch here:
torch only provides device buffers, the current stream and (in dist.py) the NCCL process group."""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import DTYPE_CODE, call, i64, stream_ptr, try_call

LOG2E = math.log2(math.e)
GATED_BOUND = 0.7310585786300049  # max of s*sigmoid(s) on [-1, 1] (at s = 1)


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.B200ClipError(
                "deepcoro_clip_b200 runs on sm_100a GPUs only: got a CPU tensor and there is no CPU fallback")
        dev = t.device
    _lib.lib()  # fail loudly if the native library is absent
    return dev


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _rowmajor(x: torch.Tensor) -> torch.Tensor:
    if x.dim() != 2:
        raise ValueError(f"expected a 2-D [rows, dim] tensor, got shape {tuple(x.shape)}")
    if x.dtype not in DTYPE_CODE:
        x = x.float()
    if x.stride(1) != 1:
        x = x.contiguous()
    return x


def l2norm_operand(x: torch.Tensor, split3_role: int = -1, normalize: bool = True):
    """x [rows, dim] -> (bf16 operand [rows, ld], inv_norm [rows] fp32, Kp). split3_role 0/1 = bf16x3 panels.
    normalize=False packs the raw features (inv_norm then holds ||x||)."""
    x = _rowmajor(x)
    rows, dim = x.shape
    Kp = round_up(dim, 64)
    ld = 3 * Kp if split3_role >= 0 else Kp
    op = torch.empty((rows, ld), dtype=torch.bfloat16, device=x.device)
    inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
    call("l2norm_fwd", x, DTYPE_CODE[x.dtype], i64(x.stride(0)), rows, dim, op, ld, Kp, split3_role, inv, None, 0,
         int(normalize), stream_ptr(x.device))
    return op, inv, Kp


def l2norm_backward(dxhat, x, inv_norm, *, other_x=None, other_inv=None, other_hi=None, diag_corr=None, usum=None,
                    gscale=1.0, ucoef=0.0, dev_omul=None, dev_gmul=None):
    x = _rowmajor(x)
    rows, dim = x.shape
    dx = torch.empty((rows, dim), dtype=torch.float32, device=x.device)
    if other_x is not None:
        other_x = _rowmajor(other_x)
    call("l2norm_bwd", dxhat, dxhat.stride(0), x, DTYPE_CODE[x.dtype], i64(x.stride(0)), inv_norm, other_x,
         DTYPE_CODE[other_x.dtype] if other_x is not None else 0, i64(other_x.stride(0) if other_x is not None else 0),
         other_inv, other_hi, other_hi.stride(0) if other_hi is not None else 0, diag_corr, usum, float(gscale),
         float(ucoef), dev_omul, dev_gmul, rows, dim, dx, i64(dim), stream_ptr(x.device))
    return dx


def _scalar_f32(t: torch.Tensor) -> torch.Tensor:
    """First element of ``t`` as an fp32 device tensor without copies in the common case (fp32 0-d / [1] parameter)."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.numel() == 1 else t.reshape(-1)[:1].contiguous()


def dyn_prep(log_temp: torch.Tensor, bias, clamp_min: float, bound: float) -> torch.Tensor:
    lt = _scalar_f32(log_temp)
    b = None if bias is None else _scalar_f32(bias)
    dyn = torch.empty(16, dtype=torch.float32, device=lt.device)
    call("dyn_prep", lt, b, float(clamp_min), float(bound), dyn, stream_ptr(lt.device))
    return dyn


def lse_fwd(A, B, Ma, Nb, K, dyn, gated, rowsum, colsum, diag=None, diag_off=0, diag_corr=None, gnorm=1.0, hp=False):
    call("logits_lse_fwd", A, B, Ma, Nb, K, A.stride(0), B.stride(0), 0.0, 0.0, int(gated), dyn, 0, rowsum, colsum,
         diag, int(diag_off), stream_ptr(A.device))


def vec_fsum(v, gated, acc_slot):
    call("vec_fsum", v, v.numel(), int(gated), acc_slot, stream_ptr(v.device))


def lse_finalize(sums, dyn, c, scale_out, acc_slot):
    call("lse_finalize", sums, sums.numel(), dyn, float(c), scale_out, acc_slot, stream_ptr(sums.device))


def diag_sum(a, b, rows, K, gated, dots, acc_slot):
    call("diag_sum", a, a.stride(0), b, b.stride(0), rows, K, int(gated), dots, acc_slot, stream_ptr(a.device))


def colsum_bf16(op, rows, dim, out=None):
    """out[c] (+)= sum_r op[r, c]; pass ``out`` to accumulate a second panel into the same sums."""
    if out is None:
        out = torch.zeros(dim, dtype=torch.float32, device=op.device)
    call("colsum_bf16", op, op.stride(0), rows, dim, out, stream_ptr(op.device))
    return out


def logits_bwd(mode, X, Y, Nx, Ny, K, Dp, D, dyn, rowscale, colscale, dX, scal, *, wneg_c=0.0, nseg=0, ydiag=0.0,
               diag_off=0, diag_corr=None, gnorm=1.0, hp=False):
    call("logits_bwd", mode, X, Y, Nx, Ny, K, Dp, D, K - Dp, X.stride(0), Y.stride(0), 0.0, 0.0, 0.0, 0.0, float(wneg_c),
         rowscale, colscale, 0.0, float(gnorm), int(hp), dyn, float(ydiag), int(diag_off), diag_corr, dX, dX.stride(0), scal, nseg, stream_ptr(X.device))



def gstore_elems(Nx: int, Ny: int) -> int:
    """bf16 elements of the blocked gradient-tile buffer of ``logits_bwd_both`` (8 KB per [64 x 64] block, include/b200clip.h K3b)."""
    return 2 * ((Nx + 127) // 128) * 4 * ((Ny + 255) // 256) * 4096


def logits_bwd_both(mode, X, Y, Nx, Ny, K, D, dyn, rowscale, colscale, dX, dY, scal, G, *, ydiag=0.0, diag_off=0, diag_corr=None,
                    gnorm=1.0, wneg_c=0.0) -> bool:
    """dX += G Y and dY += G^T X from one recompute of the logits (G tiles kept in the caller's flat bf16 buffer ``G`` of
    ``gstore_elems(Nx, Ny)`` elements). False when the shape does not qualify (nothing launched)."""
    return try_call("logits_bwd_both", mode, X, Y, Nx, Ny, K, K, D, X.stride(0), Y.stride(0), float(wneg_c), rowscale, colscale,
                    float(gnorm), dyn, float(ydiag), int(diag_off), diag_corr, dX, dX.stride(0), dY, dY.stride(0), scal, G,
                    G.numel(), stream_ptr(X.device))


call = call  # re-export for the loss module
i64 = i64
stream_ptr = stream_ptr
