"""One-recompute CLIP backward (csrc/gt_gemm.cu + the G-tile stores of csrc/logits_bwd3.cu): the transposed product alone
against a float64 matmul of the same bf16 operands, and the loss gradients with and without the stored-G path
(B200CLIP_GSTORE=0 = two passes, each with its own recompute of the logits)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("Nx,Ny,D", [(1000, 777, 200), (4097, 4099, 512), (64, 300, 256), (3000, 520, 500), (1000, 777, 700), (2050, 260, 768)])
def test_gt_gemm_matches_float64_matmul(Nx, Ny, D):
    """dY += dyn[2] / gnorm * G^T X on ragged shapes (rows / columns that are no multiples of the 256 x 64 tiles, D below the
    padded width), accumulating into a non-zero dY."""
    from deepcoro_clip_b200 import ops
    dev = torch.device("cuda", 0)
    torch.manual_seed(Nx + Ny)
    Dp = ops.round_up(D, 256)
    Gd = torch.randn(Nx, Ny, device=dev).bfloat16()
    # the blocked layout of include/b200clip.h K3b: [64 x 64] blocks, 16-byte unit u of row r at position u ^ (r & 7)
    nib, njb = 2 * ((Nx + 127) // 128), 4 * ((Ny + 255) // 256)
    assert ops.gstore_elems(Nx, Ny) == nib * njb * 4096
    pad = torch.zeros(nib * 64, njb * 64, dtype=torch.bfloat16, device=dev)
    pad[:Nx, :Ny] = Gd
    blocks = pad.view(nib, 64, njb, 8, 8).permute(0, 2, 1, 3, 4)                 # [ib, jb, r, u, 8]
    r = torch.arange(64, device=dev).view(64, 1)
    pos = torch.arange(8, device=dev).view(1, 8) ^ (r & 7)                        # unit stored at position p is u = p ^ (r & 7)
    G = blocks[:, :, r, pos, :].contiguous().view(-1)
    X = torch.zeros(Nx, Dp, dtype=torch.bfloat16, device=dev)
    X[:, :D] = torch.randn(Nx, D, device=dev).bfloat16()
    dyn = torch.zeros(16, device=dev)
    dyn[2] = 17.0
    gnorm = 4.0
    dY0 = torch.randn(Ny, D, device=dev)
    dY = dY0.clone()
    ops.call("gt_gemm", G, G.numel(), Nx, Ny, X, X.stride(0), Dp, D, dyn, gnorm, dY, dY.stride(0), ops.stream_ptr(dev))
    want = dY0.double() + (17.0 / gnorm) * Gd.double().t() @ X[:, :D].double()
    assert _rel(dY, want) <= 2e-6


@pytest.mark.parametrize("N,D,kw", [(1500, 512, {}), (777, 256, {}), (2048, 512, {"label_smoothing": 0.1}),
                                    (1029, 512, {"gated": True}), (640, 512, {"tau": 0.004}), (1500, 768, {}), (900, 768, {"label_smoothing": 0.05}),
                                    (385, 768, {"tau": 0.004})])
def test_clip_gradients_one_recompute_vs_two_passes(N, D, kw):
    from deepcoro_clip_b200 import loss as L
    dev = torch.device("cuda", 0)
    kw = dict(kw)
    tau = kw.pop("tau", 0.07)
    res = {}
    saved = os.environ.get("B200CLIP_GSTORE")
    try:
        for mode in ("1", "0"):
            os.environ["B200CLIP_GSTORE"] = mode
            torch.manual_seed(3)
            v = torch.randn(N, D, device=dev, requires_grad=True)
            t = torch.randn(N, D, device=dev, requires_grad=True)
            lt = torch.tensor([math.log(tau)], device=dev, requires_grad=True)
            out = L.clip_loss(v, t, lt, precision="bf16", **kw)
            out.backward()
            res[mode] = (out.detach(), v.grad, t.grad, lt.grad)
    finally:
        if saved is None:
            os.environ.pop("B200CLIP_GSTORE", None)
        else:
            os.environ["B200CLIP_GSTORE"] = saved
    assert torch.equal(res["1"][0], res["0"][0])
    assert _rel(res["1"][1], res["0"][1]) <= 1e-5            # same kernel, same G: only the atomics' order differs
    assert _rel(res["1"][2], res["0"][2]) <= 1e-5            # G^T V from the stored bf16 tiles vs the recomputed ones
    assert _rel(res["1"][3], res["0"][3]) <= 1e-5
