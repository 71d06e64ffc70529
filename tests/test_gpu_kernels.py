"""Kernel-level parity through the C ABI (ctypes), against float64 torch on the same bf16 operands: the logits forward
(row / column sums, captured diagonal) and the logits backward in all three modes, on shapes that exercise the tile
edges of BOTH kernel families (CTA pairs: odd tile counts, Dp = 128 / 256 / 384 / 512, several segment counts;
single-CTA: Dp not a multiple of 128, Kp > 512)."""
import math

import pytest
import torch

from deepcoro_clip_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOG2E = math.log2(math.e)


def _operands(Nx, Ny, D, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    Dp = (D + 63) // 64 * 64
    x = torch.zeros(Nx, Dp, device=DEV); y = torch.zeros(Ny, Dp, device=DEV)
    x[:, :D] = torch.nn.functional.normalize(torch.randn(Nx, D, device=DEV, generator=g), dim=-1)
    y[:, :D] = torch.nn.functional.normalize(torch.randn(Ny, D, device=DEV, generator=g), dim=-1)
    return x.bfloat16(), y.bfloat16(), Dp


@pytest.mark.parametrize("Nx,Ny,D,nseg", [(333, 200, 128, 0), (1100, 700, 384, 3), (128, 4000, 512, 0), (2049, 257, 256, 2),
                                          (640, 640, 192, 0), (300, 900, 768, 0), (129, 129, 64, 1), (4100, 1500, 512, 5)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_logits_bwd_all_modes_and_edges(Nx, Ny, D, nseg, mode):
    tau, bias, wneg_c = 0.07, -2.0, 0.41
    x, y, Dp = _operands(Nx, Ny, D, Nx + Ny + D)
    S = x.double() @ y.double().T
    rs = (torch.rand(Nx, device=DEV) * 0.5 + 0.5); cs = (torch.rand(Ny, device=DEV) * 0.5 + 0.5)
    dX = torch.zeros(Nx, D, device=DEV)
    scal = torch.zeros(4, device=DEV, dtype=torch.float64)
    L.call("logits_bwd", mode, x, y, Nx, Ny, Dp, Dp, D, 0, Dp, Dp, LOG2E / tau, LOG2E / tau, 1.0 / tau, bias, wneg_c, rs, cs,
           1.0 / tau, 1.0, 0, None, 0.0, 0, None, dX, D, scal, nseg, L.stream_ptr())
    torch.cuda.synchronize()
    if mode == 0:
        f = S; G = torch.exp((f - 1.0) / tau) * (rs.double()[:, None] + cs.double()[None, :]); GS = G
    elif mode == 1:
        sg = torch.sigmoid(S); f = S * sg; fp = sg * (1 + S * (1 - sg))
        G = torch.exp((f - 1.0) / tau) * (rs.double()[:, None] + cs.double()[None, :]); GS = G * fp
    else:
        R = S / tau + bias; Lc = R.clamp(-30, 30)
        G = wneg_c * torch.sigmoid(Lc) * (R.abs() <= 30); GS = G; f = S
    ref = (GS @ y.double()[:, :D]) / tau
    refq = (GS.float().bfloat16().double() @ y.double()[:, :D]) / tau       # same bf16 rounding of the gradient operand
    assert ((dX.double() - refq).norm() / refq.norm()).item() <= 2e-4
    assert ((dX.double() - ref).norm() / ref.norm()).item() <= 2e-3         # north_star gradient tolerance
    assert abs(scal[0].item() / (G * f).sum().item() - 1) <= 1e-5
    if mode == 2:
        sp = torch.nn.functional.softplus(Lc)
        assert abs(scal[1].item() / sp.sum().item() - 1) <= 1e-5
        assert abs(scal[2].item() / G.sum().item() - 1) <= 1e-5


@pytest.mark.parametrize("Ma,Nb,D", [(333, 200, 128), (1000, 1300, 512), (129, 4100, 256), (2049, 257, 768), (4096, 4096, 512)])
@pytest.mark.parametrize("gated", [0, 1])
def test_logits_lse_fwd_sums_and_diag(Ma, Nb, D, gated):
    tau = 0.0588
    a, b, Kp = _operands(Ma, Nb, D, Ma * 7 + Nb)
    rowsum = torch.zeros(Ma, device=DEV); colsum = torch.zeros(Nb, device=DEV)
    nd = min(Ma, Nb)
    diag = torch.zeros(Ma, device=DEV)
    bound = 0.7310585786300049 if gated else 1.0
    scale2 = LOG2E / tau
    L.call("logits_lse_fwd", a, b, Ma, Nb, Kp, Kp, Kp, scale2, scale2 * bound, gated, None, rowsum, colsum, diag, 0,
           L.stream_ptr())
    torch.cuda.synchronize()
    S = a.double() @ b.double().T
    f = S * torch.sigmoid(S) if gated else S
    P = torch.exp((f - bound) / tau)
    assert ((rowsum.double() - P.sum(1)).abs() / P.sum(1)).max().item() <= 2e-5
    assert ((colsum.double() - P.sum(0)).abs() / P.sum(0)).max().item() <= 2e-5
    assert (diag[:nd].double() - S.diagonal()[:nd]).abs().max().item() <= 1e-6
