"""Fused transformer block of EnhancedVideoAggregator (csrc/xfblock.cu) against the block evaluated with
PyTorch ops in float64 — the reference's own module structure (models/video_aggregator.py:7-54: LayerNorm,
nn.MultiheadAttention, Linear-GELU-Linear, residuals) on the same parameters: output, dx and every parameter gradient at
2e-5; one forward + five backward library launches; dropout checked by a directional finite difference with a fixed seed."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture
def env():
    saved = os.environ.get("B200CLIP_XFBLOCK")
    yield
    if saved is None:
        os.environ.pop("B200CLIP_XFBLOCK", None)
    else:
        os.environ["B200CLIP_XFBLOCK"] = saved


@pytest.mark.parametrize("B,N,D,H,masked", [(3, 4, 512, 4, False), (5, 4, 512, 4, True), (4, 7, 512, 8, True),
                                            (2, 15, 512, 4, True), (3, 3, 256, 4, False), (2, 11, 384, 8, False),
                                            (5, 1, 128, 8, False), (2, 16, 512, 2, True), (9, 2, 512, 1, False)])
def test_fused_block_matches_float64_pytorch_block(env, B, N, D, H, masked):
    from deepcoro_clip_b200 import _lib
    from deepcoro_clip_b200.video_aggregator import TransformerBlock
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    blk = TransformerBlock(D, H, 0.1).to(dev).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(B, N, D, device=dev)
    mask = None
    if masked:
        mask = torch.rand(B, N, device=dev) < 0.3
        mask[:, 0] = False
    g = torch.randn(B, N, D, device=dev)
    os.environ["B200CLIP_XFBLOCK"] = "1"
    xr = x.clone().requires_grad_(True)
    l0 = _lib.LAUNCHES
    y = blk(xr, mask)
    y.backward(g)
    torch.cuda.synchronize()
    assert _lib.LAUNCHES - l0 == 6
    os.environ["B200CLIP_XFBLOCK"] = "0"
    ref = copy.deepcopy(blk).double()
    ref.zero_grad(set_to_none=True)
    xd = x.double().requires_grad_(True)
    yr = ref(xd, mask)
    yr.backward(g.double())
    assert _rel(y, yr) <= 2e-5 and _rel(xr.grad, xd.grad) <= 2e-5
    for (k, p), (_, r) in zip(blk.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad, r.grad) <= 2e-5, k


def test_fused_block_dropout_gradient_is_consistent(env):
    from deepcoro_clip_b200.video_aggregator import TransformerBlock
    dev = torch.device("cuda", 0)
    os.environ["B200CLIP_XFBLOCK"] = "1"
    torch.manual_seed(3)
    blk = TransformerBlock(512, 4, 0.2).to(dev).train()
    x = torch.randn(4, 4, 512, device=dev)
    g = torch.randn(4, 4, 512, device=dev)
    dirx = torch.randn_like(x)

    def f(xx):
        torch.manual_seed(11)              # fixes the dropout seed the module draws
        return (blk(xx) * g).sum()
    xr = x.clone().requires_grad_(True)
    f(xr).backward()
    ana = (xr.grad * dirx).sum().item()
    h = 2e-3
    num = (f(x + h * dirx).item() - f(x - h * dirx).item()) / (2 * h)
    assert abs(ana - num) <= 2e-2 * max(abs(num), 1.0), (ana, num)
    with torch.no_grad():
        torch.manual_seed(11)
        y1 = blk(x)
        y0 = blk.eval()(x)
    assert ((y1 - y0).abs() > 1e-6).float().mean().item() > 0.5       # dropout really acts


def test_aggregator_with_fused_blocks_matches_pytorch_blocks(env):
    from deepcoro_clip_b200 import EnhancedVideoAggregator
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    agg = EnhancedVideoAggregator(512).to(dev).eval()
    xa = torch.randn(8, 4, 512, device=dev)
    mask = torch.zeros(8, 4, dtype=torch.bool, device=dev)
    mask[1, 3] = True
    mask[5, 1:] = True
    ga = torch.randn(8, 512, device=dev)
    res = {}
    for mode in ("1", "0"):
        os.environ["B200CLIP_XFBLOCK"] = mode
        agg.zero_grad(set_to_none=True)
        xr = xa.clone().requires_grad_(True)
        out = agg(xr, mask)
        out.backward(ga)
        res[mode] = (out.detach(), xr.grad, {k: v.grad.clone() for k, v in agg.named_parameters() if v.grad is not None})
    assert _rel(res["1"][0], res["0"][0]) <= 1e-4 and _rel(res["1"][1], res["0"][1]) <= 1e-4
    for k, v in res["0"][2].items():
        if v.norm() > 0:
            assert _rel(res["1"][2][k], v) <= 1e-4, k


@pytest.mark.parametrize("train", [False, True])
def test_whole_aggregator_node_equals_per_block_nodes(env, train):
    """The single autograd node over positional add + blocks + tail launches the same kernels as the per-module path
    (B200CLIP_AGG_FUSED=0): same output bit for bit, gradients up to the atomics of the tail; with dropout the same seeds
    are drawn in the same order."""
    from deepcoro_clip_b200 import EnhancedVideoAggregator
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    agg = EnhancedVideoAggregator(512, dropout=0.2, aggregator_depth=3).to(dev).train(train)
    xa = torch.randn(6, 5, 512, device=dev)
    mask = torch.zeros(6, 5, dtype=torch.bool, device=dev)
    mask[2, 3:] = True
    ga = torch.randn(6, 512, device=dev)
    res = {}
    saved = os.environ.get("B200CLIP_AGG_FUSED")
    try:
        for mode in ("1", "0"):
            os.environ["B200CLIP_AGG_FUSED"] = mode
            agg.zero_grad(set_to_none=True)
            xr = xa.clone().requires_grad_(True)
            torch.manual_seed(11)
            out = agg(xr, mask)
            out.backward(ga)
            res[mode] = (out.detach(), xr.grad, {k: v.grad.clone() for k, v in agg.named_parameters() if v.grad is not None})
    finally:
        if saved is None:
            os.environ.pop("B200CLIP_AGG_FUSED", None)
        else:
            os.environ["B200CLIP_AGG_FUSED"] = saved
    assert torch.equal(res["1"][0], res["0"][0])
    assert _rel(res["1"][1], res["0"][1]) <= 1e-6
    assert set(res["1"][2]) == set(res["0"][2])
    for k, v in res["0"][2].items():
        assert res["1"][2][k].shape == v.shape
        if v.norm() > 0:
            assert _rel(res["1"][2][k], v) <= 1e-6, k
