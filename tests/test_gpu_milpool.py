"""Gated-attention MIL pooling (csrc/milpool.cu through deepcoro_clip_b200.mil_pooling) against the goldens generated from
the reference's MultiInstanceLinearProbing._pool_instances (models/multi_instance_linear_probing.py:493-536) and against
the float64 oracle at larger shapes. Tolerances: 2e-5 relative (forward), 5e-5 (gradients) for the fp32 FMA tiles (every
golden; < 1024 rows or D not in {256, 512, 768}); 1e-4 / 2e-4 for the tensor-core variant (>= 1024 rows, D in {256, 512, 768}), whose products carry
16 mantissa bits per operand (hi + lo bf16, lo*lo dropped: ~4e-6 relative per pre-activation, amplified by |w| sqrt(hidden)
on the way to the softmax — 2.4e-5 expected at hidden = 512 with unit-variance w; B200CLIP_MIL_TC=0 keeps fp32 products)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
CASES = ["milpool_b5_n6_d64_h32_mask", "milpool_b3_n4_d128_h128", "milpool_b2_n3_l20_d64_h24_mask"]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _module(V, bV, U, bU, w, bw, dropout=0.0):
    from deepcoro_clip_b200 import GatedAttentionPooling
    mod = GatedAttentionPooling(V.shape[1], V.shape[0], dropout).cuda()
    with torch.no_grad():
        for prm, val in ((mod.attention_V.weight, V), (mod.attention_V.bias, bV), (mod.attention_U.weight, U),
                         (mod.attention_U.bias, bU), (mod.attention_w.weight, w), (mod.attention_w.bias, bw)):
            prm.copy_(torch.as_tensor(np.asarray(val), dtype=torch.float32).reshape(prm.shape))
    return mod


def _check(mod, x, mask, go, want, tol_f=2e-5, tol_g=5e-5):
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    mk = None if mask is None else torch.tensor(mask, device="cuda")
    out = mod(xt, mk)
    out.backward(torch.tensor(go, dtype=torch.float32, device="cuda"))
    assert _rel(out.detach().cpu().numpy(), want["out"]) <= tol_f
    got = {"x": xt.grad, "V": mod.attention_V.weight.grad, "bV": mod.attention_V.bias.grad, "U": mod.attention_U.weight.grad,
           "bU": mod.attention_U.bias.grad, "w": mod.attention_w.weight.grad, "bw": mod.attention_w.bias.grad}
    for k, v in got.items():
        if k == "bw":       # the softmax is shift-invariant: d/d(bw) is exactly 0 in the reference, rounding noise here
            assert abs(float(v)) <= 1e-5 * max(1.0, float(np.abs(want["w"]).max())), k
        else:
            assert _rel(v.cpu().numpy(), want[k]) <= tol_g, k


@pytest.mark.parametrize("name", CASES)
def test_golden(name):
    g = np.load(GOLD / f"{name}.npz")
    mod = _module(g["V"], g["bV"], g["U"], g["bU"], g["w"], g["bw"])
    want = {"out": g["out"], "x": g["dx"], **{k: g["g_" + k] for k in ("V", "bV", "U", "bU", "w", "bw")}}
    _check(mod, g["x"], g["mask"] if bool(g["has_mask"]) else None, g["go"], want)


@pytest.mark.parametrize("shape,hd,masked", [((7, 300, 128), 64, True), ((300, 9, 512), 128, True), ((2, 3, 700, 256), 72, True),
                                             ((3, 1, 48), 8, False), ((5, 333, 768), 40, True), ((9, 130, 512), 512, False),
                                             ((1, 1100, 256), 8, True)])
def test_against_oracle(shape, hd, masked):
    """Ragged tiles (rows, hidden units and columns that are no multiples of the tile sizes), the split pooling pass
    (L >= 256), many sequences, a single instance; >= 1024 rows with D in {256, 512, 768} run the tensor-core variant
    (split-precision operands), the others the fp32 FMA tiles."""
    from oracle import token_oracle as to
    rng = np.random.default_rng(5)
    D = shape[-1]
    x = rng.standard_normal(shape).astype(np.float32)
    V, U = (rng.standard_normal((hd, D)) / np.sqrt(D)).astype(np.float32), (rng.standard_normal((hd, D)) / np.sqrt(D)).astype(np.float32)
    bV, bU = 0.3 * rng.standard_normal(hd).astype(np.float32), 0.3 * rng.standard_normal(hd).astype(np.float32)
    w, bw = rng.standard_normal((1, hd)).astype(np.float32), rng.standard_normal(1).astype(np.float32)
    mask = None
    if masked:
        mask = rng.random(shape[:2]) > 0.3
        mask[:, 0] = True
    go = rng.standard_normal((shape[0], D)).astype(np.float32)
    if len(shape) == 4:
        out, c = to.mil_hierarchical_pool_forward(x, mask, V, bV, U, bU, w, bw, want_cache=True)
        want = to.mil_hierarchical_pool_backward(go, c)
    else:
        out, c = to.mil_gated_pool_forward(x, mask, V, bV, U, bU, w, bw, want_cache=True)
        want = to.mil_gated_pool_backward(go, c)
    want["out"] = out
    rows = int(np.prod(shape[:-1]))
    tc = rows >= 1024 and D in (256, 512, 768)
    _check(_module(V, bV, U, bU, w, bw), x, mask, go, want, *((1e-4, 2e-4) if tc else (2e-5, 5e-5)))


def test_strided_rows_and_empty_sequence():
    """A view with a row pitch (no copy) gives the same result; a sequence without a valid instance is NaN like the
    reference's softmax over an all -inf row, the others are unaffected."""
    from deepcoro_clip_b200 import GatedAttentionPooling
    torch.manual_seed(0)
    mod = GatedAttentionPooling(64, 16).cuda()
    big = torch.randn(4, 5, 128, device="cuda")
    x = big[:, :, :64]
    a, b = mod(x), mod(x.contiguous())
    assert torch.equal(a, b)
    mask = torch.ones(4, 5, dtype=torch.bool, device="cuda")
    mask[2] = False
    c = mod(x, mask)
    assert torch.isnan(c[2]).all() and torch.equal(c[[0, 1, 3]], a[[0, 1, 3]])


def test_dropout_directional_derivative():
    """Training-mode dropout of the attention weights: same seed -> same mask; the analytic gradient matches a central
    difference along a random direction; the mask actually changes the output."""
    from deepcoro_clip_b200 import GatedAttentionPooling
    torch.manual_seed(1)
    mod = GatedAttentionPooling(64, 32, dropout=0.3).cuda().train()
    x = torch.randn(6, 12, 64, device="cuda")
    go, dirx = torch.randn(6, 64, device="cuda"), torch.randn(6, 12, 64, device="cuda")

    def f(xx):
        torch.manual_seed(9)
        return (mod(xx) * go).sum()
    xr = x.clone().requires_grad_(True)
    f(xr).backward()
    ana = (xr.grad * dirx).sum().item()
    h = 1e-2
    num = (f(x + h * dirx).item() - f(x - h * dirx).item()) / (2 * h)
    assert abs(ana - num) <= 2e-2 * max(abs(num), 1.0)
    with torch.no_grad():
        torch.manual_seed(9)
        y1 = mod(x)
        y0 = mod.eval()(x)
    assert (y1 - y0).abs().max().item() > 1e-3


def test_requires_cuda_and_shapes():
    from deepcoro_clip_b200 import GatedAttentionPooling
    mod = GatedAttentionPooling(64, 16)
    with pytest.raises(Exception):
        mod(torch.randn(2, 3, 64))
    mod = mod.cuda()
    with pytest.raises(ValueError):
        mod(torch.randn(2, 3, 64, device="cuda"), torch.ones(2, 4, dtype=torch.bool, device="cuda"))
    with pytest.raises(ValueError):
        GatedAttentionPooling(60, 16).cuda()(torch.randn(2, 3, 60, device="cuda"))


def test_tensor_core_variant_matches_fma_tiles():
    """Same module, same inputs through both variants (B200CLIP_MIL_TC=0 forces the fp32 FMA tiles)."""
    import os
    from deepcoro_clip_b200 import GatedAttentionPooling
    torch.manual_seed(4)
    mod = GatedAttentionPooling(512, 128, dropout=0.25).cuda().train()
    x = torch.randn(6, 400, 512, device="cuda")
    mask = torch.rand(6, 400, device="cuda") > 0.2
    mask[:, 0] = True
    go = torch.randn(6, 512, device="cuda")
    res = {}
    saved = os.environ.get("B200CLIP_MIL_TC")
    try:
        for mode in ("1", "0"):
            os.environ["B200CLIP_MIL_TC"] = mode
            mod.zero_grad(set_to_none=True)
            xr = x.clone().requires_grad_(True)
            torch.manual_seed(8)                      # same dropout seed
            out = mod(xr, mask)
            out.backward(go)
            res[mode] = [out.detach(), xr.grad] + [p.grad.clone() for p in mod.parameters()]
    finally:
        if saved is None:
            os.environ.pop("B200CLIP_MIL_TC", None)
        else:
            os.environ["B200CLIP_MIL_TC"] = saved
    for a, b in list(zip(res["1"], res["0"]))[:-1]:       # the last one is d b_w = 0 up to rounding
        assert _rel(a.cpu().numpy(), b.cpu().numpy()) <= 5e-5
