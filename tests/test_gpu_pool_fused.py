"""AttentionPool on the tcgen05 streaming kernels + the fused [B, D] tails (csrc/attnpool_tc.cu, csrc/pooltail.cu): 16-bit x,
3 + 4 library launches per forward + backward. Checked against the reference module's own structure evaluated in float64
(nn.MultiheadAttention with one learnable query + LayerNorm + optional Linear, models/attention_pool.py:10-101, same
parameters) and against this package's unfused host path (B200CLIP_POOL_FUSED=0 / B200CLIP_POOL_TC=0)."""
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


class RefPool(nn.Module):
    def __init__(self, D, H, Do):
        super().__init__()
        self.query = nn.Parameter(torch.randn(1, 1, D))
        self.attn = nn.MultiheadAttention(D, H, batch_first=True)
        self.norm = nn.LayerNorm(D)
        self.proj = nn.Linear(D, Do) if Do != D else nn.Identity()

    def forward(self, x, mask=None):
        q = self.query.expand(x.shape[0], -1, -1)
        o, _ = self.attn(query=q, key=x, value=x, key_padding_mask=mask)
        return self.proj(self.norm(o)).squeeze(1)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture
def env():
    saved = {k: os.environ.get(k) for k in ("B200CLIP_POOL_FUSED", "B200CLIP_POOL_TC")}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("B,N,D,H,Do,dtype,masked", [
    (5, 300, 256, 8, 256, torch.bfloat16, False),
    (6, 777, 512, 8, 512, torch.bfloat16, True),
    (3, 500, 512, 8, 256, torch.bfloat16, True),        # output Linear
    (9, 200, 512, 4, 512, torch.float16, False),
    (2, 3136, 512, 8, 512, torch.bfloat16, False),      # one C3 view pair
    (7, 130, 128, 4, 128, torch.bfloat16, True),        # narrowest supported width
    (4, 333, 384, 8, 384, torch.float16, True),         # head_dim 48
    (1, 65, 512, 2, 64, torch.bfloat16, False),         # one batch row, head_dim 256, narrow output Linear
])
def test_fused_attention_pool_matches_float64_reference_module(env, B, N, D, H, Do, dtype, masked):
    from deepcoro_clip_b200 import AttentionPool, _lib
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    pool = AttentionPool(D, H, output_dim=Do).to(dev)
    with torch.no_grad():
        for p in pool.parameters():
            p.add_(0.05 * torch.randn_like(p))
        pool.query.mul_(20.0)
    x = torch.randn(B, N, D, device=dev).to(dtype)
    mask = None
    if masked:
        mask = torch.rand(B, N, device=dev) < 0.25
        mask[:, 0] = False
    gy = torch.randn(B, Do, device=dev).to(dtype)
    os.environ["B200CLIP_POOL_FUSED"] = "1"
    os.environ["B200CLIP_POOL_TC"] = "1"
    xr = x.clone().requires_grad_(True)
    l0 = _lib.LAUNCHES
    y = pool(xr, mask)
    y.backward(gy)
    torch.cuda.synchronize()
    assert _lib.LAUNCHES - l0 == 7, "the fused path is 3 forward + 4 backward library launches"
    ref = RefPool(D, H, Do).to(dev).double()
    ref.load_state_dict({k: v.double() for k, v in pool.state_dict().items()})
    xd = x.double().requires_grad_(True)
    yr = ref(xd, mask)
    yr.backward(gy.double())
    tol16 = 1.2e-2 if dtype == torch.bfloat16 else 2e-3       # y and dx are stored in the 16-bit dtype
    assert _rel(y, yr) <= tol16
    assert _rel(xr.grad, xd.grad) <= tol16
    for (k, p), (_, r) in zip(pool.named_parameters(), ref.named_parameters()):
        if k == "attn.in_proj_bias":      # the key-bias third is exactly 0 here and rounding noise in the reference
            assert _rel(p.grad[:D], r.grad[:D]) <= 5e-5 and _rel(p.grad[2 * D:], r.grad[2 * D:]) <= 5e-5
            assert p.grad[D:2 * D].abs().max().item() == 0.0
        else:
            assert _rel(p.grad, r.grad) <= 5e-5, k


def test_fused_and_unfused_paths_agree_with_attention_dropout(env):
    """Training-mode attention dropout: the counter-based keep mask depends on (seed, row, token) only, so the fused path,
    the unfused tcgen05 path and the mma.sync path draw the same mask for the same torch seed."""
    from deepcoro_clip_b200 import AttentionPool
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    pool = AttentionPool(512, 8, dropout=0.2).to(dev).train()
    x = torch.randn(4, 640, 512, device=dev, dtype=torch.bfloat16)
    gy = torch.randn(4, 512, device=dev, dtype=torch.bfloat16)
    res = []
    for fused, tc in (("1", "1"), ("0", "1"), ("0", "0")):
        os.environ["B200CLIP_POOL_FUSED"] = fused
        os.environ["B200CLIP_POOL_TC"] = tc
        torch.manual_seed(7)
        pool.zero_grad(set_to_none=True)
        xr = x.clone().requires_grad_(True)
        y = pool(xr)
        y.backward(gy)
        res.append((y.float(), xr.grad.float(), {k: v.grad.clone() for k, v in pool.named_parameters()}))
    for other in res[1:]:
        assert _rel(res[0][0], other[0]) <= 1e-2
        assert _rel(res[0][1], other[1]) <= 1e-2
        for k, g in other[2].items():
            if g.norm() > 0:
                assert _rel(res[0][2][k], g) <= 2e-3, k
