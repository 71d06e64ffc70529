"""GPU parity: Rope3D, AttentionPool and the multi-view query pool vs the reference golden vectors / the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import token_oracle as to
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("name,dtype", [("rope_f32_t2h3w4_cls", torch.float32), ("rope_f32_t3h2w2", torch.float32),
                                        ("rope_bf16_t4h7w7_cls", torch.bfloat16)])
def test_rope_bit_exact_with_reference_tables(name, dtype):
    from deepcoro_clip_b200.rope_3d import apply_rope_qk
    g = np.load(GOLDEN / f"{name}.npz")
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=DEV).to(dtype)
    q, k = t(g["q"]).requires_grad_(True), t(g["k"]).requires_grad_(True)
    qr, kr = apply_rope_qk(q, k, t(g["sin"]), t(g["cos"]))
    assert (qr.detach().float().cpu().numpy() == g["q_rot"].astype(np.float32)).all()
    assert (kr.detach().float().cpu().numpy() == g["k_rot"].astype(np.float32)).all()
    ((qr * t(g["gq"])).sum() + (kr * t(g["gk"])).sum()).backward()
    assert (q.grad.float().cpu().numpy() == g["dq"].astype(np.float32)).all()
    assert (k.grad.float().cpu().numpy() == g["dk"].astype(np.float32)).all()


@pytest.mark.parametrize("name,dtype", [("rope_f32_t2h3w4_cls", torch.float32), ("rope_bf16_t4h7w7_cls", torch.bfloat16)])
def test_rope_module_forward(name, dtype):
    from deepcoro_clip_b200.rope_3d import Rope3D
    g = np.load(GOLDEN / f"{name}.npz")
    B, heads, T, H, W, cls = [int(x) for x in g["meta"]]
    mod = Rope3D(96 * heads, heads).eval()
    q = torch.tensor(g["q"], dtype=torch.float32, device=DEV).to(dtype)
    k = torch.tensor(g["k"], dtype=torch.float32, device=DEV).to(dtype)
    qr, kr = mod(q, k, T, H, W)                       # CLS auto-detected; tables built on the GPU
    tol = 2e-6 if dtype == torch.float32 else 2 ** -7    # GPU sin/cos may differ from the CPU tables by one ulp
    assert np.abs(qr.float().cpu().numpy() - g["q_rot"]).max() <= tol * max(1.0, np.abs(g["q_rot"]).max())
    assert np.abs(kr.float().cpu().numpy() - g["k_rot"]).max() <= tol * max(1.0, np.abs(g["k_rot"]).max())
    # non-contiguous (permuted) views, as MViT hands them over
    qp = q.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
    assert not qp.is_contiguous()
    qr2, _ = mod(qp, k, T, H, W)
    assert torch.equal(qr2, qr)
    # mismatch -> unchanged
    bad = torch.randn(1, heads, 7, 96, device=DEV)
    o1, o2 = mod(bad, bad, 2, 2, 2)
    assert o1 is bad and o2 is bad


def test_rope_large_bf16_vs_oracle():
    from deepcoro_clip_b200.rope_3d import Rope3D
    T, H, W = 16, 14, 14
    mod = Rope3D(768, 8).eval()
    q = torch.randn(2, 8, T * H * W + 1, 96, device=DEV).bfloat16()
    k = torch.randn(2, 8, T * H * W + 1, 96, device=DEV).bfloat16()
    qr, kr = mod(q, k, T, H, W)
    oq, ok = to.rope3d_forward(q.float().cpu().numpy().astype(np.float64), k.float().cpu().numpy().astype(np.float64), T, H, W)
    # bf16 tables + bf16 roundings: ~1e-2 absolute at |x| ~ 4 (same as the reference module in bf16)
    assert np.abs(qr.float().cpu().numpy() - oq).max() < 0.2
    assert _rel(qr.float().cpu().numpy(), oq) < 2e-2


@pytest.mark.parametrize("name", ["attnpool_b3_n50_d128_h8", "attnpool_b4_n37_d256_h4_mask_proj"])
def test_attention_pool_golden(name):
    from deepcoro_clip_b200.attention_pool import AttentionPool
    g = np.load(GOLDEN / f"{name}.npz")
    B, N, D = g["x"].shape
    out_dim = g["out"].shape[1]
    mod = AttentionPool(D, int(g["heads"]), output_dim=None if out_dim == D else out_dim).to(DEV)
    sd = {"query": g["p_query"], "attn.in_proj_weight": g["p_in_proj_weight"], "attn.in_proj_bias": g["p_in_proj_bias"],
          "attn.out_proj.weight": g["p_out_proj_weight"], "attn.out_proj.bias": g["p_out_proj_bias"],
          "norm.weight": g["p_norm_weight"], "norm.bias": g["p_norm_bias"]}
    if out_dim != D:
        sd["proj.weight"] = g["p_proj_weight"]; sd["proj.bias"] = g["p_proj_bias"]
    mod.load_state_dict({k: torch.tensor(v, dtype=torch.float32) for k, v in sd.items()})
    x = torch.tensor(g["x"], dtype=torch.float32, device=DEV, requires_grad=True)
    mask = torch.tensor(g["mask"], device=DEV) if bool(g["has_mask"]) else None
    out = mod(x, mask)
    assert out.shape == (B, out_dim)
    assert _rel(out.detach().cpu().numpy(), g["out"]) < 2e-5
    (out * torch.tensor(g["go"], dtype=torch.float32, device=DEV)).sum().backward()
    assert _rel(x.grad.cpu().numpy(), g["dx"]) < 5e-5
    grads = {"query": mod.query.grad, "in_proj_weight": mod.attn.in_proj_weight.grad, "in_proj_bias": mod.attn.in_proj_bias.grad,
             "out_proj_weight": mod.attn.out_proj.weight.grad, "out_proj_bias": mod.attn.out_proj.bias.grad,
             "norm_weight": mod.norm.weight.grad, "norm_bias": mod.norm.bias.grad}
    for k, v in grads.items():
        ref = g["g_" + k]
        assert np.abs(v.cpu().numpy() - ref).max() <= 5e-5 * max(np.abs(ref).max(), 1e-3), k


def test_attention_pool_c3_shape_bf16_vs_oracle():
    from deepcoro_clip_b200.attention_pool import AttentionPool
    torch.manual_seed(2)
    mod = AttentionPool(512, 8, dropout=0.0).to(DEV)
    x = torch.randn(4, 3136, 512, device=DEV).bfloat16().requires_grad_(True)
    mask = torch.rand(4, 3136, device=DEV) < 0.1
    out = mod(x, mask)
    params = {"query": mod.query, "in_proj_weight": mod.attn.in_proj_weight, "in_proj_bias": mod.attn.in_proj_bias,
              "out_proj_weight": mod.attn.out_proj.weight, "out_proj_bias": mod.attn.out_proj.bias,
              "norm_weight": mod.norm.weight, "norm_bias": mod.norm.bias}
    pn = {k: v.detach().double().cpu().numpy() for k, v in params.items()}
    o, cache = to.attention_pool_forward(x.detach().float().cpu().numpy(), pn, 8, mask.cpu().numpy(), want_cache=True)
    assert out.dtype == torch.bfloat16
    assert _rel(out.float().detach().cpu().numpy(), o) < 1e-2          # bf16 output rounding
    go = torch.randn(4, 512, device=DEV)
    (out.float() * go).sum().backward()
    gr = to.attention_pool_backward(go.cpu().numpy(), cache, pn)
    assert _rel(x.grad.float().cpu().numpy(), gr["x"]) < 1e-2


@pytest.mark.parametrize("name", ["qpool_b5_n4_d64", "qpool_b6_n5_d128_mask"])
def test_query_pool_golden(name):
    from deepcoro_clip_b200.video_aggregator import EnhancedVideoAggregator
    g = np.load(GOLDEN / f"{name}.npz")
    B, N, D = g["x"].shape
    mod = EnhancedVideoAggregator(D, num_heads=4, dropout=0.0, aggregator_depth=0, max_segments=16).to(DEV)
    mod.load_state_dict({"pos_encoding": torch.tensor(g["pos"], dtype=torch.float32),
                         "final_ln.weight": torch.tensor(g["ln_w"], dtype=torch.float32),
                         "final_ln.bias": torch.tensor(g["ln_b"], dtype=torch.float32),
                         "attn_query": torch.tensor(g["attn_query"], dtype=torch.float32)})
    x = torch.tensor(g["x"], dtype=torch.float32, device=DEV, requires_grad=True)
    mask = torch.tensor(g["mask"], device=DEV) if bool(g["has_mask"]) else None
    out = mod(x, mask)
    assert _rel(out.detach().cpu().numpy(), g["out"]) < 1e-5
    (out * torch.tensor(g["go"], dtype=torch.float32, device=DEV)).sum().backward()
    assert _rel(x.grad.cpu().numpy(), g["dx"]) < 2e-5
    assert _rel(mod.pos_encoding.grad.cpu().numpy(), g["g_pos"]) < 2e-5
    assert _rel(mod.final_ln.weight.grad.cpu().numpy(), g["g_ln_w"]) < 2e-5
    assert _rel(mod.final_ln.bias.grad.cpu().numpy(), g["g_ln_b"]) < 2e-5
    assert _rel(mod.attn_query.grad.cpu().numpy(), g["g_attn_query"]) < 2e-5


def test_aggregator_with_blocks_matches_torch_composition():
    """depth > 0: blocks stay PyTorch; the fused tail must equal the unfused torch tail on the same block output."""
    from deepcoro_clip_b200.video_aggregator import EnhancedVideoAggregator
    torch.manual_seed(0)
    mod = EnhancedVideoAggregator(512, num_heads=4, dropout=0.0, aggregator_depth=2, max_segments=16).to(DEV).eval()
    x = torch.randn(8, 4, 512, device=DEV)
    mask = torch.zeros(8, 4, dtype=torch.bool, device=DEV); mask[3, 2:] = True
    out = mod(x, mask)
    h = x + mod.pos_encoding[:, :4]
    for b in mod.blocks:
        h = b(h, key_padding_mask=mask)
    h = mod.final_ln(h).masked_fill(mask.unsqueeze(-1), 0.0)
    s = (mod.attn_query.expand(8, -1, -1) @ h.transpose(1, 2)).masked_fill(mask.unsqueeze(1), float("-inf"))
    ref = (torch.softmax(s, -1) @ h).squeeze(1)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5)


def _keep_mask(seed, B, H, N, p, device):
    """torch replica of attn_keep() (csrc/common.cuh): splitmix64 of (seed, b*H + h, n), top 24 bits >= p."""
    M = (1 << 64) - 1
    import numpy as np
    rows = np.arange(B * H, dtype=np.uint64)[:, None]
    n = np.arange(N, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (rows * np.uint64(0x100000001B3) + n + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return torch.tensor(u >= np.float32(p), device=device).view(B, H, N)


@pytest.mark.parametrize("dtype,D,N", [(torch.float32, 128, 70), (torch.bfloat16, 256, 333), (torch.bfloat16, 512, 100)])
def test_attention_pool_training_dropout(dtype, D, N):
    """Training-mode attention dropout (reference default attention_pool_dropout = 0.1): our counter-based mask is not
    PyTorch's Philox stream, so the check is (1) exact: forward / backward equal a float64 autograd replica of
    nn.MultiheadAttention's math driven by the SAME mask, (2) statistical: the keep rate is 1 - p."""
    from deepcoro_clip_b200 import attention_pool as ap
    B, H, p = 3, 8, 0.25
    torch.manual_seed(7)
    mod = ap.AttentionPool(D, H, dropout=p).to("cuda:0").train()
    x = torch.randn(B, N, D, device="cuda:0").to(dtype).requires_grad_(True)
    torch.manual_seed(123)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # what forward() will draw
    torch.manual_seed(123)
    out = mod(x)
    gout = torch.randn_like(out.float())
    out.float().backward(gout)
    keep = _keep_mask(seed, B, H, N, p, "cuda:0")
    assert abs(keep.float().mean().item() - (1 - p)) < 0.03
    # float64 replica with the same mask
    xd = x.detach().double().requires_grad_(True)
    Wd = mod.attn.in_proj_weight.detach().double(); bd = mod.attn.in_proj_bias.detach().double()
    Dh = D // H
    q0 = (mod.query.detach().double().view(1, D) @ Wd[:D].T + bd[:D]).view(H, Dh)
    K = (xd @ Wd[D:2 * D].T + bd[D:2 * D]).view(B, N, H, Dh)
    V = (xd @ Wd[2 * D:].T + bd[2 * D:]).view(B, N, H, Dh)
    s_ = torch.einsum("hk,bnhk->bhn", q0, K) / math.sqrt(Dh)
    a = torch.softmax(s_, dim=-1) * keep.double() / (1 - p)
    o = torch.einsum("bhn,bnhk->bhk", a, V).reshape(B, D)
    y = o @ mod.attn.out_proj.weight.detach().double().T + mod.attn.out_proj.bias.detach().double()
    y = torch.nn.functional.layer_norm(y, (D,), mod.norm.weight.detach().double(), mod.norm.bias.detach().double(), mod.norm.eps)
    y.backward(gout.double())
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert _rel(out.float().detach().cpu().numpy(), y.detach().cpu().numpy()) < tol
    assert _rel(x.grad.float().cpu().numpy(), xd.grad.cpu().numpy()) < (5e-5 if dtype == torch.float32 else 1.5e-2)
    # eval mode ignores dropout
    mod.eval()
    o1 = mod(x.detach()); o2 = mod(x.detach())
    assert torch.equal(o1, o2)


def test_attention_pool_cuda_graph_callable():
    """The module's forward + backward are capturable (no host sync, no stream-unsafe call inside the custom ops):
    torch.cuda.make_graphed_callables(AttentionPool) reproduces the eager module. This is the recommended way to
    remove the per-call host work of the [B, D] projection tails around the streaming kernel."""
    import copy
    from deepcoro_clip_b200.attention_pool import AttentionPool
    torch.manual_seed(5)
    pool = AttentionPool(512, 8).to(DEV)
    ref = copy.deepcopy(pool)
    xs = torch.randn(4, 393, 512, device=DEV).bfloat16().requires_grad_(True)
    graphed = torch.cuda.make_graphed_callables(pool, (xs,))
    for trial in range(2):
        x = torch.randn(4, 393, 512, device=DEV).bfloat16()
        xg = x.clone().requires_grad_(True)
        xe = x.clone().requires_grad_(True)
        up = torch.randn(4, 512, device=DEV).bfloat16()
        for p in list(pool.parameters()) + list(ref.parameters()):
            p.grad = None
        yg = graphed(xg)
        yg.backward(up)
        ye = ref(xe)
        ye.backward(up)
        torch.cuda.synchronize()
        assert torch.equal(yg, ye)
        assert float((xg.grad.float() - xe.grad.float()).abs().max()) <= 1e-2 * float(xe.grad.float().abs().max())
        for (n, a), (_, b) in zip(pool.named_parameters(), ref.named_parameters()):
            assert float((a.grad - b.grad).abs().max()) <= 1e-4 * float(b.grad.abs().max()) + 1e-8, n


# ---- AttentionPoolWithCLS (reference models/attention_pool.py:104-197) ----
def _cls_pool_replica(mod, x, mask, keep_x=None, keep_c=None, p=0.0):
    """float64 autograd replica of the reference module's math (CLS row of one post-LN TransformerEncoderLayer over
    [CLS; x], final norm, proj) with the attention-weight dropout mask given explicitly: keep_x [B, H, N] for the
    tokens, keep_c [B, H] for the CLS key. Returns (out, x64, params64) with gradients attached to x64 / params64."""
    F = torch.nn.functional
    P = {k: v.detach().double().requires_grad_(True) for k, v in mod.named_parameters()}
    L = "transformer.layers.0."
    xd = x.detach().double().requires_grad_(True)
    B, N, D = xd.shape
    H = mod.num_heads
    Dh = D // H
    X = torch.cat([P["cls_token"].expand(B, 1, D), xd], 1)
    W, bias = P[L + "self_attn.in_proj_weight"], P[L + "self_attn.in_proj_bias"]
    q = (X[:, 0] @ W[:D].T + bias[:D]).view(B, H, Dh)
    K = (X @ W[D:2 * D].T + bias[D:2 * D]).view(B, N + 1, H, Dh)
    V = (X @ W[2 * D:].T + bias[2 * D:]).view(B, N + 1, H, Dh)
    s_ = torch.einsum("bhk,bnhk->bhn", q, K) / math.sqrt(Dh)
    if mask is not None:
        mk = torch.cat([torch.zeros(B, 1, dtype=torch.bool, device=mask.device), mask], 1)
        s_ = s_.masked_fill(mk[:, None, :], float("-inf"))
    a = torch.softmax(s_, dim=-1)
    if keep_x is not None:
        a = a * torch.cat([keep_c.unsqueeze(-1), keep_x], -1).double() / (1.0 - p)
    o = torch.einsum("bhn,bnhk->bhk", a, V).reshape(B, D)
    y = o @ P[L + "self_attn.out_proj.weight"].T + P[L + "self_attn.out_proj.bias"]
    y = F.layer_norm(X[:, 0] + y, (D,), P[L + "norm1.weight"], P[L + "norm1.bias"], 1e-5)
    f = torch.relu(y @ P[L + "linear1.weight"].T + P[L + "linear1.bias"]) @ P[L + "linear2.weight"].T + P[L + "linear2.bias"]
    y = F.layer_norm(y + f, (D,), P[L + "norm2.weight"], P[L + "norm2.bias"], 1e-5)
    y = F.layer_norm(y, (D,), P["norm.weight"], P["norm.bias"], 1e-5)
    if "proj.weight" in P:
        y = y @ P["proj.weight"].T + P["proj.bias"]
    return y, xd, P


@pytest.mark.parametrize("name", ["clspool_b3_n50_d128_h8", "clspool_b4_n37_d128_h4_mask_proj"])
def test_cls_pool_golden(name):
    """fp32 module on the GPU vs the imported reference's fp64 output and gradients (x, cls_token, every layer
    parameter); the masked fixture contains a sample whose tokens are ALL masked."""
    from deepcoro_clip_b200.attention_pool import AttentionPoolWithCLS
    from tests.test_host_logic import load_cls_pool
    g = np.load(GOLDEN / f"{name}.npz")
    B, N, D = g["x"].shape
    out_dim = g["out"].shape[1]
    mod = AttentionPoolWithCLS(D, int(g["heads"]), output_dim=None if out_dim == D else out_dim).to(DEV).eval()
    params = load_cls_pool(mod, g)
    x = torch.tensor(g["x"], dtype=torch.float32, device=DEV, requires_grad=True)
    mask = torch.tensor(g["mask"], device=DEV) if bool(g["has_mask"]) else None
    from deepcoro_clip_b200 import _lib
    n0 = _lib.LAUNCHES
    out = mod(x, mask)
    assert _lib.LAUNCHES >= n0 + 2                      # the streaming kernel + merge ran (no framework fallback)
    assert out.shape == (B, out_dim) and bool(torch.isfinite(out).all())
    assert _rel(out.detach().cpu().numpy(), g["out"]) < 2e-5
    (out * torch.tensor(g["go"], dtype=torch.float32, device=DEV)).sum().backward()
    assert bool(torch.isfinite(x.grad).all())
    assert _rel(x.grad.cpu().numpy(), g["dx"]) < 5e-5
    for k, prm in params.items():
        got = prm.grad.cpu().numpy()
        got = got[::16] if k == "linear1_weight" else got[:, ::16] if k == "linear2_weight" else got
        ref = g["g_" + k]
        assert np.abs(got - ref).max() <= 5e-5 * max(np.abs(ref).max(), 1e-3), k


@pytest.mark.parametrize("B,N,use_mask", [(4, 393, False), (2, 3136, True)])
def test_cls_pool_bf16_d512_vs_oracle(B, N, use_mask):
    """Real call shapes (MViT-v2-S output 393 tokens / synthetic C3 3,136 tokens, D = 512, 8 heads, bf16 tokens):
    tensor-core streaming kernels; oracle = literal numpy restatement (token_oracle.cls_pool_*)."""
    from deepcoro_clip_b200.attention_pool import AttentionPoolWithCLS
    from tests.test_host_logic import CLS_POOL_PARAMS
    torch.manual_seed(3)
    mod = AttentionPoolWithCLS(512, 8, dropout=0.1).to(DEV).eval()
    with torch.no_grad():
        mod.cls_token.normal_(std=0.5)
    x = torch.randn(B, N, 512, device=DEV).bfloat16().requires_grad_(True)
    mask = (torch.rand(B, N, device=DEV) < 0.1) if use_mask else None
    out = mod(x, mask)
    assert out.dtype == torch.bfloat16 and out.shape == (B, 512)
    named = dict(mod.named_parameters())
    pn = {}
    for k, tail in CLS_POOL_PARAMS.items():
        full = tail if tail.split(".")[0] in ("cls_token", "norm", "proj") else "transformer.layers.0." + tail
        if full in named:
            pn[k] = named[full].detach().double().cpu().numpy()
    o, cache = to.cls_pool_forward(x.detach().float().cpu().numpy(), pn, 8, None if mask is None else mask.cpu().numpy(),
                                   want_cache=True)
    assert _rel(out.float().detach().cpu().numpy(), o) < 1e-2          # bf16 output rounding
    go = torch.randn(B, 512, device=DEV)
    (out.float() * go).sum().backward()
    gr = to.cls_pool_backward(go.cpu().numpy(), cache, pn)
    assert _rel(x.grad.float().cpu().numpy(), gr["x"]) < 1e-2           # dx is written in bf16
    assert _rel(mod.cls_token.grad.cpu().numpy(), gr["cls_token"]) < 5e-3
    layer = mod.transformer.layers[0]
    assert _rel(layer.self_attn.in_proj_weight.grad.cpu().numpy(), gr["in_proj_weight"]) < 5e-3
    assert _rel(layer.linear1.weight.grad.cpu().numpy(), gr["linear1_weight"]) < 5e-3


@pytest.mark.parametrize("dtype,D,N", [(torch.float32, 128, 70), (torch.bfloat16, 512, 333)])
def test_cls_pool_training_attention_dropout(dtype, D, N, monkeypatch):
    """Training mode: attention-weight dropout over the N + 1 keys (token keys: the kernel's counter-based mask, CLS
    key: one device draw per (sample, head)) against a float64 replica driven by the SAME masks. The three
    activation dropouts of the layer (Philox, dtype-dependent) are switched off for the exact comparison and
    exercised separately below."""
    from deepcoro_clip_b200 import attention_pool as ap
    B, H, p = 3, 8, 0.25
    torch.manual_seed(11)
    mod = ap.AttentionPoolWithCLS(D, H, dropout=p).to(DEV).train()
    with torch.no_grad():
        mod.cls_token.normal_(std=0.5)
    x = torch.randn(B, N, D, device=DEV).to(dtype).requires_grad_(True)
    mask = torch.rand(B, N, device=DEV) < 0.15
    real_dropout = ap.F.dropout
    monkeypatch.setattr(ap.F, "dropout", lambda t, p_=0.5, training=True, inplace=False: t)
    torch.manual_seed(321)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # what forward() will draw
    keep_c = torch.rand(B, H, device=DEV) >= p
    torch.manual_seed(321)
    out = mod(x, mask)
    gout = torch.randn_like(out.float())
    out.float().backward(gout)
    keep_x = _keep_mask(seed, B, H, N, p, DEV)
    y, xd, P = _cls_pool_replica(mod, x, mask, keep_x, keep_c, p)
    y.backward(gout.double())
    f32 = dtype == torch.float32
    assert _rel(out.float().detach().cpu().numpy(), y.detach().cpu().numpy()) < (2e-5 if f32 else 1e-2)
    assert _rel(x.grad.float().cpu().numpy(), xd.grad.cpu().numpy()) < (5e-5 if f32 else 1.5e-2)
    assert _rel(mod.cls_token.grad.cpu().numpy(), P["cls_token"].grad.cpu().numpy()) < (5e-5 if f32 else 5e-3)
    w = "transformer.layers.0.self_attn.in_proj_weight"
    assert _rel(dict(mod.named_parameters())[w].grad.cpu().numpy(), P[w].grad.cpu().numpy()) < (5e-5 if f32 else 5e-3)
    # all dropouts on: finite, different from call to call; eval mode deterministic
    monkeypatch.setattr(ap.F, "dropout", real_dropout)
    o1 = mod(x.detach(), mask); o2 = mod(x.detach(), mask)
    assert bool(torch.isfinite(o1).all()) and not torch.equal(o1, o2)
    mod.eval()
    assert torch.equal(mod(x.detach(), mask), mod(x.detach(), mask))


@pytest.mark.parametrize("name", ["tokenmean_b2_n3_l50_d128", "tokenmean_b1_n2_l197_d256"])
def test_token_mean_pool_matches_reference_golden(name):
    """VideoEncoder._pool_video_tokens' mean branch (models/video_encoder.py:603) = the pool kernel with
    uniform weights; golden from the reference method."""
    from deepcoro_clip_b200 import token_mean_pool
    g = np.load(GOLDEN / f"{name}.npz")
    x = torch.tensor(g["x"], dtype=torch.float32, device=DEV, requires_grad=True)
    out = token_mean_pool(x)
    assert out.shape == g["out"].shape
    assert np.abs(out.detach().cpu().numpy() - g["out"]).max() <= 2e-6 * np.abs(g["out"]).max()
    out.backward(torch.tensor(g["dout"], dtype=torch.float32, device=DEV))
    assert np.abs(x.grad.cpu().numpy() - g["dx"]).max() <= 2e-6 * np.abs(g["dx"]).max()


def test_token_mean_pool_bf16_mask_and_batched_view_pooling():
    """bf16 tokens at the study-mode shape (mean over 3,136 tokens, fp32 accumulation), the masked mean, and
    pool_video_tokens: all views in ONE pass of the attention pool equal the reference's per-view Python loop."""
    import types
    from deepcoro_clip_b200 import AttentionPool, pool_video_tokens, token_mean_pool
    torch.manual_seed(0)
    x = torch.randn(2, 4, 3136, 512, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    out = token_mean_pool(x)
    ref = x.detach().float().mean(dim=2)
    assert out.dtype == torch.bfloat16 and (out.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item()
    go = torch.randn_like(out)
    out.backward(go)
    assert (x.grad.float() - (go.float() / 3136).unsqueeze(2).expand_as(x)).abs().max().item() <= 2.0 ** -8 * go.abs().max().item() / 3136
    xs = torch.randn(3, 100, 256, device=DEV)
    mask = torch.rand(3, 100, device=DEV) < 0.3
    got = token_mean_pool(xs, mask)
    want = (xs * (~mask).unsqueeze(-1)).sum(1) / (~mask).sum(1, keepdim=True)
    assert (got - want).abs().max().item() <= 1e-5
    pool = AttentionPool(256, 4).to(DEV).eval()
    tf = torch.randn(2, 3, 77, 256, device=DEV)
    enc = types.SimpleNamespace(attention_pool=pool)
    batched = pool_video_tokens(enc, tf)
    looped = torch.cat([pool(tf[:, i]).unsqueeze(1) for i in range(3)], dim=1)      # models/video_encoder.py:598-602
    assert batched.shape == (2, 3, 256) and (batched - looped).abs().max().item() <= 1e-5
    assert (pool_video_tokens(types.SimpleNamespace(attention_pool=None), tf) - tf.mean(dim=2)).abs().max().item() <= 1e-6
