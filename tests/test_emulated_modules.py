"""CPU, end to end: the package's own AttentionPool / AttentionPoolWithCLS modules (Python host code, autograd glue,
projection folding) running on top of the SHIPPED attention-pool kernels compiled for the host under the emulation
(tests/emul/pool_emul.cpp exports the b200clip_attnpool_* entry points of include/b200clip.h with the library's own
dispatch and launch geometry), checked against the golden vectors of the imported reference modules. The only stand-ins
are the device plumbing (`require_cuda`, the stream handle) and the ctypes target; no math is replaced."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN
from tests.test_host_logic import load_cls_pool

EMUL = Path(__file__).resolve().parent / "emul"
CSRC = EMUL.parents[1] / "deepcoro_clip_b200" / "csrc"


@pytest.fixture()
def on_emulated_kernels(monkeypatch):
    so = EMUL / "libpoolemul.so"
    srcs = [EMUL / "pool_emul.cpp", EMUL / "pool_mma_prims_emul.h", EMUL / "cuda_emul.h", CSRC / "attnpool_mma_kernels.cuh",
            CSRC / "attnpool_kernels.cuh", CSRC / "rope3d_kernels.cuh", CSRC / "querypool_kernels.cuh",
            CSRC / "multipos_kernels.cuh", CSRC / "dense_metrics_kernels.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", str(so), str(srcs[0])], check=True)
    emul = ctypes.CDLL(str(so))
    from deepcoro_clip_b200 import _lib, attention_pool as ap, ops
    for name, (ret, types) in _lib._prototypes().items():          # the header's prototypes, as for the real library
        fn = getattr(emul, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = ret, types
    calls = []

    def call(name, *args):
        calls.append(name)
        rc = getattr(emul, "b200clip_" + name)(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args])
        if rc != 0:
            raise _lib.B200ClipError(f"emulated b200clip_{name} failed with {rc}")

    from deepcoro_clip_b200 import multipos_loss, retrieval_metrics, rope_3d, video_aggregator
    monkeypatch.setattr(ops, "require_cuda", lambda *t: torch.device("cpu"))
    for mod in (ap, rope_3d, video_aggregator, multipos_loss, retrieval_metrics):
        monkeypatch.setattr(mod, "call", call)
        monkeypatch.setattr(mod, "stream_ptr", lambda dev=None: 0)
    monkeypatch.setattr(ap, "lib", lambda: emul)
    monkeypatch.setattr(multipos_loss, "lib", lambda: emul)
    return calls


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("name", ["attnpool_b3_n50_d128_h8", "attnpool_b4_n37_d256_h4_mask_proj"])
def test_attention_pool_module_on_emulated_kernels(on_emulated_kernels, name):
    from deepcoro_clip_b200.attention_pool import AttentionPool
    g = np.load(GOLDEN / f"{name}.npz")
    B, N, D = g["x"].shape
    out_dim = g["out"].shape[1]
    mod = AttentionPool(D, int(g["heads"]), output_dim=None if out_dim == D else out_dim)
    sd = {"query": g["p_query"], "attn.in_proj_weight": g["p_in_proj_weight"], "attn.in_proj_bias": g["p_in_proj_bias"],
          "attn.out_proj.weight": g["p_out_proj_weight"], "attn.out_proj.bias": g["p_out_proj_bias"],
          "norm.weight": g["p_norm_weight"], "norm.bias": g["p_norm_bias"]}
    if out_dim != D:
        sd["proj.weight"], sd["proj.bias"] = g["p_proj_weight"], g["p_proj_bias"]
    mod.load_state_dict({k: torch.tensor(v, dtype=torch.float32) for k, v in sd.items()})
    x = torch.tensor(g["x"], dtype=torch.float32, requires_grad=True)
    mask = torch.tensor(g["mask"]) if bool(g["has_mask"]) else None
    out = mod(x, mask)
    assert _rel(out.detach().numpy(), g["out"]) < 2e-5
    (out * torch.tensor(g["go"], dtype=torch.float32)).sum().backward()
    assert _rel(x.grad.numpy(), g["dx"]) < 5e-5
    grads = {"query": mod.query.grad, "in_proj_weight": mod.attn.in_proj_weight.grad, "in_proj_bias": mod.attn.in_proj_bias.grad,
             "out_proj_weight": mod.attn.out_proj.weight.grad, "out_proj_bias": mod.attn.out_proj.bias.grad,
             "norm_weight": mod.norm.weight.grad, "norm_bias": mod.norm.bias.grad}
    for k, v in grads.items():
        ref = g["g_" + k]
        assert np.abs(v.numpy() - ref).max() <= 5e-5 * max(np.abs(ref).max(), 1e-3), k
    assert on_emulated_kernels == ["attnpool_fwd", "attnpool_merge", "attnpool_bwd_dx", "attnpool_fwd", "attnpool_merge"]


def test_token_mean_pool_on_emulated_kernels(on_emulated_kernels):
    """SURVEY row a12 — the mean branch of VideoEncoder._pool_video_tokens (models/video_encoder.py:603) as the pool kernel's
    uniform-weights mode, against the golden produced by the reference method itself."""
    from deepcoro_clip_b200.attention_pool import token_mean_pool
    g = np.load(GOLDEN / "tokenmean_b2_n3_l50_d128.npz")
    x = torch.tensor(g["x"], dtype=torch.float32, requires_grad=True)
    out = token_mean_pool(x)
    assert out.shape == g["out"].shape and _rel(out.detach().numpy(), g["out"]) < 2e-6
    out.backward(torch.tensor(g["dout"], dtype=torch.float32))
    assert _rel(x.grad.numpy(), g["dx"]) < 2e-6


@pytest.mark.parametrize("name", ["clspool_b3_n50_d128_h8", "clspool_b4_n37_d128_h4_mask_proj"])
def test_cls_pool_module_on_emulated_kernels(on_emulated_kernels, name):
    from deepcoro_clip_b200.attention_pool import AttentionPoolWithCLS
    g = np.load(GOLDEN / f"{name}.npz")
    out_dim = g["p_proj_weight"].shape[0] if "p_proj_weight" in g.files else None
    mod = AttentionPoolWithCLS(g["x"].shape[2], int(g["heads"]), output_dim=out_dim).eval()
    params = load_cls_pool(mod, g)
    x = torch.tensor(g["x"], dtype=torch.float32, requires_grad=True)
    mask = torch.tensor(g["mask"]) if bool(g["has_mask"]) else None
    out = mod(x, mask)
    assert _rel(out.detach().numpy(), g["out"]) < 2e-5
    (out * torch.tensor(g["go"], dtype=torch.float32)).sum().backward()
    assert _rel(x.grad.numpy(), g["dx"]) < 5e-5
    for k, prm in params.items():
        got = prm.grad.numpy()
        got = got[::16] if k == "linear1_weight" else got[:, ::16] if k == "linear2_weight" else got
        ref = g["g_" + k]
        assert np.abs(got - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), k


@pytest.mark.parametrize("fused", ["0", "1"])
def test_bf16_module_matches_fp32_module_on_emulated_kernels(on_emulated_kernels, fused, monkeypatch):
    """16-bit inputs take the MMA kernels (and, with B200CLIP_POOL_FUSED_DQ=1, the fused query-gradient variant): same
    module, same parameters, bf16 x against the fp32 CUDA-core path on the rounded x."""
    from deepcoro_clip_b200.attention_pool import AttentionPool
    monkeypatch.setenv("B200CLIP_POOL_FUSED_DQ", fused)
    torch.manual_seed(3)
    mod = AttentionPool(256, 8).eval()
    x16 = torch.randn(2, 90, 256).bfloat16()
    mask = torch.rand(2, 90) < 0.15
    mask[:, 0] = False
    go = torch.randn(2, 256)
    res = []
    for x in (x16.clone().requires_grad_(True), x16.float().requires_grad_(True)):
        for p in mod.parameters():
            p.grad = None
        out = mod(x, mask)
        (out.float() * go).sum().backward()
        res.append((out.detach().float(), x.grad.float(), mod.query.grad.clone(), mod.attn.in_proj_weight.grad.clone()))
    (o16, dx16, dq16, dw16), (o32, dx32, dq32, dw32) = res
    assert _rel(o16, o32) < 1e-2 and _rel(dx16, dx32) < 1e-2              # bf16 output / dx rounding
    assert _rel(dq16, dq32) < 2e-3 and _rel(dw16, dw32) < 2e-3
    n16 = [c for c in on_emulated_kernels[:len(on_emulated_kernels) // 2]]
    assert ("attnpool_bwd_dx_dq" in n16) == (fused == "1")
    assert n16.count("attnpool_fwd") == (1 if fused == "1" else 2)         # the second pass over x is gone


@pytest.mark.parametrize("name,dtype", [("rope_f32_t3h2w2", torch.float32), ("rope_f32_t2h3w4_cls", torch.float32),
                                        ("rope_bf16_t4h7w7_cls", torch.bfloat16)])
def test_rope_on_emulated_kernels_bit_exact(on_emulated_kernels, name, dtype):
    """apply_rope_qk (forward + autograd backward) and the Rope3D module through the shipped rope3d kernels on CPU:
    BIT-EXACT against the reference module's outputs and gradients, contiguous and permuted (MViT-style) inputs."""
    from deepcoro_clip_b200.rope_3d import Rope3D, apply_rope_qk
    g = np.load(GOLDEN / f"{name}.npz")
    t = lambda a: torch.tensor(a, dtype=torch.float32).to(dtype)
    q, k = t(g["q"]).requires_grad_(True), t(g["k"]).requires_grad_(True)
    qr, kr = apply_rope_qk(q, k, t(g["sin"]), t(g["cos"]))
    assert (qr.detach().float().numpy() == g["q_rot"].astype(np.float32)).all()
    assert (kr.detach().float().numpy() == g["k_rot"].astype(np.float32)).all()
    ((qr * t(g["gq"])).sum() + (kr * t(g["gk"])).sum()).backward()
    assert (q.grad.float().numpy() == g["dq"].astype(np.float32)).all()
    assert (k.grad.float().numpy() == g["dk"].astype(np.float32)).all()
    B, heads, T, H, W, cls = [int(x) for x in g["meta"]]
    mod = Rope3D(q.shape[-1] * heads, heads).eval()
    qd = q.detach()
    qp = qd.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)           # non-contiguous view: generic-stride path
    assert not qp.is_contiguous()
    q1, k1 = mod(qd, k.detach(), T, H, W)
    q2, _ = mod(qp, k.detach(), T, H, W)
    assert torch.equal(q1, q2)
    assert (q1.float().numpy() == g["q_rot"].astype(np.float32)).all()     # CPU-built tables == the reference's
    assert on_emulated_kernels.count("rope3d_apply") == 4


@pytest.mark.parametrize("name", ["qpool_b5_n4_d64", "qpool_b6_n5_d128_mask"])
def test_query_pool_module_on_emulated_kernels(on_emulated_kernels, name):
    """EnhancedVideoAggregator's fused tail (position add, LayerNorm, masked query softmax, weighted sum) forward and
    backward through the shipped querypool kernel on CPU, against the reference module's goldens."""
    from deepcoro_clip_b200.video_aggregator import EnhancedVideoAggregator
    g = np.load(GOLDEN / f"{name}.npz")
    B, N, D = g["x"].shape
    mod = EnhancedVideoAggregator(D, num_heads=4, dropout=0.0, aggregator_depth=0, max_segments=16)
    mod.load_state_dict({"pos_encoding": torch.tensor(g["pos"], dtype=torch.float32),
                         "final_ln.weight": torch.tensor(g["ln_w"], dtype=torch.float32),
                         "final_ln.bias": torch.tensor(g["ln_b"], dtype=torch.float32),
                         "attn_query": torch.tensor(g["attn_query"], dtype=torch.float32)})
    x = torch.tensor(g["x"], dtype=torch.float32, requires_grad=True)
    mask = torch.tensor(g["mask"]) if bool(g["has_mask"]) else None
    out = mod(x, mask)
    assert _rel(out.detach().numpy(), g["out"]) < 1e-5
    (out * torch.tensor(g["go"], dtype=torch.float32)).sum().backward()
    assert _rel(x.grad.numpy(), g["dx"]) < 2e-5
    assert _rel(mod.pos_encoding.grad.numpy(), g["g_pos"]) < 2e-5
    assert _rel(mod.final_ln.weight.grad.numpy(), g["g_ln_w"]) < 2e-5
    assert _rel(mod.final_ln.bias.grad.numpy(), g["g_ln_b"]) < 2e-5
    assert _rel(mod.attn_query.grad.numpy(), g["g_attn_query"]) < 2e-5
    assert on_emulated_kernels == ["querypool", "querypool"]


@pytest.mark.parametrize("name", ["multipos_48x64"])      # 130x37 is covered on the GPU; it costs 14 s of thread spawns here
def test_multipos_loss_modules_on_emulated_kernels(on_emulated_kernels, name):
    """WeightedSigLIPLoss / MultiPositiveInfoNCELoss (the package's nn.Modules and autograd function) through the shipped
    multipos kernels on CPU, against the reference classes' fp32 autograd goldens."""
    from deepcoro_clip_b200.multipos_loss import MultiPositiveInfoNCELoss, WeightedSigLIPLoss
    g = np.load(GOLDEN / f"{name}.npz")
    logits, mask, pw = (torch.tensor(g[k], dtype=torch.float32) for k in ("logits", "mask", "pos_weights"))
    cases = {"wsl": lambda L: WeightedSigLIPLoss()(L, mask * pw - 0.2 * (1 - mask)),
             "mpi_mean": lambda L: MultiPositiveInfoNCELoss()(L, mask, pw),
             "mpi_sum_noweights": lambda L: MultiPositiveInfoNCELoss(reduction="sum")(L, mask),
             "mpi_imp_mean": lambda L: MultiPositiveInfoNCELoss(use_importance_weighting=True)(L, mask, pw),
             "mpi_imp_sum_noweights": lambda L: MultiPositiveInfoNCELoss(reduction="sum",
                                                                         use_importance_weighting=True)(L, mask)}
    for key, fn in cases.items():
        L = logits.clone().requires_grad_(True)
        loss = fn(L)
        assert loss.ndim == 0 and loss.requires_grad
        loss.backward()
        ref = float(g[key + "_loss"])
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref), key
        d = g[key + "_dlogits"]
        assert np.abs(L.grad.numpy() - d).max() <= 1e-4 * np.abs(d).max() + 1e-9, key


@pytest.mark.parametrize("name", ["dense_metrics_64x7_g3", "dense_metrics_120x90_g1"])
def test_dense_metrics_on_emulated_kernels(on_emulated_kernels, name):
    """utils/retrieval_metrics.py drop-ins (recall with ground-truth sets, MRR, MAP, NDCG, median rank) through the shipped
    rank-pass / per-row kernels on CPU, against the reference functions' goldens and the oracle on a tie-heavy matrix."""
    from deepcoro_clip_b200 import retrieval_metrics as rm
    from oracle import dense_metrics_oracle as dmo
    g = np.load(GOLDEN / f"{name}.npz")
    sim = torch.tensor(g["sim"])
    gt = [[int(c) for c in row if c >= 0] for row in g["gt"]]
    ks = [int(k) for k in g["k_values"]]
    allm = rm.compute_all_dense_metrics(sim, gt, recall_k=ks, ndcg_k=ks)          # every metric from ONE rank pass
    for i, k in enumerate(ks):
        assert allm[f"Recall@{k}"] == float(g["recall"][i])
        assert abs(allm[f"NDCG@{k}_V2T"] - float(g["ndcg"][i])) <= 1e-6
    assert abs(allm["MRR_V2T"] - float(g["mrr"])) <= 1e-15
    assert abs(allm["MAP"] - float(g["map"])) <= 1e-6
    assert allm["MedianRank_V2T"] == int(g["median_rank"])
    assert rm.compute_recall_at_k(sim, gt, ks) == {f"Recall@{k}": allm[f"Recall@{k}"] for k in ks}   # the reference API
    if name == "dense_metrics_64x7_g3":
        # ties (lowest index first), empty rows, out-of-range indices; bf16 input through the 16-bit instantiation
        rng = np.random.default_rng(5)
        q = torch.tensor((np.round(rng.standard_normal((40, 77)) * 4) / 4).astype(np.float32)).bfloat16()
        gq = [[int(c) for c in rng.choice(80, size=int(rng.integers(0, 5)), replace=False)] for _ in range(40)]
        qo = q.float().numpy()
        a = rm.compute_all_dense_metrics(q, gq, recall_k=[1, 5, 10], ndcg_k=[5])
        o = dmo.recall_at_k(qo, gq, [1, 5, 10])
        assert all(a[k] == o[k] for k in o)
        assert a["MedianRank_V2T"] == dmo.median_rank(qo, gq)
        assert abs(a["MAP"] - dmo.mean_ap(qo, gq)) <= 1e-6


@pytest.mark.parametrize("dtype,D,N,fused", [(torch.float32, 128, 70, "0"), (torch.bfloat16, 256, 90, "0"), (torch.bfloat16, 256, 90, "1")])
def test_attention_pool_training_dropout_on_emulated_kernels(on_emulated_kernels, monkeypatch, dtype, D, N, fused):
    """Training-mode attention dropout (the reference's default attention_pool_dropout is 0.1, so this is the path a training
    step takes): forward and backward equal a float64 autograd replica of nn.MultiheadAttention's math driven by the SAME
    counter-based mask — CUDA-core fp32 kernels, the MMA kernels and the opt-in fused query gradient."""
    import math
    from deepcoro_clip_b200 import attention_pool as ap
    from tests.test_gpu_tokens import _keep_mask
    monkeypatch.setenv("B200CLIP_POOL_FUSED_DQ", fused)
    B, H, p = 2, 8, 0.25
    torch.manual_seed(7)
    mod = ap.AttentionPool(D, H, dropout=p).train()
    x = torch.randn(B, N, D).to(dtype).requires_grad_(True)
    torch.manual_seed(123)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # what forward() will draw
    torch.manual_seed(123)
    out = mod(x)
    gout = torch.randn(B, D)
    out.float().backward(gout)
    keep = _keep_mask(seed, B, H, N, p, "cpu")
    xd = x.detach().double().requires_grad_(True)
    P = {k: v.detach().double().requires_grad_(True) for k, v in mod.named_parameters()}
    Wd, bd = P["attn.in_proj_weight"], P["attn.in_proj_bias"]
    Dh = D // H
    q0 = (P["query"].view(1, D) @ Wd[:D].T + bd[:D]).view(H, Dh)
    K = (xd @ Wd[D:2 * D].T + bd[D:2 * D]).view(B, N, H, Dh)
    V = (xd @ Wd[2 * D:].T + bd[2 * D:]).view(B, N, H, Dh)
    a = torch.softmax(torch.einsum("hk,bnhk->bhn", q0, K) / math.sqrt(Dh), dim=-1) * keep.double() / (1 - p)
    o = torch.einsum("bhn,bnhk->bhk", a, V).reshape(B, D)
    y = o @ P["attn.out_proj.weight"].T + P["attn.out_proj.bias"]
    y = torch.nn.functional.layer_norm(y, (D,), P["norm.weight"], P["norm.bias"], mod.norm.eps)
    y.backward(gout.double())
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    assert _rel(out.float().detach().numpy(), y.detach().numpy()) < tol
    assert _rel(x.grad.float().numpy(), xd.grad.numpy()) < (5e-5 if dtype == torch.float32 else 1.5e-2)
    for name, prm in mod.named_parameters():                  # parameter gradients incl. the query (dqt path) and biases
        assert _rel(prm.grad.numpy(), P[name].grad.numpy()) < (1e-4 if dtype == torch.float32 else 2e-2), name
    assert ("attnpool_bwd_dx_dq" in on_emulated_kernels) == (fused == "1")


@pytest.mark.parametrize("fused", ["0", "1"])
def test_cls_pool_bf16_matches_fp32_on_emulated_kernels(on_emulated_kernels, fused, monkeypatch):
    """AttentionPoolWithCLS with 16-bit tokens (MMA kernels, lse output forward, dlse input backward — also through the
    opt-in fused query gradient) against the same module on the fp32 CUDA-core kernels with the rounded tokens."""
    from deepcoro_clip_b200.attention_pool import AttentionPoolWithCLS
    monkeypatch.setenv("B200CLIP_POOL_FUSED_DQ", fused)
    torch.manual_seed(11)
    mod = AttentionPoolWithCLS(256, 8).eval()
    x16 = torch.randn(2, 75, 256).bfloat16()
    mask = torch.rand(2, 75) < 0.2
    mask[:, 0] = False
    go = torch.randn(2, 256)
    res = []
    for x in (x16.clone().requires_grad_(True), x16.float().requires_grad_(True)):
        for p in mod.parameters():
            p.grad = None
        out = mod(x, mask)
        (out.float() * go).sum().backward()
        res.append((out.detach().float(), x.grad.float(), {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None}))
    (o16, dx16, g16), (o32, dx32, g32) = res
    assert _rel(o16, o32) < 1e-2 and _rel(dx16, dx32) < 1.5e-2
    assert set(g16) == set(g32)
    for n in g32:
        assert _rel(g16[n], g32[n]) < 1e-2, n
    assert ("attnpool_bwd_dx_dq" in on_emulated_kernels) == (fused == "1")
