"""CPU: edge cases of the host code around the kernels, run end to end on the emulated library (see
tests/test_emulated_losses.py / test_emulated_retrieval.py for what is shipped device code and what is modelled) against the
numpy oracle: input dtypes and layouts, partial gradient requests, log_temp forms, ragged / degenerate shapes, the
positive-list overflow poison, tiny and clamped retrieval problems."""
import math

import numpy as np
import pytest
import torch

from oracle import contrastive_oracle as co
from oracle import retrieval_oracle as ro
from tests.test_emulated_losses import _rel, build_emul, patch_package
from tests.test_emulated_retrieval import patch_retrieval


@pytest.fixture()
def loss_mod(monkeypatch):
    return patch_package(build_emul(), monkeypatch.setattr)


def _data(B, D, seed, T=None):
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((T or B, D)).astype(np.float32)
    v = (0.5 * t[:B] if T is None else 0.5 * t[rng.integers(0, T, size=B)]) + rng.standard_normal((B, D)).astype(np.float32)
    return v.astype(np.float32), t


def test_clip_loss_input_forms(loss_mod):
    v, t = _data(40, 96, 1)
    lt = math.log(0.07)
    o = co.clip_loss(v, t, lt)
    # non-contiguous fp32 views, 0-d log_temp, python-float log_temp, log_temp without grad, one-sided gradient requests
    big = torch.zeros(40, 200)
    big[:, ::2][:, :96] = torch.tensor(v)
    vv = big[:, ::2][:, :96]
    assert not vv.is_contiguous()
    vv = vv.detach().requires_grad_(True)
    tt = torch.tensor(t, requires_grad=True)
    l0 = torch.tensor(lt, requires_grad=True)                        # 0-d parameter
    loss = loss_mod.CLIPLoss()(video_features=vv, text_features=tt, log_temp=l0)
    loss.backward()
    assert abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"])
    assert _rel(vv.grad.numpy(), o["dvideo"]) <= 2e-3 and _rel(tt.grad.numpy(), o["dtext"]) <= 2e-3
    assert l0.grad.shape == () and abs(l0.grad.item() - o["dlog_temp"]) <= 2e-3 * abs(o["dlog_temp"])
    # text frozen, temperature a plain float
    v2 = torch.tensor(v, requires_grad=True)
    loss = loss_mod.clip_loss(v2, torch.tensor(t), lt)
    loss.backward()
    assert _rel(v2.grad.numpy(), o["dvideo"]) <= 2e-3
    # video frozen
    t2 = torch.tensor(t, requires_grad=True)
    loss_mod.clip_loss(torch.tensor(v), t2, torch.tensor([lt])).backward()
    assert _rel(t2.grad.numpy(), o["dtext"]) <= 2e-3
    # upstream scale
    v3 = torch.tensor(v, requires_grad=True)
    (3.0 * loss_mod.clip_loss(v3, torch.tensor(t), lt)).backward()
    assert _rel(v3.grad.numpy(), 3.0 * o["dvideo"]) <= 2e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_clip_loss_16_bit_features(loss_mod, dtype):
    v, t = _data(33, 64, 2)
    v16, t16 = torch.tensor(v).to(dtype), torch.tensor(t).to(dtype)
    o = co.clip_loss(v16.float().numpy(), t16.float().numpy(), math.log(0.1))       # the reference casts .float() first
    vv, tt = v16.clone().requires_grad_(True), t16.clone().requires_grad_(True)
    loss = loss_mod.CLIPLoss()(video_features=vv, text_features=tt, log_temp=torch.tensor([math.log(0.1)]))
    loss.backward()
    assert abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"])
    assert vv.grad.dtype == dtype and _rel(vv.grad.float().numpy(), o["dvideo"]) <= 1e-2      # gradient rounded to 16 bits


def test_clip_loss_shape_errors_and_tiny_batches(loss_mod):
    with pytest.raises(ValueError):
        loss_mod.CLIPLoss()(video_features=torch.zeros(4, 8), text_features=torch.zeros(5, 8), log_temp=torch.zeros(1))
    for B in (1, 2, 3):
        v, t = _data(B, 64, 10 + B)
        o = co.clip_loss(v, t, math.log(0.2))
        vv = torch.tensor(v, requires_grad=True)
        loss = loss_mod.clip_loss(vv, torch.tensor(t), math.log(0.2))
        loss.backward()
        assert abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"]) + 1e-6, B     # B = 1: the loss is exactly 0
        assert np.abs(vv.grad.numpy() - o["dvideo"]).max() <= 2e-3 * np.abs(o["dvideo"]).max() + 1e-6, B   # B = 1: exactly 0
    # tau = 0.01 (the usual floor of a learnable CLIP temperature; logits up to +-100) is inside the fixed-shift range
    v, t = _data(16, 64, 20)
    lt = torch.tensor([math.log(0.01)], requires_grad=True)
    loss = loss_mod.CLIPLoss()(video_features=torch.tensor(v), text_features=torch.tensor(t), log_temp=lt)
    loss.backward()
    o = co.clip_loss(v, t, math.log(0.01))
    assert abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"])
    assert abs(lt.grad.item() - o["dlog_temp"]) <= 2e-3 * abs(o["dlog_temp"])
    # an active temperature clamp: the loss is evaluated at the floor and the temperature gradient is zero (:153)
    lt = torch.tensor([math.log(0.02)], requires_grad=True)
    loss = loss_mod.clip_loss(torch.tensor(v), torch.tensor(t), lt, clamp_min=0.05)
    loss.backward()
    o = co.clip_loss(v, t, math.log(0.02), clamp_min=0.05)
    assert abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"]) and lt.grad.item() == 0.0


def test_stable_mode_at_the_reference_clamp_floor(loss_mod):
    """The reference clamps tau only at 1e-4 (utils/loss/contrastive.py:153): logits span +-10^4 and log_softmax's own max
    subtraction keeps it finite (golden clip_clamp_b8_d64, loss 1367.8). The fixed exponent shift of the fast sweep cannot
    represent that range, so dyn_prep switches (on the device, from tau) to the stable mode: running-maximum row sweeps in the
    forward, two exponentials <= 1 in the backward. Host flow + shipped scalar kernels + contract models here; the kernels
    themselves are checked by tests/test_gpu_clip_loss.py on the same golden."""
    from tests.conftest import GOLDEN
    g = np.load(GOLDEN / "clip_clamp_b8_d64.npz")
    v = torch.tensor(g["video"], dtype=torch.float32, requires_grad=True)
    t = torch.tensor(g["text"], dtype=torch.float32, requires_grad=True)
    lt = torch.tensor(g["log_temp"].astype(np.float32), requires_grad=True)
    loss = loss_mod.CLIPLoss()(video_features=v, text_features=t, log_temp=lt)
    loss.backward()
    ref = float(g["f32_loss"])
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (loss.item(), ref)
    for got, key in ((v.grad, "f32_dvideo"), (t.grad, "f32_dtext")):
        assert np.linalg.norm(got.numpy() - g[key]) <= 2e-3 * np.linalg.norm(g[key]), key
    assert lt.grad.item() == 0.0                                  # the clamp is active: no temperature gradient
    # both sides of the switch (tau = 0.0128 for unit vectors), the unclamped legacy class far below it, label smoothing
    for tau, kw in ((0.006, {}), (0.012, {}), (0.014, {}), (0.02, {}), (0.001, {"clamp_min": 0.0}),
                    (0.004, {"label_smoothing": 0.1})):
        o = co.clip_loss(g["video"], g["text"], math.log(tau), label_smoothing=kw.get("label_smoothing", 0.0),
                         clamp_min=None if kw.get("clamp_min") == 0.0 else 1e-4)
        v = torch.tensor(g["video"], dtype=torch.float32, requires_grad=True)
        t = torch.tensor(g["text"], dtype=torch.float32, requires_grad=True)
        lt = torch.tensor([math.log(tau)], dtype=torch.float32, requires_grad=True)
        ours = loss_mod.clip_loss(v, t, lt, **kw)
        ours.backward()
        assert abs(ours.item() - o["loss"]) <= 1e-5 * abs(o["loss"]), tau
        assert np.linalg.norm(v.grad.numpy() - o["dvideo"]) <= 2e-3 * np.linalg.norm(o["dvideo"]), tau
        assert np.linalg.norm(t.grad.numpy() - o["dtext"]) <= 2e-3 * np.linalg.norm(o["dtext"]), tau
        assert abs(lt.grad.item() - o["dlog_temp"]) <= 2e-3 * max(abs(o["dlog_temp"]), 1e-3), tau


def test_label_smoothing_at_tiny_batches(loss_mod):
    """The uniform-target term (eps / N) sum_ij L_ij dominates the error budget at small N: with bf16x3 operands its column
    sums include the lo panels (hi-only sums were off by up to 3.6e-5 relative at N = 8, eps = 0.2)."""
    for seed, N, D, eps, tau in ((1, 8, 64, 0.2, 0.05), (2, 8, 64, 0.2, 0.05), (1, 8, 512, 0.2, 0.05), (3, 5, 96, 0.3, 0.1)):
        r = np.random.default_rng(seed)
        v = r.standard_normal((N, D)).astype(np.float32)
        t = r.standard_normal((N, D)).astype(np.float32)
        o = co.clip_loss(v, t, math.log(tau), label_smoothing=eps)
        vv, tt = torch.tensor(v, requires_grad=True), torch.tensor(t, requires_grad=True)
        lt = torch.tensor([math.log(tau)], requires_grad=True)
        loss = loss_mod.CLIPLoss(label_smoothing=eps)(video_features=vv, text_features=tt, log_temp=lt)
        loss.backward()
        assert abs(loss.item() - o["loss"]) <= 5e-6 * abs(o["loss"]), (seed, N, D, loss.item(), o["loss"])
        assert _rel(vv.grad.numpy(), o["dvideo"]) <= 2e-3 and _rel(tt.grad.numpy(), o["dtext"]) <= 2e-3
        assert abs(lt.grad.item() - o["dlog_temp"]) <= 2e-3 * max(abs(o["dlog_temp"]), 1e-3)


def test_siglip_ragged_rows_without_positives_and_frozen_bias(loss_mod):
    B, T, D = 21, 50, 64
    v, t = _data(B, D, 3, T=T)
    rng = np.random.default_rng(4)
    pm = (rng.random((B, T)) < 0.06).astype(np.float32)
    pm[5] = 0.0                                                       # a video without any positive text
    pw = (pm * rng.choice([1.0, 1.5, 2.5, 3.0], size=(B, T))).astype(np.float32)
    lt = math.log(0.09)
    o = co.siglip_loss(v, t, lt, bias=-6.0, pos_mask=pm, pos_weights=pw)
    mod = loss_mod.SigLIPLoss(bias_init=-6.0, learnable_bias=False)
    assert not isinstance(mod.bias, torch.nn.Parameter)
    vv, tt = torch.tensor(v, requires_grad=True), torch.tensor(t, requires_grad=True)
    l1 = torch.tensor([lt], requires_grad=True)
    loss = mod(video_features=vv, text_features=tt, log_temp=l1, pos_mask=torch.tensor(pm), pos_weights=torch.tensor(pw))
    loss.backward()
    assert abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"])
    assert _rel(vv.grad.numpy(), o["dvideo"]) <= 2e-3 and _rel(tt.grad.numpy(), o["dtext"]) <= 2e-3
    assert abs(l1.grad.item() - o["dlog_temp"]) <= 2e-3 * max(abs(o["dlog_temp"]), 1e-4)
    # boolean / integer masks are accepted like the reference's .float() cast
    loss_b = mod(video_features=torch.tensor(v), text_features=torch.tensor(t), log_temp=lt,
                 pos_mask=torch.tensor(pm).bool(), pos_weights=torch.tensor(pw))
    assert abs(loss_b.item() - o["loss"]) <= 1e-5 * abs(o["loss"])
    with pytest.raises(ValueError):
        mod(video_features=torch.tensor(v), text_features=torch.tensor(t), log_temp=lt, pos_mask=torch.zeros(B, T + 1))


def test_siglip_positive_list_overflow_poisons_the_loss(loss_mod):
    v, t = _data(8, 64, 5, T=40)
    pm = np.zeros((8, 40), np.float32)
    pm[2, :9] = 1.0                                                   # 9 positives in a row, capacity 4
    mod = loss_mod.SigLIPLoss(max_positives_per_row=4)
    loss = mod(video_features=torch.tensor(v), text_features=torch.tensor(t), log_temp=math.log(0.1), pos_mask=torch.tensor(pm))
    assert math.isnan(loss.item())                                    # never silently drops positives
    ok = loss_mod.SigLIPLoss(max_positives_per_row=16)(video_features=torch.tensor(v), text_features=torch.tensor(t),
                                                       log_temp=math.log(0.1), pos_mask=torch.tensor(pm))
    o = co.siglip_loss(v, t, math.log(0.1), bias=-10.0, pos_mask=pm)
    assert abs(ok.item() - o["loss"]) <= 1e-5 * abs(o["loss"])


def test_retrieval_degenerate_shapes(monkeypatch):
    rms = patch_retrieval(build_emul(), monkeypatch.setattr)
    v = torch.tensor(ro.exact_grid_embeddings(7, 16, 1))
    t = torch.tensor(ro.exact_grid_embeddings(3, 16, 2))
    gt = torch.tensor([0, 1, 2, 0, 1, 2, 0])
    sim = ro.similarity(v.numpy(), t.numpy())
    # k larger than the database: the reference's best_indices has min(k, M) columns, larger k report 0 (:89-93)
    r = rms.compute_recall_at_k_streaming(v, t, gt, k_values=[1, 2, 3, 5], device="cpu")
    ranks = ro.gt_ranks(sim, gt.numpy())
    assert r["Recall@5"] == 0.0
    for k in (1, 2, 3):
        assert r[f"Recall@{k}"] == float((ranks <= k).mean() * 100)
    s, i = rms.streaming_topk(v, t, 10)                               # clamped to the 3 texts
    ov, oi = ro.topk_lowest_index(sim, 3)
    assert i.shape == (7, 3) and (i.numpy() == oi).all() and (s.numpy() == ov).all()
    # one video, one text
    m = rms.compute_metrics_streaming(v[:1], t[:1], torch.tensor([0]), k_values=[1])
    assert m["Recall@1"] == 100.0 and m["MRR_V2T"] == 1.0
    with pytest.raises(ValueError):
        rms.streaming_topk(v, t, 0)
