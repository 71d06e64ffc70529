"""GPU parity: SigLIP multi-positive loss vs reference golden vectors and the numpy oracle.
Tolerances (BASELINE.json north_star): loss <= 1e-5 relative, gradients <= 2e-3."""
import math

import numpy as np
import pytest
import torch

from oracle import contrastive_oracle as co
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _run(mod, v, t, lt, pos_mask=None, pos_weights=None):
    mod = mod.to(DEV)
    v = torch.tensor(v, dtype=torch.float32, device=DEV, requires_grad=True)
    t = torch.tensor(t, dtype=torch.float32, device=DEV, requires_grad=True)
    lt = torch.tensor(np.asarray(lt, np.float32).reshape(1), device=DEV, requires_grad=True)
    kw = {}
    if pos_mask is not None:
        kw["pos_mask"] = torch.tensor(pos_mask, dtype=torch.float32, device=DEV)
    if pos_weights is not None:
        kw["pos_weights"] = torch.tensor(pos_weights, dtype=torch.float32, device=DEV)
    loss = mod(video_features=v, text_features=t, log_temp=lt, **kw)
    assert loss.ndim == 0 and loss.requires_grad
    loss.backward()
    torch.cuda.synchronize()
    bias = getattr(mod, "bias", None)
    db = bias.grad.item() if isinstance(bias, torch.nn.Parameter) and bias.grad is not None else None
    return loss.item(), v.grad.cpu().numpy(), t.grad.cpu().numpy(), lt.grad.item(), db


CASES = {
    "siglip_diag_b32_t32_d64": dict(),
    "siglip_mp_b32_t40_d64": dict(),
    "siglip_mp_noweights_b24_t50_d96": dict(positive_weight=2.0, negative_weight=0.5, use_severity_weights=False),
    "siglip_autobalance_b16_t48_d64": dict(auto_balance=True),
    "siglip_bias0_b130_t260_d512": dict(bias_init=-1.0),
    "siglip_entropy_b16_t32_d64": dict(entropy_regularization=True, bias_init=-2.0, min_entropy_threshold=5.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_siglip_matches_reference_golden(name):
    from deepcoro_clip_b200.loss import SigLIPLoss
    g = np.load(GOLDEN / f"{name}.npz")
    pm = g["in_pos_mask"] if "in_pos_mask" in g else None
    pw = g["in_pos_weights"] if "in_pos_weights" in g else None
    loss, dv, dt, dlt, db = _run(SigLIPLoss(**CASES[name]), g["video"], g["text"], g["log_temp"], pm, pw)
    ref = float(g["f32_loss"])
    assert abs(loss - ref) <= 1e-5 * abs(ref), (loss, ref)
    assert _rel(dv, g["f32_dvideo"]) <= 2e-3
    assert _rel(dt, g["f32_dtext"]) <= 2e-3
    rlt = float(g["f32_dlog_temp"].reshape(-1)[0])
    assert abs(dlt - rlt) <= 2e-3 * max(abs(rlt), 1e-4)
    rb = float(g["f32_dbias"])
    assert abs(db - rb) <= 2e-3 * max(abs(rb), 1e-4)


@pytest.mark.parametrize("B,T,D,tau,bias,prec", [(512, 640, 512, 0.087, -10.0, "auto"), (1024, 1024, 512, 0.087, -3.0, "bf16"),
                                                   (300, 2000, 768, 0.07, -5.0, "bf16x3")])
def test_siglip_vs_oracle(B, T, D, tau, bias, prec):
    from deepcoro_clip_b200.loss import SigLIPLoss
    rng = np.random.default_rng(B + T)
    t = rng.standard_normal((T, D)).astype(np.float32)
    v = (0.7 * t[rng.integers(0, T, size=B)] + rng.standard_normal((B, D))).astype(np.float32)
    pm = np.zeros((B, T), np.float32)
    for _ in range(4):
        pm[np.arange(B), rng.integers(0, T, size=B)] = 1.0
    pw = pm * rng.choice([1.0, 1.5, 2.5, 3.0], size=(B, T)).astype(np.float32)
    loss, dv, dt, dlt, db = _run(SigLIPLoss(bias_init=bias, precision=prec), v, t, math.log(tau), pm, pw)
    o = co.siglip_loss(v, t, math.log(tau), bias=bias, pos_mask=pm, pos_weights=pw)
    ltol, gtol = 1e-5, 2e-3          # north_star tolerances for every operand precision
    assert abs(loss - o["loss"]) <= ltol * abs(o["loss"]), (loss, o["loss"])
    assert _rel(dv, o["dvideo"]) <= gtol
    assert _rel(dt, o["dtext"]) <= gtol
    assert abs(dlt - o["dlog_temp"]) <= gtol * max(abs(o["dlog_temp"]), 1e-4)
    assert abs(db - o["dbias"]) <= gtol * max(abs(o["dbias"]), 1e-4)


def test_siglip_no_grad_forward_and_overflow_poison():
    from deepcoro_clip_b200.loss import SigLIPLoss
    rng = np.random.default_rng(9)
    v = rng.standard_normal((200, 128)).astype(np.float32)
    t = rng.standard_normal((300, 128)).astype(np.float32)
    pm = (rng.random((200, 300)) < 0.02).astype(np.float32)
    mod = SigLIPLoss(bias_init=-2.0).to(DEV)
    with torch.no_grad():
        l = mod(torch.tensor(v, device=DEV), torch.tensor(t, device=DEV), torch.tensor(math.log(0.1), device=DEV),
                pos_mask=torch.tensor(pm, device=DEV))
    o = co.siglip_loss(v, t, math.log(0.1), bias=-2.0, pos_mask=pm, want_grads=False)
    assert abs(l.item() - o["loss"]) <= 1e-5 * abs(o["loss"])
    # more positives in a row than the compacted list can hold => NaN, never a silently wrong loss
    small = SigLIPLoss(bias_init=-2.0, max_positives_per_row=2).to(DEV)
    pm2 = np.zeros((200, 300), np.float32)
    pm2[0, :5] = 1
    with torch.no_grad():
        l2 = small(torch.tensor(v, device=DEV), torch.tensor(t, device=DEV), torch.tensor(math.log(0.1), device=DEV),
                   pos_mask=torch.tensor(pm2, device=DEV))
    assert math.isnan(l2.item())


# ---- the SigLIP classes the reference keeps importable next to the unified loss ----
def _variant(name):
    from deepcoro_clip_b200 import loss as L
    return {
        "pairwise_mp_b24_t40_d64": lambda: L.SiglipPairwiseFeatureLoss(positive_weight=1.5, negative_weight=0.7),
        "pairwise_entropy_auto_b16_t48_d64": lambda: L.SiglipPairwiseFeatureLoss(
            auto_positive_weight=True, entropy_regularization=True, entropy_weight=0.2, min_entropy_threshold=6.0),
        "bce2_b40_d96": lambda: L.SigLIP2BCELoss(),
        "bce2_ls_noclamp_b32_d64": lambda: L.SigLIP2BCELoss(bias_init=-4.0, label_smoothing=0.1),
        "mp2_ls_b20_t36_d64": lambda: L.SigLIP2MultiPositiveBCELoss(bias_init=-3.0, positive_weight=2.0,
                                                                    negative_weight=0.5, label_smoothing=0.2),
        "mp2_diag_b18_t30_d64": lambda: L.SigLIP2MultiPositiveBCELoss(bias_init=-5.0),
    }[name]()


@pytest.mark.parametrize("name", ["pairwise_mp_b24_t40_d64", "pairwise_entropy_auto_b16_t48_d64", "bce2_b40_d96",
                                  "bce2_ls_noclamp_b32_d64", "mp2_ls_b20_t36_d64", "mp2_diag_b18_t30_d64"])
def test_siglip_variants_match_reference_golden(name):
    g = np.load(GOLDEN / f"{name}.npz")
    pm = g["in_pos_mask"] if "in_pos_mask" in g else None
    pw = g["in_pos_weights"] if "in_pos_weights" in g else None
    loss, dv, dt, dlt, db = _run(_variant(name), g["video"], g["text"], g["log_temp"], pm, pw)
    ref = float(g["f32_loss"])
    assert abs(loss - ref) <= 1e-5 * abs(ref), (loss, ref)
    assert _rel(dv, g["f32_dvideo"]) <= 2e-3
    assert _rel(dt, g["f32_dtext"]) <= 2e-3
    rlt = float(g["f32_dlog_temp"].reshape(-1)[0])
    assert abs(dlt - rlt) <= 2e-3 * max(abs(rlt), 1e-4)
    if "f32_dbias" in g:
        rb = float(g["f32_dbias"])
        assert abs(db - rb) <= 2e-3 * max(abs(rb), 1e-4)


@pytest.mark.parametrize("B,T,D,tau,bias,thr,prec", [(384, 512, 512, 0.05, -2.0, 7.0, "auto"),
                                                       (1024, 1536, 256, 0.087, -10.0, 8.0, "bf16")])
def test_siglip_entropy_vs_oracle(B, T, D, tau, bias, thr, prec):
    """Entropy regulariser with an ACTIVE deficit (threshold above ln T would always fire; here thr is chosen so the
    penalty and its gradient are non-zero) against the numpy oracle, plus the diagnostics dict."""
    from deepcoro_clip_b200.loss import SigLIPLoss
    rng = np.random.default_rng(B * 7 + T)
    t = rng.standard_normal((T, D)).astype(np.float32)
    v = (0.9 * t[rng.integers(0, T, size=B)] + 0.6 * rng.standard_normal((B, D))).astype(np.float32)
    pm = np.zeros((B, T), np.float32)
    for _ in range(3):
        pm[np.arange(B), rng.integers(0, T, size=B)] = 1.0
    mod = SigLIPLoss(bias_init=bias, precision=prec, entropy_regularization=True, entropy_weight=0.3,
                     min_entropy_threshold=thr)
    loss, dv, dt, dlt, db = _run(mod, v, t, math.log(tau), pm, None)
    o = co.siglip_loss(v, t, math.log(tau), bias=bias, pos_mask=pm, entropy_regularization_on=True, entropy_weight=0.3,
                       min_entropy_threshold=thr)
    assert o["entropy_diagnostics"]["entropy_deficit"] > 0.05          # the regulariser is really active
    # explicit plain-bf16 operands (non-default at these sizes: "auto" = bf16x3): the row-softmax of the entropy term sees the
    # 2^-9 operand rounding of every logit, which a 1024-row batch does not average away
    ltol, gtol = (1e-5, 2e-3) if prec != "bf16" else (1e-5, 6e-3)
    assert abs(loss - o["loss"]) <= ltol * abs(o["loss"]), (loss, o["loss"])
    assert _rel(dv, o["dvideo"]) <= gtol
    assert _rel(dt, o["dtext"]) <= gtol
    assert abs(dlt - o["dlog_temp"]) <= gtol * max(abs(o["dlog_temp"]), 1e-4)
    assert abs(db - o["dbias"]) <= gtol * max(abs(o["dbias"]), 1e-4)
    d = mod.get_entropy_diagnostics()
    assert set(d) == {"entropy_mean", "entropy_min", "entropy_max", "entropy_normalized", "entropy_deficit",
                      "bce_loss", "entropy_loss"}
    # per-row extrema see the operand rounding of a single row: plain bf16 operands move one row's entropy by ~1e-3
    dtol = 2e-4 if prec != "bf16" else 5e-3
    for k in ("entropy_mean", "entropy_min", "entropy_max", "entropy_normalized", "entropy_deficit"):
        assert abs(d[k] - o["entropy_diagnostics"][k]) <= dtol * max(1.0, abs(o["entropy_diagnostics"][k])), k
    assert abs(d["bce_loss"] - o["bce_loss"]) <= ltol * 10 * abs(o["bce_loss"])
    assert abs(d["entropy_loss"] - 0.3 * o["entropy_diagnostics"]["entropy_deficit"]) <= 0.3 * dtol


def test_siglip_entropy_inactive_and_no_grad():
    """Threshold below the mean entropy: zero penalty, gradients equal the plain BCE ones; no_grad forward works."""
    from deepcoro_clip_b200.loss import SigLIPLoss
    rng = np.random.default_rng(77)
    v = rng.standard_normal((150, 96)).astype(np.float32)
    t = rng.standard_normal((220, 96)).astype(np.float32)
    pm = (rng.random((150, 220)) < 0.02).astype(np.float32)
    plain = _run(SigLIPLoss(bias_init=-3.0), v, t, math.log(0.1), pm, None)
    ent = _run(SigLIPLoss(bias_init=-3.0, entropy_regularization=True, min_entropy_threshold=0.5), v, t, math.log(0.1),
               pm, None)
    assert abs(plain[0] - ent[0]) <= 1e-6 * abs(plain[0])
    assert _rel(ent[1], plain[1]) <= 1e-5 and _rel(ent[2], plain[2]) <= 1e-5
    mod = SigLIPLoss(bias_init=-3.0, entropy_regularization=True, entropy_weight=0.25, min_entropy_threshold=9.0).to(DEV)
    with torch.no_grad():
        l = mod(torch.tensor(v, device=DEV), torch.tensor(t, device=DEV), torch.tensor(math.log(0.1), device=DEV),
                pos_mask=torch.tensor(pm, device=DEV))
    o = co.siglip_loss(v, t, math.log(0.1), bias=-3.0, pos_mask=pm, entropy_regularization_on=True,
                       entropy_weight=0.25, min_entropy_threshold=9.0, want_grads=False)
    assert abs(l.item() - o["loss"]) <= 1e-5 * abs(o["loss"])
