"""GPU parity: fused CLIP / legacy / gated losses (through the module API -> C ABI -> sm_100a kernels) against
the reference's golden vectors and the numpy oracle. Tolerances from BASELINE.json north_star:
loss <= 1e-5 relative (fp32 accumulation), gradients <= 2e-3 (bf16 gradient operands)."""
import math

import numpy as np
import pytest
import torch

from oracle import contrastive_oracle as co
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

LOSS_RTOL = 1e-5
GRAD_RTOL = 2e-3


def _rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _run(mod, v, t, lt):
    dev = torch.device("cuda:0")
    v = torch.tensor(v, dtype=torch.float32, device=dev, requires_grad=True)
    t = torch.tensor(t, dtype=torch.float32, device=dev, requires_grad=True)
    lt = torch.tensor(np.asarray(lt, np.float32).reshape(1), device=dev, requires_grad=True)
    loss = mod(video_features=v, text_features=t, log_temp=lt)
    assert loss.ndim == 0 and loss.requires_grad and loss.device.type == "cuda"
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), v.grad.cpu().numpy(), t.grad.cpu().numpy(), lt.grad.item()


CASES = {
    "clip_c1_b64_d512": ("CLIPLoss", {}),
    "clip_ls_b48_d96": ("CLIPLoss", {"label_smoothing": 0.1}),
    # tau clamped at the reference's floor 1e-4 (contrastive.py:153): logits +-1e4, stable softmax mode on the device
    "clip_clamp_b8_d64": ("CLIPLoss", {}),
    "clip_b300_d200": ("CLIPLoss", {}),
    "contrastive_legacy_b32_d128": ("ContrastiveLoss", {}),
    "gated_siglip_legacy_b40_d128": ("SiglipLoss", {}),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_loss_matches_reference_golden(name):
    from deepcoro_clip_b200 import loss as L
    g = np.load(GOLDEN / f"{name}.npz")
    cls, kw = CASES[name]
    loss, dv, dt, dlt = _run(getattr(L, cls)(**kw), g["video"], g["text"], g["log_temp"])
    ref = float(g["f32_loss"])
    assert abs(loss - ref) <= LOSS_RTOL * abs(ref), (loss, ref)
    assert _rel(dv, g["f32_dvideo"]) <= GRAD_RTOL
    assert _rel(dt, g["f32_dtext"]) <= GRAD_RTOL
    rlt = float(g["f32_dlog_temp"].reshape(-1)[0])
    assert abs(dlt - rlt) <= GRAD_RTOL * max(abs(rlt), 1e-3)


@pytest.mark.parametrize("N,D,tau,prec", [(64, 512, 0.07, "auto"), (1000, 512, 0.0588, "bf16x3"),
                                           (1000, 512, 0.0588, "bf16"), (2048, 768, 0.0588, "bf16"),
                                           (333, 100, 0.1, "auto")])
def test_clip_vs_oracle(N, D, tau, prec):
    from deepcoro_clip_b200.loss import CLIPLoss
    rng = np.random.default_rng(N + D)
    v = rng.standard_normal((N, D)).astype(np.float32)
    t = (0.5 * v + rng.standard_normal((N, D))).astype(np.float32)      # correlated pairs: peaked diagonal
    lt = math.log(tau)
    loss, dv, dt, dlt = _run(CLIPLoss(precision=prec), v, t, lt)
    o = co.clip_loss(v, t, lt)
    # plain bf16 operands perturb each logit by ~1e-4/tau, but the target logits enter in fp32 and the errors inside the
    # log-sum-exps average out: 1e-5 for every precision
    assert abs(loss - o["loss"]) <= LOSS_RTOL * abs(o["loss"]), (loss, o["loss"])
    gt = GRAD_RTOL
    assert _rel(dv, o["dvideo"]) <= gt
    assert _rel(dt, o["dtext"]) <= gt
    assert abs(dlt - o["dlog_temp"]) <= gt * max(abs(o["dlog_temp"]), 1e-3)


def _small_tau_tol(ref, tau):
    """1e-5 relative (north_star) plus ONE fp32 ulp of the largest logit 1/tau: the loss of an almost separated batch is a
    difference of logits of that size, which the fp32 reference itself resolves no better than this (its log_softmax
    subtracts fp32 logits S / tau). The checker is the float64 oracle."""
    return LOSS_RTOL * abs(ref) + 2.0 ** -23 / tau


@pytest.mark.parametrize("N,D,tau,corr,kw", [
    (300, 200, 0.012, 0.1, {}), (300, 200, 0.004, 0.1, {}), (257, 64, 0.003, 0.05, {}),
    (130, 96, 0.005, 0.1, {"label_smoothing": 0.1}), (1500, 512, 0.006, 0.1, {}),
    # almost perfectly separated batches (loss ~1e-3 ... 1e-13, vanishing gradients): the loss keeps its accuracy
    (300, 200, 0.012, 0.5, {}), (64, 512, 2e-4, 0.5, {}), (257, 64, 0.001, 0.5, {}),
])
def test_small_tau_stable_mode_vs_oracle(N, D, tau, corr, kw):
    """Temperatures below the fixed-shift window (tau < ~0.0128), down to the reference's floor: the device switches to
    the running-maximum forward sweeps and the two-exponential backward. bf16x3 operands (precision='auto' at these
    sizes). Gradients are compared where they do not vanish (weakly correlated pairs) and for tau >= 0.003: the
    error-compensated operands carry 16 mantissa bits, i.e. a logit error of ~5e-6 / tau, which is a relative softmax
    error of 2e-3 at tau = 0.0025 (the fp32 reference resolves logits to ~1e-7 / tau); below that only the loss, the
    finiteness and the clamp behaviour are asserted."""
    from deepcoro_clip_b200.loss import CLIPLoss
    rng = np.random.default_rng(N + D)
    v = rng.standard_normal((N, D)).astype(np.float32)
    t = (corr * v + rng.standard_normal((N, D))).astype(np.float32)
    lt = math.log(tau)
    loss, dv, dt, dlt = _run(CLIPLoss(**kw), v, t, lt)
    o = co.clip_loss(v, t, lt, **kw)
    assert math.isfinite(loss) and abs(loss - o["loss"]) <= _small_tau_tol(o["loss"], tau), (loss, o["loss"])
    assert np.isfinite(dv).all() and np.isfinite(dt).all() and math.isfinite(dlt)
    if corr <= 0.1 and tau >= 0.003:
        assert _rel(dv, o["dvideo"]) <= GRAD_RTOL and _rel(dt, o["dtext"]) <= GRAD_RTOL
        assert abs(dlt - o["dlog_temp"]) <= GRAD_RTOL * max(abs(o["dlog_temp"]), 1e-3)


@pytest.mark.parametrize("cls,tau", [("ContrastiveLoss", 0.003), ("SiglipLoss", 0.002), ("SiglipLoss", 0.0005)])
def test_small_tau_legacy_classes(cls, tau):
    """The legacy classes apply no temperature floor at all (utils/loss/losses.py:53, 146, 198)."""
    from deepcoro_clip_b200 import loss as L
    rng = np.random.default_rng(7)
    v = rng.standard_normal((200, 128)).astype(np.float32)
    t = (0.1 * v + rng.standard_normal((200, 128))).astype(np.float32)
    loss, dv, dt, dlt = _run(getattr(L, cls)(), v, t, math.log(tau))
    o = co.clip_loss(v, t, math.log(tau), clamp_min=None, gated=cls == "SiglipLoss")
    assert math.isfinite(loss) and abs(loss - o["loss"]) <= _small_tau_tol(o["loss"], tau), (loss, o["loss"])
    assert np.isfinite(dv).all() and np.isfinite(dt).all() and math.isfinite(dlt)
    if tau >= 0.002:          # gated logits span 0.27x the plain range: the 16-bit operands resolve them at this tau
        assert _rel(dv, o["dvideo"]) <= GRAD_RTOL and _rel(dt, o["dtext"]) <= GRAD_RTOL
        assert abs(dlt - o["dlog_temp"]) <= GRAD_RTOL * max(abs(o["dlog_temp"]), 1e-3)


@pytest.mark.parametrize("N,D,prec", [(2048, 512, "bf16"), (1000, 768, "bf16"), (700, 128, "bf16"), (500, 200, "bf16x3"),
                                      (900, 384, "bf16")])
def test_stable_mode_equals_fixed_shift_mode_inside_the_window(N, D, prec):
    """Both modes are valid at tau = 0.0588: forcing the stable sweeps / two-exponential gradient through every backward
    kernel family (64-row pairs, 128-row pairs, single CTA) must reproduce the fixed-shift results on the same operands."""
    from deepcoro_clip_b200.loss import clip_loss
    g = torch.Generator().manual_seed(N)
    v0 = torch.randn(N, D, generator=g); t0 = 0.5 * v0 + torch.randn(N, D, generator=g)
    res = []
    for stable in (False, True):
        v = v0.to(DEV).requires_grad_(True); t = t0.to(DEV).requires_grad_(True)
        lt = torch.tensor([math.log(0.0588)], device=DEV, requires_grad=True)
        loss = clip_loss(v, t, lt, precision=prec, stable=stable)
        loss.backward()
        res.append((loss.item(), v.grad.clone(), t.grad.clone(), lt.grad.item()))
    (l0, dv0, dt0, g0), (l1, dv1, dt1, g1) = res
    assert abs(l0 - l1) <= 2e-6 * abs(l0), (l0, l1)
    # the same bf16 gradient operand up to the rounding of two differently formed fp32 values
    assert ((dv0 - dv1).norm() / dv0.norm()).item() <= 1e-3 and ((dt0 - dt1).norm() / dt0.norm()).item() <= 1e-3
    assert abs(g0 - g1) <= 1e-4 * abs(g0)


def test_clip_no_grad_and_bf16_inputs():
    from deepcoro_clip_b200.loss import CLIPLoss
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    v = torch.randn(256, 512, generator=g).to(dev)
    t = torch.randn(256, 512, generator=g).to(dev)
    lt = torch.tensor(math.log(0.07), device=dev)          # 0-d, no grad
    with torch.no_grad():
        l32 = CLIPLoss()(v, t, lt)
        lbf = CLIPLoss()(v.bfloat16(), t.bfloat16(), lt)
    o = co.clip_loss(v.cpu().numpy(), t.cpu().numpy(), math.log(0.07), want_grads=False)
    assert abs(l32.item() - o["loss"]) <= LOSS_RTOL * o["loss"]
    ob = co.clip_loss(v.bfloat16().float().cpu().numpy(), t.bfloat16().float().cpu().numpy(), math.log(0.07),
                      want_grads=False)
    assert abs(lbf.item() - ob["loss"]) <= LOSS_RTOL * ob["loss"]
    with torch.autocast("cuda", dtype=torch.bfloat16):      # runner calls the loss under autocast (runner.py:1316)
        la = CLIPLoss()(v, t, lt)
    assert la.dtype == torch.float32 and abs(la.item() - l32.item()) < 1e-6


def test_zero_rows_do_not_nan():
    """F.normalize eps: an all-zero embedding row stays zero (reference 'sanitise to zeros' path)."""
    from deepcoro_clip_b200.loss import CLIPLoss
    dev = torch.device("cuda:0")
    v = torch.randn(64, 128, device=dev)
    t = torch.randn(64, 128, device=dev)
    v[5] = 0
    v.requires_grad_(True)
    loss = CLIPLoss()(v, t, torch.tensor([math.log(0.07)], device=dev))
    loss.backward()
    o = co.clip_loss(v.detach().cpu().numpy(), t.cpu().numpy(), math.log(0.07), want_grads=False)
    assert torch.isfinite(loss) and abs(loss.item() - o["loss"]) <= 1e-5 * o["loss"]
    assert torch.isfinite(v.grad).all()


def test_cpu_tensor_is_rejected():
    from deepcoro_clip_b200.loss import CLIPLoss
    from deepcoro_clip_b200._lib import B200ClipError
    with pytest.raises(B200ClipError):
        CLIPLoss()(torch.randn(4, 8), torch.randn(4, 8), torch.tensor(0.0))


def test_graphed_step_matches_eager():
    """GraphedLossStep (forward + backward captured in a CUDA graph) reproduces the eager module (up to the order of fp32 atomic sums) on new
    inputs copied into its static buffers, for several replays."""
    import math
    from deepcoro_clip_b200 import GraphedLossStep
    from deepcoro_clip_b200.loss import CLIPLoss
    g = torch.Generator().manual_seed(3)
    mod = CLIPLoss(precision="bf16")
    lt = torch.tensor([math.log(0.07)], device=DEV)
    v0 = torch.randn(1024, 512, generator=g).to(DEV); t0 = torch.randn(1024, 512, generator=g).to(DEV)
    gs = GraphedLossStep(mod, v0, t0, lt)
    for trial in range(3):
        v = torch.randn(1024, 512, generator=g).to(DEV); t = (0.5 * v.cpu() + torch.randn(1024, 512, generator=g)).to(DEV)
        loss, dv, dt, dlt = gs.step(v, t)
        ve = v.clone().requires_grad_(True); te = t.clone().requires_grad_(True); le = lt.clone().requires_grad_(True)
        le_loss = mod(video_features=ve, text_features=te, log_temp=le)
        le_loss.backward()
        torch.cuda.synchronize()
        # row / column sums are accumulated with fp32 atomics: equal up to summation order, run to run
        assert abs(loss.item() - le_loss.item()) <= 1e-6 * abs(le_loss.item())
        # dX accumulates with red.global.add in a data-dependent order: equal up to fp32 summation order
        assert float((dv - ve.grad).abs().max()) <= 2e-5 * float(ve.grad.abs().max())
        assert float((dt - te.grad).abs().max()) <= 2e-5 * float(te.grad.abs().max())
        assert abs(dlt.item() - le.grad.item()) <= 1e-6 * abs(le.grad.item())
    # a learnable temperature is handed to every step (the graph holds a private copy): new value, new loss
    lt2 = torch.tensor([math.log(0.1)], device=DEV)
    loss2 = gs.step(v, t, log_temp=lt2)[0].item()
    ref2 = mod(video_features=v, text_features=t, log_temp=lt2).item()
    assert abs(loss2 - ref2) <= 1e-6 * abs(ref2) and abs(loss2 - le_loss.item()) > 1e-3
