"""Full-size parity (BASELINE.json configs 2, 4, 5) on the GPU. The numpy oracle cannot finish these sizes in seconds,
so the checker here is the RESTATED oracle of SURVEY 8c(3): the same formulas in plain torch, chunked, in float64 on the
GPU (test infrastructure only), plus size-independent properties: scale invariance (gradient rows orthogonal to their
inputs), shard additivity of the rank counts, and brute-force spot checks of individual rows."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _clip_fp64(v, t, log_tau, rows):
    """Chunked float64 restatement of CLIPLoss (contrastive.py:146-164): loss, plus dV / dT for the sampled rows."""
    vh = torch.nn.functional.normalize(v.double(), dim=-1)
    th = torch.nn.functional.normalize(t.double(), dim=-1)
    N = v.shape[0]
    tau = math.exp(log_tau)
    r = torch.empty(N, dtype=torch.float64, device=v.device)
    colmax = torch.full((N,), -float("inf"), dtype=torch.float64, device=v.device)
    colsum = torch.zeros(N, dtype=torch.float64, device=v.device)
    diag = torch.empty(N, dtype=torch.float64, device=v.device)
    step = 4096
    for a in range(0, N, step):
        L = vh[a:a + step] @ th.T / tau
        r[a:a + step] = torch.logsumexp(L, dim=1)
        diag[a:a + step] = L[torch.arange(L.shape[0]), torch.arange(a, a + L.shape[0])]
        m = torch.maximum(colmax, L.max(dim=0).values)
        colsum = colsum * torch.exp(colmax - m) + torch.exp(L - m).sum(dim=0)
        colmax = m
    c = colmax + torch.log(colsum)
    loss = 0.5 * ((r - diag).mean() + (c - diag).mean())
    # gradient rows for the sample (Appendix A.1): G = (exp(L - r) + exp(L - c) - 2I) / (2N)
    L = vh[rows] @ th.T / tau
    G = (torch.exp(L - r[rows, None]) + torch.exp(L - c[None, :])) / (2 * N)
    G[torch.arange(len(rows)), rows] -= 1.0 / N
    dvh = G @ th / tau
    vn = v[rows].double().norm(dim=1, keepdim=True)
    dv = (dvh - (dvh * vh[rows]).sum(1, keepdim=True) * vh[rows]) / vn
    Lt = th[rows] @ vh.T / tau                                     # rows of L^T
    Gt = (torch.exp(Lt - c[rows, None]) + torch.exp(Lt - r[None, :])) / (2 * N)
    Gt[torch.arange(len(rows)), rows] -= 1.0 / N
    dth = Gt @ vh / tau
    tn = t[rows].double().norm(dim=1, keepdim=True)
    dt = (dth - (dth * th[rows]).sum(1, keepdim=True) * th[rows]) / tn
    return loss.item(), dv, dt


@pytest.mark.parametrize("N,D", [(32768, 512), (32768, 768)])
def test_clip_full_size_vs_fp64_restatement(N, D):
    """BASELINE config 5 / the metric's configuration: global batch 32,768, bf16 operands."""
    from deepcoro_clip_b200.loss import CLIPLoss
    g = torch.Generator(device=DEV).manual_seed(5)
    v = torch.randn(N, D, device=DEV, generator=g)
    t = 0.3 * v + torch.randn(N, D, device=DEV, generator=g)
    v.requires_grad_(True); t.requires_grad_(True)
    log_tau = math.log(0.0588)
    lt = torch.tensor([log_tau], device=DEV, requires_grad=True)
    loss = CLIPLoss(precision="bf16")(video_features=v, text_features=t, log_temp=lt)
    loss.backward()
    rows = torch.randint(0, N, (256,), device=DEV, generator=g)
    ref, dv, dt = _clip_fp64(v.detach(), t.detach(), log_tau, rows)
    rel = abs(loss.item() - ref) / abs(ref)
    assert rel <= 1e-5, (loss.item(), ref, rel)                       # north_star: loss within 1e-5 relative
    gv = ((v.grad[rows].double() - dv).norm() / dv.norm()).item()
    gt = ((t.grad[rows].double() - dt).norm() / dt.norm()).item()
    assert gv <= 2e-3 and gt <= 2e-3, (gv, gt)                        # north_star: gradients within 2e-3 (bf16)
    # size-independent property: the loss is invariant to the scale of every row => grad_i . x_i = 0
    ov = ((v.grad * v.detach()).sum(1).abs() / (v.grad.norm(dim=1) * v.detach().norm(dim=1))).max().item()
    ot = ((t.grad * t.detach()).sum(1).abs() / (t.grad.norm(dim=1) * t.detach().norm(dim=1))).max().item()
    assert ov <= 1e-4 and ot <= 1e-4, (ov, ot)
    assert torch.isfinite(lt.grad).all()


@pytest.mark.parametrize("N", [4097, 8192, 16384])
def test_clip_auto_precision_between_4k_and_32k(N):
    """precision="auto" switches to plain bf16 operands above 4096^2 pairs (loss.py:_pick_precision): the default must
    still meet 1e-5 / 2e-3 there (the target logits enter in fp32, rowdot_raw + the finalize correction)."""
    from deepcoro_clip_b200.loss import CLIPLoss
    D = 512
    g = torch.Generator(device=DEV).manual_seed(N)
    v = torch.randn(N, D, device=DEV, generator=g)
    t = 0.3 * v + torch.randn(N, D, device=DEV, generator=g)
    v.requires_grad_(True); t.requires_grad_(True)
    log_tau = math.log(0.0588)
    lt = torch.tensor([log_tau], device=DEV, requires_grad=True)
    loss = CLIPLoss()(video_features=v, text_features=t, log_temp=lt)
    loss.backward()
    rows = torch.randint(0, N, (256,), device=DEV, generator=g)
    ref, dv, dt = _clip_fp64(v.detach(), t.detach(), log_tau, rows)
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (loss.item(), ref, abs(loss.item() - ref) / abs(ref))
    assert ((v.grad[rows].double() - dv).norm() / dv.norm()).item() <= 2e-3
    assert ((t.grad[rows].double() - dt).norm() / dt.norm()).item() <= 2e-3


def test_clip_full_size_stable_mode():
    """The stable mode at the metric's size (32,768 x 32,768, D = 512, plain bf16 operands) through bw3_kernel: forced at
    tau = 0.0588 it must agree with the float64 restatement like the fixed-shift mode does."""
    from deepcoro_clip_b200.loss import clip_loss
    N, D = 32768, 512
    g = torch.Generator(device=DEV).manual_seed(5)
    v = torch.randn(N, D, device=DEV, generator=g)
    t = 0.3 * v + torch.randn(N, D, device=DEV, generator=g)
    v.requires_grad_(True); t.requires_grad_(True)
    log_tau = math.log(0.0588)
    lt = torch.tensor([log_tau], device=DEV, requires_grad=True)
    loss = clip_loss(v, t, lt, precision="bf16", stable=True)
    loss.backward()
    rows = torch.randint(0, N, (256,), device=DEV, generator=g)
    ref, dv, dt = _clip_fp64(v.detach(), t.detach(), log_tau, rows)
    assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (loss.item(), ref)
    assert ((v.grad[rows].double() - dv).norm() / dv.norm()).item() <= 2e-3
    assert ((t.grad[rows].double() - dt).norm() / dt.norm()).item() <= 2e-3
    assert torch.isfinite(lt.grad).all()


def test_siglip_c2_size_vs_fp64_restatement():
    """BASELINE config 2: 8,192 x 8,192 pairs, D = 512, 4 positives per row with severity weights."""
    from deepcoro_clip_b200.loss import SigLIPLoss
    B = T = 8192; D = 512
    g = torch.Generator(device=DEV).manual_seed(1)
    t = torch.randn(T, D, device=DEV, generator=g)
    v = 0.5 * t + torch.randn(B, D, device=DEV, generator=g)
    pm = torch.zeros(B, T, device=DEV)
    pm[torch.arange(B), torch.arange(B)] = 1.0
    for _ in range(3):
        pm[torch.arange(B, device=DEV), torch.randint(0, T, (B,), device=DEV, generator=g)] = 1.0
    sev = torch.tensor([1.0, 1.5, 2.5, 3.0], device=DEV)
    pw = pm * sev[torch.randint(0, 4, (B, T), device=DEV, generator=g)]
    v.requires_grad_(True); t.requires_grad_(True)
    log_tau, bias = math.log(0.087), -10.0
    lt = torch.tensor([log_tau], device=DEV, requires_grad=True)
    mod = SigLIPLoss(bias_init=bias, precision="bf16").to(DEV)
    loss = mod(v, t, lt, pos_mask=pm, pos_weights=pw)
    loss.backward()
    # float64 restatement (contrastive.py:259-303) with autograd
    v2 = v.detach().double().requires_grad_(True); t2 = t.detach().double().requires_grad_(True)
    lt2 = torch.tensor(log_tau, dtype=torch.float64, device=DEV, requires_grad=True)
    b2 = torch.tensor(bias, dtype=torch.float64, device=DEV, requires_grad=True)
    vh = torch.nn.functional.normalize(v2, dim=-1); th = torch.nn.functional.normalize(t2, dim=-1)
    L = (vh @ th.T / torch.exp(lt2).clamp(min=1e-4) + b2).clamp(-30, 30)
    y = pm.double().clamp(0, 1)
    w = torch.where(y > 0.5, pw.double() * 1.0, torch.ones_like(y))
    ref = (w * torch.nn.functional.binary_cross_entropy_with_logits(L, y, reduction="none")).mean()
    ref.backward()
    # north_star tolerances: loss 1e-5 relative, gradients 2e-3
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()), (loss.item(), ref.item())
    assert ((v.grad.double() - v2.grad).norm() / v2.grad.norm()).item() <= 2e-3
    assert ((t.grad.double() - t2.grad).norm() / t2.grad.norm()).item() <= 2e-3
    assert abs(lt.grad.item() - lt2.grad.item()) <= 2e-3 * abs(lt2.grad.item())
    assert abs(mod.bias.grad.item() - b2.grad.item()) <= 2e-3 * abs(b2.grad.item())


def test_retrieval_c4_full_sweep_properties():
    """BASELINE config 4: 203,808 x 32,473, D = 512 exact-grid embeddings (every dot product exact in fp32)."""
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_recall_at_k_streaming, streaming_topk
    N, M, D = 203808, 32473, 512
    g = torch.Generator(device=DEV).manual_seed(3)
    v = torch.randint(-127, 128, (N, D), device=DEV, generator=g).float() / 128
    t = torch.randint(-127, 128, (M, D), device=DEV, generator=g).float() / 128
    t[torch.randint(0, M, (300,), device=DEV, generator=g)] = t[torch.randint(0, M, (300,), device=DEV, generator=g)]  # exact ties
    gt = torch.randint(0, M, (N,), device=DEV, generator=g)
    keep = []
    r = compute_recall_at_k_streaming(v, t, gt, k_values=[1, 5, 10], _counts_out=keep)
    counts = keep[0]
    # brute force on a random sample of rows (fp32 matmul of exact-grid values is exact)
    rows = torch.randint(0, N, (1024,), device=DEV, generator=g)
    sim = v[rows] @ t.T
    sg = sim.gather(1, gt[rows][:, None])
    cols = torch.arange(M, device=DEV)[None, :]
    ref = ((sim > sg) | ((sim == sg) & (cols < gt[rows][:, None]))).sum(1)
    assert (ref.int() == counts[rows]).all()
    # shard additivity ("checksum of checksums"): counts over two text shards add up to the full-sweep counts
    half = M // 2 + 37
    from deepcoro_clip_b200 import retrieval_metrics_streaming as rms
    vop, top, _, _, _ = rms._operands(v, t, False, "auto")
    c1, _, _ = rms._sweep(vop, top, gt, 0, False, _shard=(0, half))
    c2, _, _ = rms._sweep(vop, top, gt, 0, False, _shard=(half, M))
    assert (c1 + c2 == counts).all()
    # recall numerators are exactly the number of rows with rank <= k
    for k in (1, 5, 10):
        assert r[f"Recall@{k}"] == (counts < k).sum().item() / N * 100
    # top-10 of a slab equals torch.topk on tie-free rows and the lowest-index rule on tied rows
    s, i = streaming_topk(v[:2048], t, 10)
    st, it = torch.topk(v[:2048] @ t.T, 11, dim=1)
    tie_free = (st[:, :-1] != st[:, 1:]).all(dim=1)
    assert (s == st[:, :10]).all()
    assert (i[tie_free] == it[tie_free][:, :10]).all()
