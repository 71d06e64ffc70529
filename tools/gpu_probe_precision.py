"""Measured loss / gradient errors of the default precision policy between 4k^2 and 32k^2 pairs:
CLIP at N in {4097, 8192, 16384} and SigLIP at the C2 shape, plain bf16 vs bf16x3 operands, against the float64
restatements of tests/test_gpu_fullsize.py. Prints one JSON object; nothing here is a benchmark."""
import json
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests.test_gpu_fullsize import _clip_fp64  # noqa: E402

DEV = "cuda:0"
out = {}


def clip_case(N, D, prec, tau=0.0588, corr=0.3):
    from deepcoro_clip_b200.loss import CLIPLoss
    g = torch.Generator(device=DEV).manual_seed(N + D)
    v = torch.randn(N, D, device=DEV, generator=g)
    t = corr * v + torch.randn(N, D, device=DEV, generator=g)
    v.requires_grad_(True); t.requires_grad_(True)
    lt = torch.tensor([math.log(tau)], device=DEV, requires_grad=True)
    loss = CLIPLoss(precision=prec)(video_features=v, text_features=t, log_temp=lt)
    loss.backward()
    rows = torch.randint(0, N, (256,), device=DEV, generator=g)
    ref, dv, dt = _clip_fp64(v.detach(), t.detach(), math.log(tau), rows)
    return {"loss_rel": abs(loss.item() - ref) / abs(ref),
            "dv_rel": ((v.grad[rows].double() - dv).norm() / dv.norm()).item(),
            "dt_rel": ((t.grad[rows].double() - dt).norm() / dt.norm()).item()}


def siglip_case(B, T, D, prec):
    from deepcoro_clip_b200.loss import SigLIPLoss
    g = torch.Generator(device=DEV).manual_seed(1)
    t = torch.randn(T, D, device=DEV, generator=g)
    v = 0.5 * t[:B] + torch.randn(B, D, device=DEV, generator=g)
    pm = torch.zeros(B, T, device=DEV)
    pm[torch.arange(B), torch.arange(B)] = 1.0
    for _ in range(3):
        pm[torch.arange(B, device=DEV), torch.randint(0, T, (B,), device=DEV, generator=g)] = 1.0
    sev = torch.tensor([1.0, 1.5, 2.5, 3.0], device=DEV)
    pw = pm * sev[torch.randint(0, 4, (B, T), device=DEV, generator=g)]
    v.requires_grad_(True); t.requires_grad_(True)
    log_tau, bias = math.log(0.087), -10.0
    lt = torch.tensor([log_tau], device=DEV, requires_grad=True)
    mod = SigLIPLoss(bias_init=bias, precision=prec).to(DEV)
    loss = mod(v, t, lt, pos_mask=pm, pos_weights=pw)
    loss.backward()
    v2 = v.detach().double().requires_grad_(True); t2 = t.detach().double().requires_grad_(True)
    lt2 = torch.tensor(log_tau, dtype=torch.float64, device=DEV, requires_grad=True)
    b2 = torch.tensor(bias, dtype=torch.float64, device=DEV, requires_grad=True)
    vh = torch.nn.functional.normalize(v2, dim=-1); th = torch.nn.functional.normalize(t2, dim=-1)
    L = (vh @ th.T / torch.exp(lt2).clamp(min=1e-4) + b2).clamp(-30, 30)
    y = pm.double().clamp(0, 1)
    w = torch.where(y > 0.5, pw.double(), torch.ones_like(y))
    ref = (w * torch.nn.functional.binary_cross_entropy_with_logits(L, y, reduction="none")).mean()
    ref.backward()
    return {"loss_rel": abs(loss.item() - ref.item()) / abs(ref.item()),
            "dv_rel": ((v.grad.double() - v2.grad).norm() / v2.grad.norm()).item(),
            "dt_rel": ((t.grad.double() - t2.grad).norm() / t2.grad.norm()).item(),
            "dlt_rel": abs(lt.grad.item() - lt2.grad.item()) / abs(lt2.grad.item()),
            "dbias_rel": abs(mod.bias.grad.item() - b2.grad.item()) / abs(b2.grad.item())}


for N in (2048, 4097, 8192, 16384):
    for prec in ("bf16", "bf16x3"):
        out[f"clip_N{N}_D512_{prec}"] = clip_case(N, 512, prec)
out["clip_N8192_D768_bf16"] = clip_case(8192, 768, "bf16")
out["clip_N8192_D512_bf16_uncorrelated"] = clip_case(8192, 512, "bf16", corr=0.0)
out["clip_N8192_D512_bf16_tau0.1"] = clip_case(8192, 512, "bf16", tau=0.1)
for prec in ("bf16", "bf16x3"):
    out[f"siglip_c2_8192_{prec}"] = siglip_case(8192, 8192, 512, prec)
    out[f"siglip_4096_{prec}"] = siglip_case(4096, 4096, 512, prec)
print(json.dumps(out, indent=1))
