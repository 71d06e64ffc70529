"""Where the plain-bf16 CLIP loss error comes from: compares, against float64, the three pieces the
loss is assembled from — row / column log-sum-exps and the target logits — for plain bf16 and bf16x3 operands, several
seeds. Prints JSON. Diagnostic tool, not a benchmark."""
import json
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deepcoro_clip_b200 import ops  # noqa: E402

DEV = "cuda:0"
out = {}
for N, seed in ((4097, 0), (4097, 1), (8192, 0), (8192, 1), (16384, 0)):
    D, tau = 512, 0.0588
    g = torch.Generator(device=DEV).manual_seed(1000 * seed + N)
    v = torch.randn(N, D, device=DEV, generator=g)
    t = 0.3 * v + torch.randn(N, D, device=DEV, generator=g)
    vh = torch.nn.functional.normalize(v.double(), dim=-1)
    th = torch.nn.functional.normalize(t.double(), dim=-1)
    r64 = torch.empty(N, dtype=torch.float64, device=DEV)
    c64 = torch.zeros(N, dtype=torch.float64, device=DEV)
    for a in range(0, N, 4096):
        L = vh[a:a + 4096] @ th.T / tau
        r64[a:a + 4096] = torch.logsumexp(L, 1)
        c64 += torch.exp(L).sum(0)
    c64 = torch.log(c64)
    d64 = (vh * th).sum(1) / tau
    loss64 = 0.5 * ((r64 - d64).mean() + (c64 - d64).mean())
    lt = torch.tensor([math.log(tau)], device=DEV)
    for x3 in (False, True):
        top, tinv, Kp = ops.l2norm_operand(t, 1 if x3 else -1)
        vop, vinv, _ = ops.l2norm_operand(v, 0 if x3 else -1)
        K = vop.shape[1]
        dyn = ops.dyn_prep(lt, None, 1e-4, 1.0)
        ws = torch.zeros(3 * N, dtype=torch.float32, device=DEV)
        ops.call("logits_lse_fwd", vop, top, N, N, K, vop.stride(0), top.stride(0), 0.0, 0.0, 0, dyn, 0, ws[N:2 * N],
                 ws[:N], ws[2 * N:], 0, ops.stream_ptr(torch.device(DEV)))
        shift = dyn[6].double()
        r = torch.log(ws[N:2 * N].double()) + shift
        c = torch.log(ws[:N].double()) + shift
        d = ws[2 * N:].double() / tau
        loss = 0.5 * ((r - d).mean() + (c - d).mean())
        out[f"N{N}_s{seed}_{'x3' if x3 else 'bf16'}"] = {
            "loss_rel": (abs(loss - loss64) / loss64).item(),
            "mean_row_lse_err": (r - r64).mean().item(), "mean_col_lse_err": (c - c64).mean().item(),
            "mean_diag_err": (d - d64).mean().item(), "rms_diag_err": (d - d64).pow(2).mean().sqrt().item(),
            "rms_row_lse_err": (r - r64).pow(2).mean().sqrt().item(), "loss": loss64.item()}
print(json.dumps(out, indent=1))
