"""Logits-level multi-positive softmax losses at the C2 sizing: logits and weights [8192, 8192] fp32
(268 MB each, > 126 MB L2). CUDA-event timing of forward and forward + backward; algorithmic bytes: forward = two reads of
(logits, weights) (row pass + column pass), backward = one read of (logits, weights) + one write of dlogits. Beside it the
same function composed from stock torch ops (two log_softmax matrices, as the reference classes do): the GPU baseline.
Run under gpurun."""
import json
import sys

sys.path.insert(0, ".")
from pathlib import Path

import torch
import torch.nn.functional as F

from deepcoro_clip_b200 import MultiPositiveInfoNCELoss, WeightedSigLIPLoss

dev = torch.device("cuda:0")
peaks = json.loads(Path("MEASURED_PEAKS.json").read_text()) if Path("MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def stock_weighted(logits, w, eps=1e-6):
    """utils/loss/weighted_siglip.py:38-51 restated with stock torch ops (timing baseline only)."""
    pos = w.clamp(min=0)
    lv = -(pos * F.log_softmax(logits, dim=1)).sum(1) / pos.sum(1).clamp_min(eps)
    lt = -(pos * F.log_softmax(logits, dim=0)).sum(0) / pos.sum(0).clamp_min(eps)
    return 0.5 * (lv.mean() + lt.mean())


N = M = 8192
torch.manual_seed(1)
logits = (torch.randn(N, M, device=dev) * 4).requires_grad_(True)
mask = torch.zeros(N, M, device=dev)
mask[torch.arange(N, device=dev), torch.arange(N, device=dev)] = 1
idx = torch.randint(0, M, (N, 3), device=dev)
mask.scatter_(1, idx, 1.0)
w = mask * torch.tensor([1.0, 1.5, 2.5, 3.0], device=dev)[torch.randint(0, 4, (N, M), device=dev)]
nbytes = N * M * 4
res = {"shape": [N, M], "hbm_peak_GBps": HBM}
for name, fn in {"weighted_siglip": lambda: WeightedSigLIPLoss()(logits, w),
                 "multi_positive_infonce": lambda: MultiPositiveInfoNCELoss()(logits, mask, w)}.items():
    with torch.no_grad():
        ms_f = timeit(fn)

    def fb():
        logits.grad = None
        fn().backward()

    ms_fb = timeit(fb)
    nin = 2 if name == "weighted_siglip" else 3           # matrices read per pass (logits, weights[, mask])
    fwd_b, bwd_b = 2 * nin * nbytes, (nin + 1) * nbytes
    res[name] = {"fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "fwd_algorithmic_bytes": fwd_b, "bwd_algorithmic_bytes": bwd_b,
                 "fwd_GBps": fwd_b / ms_f / 1e6, "fwd_frac_hbm": fwd_b / ms_f / 1e6 / HBM,
                 "fwd_bwd_GBps": (fwd_b + bwd_b) / ms_fb / 1e6, "fwd_bwd_frac_hbm": (fwd_b + bwd_b) / ms_fb / 1e6 / HBM}


def fb_stock():
    logits.grad = None
    stock_weighted(logits, w).backward()


res["weighted_siglip"]["stock_torch_fwd_bwd_ms"] = timeit(fb_stock)
res["weighted_siglip"]["speedup_vs_stock"] = res["weighted_siglip"]["stock_torch_fwd_bwd_ms"] / res["weighted_siglip"]["fwd_bwd_ms"]
l_ours = float(WeightedSigLIPLoss()(logits, w))
l_stock = float(stock_weighted(logits, w))
res["weighted_siglip"]["loss_rel_diff_vs_stock"] = abs(l_ours - l_stock) / abs(l_stock)
print(json.dumps(res, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/multipos_bench.json", "w"), indent=1)
