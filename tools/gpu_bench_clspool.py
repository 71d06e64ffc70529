"""AttentionPoolWithCLS at the token shapes of BASELINE config 3: x [32, 3136, 512] bf16 (synthetic C3) and
[32, 393, 512] (what MViT-v2-S really emits per view). Times forward + backward of the module (eager and replayed from a
CUDA graph) with CUDA events; algorithmic bytes = one read of x forward, read x + write dx backward (3 * B * N * D * 2).
Beside it, the same function composed from stock torch modules the way the reference does it (CLS token concatenated,
one nn.TransformerEncoderLayer over all N + 1 rows under bf16 autocast, row 0 kept): the GPU baseline this replaces.
Run under gpurun."""
import copy
import json
import sys

sys.path.insert(0, ".")
from pathlib import Path

import torch
import torch.nn as nn

from deepcoro_clip_b200 import AttentionPoolWithCLS

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


class StockClsPool(nn.Module):
    """What models/attention_pool.py:104-197 computes, composed from stock torch modules (timing baseline only)."""

    def __init__(self, D, H):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.transformer = nn.TransformerEncoder(nn.TransformerEncoderLayer(D, H, dropout=0.0, batch_first=True), 1)
        self.norm = nn.LayerNorm(D)

    def forward(self, x):
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1).to(x.dtype), x], dim=1)
        return self.norm(self.transformer(x)[:, 0])


res = {}
for tag, (B, N) in {"c3_synthetic_3136": (32, 3136), "mvit_393": (32, 393)}.items():
    torch.manual_seed(2)
    D, H = 512, 8
    x = torch.randn(B, N, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    pool = AttentionPoolWithCLS(D, H, dropout=0.0).to(dev)
    y = pool(x)
    gy = torch.randn_like(y)
    bx = B * N * D * 2

    def fb_eager():
        for p in pool.parameters():
            p.grad = None
        x.grad = None
        pool(x).backward(gy)

    with torch.no_grad():
        ms_f = timeit(lambda: pool(x))
    ms_e = timeit(fb_eager)
    gpool = torch.cuda.make_graphed_callables(copy.deepcopy(pool), (x.detach().clone().requires_grad_(True),))
    xg = x.detach().clone().requires_grad_(True)

    def fb_graph():
        gpool(xg).backward(gy)

    ms_g = timeit(fb_graph)
    stock = StockClsPool(D, H).to(dev)
    xs = x.detach().clone().requires_grad_(True)

    def fb_stock():
        for p in stock.parameters():
            p.grad = None
        xs.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o = stock(xs)
        o.backward(gy.to(o.dtype))

    ms_s = timeit(fb_stock)
    res[tag] = {"shape": [B, N, D], "fwd_eager_ms": ms_f, "fwd_bwd_eager_ms": ms_e, "fwd_bwd_cuda_graph_ms": ms_g,
                "algorithmic_bytes_fwd_bwd": 3 * bx, "cuda_graph_GBps": 3 * bx / ms_g / 1e6,
                "cuda_graph_frac_hbm": 3 * bx / ms_g / 1e6 / HBM,
                "stock_torch_bf16_autocast_fwd_bwd_ms": ms_s, "speedup_vs_stock_eager": ms_s / ms_e,
                "speedup_vs_stock_graph": ms_s / ms_g}
    del gpool, stock, xs, xg
print(json.dumps(res, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/clspool_bench.json", "w"), indent=1)
