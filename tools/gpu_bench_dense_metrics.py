"""HBM roofline of the dense-metrics rank pass (dense_metrics.cu): one read of the [N, M] similarity matrix.
    python tools/gpu_bench_dense_metrics.py   -> gpurun_out/dense_metrics_bench.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import retrieval_metrics as rm

dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
out = {}
for (N, M, G, dt) in ((16384, 32473, 1, torch.float32), (16384, 32473, 4, torch.float32), (32768, 32473, 8, torch.bfloat16)):
    g = torch.Generator(device=dev).manual_seed(1)
    sim = torch.randn(N, M, device=dev, generator=g).to(dt)
    gt = torch.randint(0, M, (N, G), device=dev, generator=g)
    gta = gt[:, 0] if G == 1 else gt
    for _ in range(2):
        rm._row_terms(sim, gta, recall_k=[1, 5, 10], ndcg_k=[5])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rm._row_terms(sim, gta, recall_k=[1, 5, 10], ndcg_k=[5])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    by = N * M * sim.element_size()
    out[f"N{N}_M{M}_G{G}_{str(dt).split('.')[-1]}"] = {"ms": ms, "algorithmic_bytes": by, "GBps": by / ms / 1e6,
                                                      "frac_hbm": by / ms / 1e6 / peaks["hbm_gbs"]}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dense_metrics_bench.json", "w"), indent=1)
print(json.dumps(out, indent=1))
