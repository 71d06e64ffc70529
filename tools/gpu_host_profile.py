"""Where does the host time of the eager token modules go? cProfile over the aggregator and the AttentionPool (fwd + bwd)."""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import AttentionPool, EnhancedVideoAggregator
dev = torch.device("cuda", 0)
agg = EnhancedVideoAggregator(512).to(dev)
xa = torch.randn(8, 4, 512, device=dev, requires_grad=True); ga = torch.randn(8, 512, device=dev)
pool = AttentionPool(512, 8).to(dev)
x = torch.randn(32, 3136, 512, device=dev, dtype=torch.bfloat16, requires_grad=True); gy = torch.randn(32, 512, device=dev, dtype=torch.bfloat16)
def fa():
    agg.zero_grad(set_to_none=True); xa.grad = None
    agg(xa).backward(ga)
def fp():
    pool.zero_grad(set_to_none=True); x.grad = None
    pool(x).backward(gy)
for name, fn in (("aggregator", fa), ("attention pool", fp)):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): fn()
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
    print("=====", name, "(200 iterations)"); print("\n".join(s.getvalue().splitlines()[:40]))
# CLIP loss step at the per-rank size of the 8-GPU point (4096 rows): host time of the eager plugin path
import math
from deepcoro_clip_b200.loss import CLIPLoss
v = torch.randn(4096, 512, device=dev, requires_grad=True); t = torch.randn(4096, 512, device=dev, requires_grad=True)
lt = torch.tensor([math.log(0.0588)], device=dev, requires_grad=True)
mod = CLIPLoss(precision="bf16")
def fc():
    v.grad = None; t.grad = None; lt.grad = None
    mod(video_features=v, text_features=t, log_temp=lt).backward()
for _ in range(20): fc()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(200): fc()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"CLIP step 4096 x 4096: host enqueue {(t1 - t0) / 200 * 1e6:.0f} us, wall {(t2 - t0) / 200 * 1e6:.0f} us")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): fc()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print("===== CLIP step (200 iterations)"); print("\n".join(s.getvalue().splitlines()[:44]))
