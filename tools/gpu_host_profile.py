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
