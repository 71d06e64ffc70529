"""cProfile of the eager aggregator step (40 iterations: the launch queue never fills, so host time is host work)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import EnhancedVideoAggregator
dev = torch.device("cuda", 0)
agg = EnhancedVideoAggregator(512).to(dev)
xa = torch.randn(8, 4, 512, device=dev, requires_grad=True); ga = torch.randn(8, 512, device=dev)
def fa():
    agg.zero_grad(set_to_none=True); xa.grad = None
    agg(xa).backward(ga)
def fwd_only():
    with torch.no_grad():
        agg(xa)
for name, fn in (("fwd+bwd", fa), ("forward only", fwd_only)):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(40): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: host {(t1 - t0) / 40 * 1e6:.0f} us, wall {(t2 - t0) / 40 * 1e6:.0f} us per step")
    pr = cProfile.Profile(); pr.enable()
    for _ in range(40): fn()
    pr.disable(); torch.cuda.synchronize()
    for key in ("tottime", "cumulative"):
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(16)
        print("=====", name, key, "(40 iterations)"); print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:32]))
