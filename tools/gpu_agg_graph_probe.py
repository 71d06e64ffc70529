"""GPU-bound time of the aggregator / AttentionPool forward + backward: torch.cuda.make_graphed_callables replays."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import AttentionPool, EnhancedVideoAggregator
dev = torch.device("cuda", 0)
def timeit(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
agg = EnhancedVideoAggregator(512).to(dev).eval()       # eval: no host-side dropout seed inside the captured region
xa = torch.randn(8, 4, 512, device=dev, requires_grad=True); ga = torch.randn(8, 512, device=dev)
def fa_eager():
    xa.grad = None
    agg(xa).backward(ga)
print(f"aggregator eager fwd+bwd {timeit(fa_eager):.0f} us")
gagg = torch.cuda.make_graphed_callables(copy.deepcopy(agg), (xa.detach().clone().requires_grad_(True),))
xg = xa.detach().clone().requires_grad_(True)
def fa_graph():
    xg.grad = None
    gagg(xg).backward(ga)
print(f"aggregator graphed fwd+bwd {timeit(fa_graph):.0f} us")
pool = AttentionPool(512, 8).to(dev)
x = torch.randn(32, 3136, 512, device=dev, dtype=torch.bfloat16, requires_grad=True); gy = torch.randn(32, 512, device=dev, dtype=torch.bfloat16)
def fp_eager():
    x.grad = None
    pool(x).backward(gy)
print(f"attention pool eager fwd+bwd {timeit(fp_eager):.0f} us")
gpool = torch.cuda.make_graphed_callables(copy.deepcopy(pool), (x.detach().clone().requires_grad_(True),))
xp = x.detach().clone().requires_grad_(True)
def fp_graph():
    xp.grad = None
    gpool(xp).backward(gy)
print(f"attention pool graphed fwd+bwd {timeit(fp_graph):.0f} us")
