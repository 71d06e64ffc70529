"""cProfile of the eager view-level MIL pooling step ([32, 4, 512], hidden 128): host work vs device time."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import GatedAttentionPooling
dev = torch.device("cuda", 0)
mod = GatedAttentionPooling(512, 128).to(dev)
x = torch.randn(32, 4, 512, device=dev, requires_grad=True); g = torch.randn(32, 512, device=dev)
mask = torch.ones(32, 4, dtype=torch.bool, device=dev)
params = list(mod.parameters())
def fb():
    x.grad = None
    for p in params: p.grad = None
    mod(x, mask).backward(g)
for _ in range(10): fb()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(40): fb()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host {(t1 - t0) / 40 * 1e6:.0f} us, wall {(t2 - t0) / 40 * 1e6:.0f} us per step")
g_ = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    fb(); torch.cuda.synchronize()
    with torch.cuda.graph(g_, stream=s):
        fb()
g_.replay(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): g_.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay {e0.elapsed_time(e1) / 20 * 1e3:.0f} us per step")
pr = cProfile.Profile(); pr.enable()
for _ in range(40): fb()
pr.disable(); torch.cuda.synchronize()
for key in ("tottime",):
    s_ = io.StringIO(); pstats.Stats(pr, stream=s_).sort_stats(key).print_stats(14)
    print("\n".join(l[:150] for l in s_.getvalue().splitlines()[4:26]))
