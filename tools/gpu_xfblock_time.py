"""GPU-bound time of one fused transformer block (forward + backward) at a few shapes: CUDA-graph replays."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200.video_aggregator import TransformerBlock
dev = torch.device("cuda", 0)
def timeit(fn, reps=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
out = []
for B, N in ((8, 4), (32, 4), (8, 8), (8, 16)):
    blk = TransformerBlock(512, 4, 0.1).to(dev).eval()
    x = torch.randn(B, N, 512, device=dev, requires_grad=True); g = torch.randn(B, N, 512, device=dev)
    gb = torch.cuda.make_graphed_callables(blk, (x.detach().clone().requires_grad_(True),))
    def f():
        x.grad = None
        gb(x).backward(g)
    out.append(f"B={B} N={N}: {timeit(f):.0f} us")
print(sys.argv[1] if len(sys.argv) > 1 else "", " | ".join(out), flush=True)
