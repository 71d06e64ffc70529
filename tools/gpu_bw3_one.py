"""One logits_bwd launch per shape for ncu captures of the backward kernels (run under gpurun):
    python tools/gpu_bw3_one.py D [reps]      (B200CLIP_BWD3=1/0 selects logits_bwd3.cu / the 128-row kernels)"""
import math, sys
sys.path.insert(0, ".")
import torch
from deepcoro_clip_b200 import _lib as L

D = int(sys.argv[1]) if len(sys.argv) > 1 else 768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 32768
dev = torch.device("cuda:0")
st = L.stream_ptr()
LOG2E = math.log2(math.e)
torch.manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=-1).bfloat16()
y = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=-1).bfloat16()
rs = torch.rand(N, device=dev); cs = torch.rand(N, device=dev)
dX = torch.zeros(N, D, device=dev); scal = torch.zeros(4, device=dev, dtype=torch.float64)
tau = 0.0588
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for it in range(reps):
    if it == reps - 1:
        e0.record()
    L.call("logits_bwd", 0, x, y, N, N, D, D, D, 0, D, D, LOG2E / tau, LOG2E / tau, 1 / tau, 0.0, 0.0, rs, cs, 1 / tau, 1.0, 0,
           None, 0.0, 0, None, dX, D, scal, 0, st)
e1.record(); torch.cuda.synchronize()
print(f"D={D}: {e0.elapsed_time(e1):.3f} ms")
