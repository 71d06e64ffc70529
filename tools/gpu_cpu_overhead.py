"""Host-side overhead of one CLIP loss step (forward + backward): wall time at a size where the GPU work is
negligible, plus a cProfile of the Python side. Run under gpurun."""
import cProfile, math, pstats, sys, time
sys.path.insert(0, ".")
import torch
from deepcoro_clip_b200.loss import CLIPLoss
dev = torch.device("cuda:0")
N, D = 4096, 512
v = torch.randn(N, D, device=dev, requires_grad=True); t = torch.randn(N, D, device=dev, requires_grad=True)
lt = torch.tensor([math.log(0.0588)], device=dev, requires_grad=True)
mod = CLIPLoss(precision="bf16")
def step():
    v.grad = None; t.grad = None; lt.grad = None
    loss = mod(video_features=v, text_features=t, log_temp=lt)
    loss.backward()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter()          # CPU-side issue time (no sync inside)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"N={N}: CPU issue {1e6*(t1-t0)/200:.0f} us/step, incl. GPU drain {1e6*(t2-t0)/200:.0f} us/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
