"""Why is the eager e2e loop of bench.py slower than step + copy? Times the pieces separately (1 GPU)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import HostBatchPrefetcher, GraphedLossStep
from deepcoro_clip_b200.loss import CLIPLoss

dev = torch.device("cuda", 0)
N, D = 32768, 512
v = torch.randn(N, D, device=dev, requires_grad=True); t = torch.randn(N, D, device=dev, requires_grad=True)
lt = torch.tensor([math.log(0.0588)], device=dev, requires_grad=True)
mod = CLIPLoss(precision="bf16")
vh = v.detach().cpu().pin_memory(); th = t.detach().cpu().pin_memory()
lh = torch.empty(1).pin_memory()

def step(vv, tt):
    vv.grad = None; tt.grad = None; lt.grad = None
    loss = mod(video_features=vv, text_features=tt, log_temp=lt); loss.backward(); return loss

def ev_time(fn, n):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

for _ in range(5): step(v, t)
print("eager step, resident:", ev_time(lambda n: [step(v, t) for _ in range(n)], 20))
def sync_steps(n):
    for _ in range(n):
        lh.copy_(step(v, t).detach().reshape(1), non_blocking=True); torch.cuda.current_stream().synchronize()
print("eager step + loss readback + sync:", ev_time(sync_steps, 20))
dv = torch.empty_like(vh, device=dev); dt_ = torch.empty_like(th, device=dev)
def copies(n):
    for _ in range(n): dv.copy_(vh, non_blocking=True); dt_.copy_(th, non_blocking=True)
ms = ev_time(copies, 10); print("H2D 2 x 67 MB alone:", ms, "ms =", 2 * N * D * 4 / ms / 1e6, "GB/s")
side = torch.cuda.Stream()
def overlapped(n):
    for _ in range(n):
        with torch.cuda.stream(side):
            dv.copy_(vh, non_blocking=True); dt_.copy_(th, non_blocking=True)
        step(v, t)
    torch.cuda.current_stream().wait_stream(side)
print("eager step with an independent H2D on a side stream:", ev_time(overlapped, 20))
def e2e_loop(n):
    pf = HostBatchPrefetcher(((vh, th) for _ in range(n)), dev)
    for batch in pf:
        vv = batch[0].requires_grad_(True); tt = batch[1].requires_grad_(True)
        lh.copy_(step(vv, tt).detach().reshape(1), non_blocking=True)
        pf.release(batch)
        vv.requires_grad_(False); tt.requires_grad_(False)
        torch.cuda.current_stream().synchronize()
e2e_loop(3)
print("bench e2e loop:", ev_time(e2e_loop, 20))
def e2e_nosync(n):
    pf = HostBatchPrefetcher(((vh, th) for _ in range(n)), dev)
    for batch in pf:
        vv = batch[0].requires_grad_(True); tt = batch[1].requires_grad_(True)
        lh.copy_(step(vv, tt).detach().reshape(1), non_blocking=True)
        pf.release(batch)
        vv.requires_grad_(False); tt.requires_grad_(False)
print("bench e2e loop without the per-step sync:", ev_time(e2e_nosync, 20))
g = GraphedLossStep(mod, v, t, lt, warmup=2)
def e2e_graph(n):
    pf = HostBatchPrefetcher(((vh, th) for _ in range(n)), dev)
    for batch in pf:
        lh.copy_(g.step(batch[0], batch[1])[0].detach().reshape(1), non_blocking=True)
        pf.release(batch)
        torch.cuda.current_stream().synchronize()
e2e_graph(3)
print("graphed e2e loop:", ev_time(e2e_graph, 20))
# host time of one eager step (python dispatch)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): step(v, t)
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host time to enqueue one eager step: %.3f ms" % ((t1 - t0) / 20 * 1e3))
