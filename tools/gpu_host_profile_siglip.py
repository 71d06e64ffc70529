"""cProfile of the eager multi-GPU SigLIP config-2 step on rank 0 (torchrun): where the host time goes."""
import cProfile, io, math, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); W = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from deepcoro_clip_b200.loss import SigLIPLoss
B, T, D = 1024, 8192, 512
g = torch.Generator(device=dev).manual_seed(1)
t = torch.randn(T, D, device=dev, generator=g).requires_grad_(True)
v = torch.randn(B, D, device=dev).requires_grad_(True)
pm = torch.zeros(B, T, device=dev); pm[torch.arange(B), torch.arange(B)] = 1.0
pw = pm * 2.0
lt = torch.tensor([math.log(0.087)], device=dev, requires_grad=True)
mod = SigLIPLoss(precision="bf16", text_replicated=True).to(dev)
def f():
    v.grad = None; t.grad = None; lt.grad = None; mod.bias.grad = None
    mod(v, t, lt, pos_mask=pm, pos_weights=pw).backward()
for _ in range(20): f()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(200): f()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
if rank == 0: print(f"W={W} rows/rank {B}: host enqueue {(t1 - t0) / 200 * 1e6:.0f} us, wall {(t2 - t0) / 200 * 1e6:.0f} us per step")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): f()
pr.disable(); torch.cuda.synchronize()
if rank == 0:
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(24)
    print("\n".join(s.getvalue().splitlines()[:44]))
dist.barrier(); dist.destroy_process_group()
