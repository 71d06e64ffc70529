"""cProfile of the eager multi-GPU CLIP step on rank 0 (torchrun, 2+ ranks): where the host time of the plugin path goes."""
import cProfile, io, math, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); W = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from deepcoro_clip_b200.loss import CLIPLoss
B = 4096
v = torch.randn(B, 512, device=dev, requires_grad=True); t = torch.randn(B, 512, device=dev, requires_grad=True)
lt = torch.tensor([math.log(0.0588)], device=dev, requires_grad=True)
mod = CLIPLoss(precision="bf16")
def fc():
    v.grad = None; t.grad = None; lt.grad = None
    mod(video_features=v, text_features=t, log_temp=lt).backward()
for _ in range(20): fc()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(200): fc()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
from deepcoro_clip_b200 import symm as _symm
if rank == 0: print("symmetric plans:", [(k[1], k[2], "multicast" if p.mc_op is not None else "unicast", "own barrier" if p.own_barrier else "torch barrier") for k, p in _symm._PLANS.items()])
if rank == 0: print(f"W={W} rows/rank {B}: host enqueue {(t1 - t0) / 200 * 1e6:.0f} us, wall {(t2 - t0) / 200 * 1e6:.0f} us per step (SYMM={os.environ.get('B200CLIP_SYMM', '1')})")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): fc()
pr.disable(); torch.cuda.synchronize()
if rank == 0:
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(26)
    print("\n".join(s.getvalue().splitlines()[:48]))
dist.barrier(); dist.destroy_process_group()
