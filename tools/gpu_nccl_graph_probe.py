"""Probe: can NCCL collectives (async all_gather + wait, all_reduce) be captured in a torch CUDA graph on this box?
    timeout 60 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29570 tools/gpu_nccl_graph_probe.py [full]"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); W = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def log(*a):
    print(f"[rank {rank} t={time.time() % 1000:.1f}]", *a, flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "full":
    from deepcoro_clip_b200 import GraphedLossStep
    from deepcoro_clip_b200.loss import CLIPLoss
    B, D = 2048, 512
    v = torch.randn(B, D, device=dev); t = torch.randn(B, D, device=dev)
    lt = torch.tensor([math.log(0.07)], device=dev)
    mod = CLIPLoss(precision="bf16")
    log("eager step")
    vv = v.clone().requires_grad_(True); tt = t.clone().requires_grad_(True); ll = lt.clone().requires_grad_(True)
    le = mod(video_features=vv, text_features=tt, log_temp=ll); le.backward(); torch.cuda.synchronize()
    log("capturing")
    gs = GraphedLossStep(mod, v, t, lt, warmup=2)
    log("captured; replaying")
    for i in range(5):
        loss, dv, dt, dlt = gs.step()
        torch.cuda.synchronize()
        log("replay", i, loss.item(), le.item(), float((dv - vv.grad).abs().max()))
else:
    x = torch.full((1024,), float(rank + 1), device=dev)
    out = torch.empty(W * 1024, device=dev)
    s = torch.zeros(4, device=dev)
    # warm-up (communicator creation) on a side stream
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            w = dist.all_gather_into_tensor(out, x, async_op=True); w.wait()
            s.copy_(out[:4]); dist.all_reduce(s)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    log("warm-up done; capturing")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        w = dist.all_gather_into_tensor(out, x, async_op=True)
        y = x * 2
        w.wait()
        s.copy_(out[:4] + y[:4]); dist.all_reduce(s)
    log("captured; replaying")
    for i in range(5):
        g.replay(); torch.cuda.synchronize()
        log("replay", i, s.tolist())
dist.barrier()
torch.cuda.synchronize()
log("done")
# a live CUDA graph that holds NCCL kernels keeps destroy_process_group() waiting: drop the graphs first
gs = g = None
import gc; gc.collect()
torch.cuda.synchronize()
dist.destroy_process_group()
log("destroyed")
