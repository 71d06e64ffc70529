"""Phase timing of one CLIP loss step under torchrun (one rank per GPU): CUDA events around every kernel / collective of
deepcoro_clip_b200.loss._ClipLossFn, issued exactly as the module issues them. Rank 0 prints the mean over the steps.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 tools/gpu_dist_phases.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from deepcoro_clip_b200 import dist_plan, ops

rank = int(os.environ.get("RANK", 0)); W = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if W > 1:
    dist.init_process_group("nccl", device_id=dev)
N, D = 32768, 512
B = N // W
g = torch.Generator().manual_seed(rank)
video = torch.randn(B, D, generator=g).to(dev); text = torch.randn(B, D, generator=g).to(dev)
log_temp = torch.tensor([math.log(0.0588)], device=dev)
names, evs = [], []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); names.append(name); evs.append(e)

def step():
    names.clear(); evs.clear()
    st = ops.stream_ptr(dev)
    mark("start")
    top, tinv, Kp = ops.l2norm_operand(text, -1); mark("l2norm_t")
    tall, tw = dist_plan.gather_rows_async(top, W, None)
    vop, vinv, _ = ops.l2norm_operand(video, -1); mark("l2norm_v")
    vall, vw = dist_plan.gather_rows_async(vop, W, None)
    K = vop.shape[1]
    dyn = ops.dyn_prep(log_temp, None, 1e-4, 1.0)
    ws = torch.zeros(5 * N + 2, dtype=torch.float32, device=dev); mark("prep+zero")
    lo = rank * B
    if tw is not None: tw.wait()
    mark("wait_gather_t")
    ops.call("logits_lse_fwd", vop, tall, B, N, K, vop.stride(0), tall.stride(0), 0.0, 0.0, 0, dyn, 0, ws[N + lo:N + lo + B], ws[:N], ws[2 * N + lo:2 * N + lo + B], lo, st); mark("fwd")
    if W > 1: dist.all_reduce(ws[:3 * N])
    mark("allreduce_sums")
    loss = torch.empty(1, device=dev)
    ops.call("clip_finalize", ws[:3 * N], N, 3, dyn, 0.0, 0, None, ws[3 * N:4 * N], ws[4 * N:5 * N], loss, None, st); mark("finalize")
    if vw is not None: vw.wait()
    mark("wait_gather_v")
    nbd = B * D
    w2 = torch.zeros(2 * nbd + 4 * B + 8, dtype=torch.float32, device=dev); mark("zero_bwd")
    scal = w2[2 * nbd + 4 * B:].view(torch.float64)
    ops.logits_bwd(0, vop, tall, B, N, K, Kp, D, dyn, ws[3 * N + lo:3 * N + lo + B], ws[4 * N:5 * N], w2[:nbd].view(B, D), scal, ydiag=1.0 / N, diag_off=lo, diag_corr=w2[2 * nbd:2 * nbd + 2 * B], gnorm=2.0 * N); mark("bwd_v")
    lw = dist.all_reduce(scal[0:1], async_op=True) if W > 1 else None
    dV = ops.l2norm_backward(w2[:nbd].view(B, D), video, vinv, other_x=text, other_inv=tinv, other_hi=top, diag_corr=w2[2 * nbd:2 * nbd + 2 * B], dev_omul=dyn[2:3]); mark("l2norm_bwd_v")
    ops.logits_bwd(0, top, vall, B, N, K, Kp, D, dyn, ws[4 * N + lo:4 * N + lo + B], ws[3 * N:4 * N], w2[nbd:2 * nbd].view(B, D), None, ydiag=1.0 / N, diag_off=lo, diag_corr=w2[2 * nbd + 2 * B:2 * nbd + 4 * B], gnorm=2.0 * N); mark("bwd_t")
    dT = ops.l2norm_backward(w2[nbd:2 * nbd].view(B, D), text, tinv, other_x=video, other_inv=vinv, other_hi=vop, diag_corr=w2[2 * nbd + 2 * B:2 * nbd + 4 * B], dev_omul=dyn[2:3]); mark("l2norm_bwd_t")
    if lw is not None: lw.wait()
    mark("wait_allreduce_lt")

for _ in range(5): step()
torch.cuda.synchronize()
if W > 1: dist.barrier()
acc = {}
R = 20
import time
t0 = time.perf_counter()
for _ in range(R):
    step()
    torch.cuda.synchronize()
    for i in range(1, len(evs)):
        acc[names[i]] = acc.get(names[i], 0.0) + evs[i - 1].elapsed_time(evs[i])
    acc["total"] = acc.get("total", 0.0) + evs[0].elapsed_time(evs[-1])
t1 = time.perf_counter()
if rank == 0:
    print(f"W={W} B={B}: " + "  ".join(f"{k}={v / R * 1e3:.0f}us" for k, v in acc.items()), flush=True)
if W > 1: dist.destroy_process_group()
