"""Top-k lists at the C4 shape (203,808 x 32,473, D = 512, exact-grid embeddings): register-list sweep (default) against
the opt-in two-sweep threshold / collect path (B200CLIP_TOPK2=1). Prints the timings and checks the two agree bit for
bit; per-phase times of the two-sweep path come from CUDA events around each C-ABI call. Run under gpurun."""
import json, os, sys
sys.path.insert(0, ".")
import torch
from deepcoro_clip_b200 import ops
from deepcoro_clip_b200._lib import call, lib, stream_ptr
from deepcoro_clip_b200.retrieval_metrics_streaming import _sweep

dev = torch.device("cuda:0")
N, M, D = 203808, 32473, 512
g = torch.Generator(device="cuda").manual_seed(3)
v = torch.randint(-127, 128, (N, D), device=dev, generator=g).float() / 128
t = torch.randint(-127, 128, (M, D), device=dev, generator=g).float() / 128
vop, _, _ = ops.l2norm_operand(v, -1, False)
top, _, _ = ops.l2norm_operand(t, -1, False)
del v, t


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
for k in (5, 10, 16):
    os.environ["B200CLIP_TOPK2"] = "0"
    a = _sweep(vop, top, None, k, False)
    res[f"register_lists_k{k}_ms"] = timeit(lambda: _sweep(vop, top, None, k, False))
    os.environ["B200CLIP_TOPK2"] = "1"
    b = _sweep(vop, top, None, k, False)
    res[f"two_sweeps_k{k}_ms"] = timeit(lambda: _sweep(vop, top, None, k, False))
    res[f"equal_k{k}"] = bool((a[1] == b[1]).all().item() and (a[2] == b[2]).all().item())
    print(k, res[f"register_lists_k{k}_ms"], res[f"two_sweeps_k{k}_ms"], res[f"equal_k{k}"], flush=True)

# phases of the two-sweep path at k = 10
k, st = 10, stream_ptr(dev)
K = vop.shape[1]
segs = lib().b200clip_retrieval_segments(N, M)
pm = torch.full((N, 2 * segs, 32), float("-inf"), device=dev)
thr = torch.empty(N, device=dev)
cap = 80
cnt = torch.zeros(N + 1, dtype=torch.int32, device=dev)
bs = torch.empty((N, cap), device=dev); bi = torch.full((N, cap), 0x7FFFFFFF, dtype=torch.int32, device=dev)
os_ = torch.empty((N, k), device=dev); oi = torch.empty((N, k), dtype=torch.int64, device=dev)
res["colmax_ms"] = timeit(lambda: call("retrieval_colmax", vop, top, N, M, K, vop.stride(0), top.stride(0), segs, pm, st))
res["kth_largest_ms"] = timeit(lambda: call("kth_largest", pm, N, 2 * segs * 32, k, thr, st))


def collect():
    cnt.zero_(); bi.fill_(0x7FFFFFFF)
    call("retrieval_collect", vop, top, N, M, K, vop.stride(0), top.stride(0), thr, 0, segs, cnt, bs, bi, cap, cnt[N:], st)


res["collect_ms_incl_buffer_reset"] = timeit(collect)
res["merge_ms"] = timeit(lambda: call("topk_merge", bs, bi, N, cap, k, os_, oi, st))
res["candidates_mean"] = float(cnt[:N].float().mean().item()); res["candidates_max"] = int(cnt[:N].max().item())
res["overflow"] = int(cnt[N].item())
print(json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/topk_bench.json", "w"))
