"""Kernel-only timings of the tcgen05 attention-pool kernels at the C3 shape (x [32, 3136, 512] bf16, 8 heads) for
several token splits; CUDA events around back-to-back launches (x = 103 MB, three rotating copies so that nothing stays in
the 126 MB L2). Usage: python tools/gpu_bench_pool_tc.py [B N D]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pathlib import Path
from deepcoro_clip_b200._lib import call, lib, i64, stream_ptr, DTYPE_CODE

dev = torch.device("cuda:0")
peaks = json.loads(Path("MEASURED_PEAKS.json").read_text()) if Path("MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6516.7)
B, N, D = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 3136, 512)
H = 8
xs = [torch.randn(B, N, D, device=dev).to(torch.bfloat16) for _ in range(3)]
qt = torch.randn(H, D, device=dev) * 0.05
dxbar = torch.randn(B, H, D, device=dev)
dxs = [torch.empty_like(xs[0]) for _ in range(3)]
st = stream_ptr(dev)
bx = B * N * D * 2


def timeit(fn, reps=30, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


auto = lib().b200clip_attnpool_tc_splits(xs[0].data_ptr(), 1, i64(N * D), i64(D), B, N, D, H)
res = {"shape": [B, N, D, H], "auto_splits": auto}
cands = sorted({auto, 4, 5, 9, 13} if len(sys.argv) < 5 else {int(a) for a in sys.argv[4:]})
for S in cands:
    if S * 1 > (N + 63) // 64:
        continue
    pm = torch.empty((B, S, H), device=dev); pl = torch.empty((B, S, H), device=dev); pa = torch.empty((B, S, H, D), device=dev)
    xbar = torch.empty((B, H, D), device=dev); m = torch.empty((B, H), device=dev); l = torch.empty((B, H), device=dev)
    pdq = torch.empty((B, S, H, D), device=dev); dqt = torch.zeros((H, D), device=dev)
    f = lambda i: call("attnpool_tc_fwd", xs[i % 3], 1, None, i64(0), qt, None, B, N, D, H, S, pm, pl, pa, 0.0, 0, None, st)
    mg = lambda i: call("attnpool_merge", pm, pl, pa, B, S, H, D, xbar, m, l, 0, None, None, st)
    t_f = timeit(f); t_m = timeit(mg)
    bw = lambda i: call("attnpool_tc_bwd", xs[i % 3], 1, None, i64(0), qt, dxbar, xbar, None, None, m, l, B, N, D, H, S, dxs[i % 3], None, None,
                        0.0, 0, None, pdq, st)
    bw_nodq = lambda i: call("attnpool_tc_bwd", xs[i % 3], 1, None, i64(0), qt, dxbar, xbar, None, None, m, l, B, N, D, H, S, dxs[i % 3], None,
                             None, 0.0, 0, None, None, st)
    mq = lambda i: call("attnpool_merge", None, None, pdq, B, S, H, D, dqt, None, None, 1, None, None, st)
    t_b = timeit(bw); t_b0 = timeit(bw_nodq); t_mq = timeit(mq)
    res[f"S={S}"] = {"fwd_us": t_f * 1e3, "fwd_frac_hbm": bx / t_f / 1e6 / HBM, "merge_us": t_m * 1e3,
                     "bwd_us": t_b * 1e3, "bwd_frac_hbm": 2 * bx / t_b / 1e6 / HBM, "bwd_no_dq_us": t_b0 * 1e3,
                     "merge_dq_us": t_mq * 1e3}
    print(f"S={S}: fwd {t_f * 1e3:.1f} us ({bx / t_f / 1e6:.0f} GB/s, {bx / t_f / 1e6 / HBM:.2f} of HBM)  merge {t_m * 1e3:.1f} us  "
          f"bwd {t_b * 1e3:.1f} us ({2 * bx / t_b / 1e6:.0f} GB/s, {2 * bx / t_b / 1e6 / HBM:.2f})  bwd without dq {t_b0 * 1e3:.1f} us  "
          f"merge dq {t_mq * 1e3:.1f} us", flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/pool_tc_bench.json", "w"), indent=1)
