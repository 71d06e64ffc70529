"""Eager back-to-back launches against one CUDA graph of the same launches: where does the eager aggregator step lose
~280 us to its graph replay? Forward block kernel (cluster launch, 190 KB smem) and a weight-gradient kernel."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200.video_aggregator import TransformerBlock, EnhancedVideoAggregator
from deepcoro_clip_b200._lib import call, i64, stream_ptr
dev = torch.device("cuda", 0)
blk = TransformerBlock(512, 4, 0.1).to(dev).eval()
x = torch.randn(8, 4, 512, device=dev)
a = torch.randn(32, 2048, device=dev); b = torch.randn(32, 512, device=dev); dw = torch.empty(2048, 512, device=dev); db = torch.empty(2048, device=dev)
def fwd():
    with torch.no_grad():
        return blk(x)
def wg():
    call("xfblock_wgrad", a, i64(2048), b, i64(512), dw, db, 2048, 512, 32, None, None, None, None, 0, stream_ptr(dev))
def ev(fn, n):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (t1 - t0) / n * 1e6
for name, fn in (("xfblock fwd", fwd), ("xfblock_wgrad", wg)):
    for _ in range(10): fn()
    dev_us, host_us = ev(fn, 50)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(50): fn()
    g.replay(); torch.cuda.synchronize()
    gdev, _ = ev(g.replay, 5)
    print(f"{name}: eager {dev_us:.1f} us per launch on the device (host {host_us:.1f} us); graph of 50: {gdev / 50:.1f} us per launch", flush=True)
