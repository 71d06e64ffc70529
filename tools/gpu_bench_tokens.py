"""C3 study-mode token kernels (BASELINE config 3): 8 studies x 4 views x 16 frames x 196 patch tokens, bf16.
Times each kernel with CUDA events (inputs 0.6-1.9 GB > 126 MB L2, no flush needed) and reports the fraction of
the measured HBM copy bandwidth against the ALGORITHMIC bytes of SURVEY 8d. Run under gpurun."""
import json, sys
sys.path.insert(0, ".")
import torch
from pathlib import Path
from deepcoro_clip_b200 import AttentionPool, Rope3D, EnhancedVideoAggregator

dev = torch.device("cuda:0")
peaks = json.loads(Path("MEASURED_PEAKS.json").read_text()) if Path("MEASURED_PEAKS.json").exists() else {}
HBM = peaks.get("hbm_gbs", 6650.0)
torch.manual_seed(2)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
# ---- RoPE: q, k [32, 8, 3136, 96] bf16, T=16, H=W=14 ----
B, Hh, N, Dh = 32, 8, 3136, 96
q = torch.randn(B, Hh, N, Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
k = torch.randn(B, Hh, N, Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
rope = Rope3D(Hh * Dh, Hh).to(dev).eval()
gq = torch.randn_like(q); gk = torch.randn_like(k)
with torch.no_grad():
    ms_f = timeit(lambda: rope(q, k, 16, 14, 14))
bytes_rope = 4 * B * Hh * N * Dh * 2
res["rope_fwd"] = {"ms": ms_f, "algorithmic_bytes": bytes_rope, "GBps": bytes_rope / ms_f / 1e6, "frac_hbm": bytes_rope / ms_f / 1e6 / HBM}
qo, ko = rope(q, k, 16, 14, 14)
ms_b = timeit(lambda: torch.autograd.grad((qo, ko), (q, k), (gq, gk), retain_graph=True))
res["rope_bwd"] = {"ms": ms_b, "algorithmic_bytes": bytes_rope, "GBps": bytes_rope / ms_b / 1e6, "frac_hbm": bytes_rope / ms_b / 1e6 / HBM}
del q, k, qo, ko, gq, gk
# ---- AttentionPool: x [32, 3136, 512] bf16, 8 heads ----
Bp, Np, D = 32, 3136, 512
x = torch.randn(Bp, Np, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
pool = AttentionPool(D, 8, dropout=0.0).to(dev)
with torch.no_grad():
    ms_f = timeit(lambda: pool(x))
bx = Bp * Np * D * 2
res["attnpool_fwd"] = {"ms": ms_f, "algorithmic_bytes": bx, "GBps": bx / ms_f / 1e6, "frac_hbm": bx / ms_f / 1e6 / HBM}
y = pool(x)
gy = torch.randn_like(y)
params = [p for p in pool.parameters() if p.requires_grad]
ms_b = timeit(lambda: torch.autograd.grad(y, [x] + params, gy, retain_graph=True, allow_unused=True))
res["attnpool_bwd"] = {"ms": ms_b, "algorithmic_bytes": 3 * bx, "GBps": 3 * bx / ms_b / 1e6, "frac_hbm": 3 * bx / ms_b / 1e6 / HBM,
                       "note": "includes the second pass over x for d(query) (4*bx executed) and the dense [B,D] tails"}
# ---- the same module forward + backward replayed from a CUDA graph (torch.cuda.make_graphed_callables) ----
import copy
gpool = torch.cuda.make_graphed_callables(copy.deepcopy(pool), (x.detach().clone().requires_grad_(True),))
xg = x.detach().clone().requires_grad_(True)
def fb_graphed():
    yy = gpool(xg)
    yy.backward(gy)
def fb_eager():
    yy = pool(x)
    yy.backward(gy)
ms_fb_g = timeit(fb_graphed)
ms_fb_e = timeit(fb_eager)
res["attnpool_fwd_bwd_module"] = {"eager_ms": ms_fb_e, "cuda_graph_ms": ms_fb_g, "algorithmic_bytes": 4 * bx,
                                  "cuda_graph_GBps": 4 * bx / ms_fb_g / 1e6, "cuda_graph_frac_hbm": 4 * bx / ms_fb_g / 1e6 / HBM}
mask = torch.rand(Bp, Np, device=dev) < 0.1
with torch.no_grad():
    ms_m = timeit(lambda: pool(x, mask))
res["attnpool_fwd_masked"] = {"ms": ms_m, "GBps": bx / ms_m / 1e6, "frac_hbm": bx / ms_m / 1e6 / HBM}
# ---- the streaming kernels alone (no host-side dense tails) ----
from deepcoro_clip_b200.attention_pool import _StreamPool
from deepcoro_clip_b200._lib import call, lib, i64, stream_ptr, DTYPE_CODE
qt = torch.randn(8, D, device=dev) * 0.05
xd = x.detach()
S = lib().b200clip_attnpool_splits(Bp, Np)
pm = torch.empty((Bp, S, 8), device=dev); pl = torch.empty((Bp, S, 8), device=dev); pa = torch.empty((Bp, S, 8, D), device=dev)
xbar = torch.empty((Bp, 8, D), device=dev); m = torch.empty((Bp, 8), device=dev); l = torch.empty((Bp, 8), device=dev)
st = stream_ptr(dev)
def k_fwd():
    call("attnpool_fwd", xd, DTYPE_CODE[xd.dtype], i64(xd.stride(0)), i64(xd.stride(1)), None, i64(0), qt, None, i64(0), i64(0), Bp, Np, D, 8, S, pm, pl, pa, 0.0, 0, None, st)
    call("attnpool_merge", pm, pl, pa, Bp, S, 8, D, xbar, m, l, 0, None, None, st)
ms_k = timeit(k_fwd, reps=20)
res["attnpool_fwd_kernels_only"] = {"ms": ms_k, "splits": S, "GBps": bx / ms_k / 1e6, "frac_hbm": bx / ms_k / 1e6 / HBM}
dxbar = torch.randn(Bp, 8, D, device=dev); dx = torch.empty_like(xd); ds = torch.empty((Bp, 8, Np), device=dev)
def k_bwd():
    call("attnpool_bwd_dx", xd, DTYPE_CODE[xd.dtype], i64(xd.stride(0)), i64(xd.stride(1)), None, i64(0), qt, dxbar, xbar, m, l, Bp, Np, D, 8, dx, ds, None, None, 0.0, 0, None, st)
ms_k = timeit(k_bwd, reps=20)
res["attnpool_bwd_dx_kernel_only"] = {"ms": ms_k, "GBps": 2 * bx / ms_k / 1e6, "frac_hbm": 2 * bx / ms_k / 1e6 / HBM, "algorithmic_bytes": 2 * bx}
# ---- multi-view query pool: [8, 4, 512] fp32 ----
agg = EnhancedVideoAggregator(embedding_dim=512).to(dev) if "embedding_dim" in EnhancedVideoAggregator.__init__.__code__.co_varnames else None
print(json.dumps(res, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/tokens_bench.json", "w"), indent=1)
