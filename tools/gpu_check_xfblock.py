"""Fused transformer block of the aggregator (csrc/xfblock.cu) against the PyTorch block in float64 (the reference's own
module structure, models/video_aggregator.py:7-54): output, dx and every parameter gradient; dropout by a directional
finite difference with a fixed seed; then the whole EnhancedVideoAggregator and timings at the C3 shape. 1 GPU."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import EnhancedVideoAggregator, _lib
from deepcoro_clip_b200.video_aggregator import TransformerBlock

dev = torch.device("cuda", 0)
ok = True


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def case(B, N, D, H, masked, seed=0):
    global ok
    torch.manual_seed(seed)
    blk = TransformerBlock(D, H, 0.1).to(dev).eval()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x = torch.randn(B, N, D, device=dev)
    mask = None
    if masked:
        mask = torch.rand(B, N, device=dev) < 0.3
        mask[:, 0] = False
    g = torch.randn(B, N, D, device=dev)
    os.environ["B200CLIP_XFBLOCK"] = "1"
    xr = x.clone().requires_grad_(True)
    l0 = _lib.LAUNCHES
    y = blk(xr, mask)
    y.backward(g)
    torch.cuda.synchronize()
    nl = _lib.LAUNCHES - l0
    os.environ["B200CLIP_XFBLOCK"] = "0"
    ref = copy.deepcopy(blk).double()
    ref.zero_grad(set_to_none=True)
    xd = x.double().requires_grad_(True)
    yr = ref(xd, mask)
    yr.backward(g.double())
    errs = {"y": rel(y, yr), "dx": rel(xr.grad, xd.grad)}
    for (k, p), (_, r) in zip(blk.named_parameters(), ref.named_parameters()):
        errs[k] = rel(p.grad, r.grad)
    worst = max(errs.values())
    good = worst <= 2e-5 and nl == 6
    ok &= good
    print(f"B={B} N={N} D={D} H={H} masked={masked}: y {errs['y']:.1e} dx {errs['dx']:.1e} worst param {max(v for k, v in errs.items() if k not in ('y', 'dx')):.1e} "
          f"launches {nl}" + ("  ok" if good else "  MISMATCH " + str({k: f'{v:.1e}' for k, v in errs.items() if v > 2e-5})), flush=True)


case(3, 4, 512, 4, False)
case(5, 4, 512, 4, True)
case(4, 7, 512, 8, True)
case(2, 15, 512, 4, True)
case(3, 3, 256, 4, False)
case(2, 11, 384, 8, False)

# dropout: directional derivative with the same seed (torch.manual_seed fixes the seed the module draws)
os.environ["B200CLIP_XFBLOCK"] = "1"
torch.manual_seed(3)
blk = TransformerBlock(512, 4, 0.2).to(dev).train()
x = torch.randn(4, 4, 512, device=dev); g = torch.randn(4, 4, 512, device=dev); dirx = torch.randn_like(x)
def f(xx):
    torch.manual_seed(11)
    return (blk(xx) * g).sum()
xr = x.clone().requires_grad_(True)
f(xr).backward()
ana = (xr.grad * dirx).sum().item()
h = 2e-3
num = (f(x + h * dirx).item() - f(x - h * dirx).item()) / (2 * h)
good = abs(ana - num) <= 2e-2 * max(abs(num), 1.0)
with torch.no_grad():
    torch.manual_seed(11); y1 = blk(x)
    blk.eval(); y0 = blk(x); blk.train()
frac_changed = ((y1 - y0).abs() > 1e-6).float().mean().item()
good &= frac_changed > 0.5
ok &= good
print(f"dropout 0.2: directional derivative analytic {ana:.4f} numeric {num:.4f}; outputs changed by dropout {frac_changed:.2f}" + ("  ok" if good else "  MISMATCH"), flush=True)

# the aggregator end to end (2 blocks + query-pool tail) against the same module on PyTorch blocks, then timing at C3
torch.manual_seed(5)
agg = EnhancedVideoAggregator(512).to(dev).eval()
xa = torch.randn(8, 4, 512, device=dev)
ga = torch.randn(8, 512, device=dev)
res = {}
for mode in ("1", "0"):
    os.environ["B200CLIP_XFBLOCK"] = mode
    agg.zero_grad(set_to_none=True)
    xr = xa.clone().requires_grad_(True)
    out = agg(xr)
    out.backward(ga)
    res[mode] = (out.detach(), xr.grad, {k: v.grad.clone() for k, v in agg.named_parameters() if v.grad is not None})
d = max([rel(res["1"][0], res["0"][0]), rel(res["1"][1], res["0"][1])] + [rel(res["1"][2][k], v) for k, v in res["0"][2].items() if v.norm() > 0])
good = d <= 1e-4
ok &= good
print(f"EnhancedVideoAggregator fused vs PyTorch blocks (fp32 both): worst relative difference {d:.1e}" + ("  ok" if good else "  MISMATCH"), flush=True)
agg.train()
for mode in ("1", "0"):
    os.environ["B200CLIP_XFBLOCK"] = mode
    def fb():
        agg.zero_grad(set_to_none=True)
        xr = xa.clone().requires_grad_(True)
        agg(xr).backward(ga)
    for _ in range(5): fb()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20): fb()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"XFBLOCK={mode}: aggregator fwd+bwd (8 studies x 4 views, depth 2, train mode) {e0.elapsed_time(e1) / 20 * 1e3:.0f} us "
          f"(host enqueue {(t1 - t0) / 20 * 1e6:.0f} us)", flush=True)
print("xfblock check", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
