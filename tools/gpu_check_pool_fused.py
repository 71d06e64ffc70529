"""Fused AttentionPool (pool_prep / attnpool_tc / pool_tail kernels, 3 + 4 launches) against (a) the reference module
structure in float64 (nn.MultiheadAttention + LayerNorm + Linear on the same parameters) and (b) the unfused host path of
this package (B200CLIP_POOL_FUSED=0), output and every gradient. 1 GPU."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from deepcoro_clip_b200 import AttentionPool
from deepcoro_clip_b200 import _lib

dev = torch.device("cuda", 0)
ok = True


class RefPool(nn.Module):          # models/attention_pool.py:10-101, restated
    def __init__(self, D, H, Do):
        super().__init__()
        self.query = nn.Parameter(torch.randn(1, 1, D))
        self.attn = nn.MultiheadAttention(D, H, batch_first=True)
        self.norm = nn.LayerNorm(D)
        self.proj = nn.Linear(D, Do) if Do != D else nn.Identity()

    def forward(self, x, mask=None):
        q = self.query.expand(x.shape[0], -1, -1)
        o, _ = self.attn(query=q, key=x, value=x, key_padding_mask=mask)
        return self.proj(self.norm(o)).squeeze(1)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def case(B, N, D, H, Do, dtype, masked, seed=0):
    global ok
    torch.manual_seed(seed)
    pool = AttentionPool(D, H, output_dim=Do).to(dev)
    with torch.no_grad():
        for p in pool.parameters():
            p.add_(0.05 * torch.randn_like(p))
        pool.query.mul_(20.0)
    x = torch.randn(B, N, D, device=dev).to(dtype)
    mask = None
    if masked:
        mask = torch.rand(B, N, device=dev) < 0.25
        mask[:, 0] = False
    gy = torch.randn(B, Do, device=dev).to(dtype)
    outs = {}
    for mode in ("1", "0"):
        os.environ["B200CLIP_POOL_FUSED"] = mode
        pool.zero_grad(set_to_none=True)
        xr = x.clone().requires_grad_(True)
        l0 = _lib.LAUNCHES
        y = pool(xr, mask)
        y.backward(gy)
        torch.cuda.synchronize()
        outs[mode] = (y.detach().float(), xr.grad.float(), {k: v.grad.clone() for k, v in pool.named_parameters() if v.grad is not None},
                      _lib.LAUNCHES - l0)
    ref = RefPool(D, H, Do).to(dev).double()
    ref.load_state_dict({k: v.double() for k, v in pool.state_dict().items()})
    xd = x.double().requires_grad_(True)
    yr = ref(xd, mask)
    yr.backward(gy.double())
    rg = {k: v.grad for k, v in ref.named_parameters()}
    yf, dxf, gf, nl = outs["1"]
    tol_y = 1.2e-2 if dtype == torch.bfloat16 else 2e-3          # the output is rounded to the 16-bit dtype
    errs = {"y": rel(yf, yr), "dx": rel(dxf, xd.grad)}
    for k, v in rg.items():
        if k == "attn.in_proj_bias":
            D_ = D
            errs[k] = max(rel(gf[k][:D_], v[:D_]), rel(gf[k][2 * D_:], v[2 * D_:]))      # the key-bias gradient is ~0 in the reference
        else:
            errs[k] = rel(gf[k], v)
    worst_param = max(v for k, v in errs.items() if k not in ("y", "dx"))
    good = errs["y"] <= tol_y and errs["dx"] <= tol_y and worst_param <= 2e-3
    # fused vs unfused host path of this package
    yu, dxu, gu, nlu = outs["0"]
    d2 = max([rel(yf, yu), rel(dxf, dxu)] + [rel(gf[k], gu[k]) for k in gu if gu[k].norm() > 0])
    good &= d2 <= 1e-2
    ok &= good
    print(f"B={B} N={N} D={D} H={H} Do={Do} {str(dtype)[6:]} masked={masked}: vs fp64 reference y {errs['y']:.1e} dx {errs['dx']:.1e} "
          f"params (worst) {worst_param:.1e} | vs unfused {d2:.1e} | launches fused {nl} unfused {nlu}" + ("  ok" if good else "  MISMATCH " + str({k: f'{v:.1e}' for k, v in errs.items()})), flush=True)


case(5, 300, 256, 8, 256, torch.bfloat16, False)
case(6, 777, 512, 8, 512, torch.bfloat16, True)
case(3, 500, 512, 8, 256, torch.bfloat16, True)          # output Linear
case(9, 200, 512, 4, 512, torch.float16, False)
case(2, 3136, 512, 8, 512, torch.bfloat16, False)

# timing at C3: eager forward + backward of the module
import time
B, N, D = 32, 3136, 512
x = torch.randn(B, N, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
gy = torch.randn(B, D, device=dev, dtype=torch.bfloat16)
pool = AttentionPool(D, 8).to(dev)
for mode in ("1", "0"):
    os.environ["B200CLIP_POOL_FUSED"] = mode
    def fb():
        x.grad = None
        pool.zero_grad(set_to_none=True)
        pool(x).backward(gy)
    for _ in range(5): fb()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20): fb()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    with torch.no_grad():
        e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(20): pool(x)
        e3.record(); torch.cuda.synchronize()
    print(f"POOL_FUSED={mode}: eager fwd+bwd {e0.elapsed_time(e1) / 20 * 1e3:.0f} us (host enqueue {(t1 - t0) / 20 * 1e6:.0f} us), fwd only "
          f"{e2.elapsed_time(e3) / 20 * 1e3:.0f} us", flush=True)
print("pool fused check", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
