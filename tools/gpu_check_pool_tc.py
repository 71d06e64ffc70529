"""tcgen05 attention-pool kernels (csrc/attnpool_tc.cu) against a float64 torch evaluation of the same closed forms and
against the mma.sync path (B200CLIP_POOL_TC=0), plus kernel-only timings at the C3 shape. 1 GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200.attention_pool import _StreamPool, _tc_splits

dev = torch.device("cuda", 0)
ok = True


def ref(x, qt, mask, dxbar, dlse=None):
    x64 = x.double().requires_grad_(True); q64 = qt.double().requires_grad_(True)
    s = torch.einsum("bnd,hd->bhn", x64, q64)
    if mask is not None:
        s = s.masked_fill(mask[:, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    xbar = torch.einsum("bhn,bnd->bhd", a, x64)
    lse = torch.logsumexp(s, dim=-1)
    obj = (xbar * dxbar.double()).sum()
    if dlse is not None:
        obj = obj + (lse * dlse.double()).sum()
    obj.backward()
    return xbar.detach(), lse.detach(), x64.grad, q64.grad


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def case(B, N, D, H, dtype, masked=False, with_lse=False, scale=1.0, seed=0):
    global ok
    g = torch.Generator(device=dev).manual_seed(seed)
    x = (scale * torch.randn(B, N, D, device=dev, generator=g)).to(dtype)
    qt = 0.05 * torch.randn(H, D, device=dev, generator=g)
    mask = None
    if masked:
        mask = torch.rand(B, N, device=dev, generator=g) < 0.3
        mask[:, 0] = False
        if B > 1:
            mask[1, : min(N - 1, 130)] = True        # whole leading tiles masked
    dxbar = torch.randn(B, H, D, device=dev, generator=g)
    dlse = torch.randn(B, H, device=dev, generator=g) if with_lse else None
    assert _tc_splits(x, B, N, D, H) > 0, "tc path not selected"
    xr = x.clone().requires_grad_(True); qr = qt.clone().requires_grad_(True)
    xbar, sa, lse = _StreamPool.apply(xr, qr, mask, 0.0, 0, with_lse)
    obj = (xbar * dxbar).sum()
    if with_lse:
        obj = obj + (lse * dlse).sum()
    obj.backward()
    torch.cuda.synchronize()
    rx, rl, rdx, rdq = ref(x, qt, mask, dxbar, dlse)
    e = [rel(xbar, rx), rel(xr.grad, rdx), rel(qr.grad, rdq)]
    if with_lse:
        e.append(rel(lse, rl))
    tol = [2e-5, 6e-3 if dtype == torch.bfloat16 else 1.5e-3, 5e-5, 1e-6]
    good = all(v <= t for v, t in zip(e, tol)) and torch.isfinite(xr.grad).all().item()
    ok &= good
    print(f"B={B} N={N} D={D} H={H} {str(dtype)[6:]} masked={masked} lse={with_lse} scale={scale}: xbar {e[0]:.2e} dx {e[1]:.2e} dqt {e[2]:.2e}"
          + (f" lse {e[3]:.2e}" if with_lse else "") + ("  ok" if good else "  MISMATCH"), flush=True)


case(2, 64, 128, 8, torch.bfloat16)
case(3, 200, 256, 8, torch.bfloat16)
case(3, 200, 256, 8, torch.bfloat16, masked=True)
case(2, 777, 512, 8, torch.bfloat16, masked=True, with_lse=True)
case(2, 1000, 384, 4, torch.bfloat16)
case(2, 640, 512, 8, torch.bfloat16, scale=6.0, seed=3)          # large scores: the reference maximum must move
case(4, 3136, 512, 8, torch.bfloat16)
if "--fp16" in sys.argv:
    case(2, 1000, 384, 4, torch.float16)
    case(2, 777, 512, 8, torch.float16, masked=True, with_lse=True)
    print("fp16 cases", "ok" if ok else "FAILED"); sys.exit(0 if ok else 1)

# kernel-only timing at C3
B, N, D, H = 32, 3136, 512, 8
x = torch.randn(B, N, D, device=dev).to(torch.bfloat16)
qt = 0.05 * torch.randn(H, D, device=dev)
dxbar = torch.randn(B, H, D, device=dev)
for tc in ("1", "0"):
    os.environ["B200CLIP_POOL_TC"] = tc
    xr = x.clone().requires_grad_(True); qr = qt.clone().requires_grad_(True)
    def f():
        xbar, _, _ = _StreamPool.apply(xr, qr, None, 0.0, 0, False)
        return xbar
    def fb():
        xr.grad = None; qr.grad = None
        f().backward(dxbar)
    for _ in range(3): fb()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
    tf = e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20): fb()
    e1.record(); torch.cuda.synchronize()
    tfb = e0.elapsed_time(e1) / 20
    nb = B * N * D * 2
    print(f"POOL_TC={tc}: stream fwd (kernel + merge) {tf * 1e3:.1f} us = {nb / tf / 1e6:.0f} GB/s; fwd+bwd {tfb * 1e3:.1f} us "
          f"(bwd alone ~{(tfb - tf) * 1e3:.1f} us = {2 * nb / (tfb - tf) / 1e6:.0f} GB/s for read + write)", flush=True)
print("pool tc check", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
