"""Gated-attention MIL pooling (csrc/milpool.cu): parity against the same formulas in PyTorch float64 at a mid-size shape,
then timings against the reference's op sequence in PyTorch fp32 on the GPU at the two-level token shape. 1 GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from deepcoro_clip_b200 import GatedAttentionPooling

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False


def ref_level(mod, x, mask):        # models/multi_instance_linear_probing.py:499-507 restated
    a = mod.attention_w(torch.tanh(mod.attention_V(x)) * torch.sigmoid(mod.attention_U(x)))
    if mask is not None:
        a = a.masked_fill(~mask.unsqueeze(-1), float("-inf"))
    return (F.softmax(a, dim=1) * x).sum(dim=1)


def ref(mod, x, mask):
    if x.dim() == 4:
        B, N, L, D = x.shape
        return ref_level(mod, ref_level(mod, x.view(B * N, L, D), None).view(B, N, D), mask)
    return ref_level(mod, x, mask)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


ok = True
for shape, hd in (((6, 300, 256), 64), ((2, 3, 500, 128), 40), ((16, 5, 512), 128), ((4, 3, 600, 512), 512), ((3, 500, 768), 96)):
    torch.manual_seed(0)
    mod = GatedAttentionPooling(shape[-1], hd).to(dev)
    with torch.no_grad():
        mod.attention_V.bias.normal_(std=0.3); mod.attention_U.bias.normal_(std=0.3); mod.attention_w.weight.mul_(3.0)
    x = torch.randn(*shape, device=dev)
    mask = torch.rand(shape[:2], device=dev) > 0.3
    mask[:, 0] = True
    g = torch.randn(shape[0], shape[-1], device=dev)
    xr = x.clone().requires_grad_(True)
    out = mod(xr, mask); out.backward(g)
    got = [out, xr.grad] + [p.grad.clone() for p in mod.parameters()]
    md = __import__("copy").deepcopy(mod).double(); md.zero_grad()
    xd = x.double().requires_grad_(True)
    od = ref(md, xd, mask); od.backward(g.double())
    want = [od, xd.grad] + [p.grad for p in md.parameters()]
    errs = [rel(a, b) for a, b in zip(got[:-1], want[:-1])]      # the last one is d/d(bw) = 0 (softmax shift invariance)
    errs.append(abs(got[-1].item()) / max(1.0, want[-2].abs().max().item()))
    good = max(errs) <= (1e-4 if (x.numel() // shape[-1] >= 1024 and shape[-1] in (256, 512, 768)) else 2e-5)   # tensor-core variant: 16-bit-mantissa operands
    ok &= good
    print(f"{shape} hidden {hd}: out {errs[0]:.1e} dx {errs[1]:.1e} worst param {max(errs[2:]):.1e}" + ("  ok" if good else "  MISMATCH " + str(errs)), flush=True)

for shape, hd in (((8, 4, 1568, 512), 128), ((8, 4, 1568, 512), 512), ((32, 4, 512), 128)):
    mod = GatedAttentionPooling(shape[-1], hd).to(dev)
    x = torch.randn(*shape, device=dev, requires_grad=True)
    mask = torch.ones(shape[:2], dtype=torch.bool, device=dev)
    g = torch.randn(shape[0], shape[-1], device=dev)
    def ours():
        x.grad = None; mod.zero_grad(set_to_none=True)
        mod(x, mask).backward(g)
    def theirs():
        x.grad = None; mod.zero_grad(set_to_none=True)
        ref(mod, x, mask).backward(g)
    with torch.no_grad():
        tf_o, tf_r = timeit(lambda: mod(x, mask)), timeit(lambda: ref(mod, x, mask))
    print(f"{shape} hidden {hd}: forward {tf_o:.0f} us (PyTorch fp32 ops {tf_r:.0f} us); fwd+bwd {timeit(ours):.0f} us ({timeit(theirs):.0f} us)", flush=True)
print("milpool check", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
