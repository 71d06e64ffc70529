"""C2: SigLIP multi-positive loss fwd+bwd at 8,192 x 8,192 pairs, D = 512 (single GPU holds the whole global batch)."""
import math, sys, time
sys.path.insert(0, ".")
import torch
from deepcoro_clip_b200.loss import SigLIPLoss
dev = torch.device("cuda:0")
B = T = 8192; D = 512
g = torch.Generator(device=dev).manual_seed(1)
t = torch.randn(T, D, device=dev, generator=g).bfloat16().float().requires_grad_(True)
v = torch.randn(B, D, device=dev, generator=g).bfloat16().float().requires_grad_(True)
pm = torch.zeros(B, T, device=dev); pm[torch.arange(B), torch.arange(B)] = 1.0
for _ in range(3):
    pm[torch.arange(B, device=dev), torch.randint(0, T, (B,), device=dev, generator=g)] = 1.0
pw = pm * torch.tensor([1.0, 1.5, 2.5, 3.0], device=dev)[torch.randint(0, 4, (B, T), device=dev, generator=g)]
lt = torch.tensor([math.log(0.087)], device=dev, requires_grad=True)
mod = SigLIPLoss(precision="bf16").to(dev)
def step(mask=True):
    v.grad = None; t.grad = None; lt.grad = None; mod.bias.grad = None
    loss = mod(v, t, lt, pos_mask=pm if mask else None, pos_weights=pw if mask else None)
    loss.backward()
    return loss
for name, m in (("multi-positive masks", True), ("diagonal targets", False)):
    for _ in range(5): step(m)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step(m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"siglip 8192x8192x512 {name}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} samples/s  algorithmic {6.0 * B * T * D / ms / 1e9:.0f} TFLOP/s")
