"""A few SigLIP config-2 steps (8,192 x 8,192 pairs, D = 512, dense fp32 masks + weights) for an ncu launch list."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200.loss import SigLIPLoss
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
B = T = 8192; D = 512
t = torch.randn(T, D, device=dev, generator=g).bfloat16().float().requires_grad_(True)
v = torch.randn(B, D, device=dev, generator=g).requires_grad_(True)
pm = torch.zeros(B, T, device=dev); pm[torch.arange(B), torch.arange(B)] = 1.0
for _ in range(3):
    pm[torch.arange(B, device=dev), torch.randint(0, T, (B,), device=dev, generator=g)] = 1.0
pw = pm * torch.tensor([1.0, 1.5, 2.5, 3.0], device=dev)[torch.randint(0, 4, (B, T), device=dev, generator=g)]
lt = torch.tensor([math.log(0.087)], device=dev, requires_grad=True)
mod = SigLIPLoss().to(dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    v.grad = None; t.grad = None; lt.grad = None
    mod(v, t, lt, pos_mask=pm, pos_weights=pw).backward()
torch.cuda.synchronize(); print("done")
