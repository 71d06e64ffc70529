"""One forward + backward of the two-level gated-attention MIL pooling at the token shape (profiling target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import GatedAttentionPooling
dev = torch.device("cuda", 0)
hd = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mod = GatedAttentionPooling(512, hd).to(dev)
x = torch.randn(8, 4, 1568, 512, device=dev, requires_grad=True)
g = torch.randn(8, 512, device=dev)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    mod(x).backward(g)
torch.cuda.synchronize()
print("ok")
