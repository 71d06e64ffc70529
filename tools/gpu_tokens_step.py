"""One C3 token step (RoPE, AttentionPool, aggregator; forward + backward) a few times — the command profiled for
profiles/r02_tokens_launches.csv (ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcoro_clip_b200 import AttentionPool, EnhancedVideoAggregator, Rope3D
dev = torch.device("cuda", 0)
torch.manual_seed(2)
S, V, L, D, Hh, Dh = 8, 4, 3136, 512, 8, 96
q = torch.randn(S * V, Hh, L, Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
k = torch.randn(S * V, Hh, L, Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
gq = torch.randn_like(q); gk = torch.randn_like(k)
rope = Rope3D(Hh * Dh, Hh).to(dev).eval()
x = torch.randn(S * V, L, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
pool = AttentionPool(D, 8).to(dev)
agg = EnhancedVideoAggregator(D).to(dev)
gy = torch.randn(S * V, D, device=dev, dtype=torch.bfloat16)
xa = torch.randn(S, V, D, device=dev, requires_grad=True); ga = torch.randn(S, D, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for _ in range(n):
    qo, ko = rope(q, k, 16, 14, 14)
    torch.autograd.backward((qo, ko), (gq, gk))
    pool(x).backward(gy)
    agg(xa).backward(ga)
torch.cuda.synchronize()
print("done")
