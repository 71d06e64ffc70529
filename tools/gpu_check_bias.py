import sys; sys.path.insert(0, ".")
import torch, math
from deepcoro_clip_b200 import _lib as L
dev = torch.device("cuda:0"); st = L.stream_ptr()
torch.manual_seed(0)
for K in (64, 256, 512, 1536):
    A = torch.rand(256, K, device=dev).bfloat16(); B = torch.rand(256, K, device=dev).bfloat16()
    out = torch.empty(256, 256, device=dev)
    L.call("logits_dump", A, B, 256, 256, K, K, K, out, 256, 0, st); torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    rel = ((out.double() - ref) / ref)
    t32 = (A.float() @ B.float().t()).double()
    print("K", K, "mean signed rel err mma", rel.mean().item(), "max abs rel", rel.abs().max().item(), "| torch fp32 matmul mean", ((t32 - ref) / ref).mean().item())
# normalized correlated vectors as in the failing test
N, D = 1000, 512
v = torch.randn(N, D, device=dev); t = 0.5 * v + torch.randn(N, D, device=dev)
vh = torch.nn.functional.normalize(v, dim=-1); th = torch.nn.functional.normalize(t, dim=-1)
def split3(x, role):
    hi = x.bfloat16(); lo = (x - hi.float()).bfloat16()
    return torch.cat([hi, hi, lo] if role == 0 else [hi, lo, hi], dim=1).contiguous()
A3 = split3(vh, 0); B3 = split3(th, 1)
out = torch.empty(N, N, device=dev)
L.call("logits_dump", A3, B3, N, N, 3 * D, 3 * D, 3 * D, out, N, 0, st); torch.cuda.synchronize()
ref = vh.double() @ th.double().t()
d = out.double() - ref
print("x3 S err: diag mean signed", d.diag().mean().item(), "diag mean rel", (d.diag() / ref.diag()).mean().item(), "offdiag rms", d.pow(2).mean().sqrt().item(), "offdiag mean", d.mean().item())
# what the 3-term formula gives in exact arithmetic
hi_v = vh.bfloat16().double(); lo_v = (vh - vh.bfloat16().float()).bfloat16().double()
hi_t = th.bfloat16().double(); lo_t = (th - th.bfloat16().float()).bfloat16().double()
ex3 = hi_v @ hi_t.t() + hi_v @ lo_t.t() + lo_v @ hi_t.t()
d3 = ex3 - ref
print("exact 3-term formula err: diag mean", d3.diag().mean().item(), "rms", d3.pow(2).mean().sqrt().item())
print("mma vs exact 3-term: diag mean", (out.double() - ex3).diag().mean().item(), "rms", (out.double() - ex3).pow(2).mean().sqrt().item())
