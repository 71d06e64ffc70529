"""GPU bring-up check for the tcgen05 tile engine (run under gpurun). Not part of the test-suite."""
import sys, time, json, math
sys.path.insert(0, ".")
import torch
from deepcoro_clip_b200 import _lib as L

torch.manual_seed(0)
dev = torch.device("cuda:0")
st = L.stream_ptr()
res = {}

def dump(Ma, Nb, K, max_ctas=0):
    A = torch.randn(Ma, K, device=dev).bfloat16()
    B = torch.randn(Nb, K, device=dev).bfloat16()
    out = torch.full((Ma, Nb), float("nan"), device=dev)
    L.call("logits_dump", A, B, Ma, Nb, K, K, K, out, Nb, max_ctas, st)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (out - ref).abs().max().item()
    return err, ref.abs().max().item()

for shape in [(128, 256, 64, 0), (128, 256, 512, 0), (256, 512, 128, 0), (1024, 1024, 512, 0), (1024, 2048, 512, 3),
              (64, 64, 512, 0), (200, 300, 192, 0), (4096, 4096, 768, 0)]:
    try:
        e, m = dump(*shape)
        print("dump", shape, "maxerr", e, "refmax", m, flush=True)
        res[str(shape)] = e
    except Exception as ex:
        print("dump", shape, "FAILED", ex, flush=True)
        res[str(shape)] = str(ex)
        break

# LSE forward check
def lse(N, M, K, tau, gated=0):
    a = torch.nn.functional.normalize(torch.randn(N, K, device=dev), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(M, K, device=dev), dim=-1).bfloat16()
    rs = torch.zeros(N, device=dev); cs = torch.zeros(M, device=dev)
    scale2 = math.log2(math.e) / tau
    mx = (0.7311 if gated else 1.0) / tau
    shift2 = mx * math.log2(math.e)
    L.call("logits_lse_fwd", a, b, N, M, K, K, K, scale2, shift2, gated, None, 0, rs, cs, None, 0, st)
    torch.cuda.synchronize()
    S = a.double() @ b.double().t()
    if gated: S = S * torch.sigmoid(S)
    Lg = S / tau
    r_ref = torch.logsumexp(Lg, dim=1); c_ref = torch.logsumexp(Lg, dim=0)
    r = (torch.log2(rs.double()) + shift2) * math.log(2); c = (torch.log2(cs.double()) + shift2) * math.log(2)
    return (r - r_ref).abs().max().item(), (c - c_ref).abs().max().item()

for args in [(64, 64, 512, 0.07, 0), (1000, 777, 512, 0.0588, 0), (4096, 4096, 512, 0.0588, 0), (2048, 2048, 512, 0.1, 1)]:
    try:
        print("lse", args, lse(*args), flush=True)
    except Exception as ex:
        print("lse", args, "FAILED", ex, flush=True); break

# timing of the forward at the headline size
N, K = 32768, 512
a = torch.nn.functional.normalize(torch.randn(N, K, device=dev), dim=-1).bfloat16()
b = torch.nn.functional.normalize(torch.randn(N, K, device=dev), dim=-1).bfloat16()
rs = torch.zeros(N, device=dev); cs = torch.zeros(N, device=dev)
tau = 0.0588
for it in range(3):
    L.call("logits_lse_fwd", a, b, N, N, K, K, K, 1.4427/tau, 1.4427/tau, 0, None, 0, rs, cs, None, 0, st)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    L.call("logits_lse_fwd", a, b, N, N, K, K, K, 1.4427/tau, 1.4427/tau, 0, None, 0, rs, cs, None, 0, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("lse_fwd 32k x 32k x 512: %.3f ms  -> %.1f TFLOP/s" % (ms, 2 * N * N * K / ms / 1e9), flush=True)
t0 = time.time()
c = a @ b.t(); torch.cuda.synchronize()
e0.record()
for it in range(5): c = a @ b.t()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("torch bf16 matmul same shape: %.3f ms -> %.1f TFLOP/s" % (ms, 2 * N * N * K / ms / 1e9), flush=True)
