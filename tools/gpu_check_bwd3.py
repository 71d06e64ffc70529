"""GPU bring-up check for the 64-row CTA-pair backward (logits_bwd3.cu; run under gpurun): forced on for every
supported shape with B200CLIP_BWD3=1, compared with a float64 torch evaluation of the same tile math."""
import os, sys, math
os.environ["B200CLIP_BWD3"] = "1"
sys.path.insert(0, ".")
import torch
from deepcoro_clip_b200 import _lib as L

torch.manual_seed(0)
dev = torch.device("cuda:0")
st = L.stream_ptr()
LOG2E = math.log2(math.e)

def run(mode, Nx, Ny, D, tau=0.0588, nseg=0, bias=-3.0, ydiag=0.0, diag_off=0):
    Dp = (D + 63) // 64 * 64
    x = torch.zeros(Nx, Dp, device=dev); y = torch.zeros(Ny, Dp, device=dev)
    x[:, :D] = torch.nn.functional.normalize(torch.randn(Nx, D, device=dev), dim=-1)
    y[:, :D] = torch.nn.functional.normalize(torch.randn(Ny, D, device=dev), dim=-1)
    x = x.bfloat16(); y = y.bfloat16()
    S = x.double() @ y.double().t()
    scale2 = LOG2E / tau; shift2 = scale2
    rs = torch.rand(Nx, device=dev) * 0.5 + 0.5
    cs = torch.rand(Ny, device=dev) * 0.5 + 0.5
    dX = torch.zeros(Nx, D, device=dev)
    scal = torch.zeros(4, device=dev, dtype=torch.float64)
    wneg_c = 0.37
    dc = torch.zeros(Nx, 2, device=dev) if ydiag else None
    L.call("logits_bwd", mode, x, y, Nx, Ny, Dp, Dp, D, 0, Dp, Dp, scale2, shift2, 1.0 / tau, bias, wneg_c, rs, cs,
           1.0 / tau, 1.0, 0, None, ydiag, diag_off, dc, dX, D, scal, nseg, st)
    torch.cuda.synchronize()
    if mode == 0:
        f = S; G = torch.exp((f - 1.0) / tau) * (rs.double()[:, None] + cs.double()[None, :]); GS = G
    elif mode == 1:
        sg = torch.sigmoid(S); f = S * sg; fp = sg * (1 + S * (1 - sg))
        G = torch.exp((f - 1.0) / tau) * (rs.double()[:, None] + cs.double()[None, :]); GS = G * fp
    else:
        R = S / tau + bias; Lc = R.clamp(-30, 30)
        G = wneg_c * torch.sigmoid(Lc) * (R.abs() <= 30); GS = G; f = S
        sp = torch.nn.functional.softplus(Lc)
    if ydiag:
        ii = torch.arange(Nx, device=dev); jj = ii + diag_off
        okd = jj < Ny
        G = G.clone(); G[ii[okd], jj[okd]] -= ydiag; GS = G
    ref = (GS @ y.double()[:, :D]) / tau
    refq = (GS.float().bfloat16().double() @ y.double()[:, :D]) / tau
    rel = ((dX.double() - ref).norm() / ref.norm()).item()
    relq = ((dX.double() - refq).norm() / refq.norm()).item()
    out = {}
    if ydiag:
        gd = GS[ii[okd], jj[okd]]
        out["diag_corr_sum_rel"] = ((dc[okd.nonzero().squeeze(1)].double().sum(1) - gd).abs().max() / gd.abs().max()).item()
    out.update({"rel": rel, "rel_vs_bf16G": relq, "scal0_rel": abs(scal[0].item() / (G * f).sum().item() - 1)})
    if mode == 2:
        out["loss_rel"] = abs(scal[1].item() / sp.sum().item() - 1)
        out["dbias_rel"] = abs(scal[2].item() / G.sum().item() - 1)
    return out

cases = [(0, 128, 128, 256, 0.0588, 0), (0, 128, 256, 256, 0.0588, 1), (0, 256, 384, 512, 0.0588, 0), (0, 200, 300, 250, 0.07, 0),
         (0, 1024, 2048, 512, 0.0588, 2), (0, 1024, 1024, 768, 0.0588, 0), (0, 333, 777, 768, 0.0588, 0),
         (1, 512, 512, 512, 0.1, 0), (2, 512, 640, 512, 0.087, 0), (2, 200, 300, 768, 0.087, 1),
         (0, 640, 640, 768, 0.0588, 0, -3.0, 1.0 / 640, 0), (0, 300, 900, 512, 0.0588, 0, -3.0, 0.5 / 900, 300),
         (0, 4096, 4096, 768, 0.0588, 0), (0, 8192, 8192, 512, 0.0588, 0), (2, 4096, 8192, 768, 0.087, 0)]
for c in cases:
    try:
        print("bwd", c, run(*c), flush=True)
    except Exception as ex:
        print("bwd", c, "FAILED", ex, flush=True); break

def timing(N, D):
    x = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=-1).bfloat16()
    y = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=-1).bfloat16()
    rs = torch.rand(N, device=dev); cs = torch.rand(N, device=dev)
    dX = torch.zeros(N, D, device=dev); scal = torch.zeros(4, device=dev, dtype=torch.float64)
    tau = 0.0588
    for nseg in (0, 2, 4):
        for it in range(2):
            L.call("logits_bwd", 0, x, y, N, N, D, D, D, 0, D, D, LOG2E / tau, LOG2E / tau, 1 / tau, 0.0, 0.0, rs, cs, 1 / tau, 1.0, 0, None, 0.0, 0, None, dX, D, scal, nseg, st)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(3):
            L.call("logits_bwd", 0, x, y, N, N, D, D, D, 0, D, D, LOG2E / tau, LOG2E / tau, 1 / tau, 0.0, 0.0, rs, cs, 1 / tau, 1.0, 0, None, 0.0, 0, None, dX, D, scal, nseg, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("bw3 pass 32k D=%d nseg=%d: %.3f ms  algorithmic %.1f TFLOP/s executed %.1f TFLOP/s" % (D, nseg, ms, 2 * N * N * D / ms / 1e9, 4 * N * N * D / ms / 1e9), flush=True)


for N, D in ((32768, 512), (32768, 768)):
    timing(N, D)
