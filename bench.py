#!/usr/bin/env python
"""bench.py — benchmark of the B200-native contrastive head (one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--legs a,b,...]

Headline workload (BASELINE.json `metric`):
  clip32k  (default)  CLIP InfoNCE loss fwd+bwd, global batch N = 32,768, D = 512, bf16 operands / fp32 accumulate.
                      N GPUs: STRONG scaling — the global batch is fixed and row-sharded, each rank owns a row slab of
                      the logits (embeddings all-gathered, per-column statistics all-reduced; NCCL).
  A "step" = one loss forward + backward over the whole global batch (normalise, logits tiles, LSE, both gradient passes).
  `value`   = the PLUGIN path a drop-in user runs (`CLIPLoss.forward` -> `loss.backward()`, eager module calls), inputs
              resident in HBM;   `graphed` = the same step replayed from a CUDA graph (GraphedLossStep);
  `e2e`     = the plugin path with pinned-host inputs copied in and the loss copied out every step;
  `parity`  = the measured step against a float64 restatement of the reference formulas on the SAME global batch (every
              rank generates the same global batch, so the N-GPU result is checked, not just timed).
Extra legs in the same line (each with its own roofline and CPU baseline): BASELINE configs 2-5
  siglip_c2     SigLIP multi-positive loss, global 8 x 1024 rows x 8,192 texts, D = 512, masks + severity weights
  tokens_c3     study mode: RoPE3D + AttentionPool x 4 views + aggregator tail fwd+bwd, 8 studies, bf16 (replicas per GPU)
  retrieval     streaming recall@1/5/10 + MRR over 203,808 x 32,473 (text shards across GPUs)
  topk10        explicit top-10 lists over the same sweep
  clip32k_d768  InfoNCE fwd+bwd at N = 32,768, D = 768
`--impl reference` times the reference's own CPU implementation of the headline step on the host cores (the unmodified
reference class from oracle/_ref when the build container provided it, else the op-for-op torch transcription).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "contrastive loss fwd+bwd samples/s @B=32k,D=512"
UNIT = "samples/s"
WORKLOADS = {"clip32k": (32768, 512), "clip32k_d768": (32768, 768), "clip8k": (8192, 512)}
ALL_LEGS = ["retrieval", "topk10", "siglip_c2", "tokens_c3", "clip32k_d768"]
TAU = 0.0588   # config/clip/base_config.yaml:46
L2_BYTES = 126e6


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_first_sample(self, timeout: float = 3.0) -> None:
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")) -> dict:
        """Statistics over the samples taken inside [t_begin, t_end] (host clock around the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        inside = [(ts, ln) for ts, ln in self.lines if t_begin <= ts <= t_end + 0.025]
        note = None
        if not inside and self.lines:      # region shorter than the 20 ms sampling period: the nearest samples
            mid = 0.5 * (t_begin + min(t_end, time.time()))
            inside = sorted(self.lines, key=lambda x: abs(x[0] - mid))[:3]
            note = "timed region shorter than the 20 ms sampling period: nearest samples"
        for ts, ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# =====================================================================================================================
# CPU baselines: the reference's own code on the host cores (oracle/_ref = the unmodified reference files copied by
# oracle/make_ref.py in the build container; otherwise the op-for-op torch transcription oracle/reference_torch_port.py)
# =====================================================================================================================
def _reference_available() -> bool:
    from oracle.make_ref import import_ref
    sys.dont_write_bytecode = True
    return import_ref()


def _best_of(fn, steps: int, warmup: int):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), sum(ts) / len(ts)


def cpu_clip_step(N: int, D: int, steps: int, warmup: int):
    """CLIPLoss fwd+bwd on torch CPU, fp32, every host thread, on a bounded sample of N rows.
    Returns (best s, mean s, threads, kind)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(5)
    v = torch.randn(N, D, generator=g)
    t = torch.randn(N, D, generator=g)
    if _reference_available():
        from utils.loss.contrastive import CLIPLoss           # oracle/_ref: the unmodified reference class
        mod = CLIPLoss()

        def fn():
            vv = v.clone().requires_grad_(True); tt = t.clone().requires_grad_(True)
            lt = torch.tensor([math.log(TAU)], requires_grad=True)
            mod(vv, tt, lt).backward()
        kind = "reference"
    else:
        from oracle import reference_torch_port as tp

        def fn():
            tp.clip_loss_step(v, t, math.log(TAU))
        kind = "port"
    best, mean = _best_of(fn, steps, warmup)
    return best, mean, torch.get_num_threads(), kind


def cpu_clip_baseline(N: int, D: int, Ns: int = 4096):
    best, _, cores, kind = cpu_clip_step(Ns, D, 3, 1)
    what = ("the unmodified reference CLIPLoss (oracle/_ref/utils/loss/contrastive.py)" if kind == "reference" else
            "reference op sequence (oracle/reference_torch_port.py)")
    return {"value": (Ns / best) * (Ns / N), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{what} on torch CPU, fp32, all host threads: fwd+bwd at N={Ns}, D={D}: {Ns / best:.0f} samples/s "
                      f"({best * 1e3:.0f} ms), N^2-extrapolated to N={N}"}


def cpu_siglip_baseline(Bg: int, T: int, D: int, Bs: int = 2048):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1)
    t = torch.randn(Bs, D, generator=g)
    v = torch.randn(Bs, D, generator=g)
    pm = torch.zeros(Bs, Bs); pm[torch.arange(Bs), torch.arange(Bs)] = 1.0
    for _ in range(3):
        pm[torch.arange(Bs), torch.randint(0, Bs, (Bs,), generator=g)] = 1.0
    pw = pm * torch.tensor([1.0, 1.5, 2.5, 3.0])[torch.randint(0, 4, (Bs, Bs), generator=g)]
    if _reference_available():
        from utils.loss.contrastive import SigLIPLoss
        mod = SigLIPLoss()

        def fn():
            vv = v.clone().requires_grad_(True); tt = t.clone().requires_grad_(True)
            lt = torch.tensor([math.log(0.087)], requires_grad=True)
            mod(vv, tt, lt, pos_mask=pm, pos_weights=pw).backward()
        kind = "reference"
    else:
        from oracle import reference_torch_port as tp

        def fn():
            tp.siglip_loss_step(v, t, math.log(0.087), pm, pw)
        kind = "port"
    best, _ = _best_of(fn, 3, 1)
    scale = (Bs * Bs) / (Bg * T)
    return {"value": (Bs / best) * scale * (Bg / Bs), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"SigLIPLoss fwd+bwd on torch CPU (fp32, all host threads) at {Bs} x {Bs} pairs, D={D}: "
                      f"{best * 1e3:.0f} ms, pair-count-extrapolated to {Bg} x {T}"}


def cpu_retrieval_baseline(M: int, D: int, rows: int = 8192):
    """recall@1/5/10 + MRR on a bounded slice of the C4 sweep: `rows` videos against the full text database (the cost is
    linear in the video rows, so the slice's Gsim/s is the sweep's)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(3)
    v = torch.randint(-127, 128, (rows, D), generator=g).float() / 128
    t = torch.randint(-127, 128, (M, D), generator=g).float() / 128
    gt = torch.randint(0, M, (rows,), generator=g)
    if _reference_available():
        from utils.retrieval_metrics_streaming import compute_metrics_streaming

        def fn(n):
            compute_metrics_streaming(v[:n], t, gt[:n], k_values=[1, 5, 10], device="cpu")
        kind, what = "reference", "the unmodified reference compute_metrics_streaming (oracle/_ref)"
    else:
        from oracle import reference_torch_port as tp

        def fn(n):
            tp.retrieval_metrics_step(v[:n], t, gt[:n])
        kind, what = "port", "reference op sequence (oracle/reference_torch_port.py)"
    fn(1024)
    t0 = time.perf_counter()
    fn(rows)
    dt = time.perf_counter() - t0
    return {"value": rows * M / dt / 1e9, "unit": "Gsim/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{what} on torch CPU: {rows} videos x {M} texts x {D} in {dt:.2f} s (cost linear in the video rows)"}


def cpu_tokens_baseline():
    """One study (4 views x 3,136 tokens) through the reference's Rope3D + AttentionPool + EnhancedVideoAggregator fwd+bwd
    on the host (bf16 autocast off: fp32, what CPU autocast leaves these modules in)."""
    import torch
    if not _reference_available():
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    from models.attention_pool import AttentionPool
    from models.rope_3d import Rope3D
    from models.video_aggregator import EnhancedVideoAggregator
    torch.manual_seed(2)
    V, L, D, Hh, Dh = 4, 3136, 512, 8, 96
    rope = Rope3D(Hh * Dh, Hh).eval()
    pool = AttentionPool(D, 8, dropout=0.0)
    agg = EnhancedVideoAggregator(embedding_dim=D)
    q = torch.randn(V, Hh, L, Dh, requires_grad=True); k = torch.randn(V, Hh, L, Dh, requires_grad=True)
    x = torch.randn(V, L, D, requires_grad=True)

    def fn():
        qo, ko = rope(q, k, 16, 14, 14)
        (qo.sum() + ko.sum()).backward()
        y = pool(x)                                  # [V, D]
        out = agg(y.unsqueeze(0).float())
        out.sum().backward()
    best, _ = _best_of(fn, 2, 1)
    return {"value": 1.0 / best, "unit": "studies/s", "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"unmodified reference Rope3D + AttentionPool x 4 views + EnhancedVideoAggregator (oracle/_ref/models) "
                      f"fwd+bwd on torch CPU, fp32, ONE study (4 x 3136 tokens): {best * 1e3:.0f} ms"}


def run_reference(args, N, D):
    """The reference arm: the reference's own CPU implementation of the headline step on this box's host cores. Each step is
    a bounded sample (N = 4096 rows; the loss is O(N^2 D)), extrapolated to the arm's N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Ns = 4096
    best, mean, cores, kind = cpu_clip_step(Ns, D, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)))
    sps_sample = Ns / best
    value = sps_sample * (Ns / N)               # samples/s the CPU path would reach at the full N (N^2 scaling)
    what = ("the unmodified reference CLIPLoss (oracle/_ref)" if kind == "reference" else
            "reference op sequence (oracle/reference_torch_port.py)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best * 1e3 * (N / Ns) ** 2,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "global_batch": N, "dim": D, "tau": TAU},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{what} on torch CPU, fp32, all host threads: fwd+bwd at N={Ns}: {sps_sample:.0f} "
                                   f"samples/s, N^2-extrapolated to N={N}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# device-side helpers
# =====================================================================================================================
class Ctx:
    def __init__(self, dev, world, rank):
        self.dev, self.world, self.rank = dev, world, rank

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()


def timed(ctx: Ctx, fn, steps: int, warmup: int, flush_buf=None) -> float:
    """ms per call: CUDA events on the current stream, barrier + synchronize on both sides, MAX over ranks. With
    ``flush_buf`` a 256 MB buffer is rewritten between the calls (outside the per-call event pairs)."""
    import torch
    for _ in range(warmup):
        fn()
    ctx.barrier()
    if flush_buf is not None:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush_buf.fill_(float(i))
            evs[i][0].record()
            fn()
            evs[i][1].record()
        ctx.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        ctx.barrier()
        ms = e0.elapsed_time(e1)
    return ctx.max_over_ranks(ms) / steps


def clip_fp64(v, t, log_tau: float, rows):
    """float64 restatement of CLIPLoss (utils/loss/contrastive.py:146-164) in plain torch on the GPU, chunked: the loss and
    the gradient rows ``rows`` (global indices) of dvideo / dtext. Checker only."""
    import torch
    vh = torch.nn.functional.normalize(v.double(), dim=-1)
    th = torch.nn.functional.normalize(t.double(), dim=-1)
    N = v.shape[0]
    tau = math.exp(log_tau)
    r = torch.empty(N, dtype=torch.float64, device=v.device)
    colmax = torch.full((N,), -float("inf"), dtype=torch.float64, device=v.device)
    colsum = torch.zeros(N, dtype=torch.float64, device=v.device)
    diag = torch.empty(N, dtype=torch.float64, device=v.device)
    step = 2048
    for a in range(0, N, step):
        L = vh[a:a + step] @ th.T / tau
        r[a:a + step] = torch.logsumexp(L, dim=1)
        diag[a:a + step] = L[torch.arange(L.shape[0]), torch.arange(a, a + L.shape[0])]
        m = torch.maximum(colmax, L.max(dim=0).values)
        colsum = colsum * torch.exp(colmax - m) + torch.exp(L - m).sum(dim=0)
        colmax = m
        del L
    c = colmax + torch.log(colsum)
    loss = 0.5 * ((r - diag).mean() + (c - diag).mean())
    L = vh[rows] @ th.T / tau
    G = (torch.exp(L - r[rows, None]) + torch.exp(L - c[None, :])) / (2 * N)
    G[torch.arange(len(rows)), rows] -= 1.0 / N
    dvh = G @ th / tau
    vn = v[rows].double().norm(dim=1, keepdim=True)
    dv = (dvh - (dvh * vh[rows]).sum(1, keepdim=True) * vh[rows]) / vn
    Lt = th[rows] @ vh.T / tau
    Gt = (torch.exp(Lt - c[rows, None]) + torch.exp(Lt - r[None, :])) / (2 * N)
    Gt[torch.arange(len(rows)), rows] -= 1.0 / N
    dth = Gt @ vh / tau
    tn = t[rows].double().norm(dim=1, keepdim=True)
    dt = (dth - (dth * th[rows]).sum(1, keepdim=True) * th[rows]) / tn
    return loss.item(), dv, dt


def bwd_kernel_roofline(ctx: Ctx, B: int, N: int, D: int, log_temp, step_ms: float, ncu_summary: str | None):
    """The dominant kernel (one gradient GEMM launch of logits_bwd) timed alone on its stream with CUDA events."""
    import torch
    from deepcoro_clip_b200 import ops
    dev = ctx.dev
    bf16_burst, bf16_sust, _, src = peaks()
    Kp = ops.round_up(D, 64)
    x = torch.nn.functional.normalize(torch.randn(B, Kp, device=dev), dim=-1).bfloat16()
    y = torch.nn.functional.normalize(torch.randn(N, Kp, device=dev), dim=-1).bfloat16()
    rs = torch.full((B,), 0.5 / N, device=dev); cs = torch.full((N,), 0.5 / N, device=dev)
    dX = torch.zeros(B, D, device=dev); scal = torch.zeros(4, dtype=torch.float64, device=dev)
    dyn = ops.dyn_prep(log_temp, None, 1e-4, 1.0)
    for _ in range(3):
        ops.logits_bwd(0, x, y, B, N, Kp, Kp, D, dyn, rs, cs, dX, scal, gnorm=2.0 * N)
    torch.cuda.synchronize()
    reps = 10
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.logits_bwd(0, x, y, B, N, Kp, Kp, D, dyn, rs, cs, dX, scal, gnorm=2.0 * N)
    e1.record(); torch.cuda.synchronize()
    kms = e0.elapsed_time(e1) / reps
    alg = 2.0 * B * N * D                      # algorithmic FLOPs of one launch (one gradient GEMM; recompute not counted)
    achieved = alg / (kms * 1e-3) / 1e12
    bw3 = Kp == D and Kp % 256 == 0 and Kp <= 768 and os.environ.get("B200CLIP_BWD3", "1") != "0"
    if bw3:
        kname = (f"bw3_kernel<CLIP, {256 if Kp <= 512 else 128}> (logits_bwd3.cu: 64-row CTA pairs, cta_group::2 M=128, whole "
                 "output width in TMEM)")
        executed = 2.0 * achieved                    # S once + the output product
    else:
        kname = "bw2_kernel / bw_kernel<CLIP> (128-row kernels: S recomputed per 256-column slice of D)"
        executed = (1 + (Kp + 255) // 256) * achieved
    # the step's backward as it runs on one GPU (loss.py): ONE recompute — bw3 storing its G tiles + the transposed product
    # of gt_gemm.cu — timed as the pair and the product alone
    pair = None
    if bw3 and B == N and ctx.world == 1 and os.environ.get("B200CLIP_GSTORE", "1") != "0":
        try:
            G = torch.empty(ops.gstore_elems(B, N), dtype=torch.bfloat16, device=dev)
            dY = torch.zeros(N, D, device=dev)

            def both():
                ops.logits_bwd_both(0, x, y, B, N, Kp, D, dyn, rs, cs, dX, dY, scal, G, gnorm=2.0 * N)

            def gemm():
                ops.call("gt_gemm", G, G.numel(), B, N, x, x.stride(0), Kp, D, dyn, 2.0 * N, dY, dY.stride(0), ops.stream_ptr(dev))
            tm = {}
            for nm, fn in (("pair", both), ("gemm", gemm)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record(); torch.cuda.synchronize()
                tm[nm] = e0.elapsed_time(e1) / reps
            pair = {"kernels": f"bw3_kernel<CLIP, {256 if Kp <= 512 else 128}> with TMA stores of its G tiles (bf16, N^2 elements in [64 x 64] blocks) + gt_gemm_kernel (dY = G^T X, "
                               "cta_group::2 M=256, both operands MN-major)",
                    "ms_pair": tm["pair"], "ms_gt_gemm": tm["gemm"], "ms_bw3_with_store": tm["pair"] - tm["gemm"],
                    "pair_algorithmic_tflops": 2 * alg / (tm["pair"] * 1e-3) / 1e12,
                    "pair_frac": 2 * alg / (tm["pair"] * 1e-3) / 1e12 / bf16_burst,
                    "gt_gemm_tflops": alg / (tm["gemm"] * 1e-3) / 1e12, "gt_gemm_frac": alg / (tm["gemm"] * 1e-3) / 1e12 / bf16_burst,
                    "note": "both gradients from one recompute: executed 6*B*N*D for 4*B*N*D algorithmic (two passes: 8)"}
            del G, dY
        except Exception as e:
            pair = {"error": f"{type(e).__name__}: {e}"}
    traffic = None
    prof = ROOT / "profiles" / ncu_summary if ncu_summary else None
    if prof is not None and prof.exists() and ctx.world == 1:
        traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
    roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": bf16_burst, "unit": "TFLOP/s",
            "frac": achieved / bf16_burst, "traffic": traffic, "peak_source": src, "ms_per_launch": kms,
            "executed_tflops": executed,
            "note": "algorithmic = 2*B*N*D FLOP per launch (SURVEY 8d: one gradient GEMM; the S recompute is executed but not "
                    "counted); achieved = algorithmic / CUDA-event time of the kernel alone on its stream; peak = the BURST "
                    "cuBLAS bf16 figure of MEASURED_PEAKS.json; traffic = dram read+write bytes per launch from the "
                    f"committed ncu --set full capture ({prof.name if prof is not None else 'none'})"}
    step_alg = 6.0 * B * N * D
    roof["step_algorithmic_tflops"] = step_alg / (step_ms * 1e-3) / 1e12
    roof["step_frac_of_peak"] = roof["step_algorithmic_tflops"] / bf16_burst
    roof["step_peak"] = bf16_burst
    if pair is not None:
        roof["backward_one_recompute"] = pair
    if pair is not None and "error" not in pair:
        roof["step_note"] = ("one GPU: the step executes 8*N^2*D FLOP for 6*N^2*D algorithmic (S in the forward and once in the "
                             "backward, whose G tiles feed both gradients), so 0.75 of the burst peak is the ceiling of the step "
                             "fraction; `frac` above is the dominant kernel alone WITHOUT the G stores")
    else:
        roof["step_note"] = ("the step executes 10*N^2*D FLOP for 6*N^2*D algorithmic (S is formed in the forward and once per "
                             "gradient side), so 0.6 of the burst peak is the ceiling of the step fraction")
    return roof


def clip_leg(ctx: Ctx, N: int, D: int, steps: int, warmup: int, precision: str, graph: bool, e2e: bool, clocks: bool,
             roofline: bool = True, roofline_profile=None):
    """CLIP InfoNCE fwd+bwd at global batch N: plugin (eager) path, CUDA-graph replay, e2e and parity on the same batch."""
    import torch
    import torch.distributed as dist
    from deepcoro_clip_b200 import GraphedLossStep, HostBatchPrefetcher, _lib
    from deepcoro_clip_b200.loss import CLIPLoss
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    B = N // world
    lo, hi = rank * B, (rank + 1) * B
    # the SAME global batch on every rank (device generator, fixed seed): the N-GPU result can be checked against the
    # single-process float64 restatement, and the loss is the same number at every N
    g = torch.Generator(device=dev).manual_seed(5)
    v_glob = torch.randn(N, D, device=dev, generator=g)
    t_glob = 0.3 * v_glob + torch.randn(N, D, device=dev, generator=g)
    v = v_glob[lo:hi].clone().requires_grad_(True)
    t = t_glob[lo:hi].clone().requires_grad_(True)
    log_temp = torch.tensor([math.log(TAU)], device=dev, requires_grad=True)
    loss_mod = CLIPLoss(precision=precision)
    box = {}

    def step(vv=v, tt=t):
        vv.grad = None; tt.grad = None; log_temp.grad = None
        loss = loss_mod(video_features=vv, text_features=tt, log_temp=log_temp)
        loss.backward()
        box["loss"] = loss
        return loss

    # L2 rule: the per-step working set must not stay L2-resident from one timed step to the next. At 1 / 2 / 4 GPUs it
    # is 470 / 268 / 168 MB at D = 512 (> the 126 MB L2); at 8 GPUs a rank's share is 117 MB, so there a 256 MB buffer is
    # rewritten between the timed steps and every step is bracketed by its own pair of events.
    ws_bytes = 4 * B * D * 4 + 2 * N * D * 2 + 2 * B * D * 4
    l2_flush = ws_bytes <= 1.25 * L2_BYTES
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev) if l2_flush else None

    sampler = ClockSampler(dev.index) if (clocks and rank == 0) else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        step()
    if sampler:
        sampler.wait_first_sample()
    t_begin = time.time()
    l0 = _lib.LAUNCHES
    eager_ms = timed(ctx, step, steps, 0, flush_buf)
    launches = _lib.LAUNCHES - l0
    clk = sampler.stop(t_begin, time.time()) if sampler else None
    loss_val = box["loss"].item()

    # ---- parity of the measured step (float64 restatement, same global batch): evaluated AFTER every timed region of the
    #      run (seconds of fp64 GEMMs would otherwise push the chip into its power cap right before the next timing) ----
    rows = torch.randint(lo, hi, (128,), device=dev, generator=g)
    dv_own, dt_own = v.grad[rows - lo].double(), t.grad[rows - lo].double()

    def parity_fn():
        ref_loss, dv_ref, dt_ref = clip_fp64(v_glob, t_glob, math.log(TAU), rows)
        gv = ((dv_own - dv_ref).norm() / dv_ref.norm()).item()
        gt = ((dt_own - dt_ref).norm() / dt_ref.norm()).item()
        return {"loss": loss_val, "loss_fp64": ref_loss, "loss_rel_vs_fp64": abs(loss_val - ref_loss) / abs(ref_loss),
                "grad_rel_sampled_rows": ctx.max_over_ranks(max(gv, gt)), "sampled_rows_per_rank": 128,
                "tolerance": "north_star: loss 1e-5 relative, gradients 2e-3",
                "checker": "float64 restatement of utils/loss/contrastive.py:146-164 in plain torch on the GPU, same global "
                           "batch on every rank"}

    out = {"eager_ms": eager_ms, "launches": launches, "loss": loss_val, "clocks": clk, "parity_fn": parity_fn,
           "l2_flush": l2_flush, "ws_bytes": ws_bytes, "B": B, "log_temp": log_temp}
    if roofline:
        out["roofline"] = bwd_kernel_roofline(ctx, B, N, D, log_temp, eager_ms, roofline_profile) if rank == 0 else None
        ctx.barrier()

    # ---- the same step replayed from a CUDA graph (public GraphedLossStep) ----
    gstep = None
    if graph:
        gstep = GraphedLossStep(loss_mod, v, t, log_temp, warmup=2)
        out["graphed_ms"] = timed(ctx, lambda: gstep.step(), steps, 3, flush_buf)

    # ---- e2e: the plugin path with pinned host inputs in and the loss out, every step ----
    if e2e:
        v_host = v.detach().cpu().pin_memory(); t_host = t.detach().cpu().pin_memory()
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def e2e_loop(n):
            pf = HostBatchPrefetcher(((v_host, t_host) for _ in range(n)), dev)
            for batch in pf:
                vv = batch[0].requires_grad_(True); tt = batch[1].requires_grad_(True)
                loss_host.copy_(step(vv, tt).detach().reshape(1), non_blocking=True)
                pf.release(batch)
                vv.requires_grad_(False); tt.requires_grad_(False)
                torch.cuda.current_stream().synchronize()          # the step's result is on the host before the next step
        e2e_loop(3)
        ctx.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_loop(steps)
        e1.record()
        ctx.barrier()
        out["e2e_ms"] = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
        out["h2d_bytes_per_step"] = 2 * B * D * 4 * world
    if world > 1 and gstep is not None:
        # a live CUDA graph that holds NCCL kernels keeps destroy_process_group() waiting forever: drop it now
        del gstep
        import gc
        gc.collect()
        torch.cuda.synchronize()
    return out


# =====================================================================================================================
# legs
# =====================================================================================================================
def retrieval_problem(dev):
    import torch
    Nv, M, D = 203808, 32473, 512
    g = torch.Generator(device=dev).manual_seed(3)
    v = torch.randint(-127, 128, (Nv, D), device=dev, generator=g).float() / 128
    t = torch.randint(-127, 128, (M, D), device=dev, generator=g).float() / 128
    gt = torch.randint(0, M, (Nv,), device=dev, generator=g)
    return v, t, gt


def retrieval_leg(ctx: Ctx, prob, cpu_baseline: bool):
    """recall@1/5/10 + MRR over the 203,808 x 32,473 x 512 sweep (BASELINE config 4): exact-grid embeddings (entries
    k/128, exact in bf16, every dot product exact in fp32 in any order), text database sharded by rows across ranks.
    Timed through the public API (operand packing, tensor-core ground-truth similarity, sweep, rank counts all-reduced,
    recall hits and the fp64 MRR sum on the device, results read back every sweep)."""
    import torch
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_recall_at_k_streaming, mrr_sum_from_counts
    v, t, gt = prob
    Nv, M, D = v.shape[0], t.shape[0], v.shape[1]
    box = {}

    def once():
        keep = []
        r = compute_recall_at_k_streaming(v, t, gt, k_values=[1, 5, 10], precision="bf16", use_ddp=True, _counts_out=keep)
        r["MRR_V2T"] = float(mrr_sum_from_counts(keep[0], M).item() / Nv)
        box["r"], box["counts"] = r, keep[0]
    ms = timed(ctx, once, 5, 3)
    r = box["r"]
    # parity at full size: brute-force ranks of a sample of rows (fp32 matmul of exact-grid values is exact)
    rows = torch.randint(0, Nv, (512,), device=ctx.dev, generator=torch.Generator(device=ctx.dev).manual_seed(7))
    sim = v[rows] @ t.T
    sg = sim.gather(1, gt[rows][:, None])
    cols = torch.arange(M, device=ctx.dev)[None, :]
    ref = ((sim > sg) | ((sim == sg) & (cols < gt[rows][:, None]))).sum(1)
    exact = bool((ref.int() == box["counts"][rows]).all().item())
    bf16_burst, _, _, src = peaks()
    tf = 2.0 * Nv * M * D / (ms * 1e-3) / 1e12
    out = {"metric": "streaming retrieval recall@1/5/10 + MRR", "value": Nv * M / (ms * 1e-3) / 1e9, "unit": "Gsim/s",
           "ms_per_sweep": ms, "n_video": Nv, "n_text": M, "dim": D, "text_shards": ctx.world,
           "roofline": {"bound": "tensor", "kernel": "te2_kernel<RetrEpi<0>> (retrieval.cu: CTA pairs, rank-count epilogue)",
                        "achieved": tf / ctx.world, "peak": bf16_burst, "unit": "TFLOP/s",
                        "frac": tf / bf16_burst / ctx.world, "peak_source": src,
                        "note": "algorithmic 2*N*M*D FLOP per sweep over the WHOLE API call (packing, ground-truth dots, sweep, "
                                "hits, MRR, read-back), per GPU"},
           "recall@1": r["Recall@1"], "mrr": r["MRR_V2T"],
           "parity": {"rank_counts_bit_exact_on_sampled_rows": exact, "sampled_rows": 512,
                      "checker": "brute-force fp32 similarity rows (exact for grid embeddings), lowest-index tie rule"}}
    if cpu_baseline and ctx.rank == 0 and ctx.world == 1:
        out["cpu_baseline_fn"] = lambda: cpu_retrieval_baseline(M, D)
    return out


def topk_leg(ctx: Ctx, prob):
    """Explicit top-10 lists (score desc, index asc) for every video of the C4 sweep."""
    import torch
    from deepcoro_clip_b200.retrieval_metrics_streaming import streaming_topk
    v, t, _ = prob
    Nv, M, D = v.shape[0], t.shape[0], v.shape[1]
    box = {}

    def once():
        box["s"], box["i"] = streaming_topk(v, t, 10, precision="bf16", use_ddp=True)
    ms = timed(ctx, once, 5, 2)
    st, it = torch.topk(v[:1024] @ t.T, 11, dim=1)
    tie_free = (st[:, :-1] != st[:, 1:]).all(dim=1)
    exact = bool((box["s"][:1024] == st[:, :10]).all().item() and
                 (box["i"][:1024][tie_free] == it[tie_free][:, :10]).all().item())
    bf16_burst, _, _, src = peaks()
    tf = 2.0 * Nv * M * D / (ms * 1e-3) / 1e12
    return {"metric": "streaming top-10 lists", "value": Nv * M / (ms * 1e-3) / 1e9, "unit": "Gsim/s", "ms_per_sweep": ms,
            "k": 10, "text_shards": ctx.world,
            "roofline": {"bound": "tensor", "achieved": tf / ctx.world, "peak": bf16_burst, "unit": "TFLOP/s",
                         "frac": tf / bf16_burst / ctx.world, "peak_source": src,
                         "note": "the selection epilogue, not the tensor pipe, bounds this sweep (DESIGN §5.4)"},
            "parity": {"top10_bit_exact_vs_torch_topk_first_1024_rows": exact}}


def siglip_leg(ctx: Ctx, steps: int, warmup: int, cpu_baseline: bool):
    """BASELINE config 2: SigLIP multi-positive sigmoid loss, global 8 x 1024 video rows x 8,192 texts, D = 512, dense fp32
    pos_mask (diagonal + 3 random positives per row) and severity weights, bias -10, tau 0.087. The global
    text batch is the same on every rank (all-gathered by the caller): text_replicated=True, row slab per rank."""
    import torch
    from deepcoro_clip_b200.loss import SigLIPLoss
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    Bg = T = 8192; D = 512
    B = Bg // world
    lo, hi = rank * B, (rank + 1) * B
    g = torch.Generator(device=dev).manual_seed(1)
    t_all = torch.randn(T, D, device=dev, generator=g).bfloat16().float()
    v_all = (0.5 * t_all + torch.randn(Bg, D, device=dev, generator=g)).bfloat16().float()
    pm = torch.zeros(Bg, T, device=dev); pm[torch.arange(Bg), torch.arange(Bg)] = 1.0
    for _ in range(3):
        pm[torch.arange(Bg, device=dev), torch.randint(0, T, (Bg,), device=dev, generator=g)] = 1.0
    pw = pm * torch.tensor([1.0, 1.5, 2.5, 3.0], device=dev)[torch.randint(0, 4, (Bg, T), device=dev, generator=g)]
    v = v_all[lo:hi].clone().requires_grad_(True)
    t = t_all.clone().requires_grad_(True)
    pm_l, pw_l = pm[lo:hi].contiguous(), pw[lo:hi].contiguous()
    lt = torch.tensor([math.log(0.087)], device=dev, requires_grad=True)
    mod = SigLIPLoss(precision="bf16", text_replicated=True).to(dev)
    box = {}

    def step():
        v.grad = None; t.grad = None; lt.grad = None; mod.bias.grad = None
        loss = mod(v, t, lt, pos_mask=pm_l, pos_weights=pw_l)
        loss.backward()
        box["loss"] = loss
    ws_bytes = 2 * B * T * 4 + (B + T) * D * 14
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev) if ws_bytes <= 1.25 * L2_BYTES else None
    ms = timed(ctx, step, steps, warmup, flush)
    loss_val = box["loss"].item()
    dv_own, dt_own = v.grad.double(), t.grad.double()
    def parity_fn():
        # float64 autograd restatement of utils/loss/contrastive.py:259-303 on the global problem
        v2 = v_all.double().requires_grad_(True); t2 = t_all.double().requires_grad_(True)
        vh = torch.nn.functional.normalize(v2, dim=-1); th = torch.nn.functional.normalize(t2, dim=-1)
        L = (vh @ th.T / math.exp(math.log(0.087)) + (-10.0)).clamp(-30, 30)
        y = pm.double()
        w = torch.where(y > 0.5, pw.double(), torch.ones_like(y))
        ref = (w * torch.nn.functional.binary_cross_entropy_with_logits(L, y, reduction="none")).mean()
        ref.backward()
        gv = ((dv_own - v2.grad[lo:hi]).norm() / v2.grad[lo:hi].norm()).item()
        gt = ((dt_own - t2.grad).norm() / t2.grad.norm()).item()
        return {"loss": loss_val, "loss_fp64": ref.item(), "loss_rel_vs_fp64": abs(loss_val - ref.item()) / abs(ref.item()),
                "grad_rel": ctx.max_over_ranks(max(gv, gt)),
                "tolerance": "north_star: loss 1e-5 relative, gradients 2e-3",
                "checker": "float64 autograd restatement of utils/loss/contrastive.py:259-303 in plain torch on the GPU, same "
                           "global problem on every rank"}
    bf16_burst, _, hbm, src = peaks()
    tf = 6.0 * B * T * D / (ms * 1e-3) / 1e12
    out = {"metric": "SigLIP multi-positive loss fwd+bwd (config 2)", "value": Bg / (ms * 1e-3), "unit": "samples/s",
           "ms_per_step": ms, "global_rows": Bg, "texts": T, "dim": D, "rows_per_gpu": B, "n_gpus": world,
           "l2": "explicit 256 MB flush between steps" if flush is not None else f"working set {ws_bytes / 1e6:.0f} MB > L2",
           "roofline": {"bound": "tensor", "kernel": "bw3_kernel<SIGLIP, 256> x 2 (both gradient passes; no separate forward)",
                        "achieved": tf, "peak": bf16_burst, "unit": "TFLOP/s", "frac": tf / bf16_burst, "peak_source": src,
                        "note": "algorithmic 6*B*T*D FLOP per rank and step over the WHOLE step (normalise, one streaming pass "
                                f"over the dense fp32 mask + weights = {2 * B * T * 4 / 1e6:.0f} MB, two tile passes, positives, "
                                "all-reduces); the mask pass alone is HBM-bound"},
           "parity_fn": parity_fn}
    if cpu_baseline and rank == 0 and world == 1:
        out["cpu_baseline_fn"] = lambda: cpu_siglip_baseline(Bg, T, D)
    return out


def tokens_leg(ctx: Ctx, steps: int, warmup: int, cpu_baseline: bool):
    """BASELINE config 3 (study mode): 8 studies x 4 views x 16 frames x 196 patch tokens per GPU (replicas only):
    Rope3D on q, k [32, 8, 3136, 96] bf16, AttentionPool on x [32, 3136, 512] bf16 (4 views folded into the batch), the
    aggregator's query-pool tail on [8, 4, 512] fp32 — forward + backward of each, eager module calls."""
    import torch
    from deepcoro_clip_b200 import AttentionPool, EnhancedVideoAggregator, Rope3D
    dev = ctx.dev
    _, _, hbm, src = peaks()
    torch.manual_seed(2)
    S, V, L, D, Hh, Dh = 8, 4, 3136, 512, 8, 96
    q = torch.randn(S * V, Hh, L, Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(S * V, Hh, L, Dh, device=dev, dtype=torch.bfloat16, requires_grad=True)
    gq = torch.randn_like(q); gk = torch.randn_like(k)
    rope = Rope3D(Hh * Dh, Hh).to(dev).eval()
    x = torch.randn(S * V, L, D, device=dev, dtype=torch.bfloat16, requires_grad=True)
    pool = AttentionPool(D, 8, dropout=0.0).to(dev)
    agg = EnhancedVideoAggregator(D).to(dev)
    gy = torch.randn(S * V, D, device=dev)
    res = {}

    def t_ms(fn, reps=10):
        return timed(ctx, fn, reps, 3)

    def rope_f():
        with torch.no_grad():
            rope(q, k, 16, 14, 14)
    b_rope = 4 * S * V * Hh * L * Dh * 2
    ms = t_ms(rope_f)
    res["rope_fwd"] = {"ms": ms, "algorithmic_bytes": b_rope, "GBps": b_rope / ms / 1e6, "frac_hbm": b_rope / ms / 1e6 / hbm}
    qo, ko = rope(q, k, 16, 14, 14)
    ms = t_ms(lambda: torch.autograd.grad((qo, ko), (q, k), (gq, gk), retain_graph=True))
    res["rope_bwd"] = {"ms": ms, "algorithmic_bytes": b_rope, "GBps": b_rope / ms / 1e6, "frac_hbm": b_rope / ms / 1e6 / hbm}
    del qo, ko
    bx = S * V * L * D * 2

    def pool_f():
        with torch.no_grad():
            pool(x)
    ms = t_ms(pool_f)
    res["attnpool_fwd"] = {"ms": ms, "algorithmic_bytes": bx, "GBps": bx / ms / 1e6, "frac_hbm": bx / ms / 1e6 / hbm}

    # parameter gradients are reset like optimizer.zero_grad(set_to_none=True) does every step (otherwise autograd adds into
    # the old .grad: one extra kernel per parameter that no training step pays)
    pool_params, agg_params = list(pool.parameters()), list(agg.parameters())

    def reset(params):
        for prm in params:
            prm.grad = None

    def pool_fb():
        x.grad = None
        reset(pool_params)
        pool(x).backward(gy)
    ms = t_ms(pool_fb)
    res["attnpool_fwd_bwd"] = {"ms": ms, "algorithmic_bytes": 4 * bx, "GBps": 4 * bx / ms / 1e6,
                               "frac_hbm": 4 * bx / ms / 1e6 / hbm,
                               "note": "algorithmic: x read once forward, read + dx written backward (SURVEY 8d: 1 + 3 passes)"}

    # the two streaming kernels alone (C ABI, back-to-back launches on three rotating copies of x so that nothing stays in L2)
    try:
        from deepcoro_clip_b200._lib import call, i64, lib, stream_ptr
        xs = [x.detach()] + [torch.randn_like(x) for _ in range(2)]
        Sp = lib().b200clip_attnpool_tc_splits(xs[0].data_ptr(), 1, i64(L * D), i64(D), S * V, L, D, 8)
        if Sp > 0:
            qt = torch.randn(8, D, device=dev) * 0.05
            pm = torch.empty((S * V, Sp, 8), device=dev); pl = torch.empty_like(pm); pa = torch.empty((S * V, Sp, 8, D), device=dev)
            xb_ = torch.randn(S * V, 8, D, device=dev); mm = torch.zeros((S * V, 8), device=dev); ll = torch.ones((S * V, 8), device=dev)
            dxb = torch.randn(S * V, 8, D, device=dev); pdq = torch.empty((S * V, Sp, 8, D), device=dev)
            dxs = [torch.empty_like(xs[0]) for _ in range(3)]
            st = stream_ptr(dev)
            it = {"i": 0}
            # operand images exactly as the module prepares them: pool_prep (query side) and pool_tail_bwd (per-row gradient side)
            R = S * V
            prm = {k_: torch.randn(n_, device=dev) * 0.05 for k_, n_ in (("query", D), ("w_in", 3 * D * D), ("b_in", 3 * D),
                                                                       ("w_o", D * D), ("gamma", D))}
            q0 = torch.empty(D, device=dev); qimg = torch.empty(D * 8, device=dev)
            call("pool_prep", prm["query"], prm["w_in"], prm["b_in"], D, 8, q0, qt, qimg, 0, st)
            yh = torch.randn(R, D, device=dev); rs = torch.ones(R, device=dev); sa1 = torch.ones(R, 8, device=dev)
            dy_ = torch.randn(R, D, device=dev).bfloat16()
            tmp = [torch.empty(R * D, device=dev) for _ in range(3)]
            cdot = torch.empty(R, 8, device=dev); wimg = torch.empty(R * D * 16, device=dev)
            call("pool_tail_bwd", dy_, 1, yh, rs, xb_, sa1, prm["w_in"].data_ptr() + 8 * D * D, prm["b_in"].data_ptr() + 8 * D,
                 prm["w_o"], prm["gamma"], None, 0, qt, R, 8, D, tmp[0], tmp[1], tmp[2], dxb, None, cdot, wimg, 0, st)

            def k_fwd():
                it["i"] += 1
                call("attnpool_tc_fwd", xs[it["i"] % 3], 1, None, i64(0), None, qimg, R, L, D, 8, Sp, pm, pl, pa, 0.0, 0, None, st)

            def k_bwd():
                it["i"] += 1
                call("attnpool_tc_bwd", xs[it["i"] % 3], 1, None, i64(0), None, None, None, wimg, cdot, mm, ll, R, L, D, 8, Sp,
                     dxs[it["i"] % 3], None, None, 0.0, 0, None, pdq, st)
            ms = t_ms(k_fwd, 20)
            res["pool_fwd_tc_kernel"] = {"ms": ms, "algorithmic_bytes": bx, "GBps": bx / ms / 1e6, "frac_hbm": bx / ms / 1e6 / hbm,
                                         "note": "csrc/attnpool_tc.cu as the module launches it (operand image from pool_prep)"}
            ms = t_ms(k_bwd, 20)
            res["pool_bwd_tc_kernel"] = {"ms": ms, "algorithmic_bytes": 2 * bx, "GBps": 2 * bx / ms / 1e6,
                                         "frac_hbm": 2 * bx / ms / 1e6 / hbm,
                                         "note": "as the module launches it (operand images + c_h from pool_tail_bwd): reads x once, "
                                                 "writes dx and the query-gradient partials in the same pass"}
            del xs, dxs, pa, pdq
    except Exception as e:      # the module-level numbers above stand on their own
        res["pool_tc_kernels"] = {"error": f"{type(e).__name__}: {e}"}

    xa = torch.randn(S, V, D, device=dev, requires_grad=True)
    ga = torch.randn(S, D, device=dev)

    def whole():
        q.grad = None; k.grad = None; x.grad = None; xa.grad = None
        reset(pool_params); reset(agg_params)
        qo, ko = rope(q, k, 16, 14, 14)
        torch.autograd.backward((qo, ko), (gq, gk))
        pool(x).backward(gy)
        agg(xa).backward(ga)
    def agg_fb():
        xa.grad = None
        reset(agg_params)
        agg(xa).backward(ga)
    ms = t_ms(agg_fb)
    res["aggregator_fwd_bwd"] = {"ms": ms, "note": "EnhancedVideoAggregator (depth 2, train mode): one library call per direction "
                                                   "(positional add, 2 x (1 + 5) block launches, query-pool tail); "
                                                   "latency-bound, [8, 4, 512] fp32"}
    # gated-attention MIL pooling of the probing head on patch tokens, two levels (not part of `whole`)
    try:
        from deepcoro_clip_b200 import GatedAttentionPooling
        Lm, Hd = 1568, 128
        mil = GatedAttentionPooling(D, Hd).to(dev)
        xm = torch.randn(S, V, Lm, D, device=dev, requires_grad=True)
        gm = torch.randn(S, D, device=dev)

        mil_params = list(mil.parameters())

        def mil_fb():
            xm.grad = None
            reset(mil_params)
            mil(xm).backward(gm)
        ms = t_ms(mil_fb)
        fl = 3 * 2.0 * S * V * Lm * D * 2 * Hd
        res["mil_gated_pool_fwd_bwd"] = {
            "ms": ms, "algorithmic_flops": fl, "TFLOPps": fl / ms / 1e9, "frac_fp32_fma": fl / ms / 1e9 / (148 * 128 * 2 * 1.965e-3),
            "note": f"[{S}, {V}, {Lm}, {D}] fp32, hidden {Hd}, two levels: three [R x 512] x [512 x 256] products on tcgen05 with "
                    "split-precision operands (bf16 hi + lo, 3x the bf16 FLOP) + the streaming passes around them (operand split, "
                    "gate epilogue, pooling, dpre); frac_fp32_fma = algorithmic rate over the CUDA-core FMA peak 148 x 128 x "
                    "1.965 GHz, the ceiling of an fp32 implementation"}
        del xm, mil
    except Exception as e:
        res["mil_gated_pool_fwd_bwd"] = {"error": f"{type(e).__name__}: {e}"}
    ms = t_ms(whole, steps)
    total_bytes = 2 * b_rope + 4 * bx
    out = {"metric": "study-mode token path fwd+bwd (config 3)", "value": S * ctx.world / (ms * 1e-3), "unit": "studies/s",
           "ms_per_step": ms, "execution": "eager module calls", "studies_per_gpu": S, "views": V, "tokens_per_view": L, "n_gpus": ctx.world,
           "parallelism": "replicas only (no collective on this path)", "kernels": res,
           "roofline": {"bound": "hbm", "achieved": total_bytes / ms / 1e6, "peak": hbm, "unit": "GB/s",
                        "frac": total_bytes / ms / 1e6 / hbm, "peak_source": src,
                        "note": "algorithmic bytes of the whole step (RoPE 2 x 4*B*H*N*Dh*2, pool 4 x B*N*D*2) over its time; "
                                "per-kernel fractions under `kernels`"}}
    if cpu_baseline and ctx.rank == 0 and ctx.world == 1:
        out["cpu_baseline_fn"] = cpu_tokens_baseline
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="clip32k", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "auto"])
    ap.add_argument("--legs", default="all", help="comma list of extra legs (retrieval,topk10,siglip_c2,tokens_c3,clip32k_d768), "
                                                  "'all' or 'none'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay of the step")
    args = ap.parse_args()
    N, D = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, N, D)
        return
    legs = ALL_LEGS if args.legs == "all" else ([] if args.legs == "none" else [x for x in args.legs.split(",") if x])
    if args.no_retrieval:
        legs = [x for x in legs if x not in ("retrieval", "topk10")]
    if args.workload != "clip32k":
        legs = [x for x in legs if x != args.workload]

    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter) is sent to stderr for the duration of the run
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert N % world == 0
    ctx = Ctx(dev, world, rank)
    W = max(3, args.warmup)
    want_cpu = not args.no_cpu_baseline

    head = clip_leg(ctx, N, D, args.steps, W, args.precision, graph=not args.no_graph, e2e=True, clocks=True,
                    roofline_profile="r02_bw3_kernel_ncu_full_summary.json" if args.workload == "clip32k" else None)
    B = head["B"]
    roof = head.get("roofline")
    extras = {}
    prob = None
    for leg in legs:
        try:
            if leg in ("retrieval", "topk10"):
                if prob is None:
                    prob = retrieval_problem(dev)
                extras[leg] = retrieval_leg(ctx, prob, want_cpu) if leg == "retrieval" else topk_leg(ctx, prob)
            elif leg == "siglip_c2":
                prob = None
                extras[leg] = siglip_leg(ctx, 10, 3, want_cpu)
            elif leg == "tokens_c3":
                prob = None
                extras[leg] = tokens_leg(ctx, 10, 3, want_cpu)
            elif leg == "clip32k_d768":
                prob = None
                torch.cuda.empty_cache()
                d = clip_leg(ctx, 32768, 768, 5, 3, "bf16", graph=False, e2e=False, clocks=False)
                extras[leg] = {"metric": "contrastive loss fwd+bwd samples/s @B=32k,D=768 (config 5)",
                               "value": 32768 / (d["eager_ms"] * 1e-3), "unit": UNIT, "ms_per_step": d["eager_ms"],
                               "n_gpus": world, "parity_fn": d["parity_fn"], "roofline": d.get("roofline")}
                if want_cpu and rank == 0 and world == 1:
                    extras[leg]["cpu_baseline_fn"] = lambda: cpu_clip_baseline(32768, 768)
        except Exception as e:      # a failing extra leg must not take the headline line down with it
            if world > 1:
                raise
            extras[leg] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    # ---- parity phase: every fp64 check runs after the last timed region ----
    head["parity"] = head.pop("parity_fn")()
    for leg in extras.values():
        if "parity_fn" in leg:
            leg["parity"] = leg.pop("parity_fn")()
    # ---- CPU baselines last: seconds of host work between GPU legs would let the GPU clocks fall before the next timing ----
    for leg in extras.values():
        if "cpu_baseline_fn" in leg:
            leg["cpu_baseline"] = leg.pop("cpu_baseline_fn")()

    cpu = cpu_clip_baseline(N, D) if (rank == 0 and world == 1 and want_cpu) else None

    if rank == 0:
        eager_ms = head["eager_ms"]
        line = {
            "metric": METRIC, "value": N / (eager_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": eager_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "global_batch": N, "per_gpu_rows": B, "dim": D, "tau": TAU,
                       "precision": args.precision, "parallelism": f"row-slab x{world}",
                       "execution": "plugin path: eager CLIPLoss.forward + loss.backward() (what LossRegistry / Loss.run calls)",
                       "inputs": "the same global batch on every rank (seeded device generator), row-sharded",
                       "l2": (f"explicit flush: a 256 MB buffer is rewritten between the timed steps, each step timed by its "
                              f"own event pair (per-rank working set {head['ws_bytes'] / 1e6:.0f} MB would fit the 126 MB L2); "
                              "e2e inputs arrive from pinned host memory every step") if head["l2_flush"] else
                             ("no explicit flush: per-step working set (fp32 inputs+grads, bf16 operands, fp32 dXhat) "
                              f"= {head['ws_bytes'] / 1e6:.0f} MB > 126 MB L2")},
            "e2e": {"value": N / (head["e2e_ms"] * 1e-3), "unit": UNIT, "ms_per_step": head["e2e_ms"],
                    "h2d_bytes_per_step": head["h2d_bytes_per_step"], "d2h_bytes_per_step": 4 * world,
                    "note": "plugin path; inputs from pinned host memory (double-buffered H2D), the 4-byte loss read back and "
                            "waited for every step; gradients stay on the device (they feed the encoders' backward)"},
            "graphed": ({"value": N / (head["graphed_ms"] * 1e-3), "unit": UNIT, "ms_per_step": head["graphed_ms"],
                         "note": "GraphedLossStep: forward + backward + collectives replayed from one CUDA graph"}
                        if "graphed_ms" in head else None),
            "gpu_launches": head["launches"], "loss": head["loss"], "parity": head["parity"], "clocks": head["clocks"],
            "roofline": roof, "cpu_baseline": cpu,
        }
        line.update(extras)
        real_out.write(json.dumps(line) + "\n")
        real_out.flush()
    if world > 1:
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
