#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native contrastive head.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (BASELINE.json):
  clip32k  (default)  CLIP InfoNCE loss fwd+bwd, global batch N=32,768, D=512 (the configuration the metric
                      "contrastive loss fwd+bwd samples/s @B=32k,D=512" is quoted on), bf16 operands / fp32 accumulate.
                      N GPUs: STRONG scaling — the global batch is fixed and row-sharded, each rank owns a slab of
                      the logits; embeddings all-gathered, column sums all-reduced, row sums all-gathered (NCCL).
  clip32k_d768        same at D=768 (BASELINE config 5)
A "step" = one loss forward + backward over the whole global batch (normalise, logits tiles, LSE, gradients).
One JSON line on rank 0. `value` = device-resident inputs; `e2e` = same step through the public module API with
pinned-host inputs copied in and the loss copied out every step.
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
TAU = 0.0588   # config/clip/base_config.yaml:46
L2_BYTES = 126e6
CPU_PORT = "reference op sequence on torch CPU (fp32, all host threads; oracle/reference_torch_port.py)"


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
        self.lines: list[str] = []

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
        """nvidia-smi needs a few hundred ms before its first line: the sampler is started ahead of the warm-up steps and
        the timed region begins only once it is producing."""
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


def cpu_reference_step(N: int, D: int, steps: int, warmup: int):
    """The reference's CPU path on a bounded sample of the workload: the reference is PyTorch, so this is its own op
    sequence (normalize, matmul, 2x cross_entropy, autograd backward) on ATen's CPU kernels with every host thread
    (oracle/reference_torch_port.py, pinned to the imported reference's golden vectors). Returns (best s, mean s, threads)."""
    import torch
    from oracle import reference_torch_port as tp
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(5)
    v = torch.randn(N, D, generator=g)
    t = torch.randn(N, D, generator=g)
    for _ in range(warmup):
        tp.clip_loss_step(v, t, math.log(TAU))
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tp.clip_loss_step(v, t, math.log(TAU))
        ts.append(time.perf_counter() - t0)
    return min(ts), sum(ts) / len(ts), torch.get_num_threads()


def cpu_reference_retrieval(M: int, D: int, rows: int = 8192):
    """The reference's CPU path for recall@1/5/10 + MRR (oracle/reference_torch_port.retrieval_metrics_step: chunked matmul /
    topk merges, then matmul + argsort + per-row Python loop) on a bounded slice of the C4 sweep: `rows` videos against the
    full text database. The cost is linear in the video rows, so the slice's Gsim/s is the sweep's."""
    import torch
    from oracle import reference_torch_port as tp
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(3)
    v = torch.randint(-127, 128, (rows, D), generator=g).float() / 128
    t = torch.randint(-127, 128, (M, D), generator=g).float() / 128
    gt = torch.randint(0, M, (rows,), generator=g)
    tp.retrieval_metrics_step(v[:2048], t, gt[:2048])
    t0 = time.perf_counter()
    tp.retrieval_metrics_step(v, t, gt)
    dt = time.perf_counter() - t0
    return {"value": rows * M / dt / 1e9, "unit": "Gsim/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"reference op sequence on torch CPU (oracle/reference_torch_port.py): {rows} videos x {M} texts x {D} "
                      f"in {dt:.2f} s (cost linear in the video rows)"}


def retrieval_leg(dev, world, rank, cpu_baseline=False):
    """recall@1/5/10 + MRR over the 203,808 x 32,473 x 512 sweep (BASELINE config 4): exact-grid embeddings (entries
    k/128, exact in bf16, every dot product exact in fp32 in any order), text database sharded by rows across ranks.
    Timed through the public API (operand packing, tensor-core ground-truth similarity, sweep, rank counts all-reduced,
    recall hits and the fp64 MRR sum on the device, results read back every sweep)."""
    import torch
    import torch.distributed as dist
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_recall_at_k_streaming, mrr_sum_from_counts
    Nv, M, D = 203808, 32473, 512
    g = torch.Generator(device=dev).manual_seed(3)
    v = torch.randint(-127, 128, (Nv, D), device=dev, generator=g).float() / 128
    t = torch.randint(-127, 128, (M, D), device=dev, generator=g).float() / 128
    gt = torch.randint(0, M, (Nv,), device=dev, generator=g)

    def once():
        keep = []
        r = compute_recall_at_k_streaming(v, t, gt, k_values=[1, 5, 10], precision="bf16", use_ddp=True, _counts_out=keep)
        r["MRR_V2T"] = float(mrr_sum_from_counts(keep[0], M).item() / Nv)
        return r
    once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        r = once()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    bf16_burst, _, _, src = peaks()
    tf = 2.0 * Nv * M * D / (ms * 1e-3) / 1e12
    out = {"metric": "streaming retrieval recall@1/5/10 + MRR", "value": Nv * M / (ms * 1e-3) / 1e9, "unit": "Gsim/s",
           "ms_per_sweep": ms, "n_video": Nv, "n_text": M, "dim": D, "text_shards": world, "achieved_tflops": tf,
           "frac_of_bf16_peak": tf / bf16_burst / world, "peak_source": src, "recall@1": r["Recall@1"], "mrr": r["MRR_V2T"]}
    if cpu_baseline and rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_reference_retrieval(M, D)
    return out


def run_reference(args, N, D):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Ns = 4096                                  # bounded sample; the loss is O(N^2 D): extrapolate by (Ns/N)
    best, mean, cores = cpu_reference_step(Ns, D, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)))
    sps_sample = Ns / best
    value = sps_sample * (Ns / N)               # samples/s the CPU path would reach at the full N (N^2 scaling)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best * 1e3 * (N / Ns) ** 2,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "global_batch": N, "dim": D, "tau": TAU},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{CPU_PORT} fwd+bwd at N={Ns}: {sps_sample:.0f} samples/s, "
                                   f"N^2-extrapolated to N={N}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="clip32k", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager module calls instead of the captured CUDA graph")
    args = ap.parse_args()
    N, D = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, N, D)
        return

    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter) is sent to stderr for the duration of the run
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from deepcoro_clip_b200 import _lib, ops
    from deepcoro_clip_b200.loss import CLIPLoss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert N % world == 0
    B = N // world
    W = max(3, args.warmup)

    g = torch.Generator().manual_seed(5 + rank)
    v_host = torch.randn(B, D, generator=g).pin_memory()
    t_host = torch.randn(B, D, generator=g).pin_memory()
    v = v_host.to(dev).requires_grad_(True)
    t = t_host.to(dev).requires_grad_(True)
    log_temp = torch.tensor([math.log(TAU)], device=dev, requires_grad=True)
    loss_mod = CLIPLoss(precision=args.precision)

    def step(vv, tt):
        vv.grad = None; tt.grad = None; log_temp.grad = None
        loss = loss_mod(video_features=vv, text_features=tt, log_temp=log_temp)
        loss.backward()
        return loss

    # Public API for launch-bound steps: the whole forward + backward (kernels and NCCL collectives) captured once in a
    # CUDA graph and replayed (deepcoro_clip_b200.GraphedLossStep). At 8 ranks the eager step is host-bound (~0.8 ms of
    # Python / launch / collective-enqueue work against ~0.55 ms of kernels).
    from deepcoro_clip_b200 import GraphedLossStep
    gstep = None
    launches_per_step = None
    if not args.no_graph:
        step(v, t)                                   # library attribute calls / communicators before the capture
        l0 = _lib.LAUNCHES
        gstep = GraphedLossStep(loss_mod, v, t, log_temp, warmup=2)
        launches_per_step = (_lib.LAUNCHES - l0) // 3      # 2 warm-up passes + the captured one

    def run_step(vv=None, tt=None):
        if gstep is not None:
            return gstep.step(vv, tt)[0]
        return step(v if vv is None else vv, t if tt is None else tt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        run_step()
    if rank == 0:
        sampler.wait_first_sample()
    # L2 rule: the per-step working set must not stay L2-resident from one timed step to the next. At 1 / 2 / 4 GPUs it
    # is 470 / 268 / 168 MB (> the 126 MB L2); at 8 GPUs a rank's share is 117 MB, so there a 256 MB buffer is rewritten
    # between the timed steps and every step is bracketed by its own pair of events (the flush is outside the brackets).
    ws_bytes = 4 * B * D * 4 + 2 * N * D * 2 + 2 * B * D * 4
    l2_flush = ws_bytes <= 1.25 * L2_BYTES
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev) if l2_flush else None
    barrier()
    t_begin = time.time()
    l0 = _lib.LAUNCHES
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    if l2_flush:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush_buf.fill_(float(i))
            evs[i][0].record()
            loss = run_step()
            evs[i][1].record()
        barrier()
        timed_ms = sum(a.elapsed_time(b) for a, b in evs)
    else:
        e0.record()
        for _ in range(args.steps):
            loss = run_step()
        e1.record()
        barrier()
        timed_ms = e0.elapsed_time(e1)
    launches = _lib.LAUNCHES - l0 if gstep is None else launches_per_step * args.steps
    ms = torch.tensor([timed_ms], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    loss_val = loss.item()

    # ---------------- e2e: pinned host inputs in, loss out, every step ----------------
    # Public API: HostBatchPrefetcher double-buffers the H2D copies of step k+1 behind step k's kernels; every step's
    # copies (2 x B x D fp32 from pinned memory) and its 4-byte loss read-back are inside the timed region.
    from deepcoro_clip_b200 import HostBatchPrefetcher
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def e2e_loop(n):
        pf = HostBatchPrefetcher(((v_host, t_host) for _ in range(n)), dev)
        for batch in pf:
            if gstep is not None:
                # device copy of the prefetched batch into the graph's static inputs, then one replay
                loss_host.copy_(gstep.step(batch[0], batch[1])[0].detach().reshape(1), non_blocking=True)
                pf.release(batch)
            else:
                vv = batch[0].requires_grad_(True); tt = batch[1].requires_grad_(True)
                loss_host.copy_(step(vv, tt).detach().reshape(1), non_blocking=True)
                pf.release(batch)
                vv.requires_grad_(False); tt.requires_grad_(False)
            torch.cuda.current_stream().synchronize()          # the step's result is on the host before the next step
    e2e_loop(3)
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = ms2.item() / args.steps

    # ---------------- roofline of the dominant kernel (logits_bwd), timed alone on its stream ----------------
    roof = None
    if rank == 0:
        bf16_burst, bf16_sust, hbm, src = peaks()
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
        e0.record()
        for _ in range(reps):
            ops.logits_bwd(0, x, y, B, N, Kp, Kp, D, dyn, rs, cs, dX, scal, gnorm=2.0 * N)
        e1.record(); torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / reps
        alg = 2.0 * B * N * D                      # algorithmic FLOPs of one launch (one gradient GEMM; recompute not counted)
        achieved = alg / (kms * 1e-3) / 1e12
        bw3 = Kp == D and Kp % 256 == 0 and Kp <= 768 and os.environ.get("B200CLIP_BWD3", "1") != "0"
        pair = Kp <= 512 and Kp % 128 == 0 and os.environ.get("B200CLIP_BWD_PAIR", "1") != "0"
        if bw3:
            kname = f"bw3_kernel<CLIP, {256 if Kp <= 512 else 128}> (logits_bwd3.cu: 64-row CTA pairs, cta_group::2 M=128, whole output width in TMEM)"
            executed = 2.0 * achieved                    # S once + the output product
            prof = ROOT / "profiles" / "r01f_bw3_kernel_ncu_full_summary.json"
        elif pair:
            kname = "bw2_kernel<CLIP> (logits_bwd2.cu: 128-row CTA pairs, cta_group::2)"
            executed = (1 + (Kp + 255) // 256) * achieved
            prof = ROOT / "profiles" / "r01c_bw2_kernel_ncu_full_summary.json"
        else:
            kname = "bw_kernel<CLIP> (logits_bwd.cu: single CTA)"
            executed = (1 + (Kp + 255) // 256) * achieved
            prof = None
        traffic = None
        if prof is not None and world == 1 and args.workload == "clip32k" and prof.exists():
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": bf16_burst,
                "unit": "TFLOP/s", "frac": achieved / bf16_burst, "traffic": traffic, "peak_source": src,
                "ms_per_launch": kms, "executed_tflops": executed,
                "note": "algorithmic = 2*B*N*D FLOP per launch (SURVEY 8d: one gradient GEMM; the S recompute is executed "
                        "but not counted); achieved = algorithmic / CUDA-event time of the kernel alone on its stream; "
                        "traffic = dram read+write bytes per launch from the committed ncu --set full capture "
                        f"({prof.name if prof is not None else 'none'})"}
        step_alg = 6.0 * B * N * D
        roof["step_algorithmic_tflops"] = step_alg / (ms_per_step * 1e-3) / 1e12
        roof["step_frac_of_peak"] = roof["step_algorithmic_tflops"] / bf16_sust
        roof["step_peak"] = bf16_sust

    # ---------------- second half of BASELINE's metric: streaming retrieval Gsim/s (C4 sweep) ----------------
    retr = None
    if not args.no_retrieval:
        retr = retrieval_leg(dev, world, rank, cpu_baseline=not args.no_cpu_baseline)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Ns = 4096
        best, _, cores = cpu_reference_step(Ns, D, 3, 1)
        cpu = {"value": (Ns / best) * (Ns / N), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{CPU_PORT} fwd+bwd at N={Ns}: {Ns / best:.0f} samples/s ({best * 1e3:.0f} ms), "
                         f"N^2-extrapolated to N={N}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": N / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "global_batch": N, "per_gpu_rows": B, "dim": D, "tau": TAU,
                       "precision": args.precision, "parallelism": f"row-slab x{world}",
                       "execution": "cuda_graph (GraphedLossStep: forward+backward+collectives replayed)" if gstep is not None
                       else "eager module calls",
                       "l2": (f"explicit flush: a 256 MB buffer is rewritten between the timed steps, each step timed by its "
                              f"own event pair (per-rank working set {ws_bytes / 1e6:.0f} MB would fit the 126 MB L2); e2e "
                              "inputs arrive from pinned host memory every step") if l2_flush else
                             ("no explicit flush: per-step working set (fp32 inputs+grads, bf16 operands, fp32 dXhat) "
                              f"= {ws_bytes / 1e6:.0f} MB > 126 MB L2")},
            "e2e": {"value": N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * B * D * 4 * world, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": launches, "loss": loss_val, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "retrieval": retr,
        }
        real_out.write(json.dumps(line) + "\n")
        real_out.flush()
    if world > 1:
        # a live CUDA graph that holds NCCL kernels keeps destroy_process_group() waiting forever: drop it first
        gstep = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
