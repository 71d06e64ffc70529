"""Multi-rank GPU parity of the sharded paths (run under torchrun, one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/gpu_check_dist.py
Every rank holds a row slab of the global batch; results are checked against the single-process numpy oracle on the
full batch (reference DDP semantics, SURVEY 8c: every rank returns the FULL loss, local grad rows = rows of the
full-batch gradient, log_temp.grad identical everywhere). Exits non-zero on any mismatch."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import contrastive_oracle as co      # checker only
from oracle import retrieval_oracle as ro


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from deepcoro_clip_b200.loss import CLIPLoss, SigLIPLoss
    from deepcoro_clip_b200.retrieval_metrics_streaming import (compute_metrics_streaming, compute_recall_at_k_streaming,
                                                                streaming_topk)
    ok = True

    def check(name, cond, info):
        nonlocal ok
        if not cond:
            ok = False
        if rank == 0 or not cond:
            print(f"[rank {rank}] {name}: {'ok' if cond else 'MISMATCH'} {info}", flush=True)

    # ---------------- CLIP, row slabs ----------------
    for (B, D, tau, prec, ltol, gtol) in [(192, 128, 0.07, "bf16x3", 1e-5, 2e-3), (1024, 512, 0.0588, "bf16", 1e-5, 2e-3),
                                          (160, 96, 0.004, "bf16x3", 1e-5, 2e-3)]:       # last: stable softmax mode
        N = B * world
        rng = np.random.default_rng(11)
        v = rng.standard_normal((N, D)).astype(np.float32)
        t = ((0.1 if tau < 0.01 else 0.5) * v + rng.standard_normal((N, D))).astype(np.float32)
        lo, hi = rank * B, (rank + 1) * B
        vt = torch.tensor(v[lo:hi], device=dev, requires_grad=True)
        tt = torch.tensor(t[lo:hi], device=dev, requires_grad=True)
        lt = torch.tensor([math.log(tau)], device=dev, requires_grad=True)
        loss = CLIPLoss(precision=prec)(video_features=vt, text_features=tt, log_temp=lt)
        loss.backward()
        o = co.clip_loss(v, t, math.log(tau))
        check(f"clip N={N} D={D} {prec} tau={tau} loss", abs(loss.item() - o["loss"]) <= ltol * abs(o["loss"]) + 2.0 ** -23 / tau,
              (loss.item(), o["loss"]))
        check("  dvideo rows", rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]) <= gtol, rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]))
        check("  dtext rows", rel(tt.grad.cpu().numpy(), o["dtext"][lo:hi]) <= gtol, rel(tt.grad.cpu().numpy(), o["dtext"][lo:hi]))
        check("  dlog_temp", abs(lt.grad.item() - o["dlog_temp"]) <= gtol * max(abs(o["dlog_temp"]), 1e-3), (lt.grad.item(), o["dlog_temp"]))

    # ---------------- SigLIP, video rows sharded, text replicated ----------------
    B, T, D, tau, bias = 160, 448, 128, 0.087, -6.0
    Bg = B * world
    rng = np.random.default_rng(12)
    t = rng.standard_normal((T, D)).astype(np.float32)
    v = (0.7 * t[rng.integers(0, T, size=Bg)] + rng.standard_normal((Bg, D))).astype(np.float32)
    pm = np.zeros((Bg, T), np.float32)
    for _ in range(3):
        pm[np.arange(Bg), rng.integers(0, T, size=Bg)] = 1.0
    pw = pm * rng.choice([1.0, 1.5, 2.5, 3.0], size=(Bg, T)).astype(np.float32)
    lo, hi = rank * B, (rank + 1) * B
    vt = torch.tensor(v[lo:hi], device=dev, requires_grad=True)
    tt = torch.tensor(t, device=dev, requires_grad=True)
    lt = torch.tensor([math.log(tau)], device=dev, requires_grad=True)
    o = co.siglip_loss(v, t, math.log(tau), bias=bias, pos_mask=pm, pos_weights=pw)
    # sharded fast path (text_replicated=True) and the reference-faithful gathered default: same text on every rank here
    for replicated in (True, False):
        vt = torch.tensor(v[lo:hi], device=dev, requires_grad=True)
        tt = torch.tensor(t, device=dev, requires_grad=True)
        lt = torch.tensor([math.log(tau)], device=dev, requires_grad=True)
        mod = SigLIPLoss(bias_init=bias, precision="bf16x3", text_replicated=replicated).to(dev)
        loss = mod(vt, tt, lt, pos_mask=torch.tensor(pm[lo:hi], device=dev), pos_weights=torch.tensor(pw[lo:hi], device=dev))
        loss.backward()
        check(f"siglip loss (text_replicated={replicated})", abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"]), (loss.item(), o["loss"]))
        check("  dvideo rows", rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]) <= 2e-3, rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]))
        check("  dtext (full, every rank)", rel(tt.grad.cpu().numpy(), o["dtext"]) <= 2e-3, rel(tt.grad.cpu().numpy(), o["dtext"]))
        check("  dlog_temp", abs(lt.grad.item() - o["dlog_temp"]) <= 2e-3 * max(abs(o["dlog_temp"]), 1e-4), (lt.grad.item(), o["dlog_temp"]))
        check("  dbias", abs(mod.bias.grad.item() - o["dbias"]) <= 2e-3 * max(abs(o["dbias"]), 1e-4), (mod.bias.grad.item(), o["dbias"]))
    # per-rank texts (what the reference's runner passes): rank r holds T/world texts and its rows' matching mask columns
    Th = T // world
    cols = slice(rank * Th, (rank + 1) * Th)
    gm = np.concatenate([pm[q * B:(q + 1) * B, q * Th:(q + 1) * Th] for q in range(world)], axis=0)
    gw = np.concatenate([pw[q * B:(q + 1) * B, q * Th:(q + 1) * Th] for q in range(world)], axis=0)
    o = co.siglip_loss(v, t[cols], math.log(tau), bias=bias, pos_mask=gm, pos_weights=gw)
    vt = torch.tensor(v[lo:hi], device=dev, requires_grad=True)
    tt = torch.tensor(t[cols], device=dev, requires_grad=True)
    lt = torch.tensor([math.log(tau)], device=dev, requires_grad=True)
    mod = SigLIPLoss(bias_init=bias, precision="bf16x3").to(dev)
    loss = mod(vt, tt, lt, pos_mask=torch.tensor(pm[lo:hi, cols], device=dev), pos_weights=torch.tensor(pw[lo:hi, cols], device=dev))
    loss.backward()
    check("siglip per-rank texts loss", abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"]), (loss.item(), o["loss"]))
    check("  dvideo rows", rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]) <= 2e-3, rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]))
    check("  dtext (own texts)", rel(tt.grad.cpu().numpy(), o["dtext"]) <= 2e-3, rel(tt.grad.cpu().numpy(), o["dtext"]))
    with torch.no_grad():
        bad = SigLIPLoss(bias_init=bias, text_replicated=True).to(dev)(vt.detach(), tt.detach(), lt.detach())
    check("  text_replicated=True with different texts is NaN", bool(torch.isnan(bad)), bad.item())

    # ---------------- SigLIP + entropy regulariser: row statistics local, mean entropy over the GLOBAL rows ----------------
    ent = SigLIPLoss(bias_init=-2.0, precision="bf16x3", entropy_regularization=True, entropy_weight=0.3,
                     min_entropy_threshold=7.0, text_replicated=True).to(dev)
    vt = torch.tensor(v[lo:hi], device=dev, requires_grad=True)
    tt = torch.tensor(t, device=dev, requires_grad=True)
    lt = torch.tensor([math.log(0.05)], device=dev, requires_grad=True)
    loss = ent(vt, tt, lt, pos_mask=torch.tensor(pm[lo:hi], device=dev))
    loss.backward()
    o = co.siglip_loss(v, t, math.log(0.05), bias=-2.0, pos_mask=pm, entropy_regularization_on=True, entropy_weight=0.3,
                       min_entropy_threshold=7.0)
    check("siglip+entropy loss", abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"]) and
          o["entropy_diagnostics"]["entropy_deficit"] > 0.05, (loss.item(), o["loss"]))
    check("  dvideo rows", rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]) <= 2e-3, rel(vt.grad.cpu().numpy(), o["dvideo"][lo:hi]))
    check("  dtext (full)", rel(tt.grad.cpu().numpy(), o["dtext"]) <= 2e-3, rel(tt.grad.cpu().numpy(), o["dtext"]))
    check("  dlog_temp", abs(lt.grad.item() - o["dlog_temp"]) <= 2e-3 * max(abs(o["dlog_temp"]), 1e-4), (lt.grad.item(), o["dlog_temp"]))
    dg = ent.get_entropy_diagnostics()
    check("  entropy_mean (global rows)", abs(dg["entropy_mean"] - o["entropy_diagnostics"]["entropy_mean"]) <= 2e-4,
          (dg["entropy_mean"], o["entropy_diagnostics"]["entropy_mean"]))

    # ---------------- SigLIP2BCELossDDP: video AND text gathered, identity labels on the global batch ----------------
    from deepcoro_clip_b200.loss import SigLIP2BCELossDDP
    B2, D2 = 96, 128
    N2 = B2 * world
    rng = np.random.default_rng(13)
    v2 = rng.standard_normal((N2, D2)).astype(np.float32)
    t2 = (0.6 * v2 + rng.standard_normal((N2, D2))).astype(np.float32)
    lo2, hi2 = rank * B2, (rank + 1) * B2
    vt = torch.tensor(v2[lo2:hi2], device=dev, requires_grad=True)
    tt = torch.tensor(t2[lo2:hi2], device=dev, requires_grad=True)
    lt = torch.tensor([math.log(0.1)], device=dev, requires_grad=True)
    m2 = SigLIP2BCELossDDP(bias_init=-4.0, label_smoothing=0.1, precision="bf16x3").to(dev)
    loss = m2(vt, tt, lt)
    loss.backward()
    o = co.siglip_loss(v2, t2, math.log(0.1), bias=-4.0, variant="bce2", label_smoothing=0.1)
    check("siglip2 bce ddp loss", abs(loss.item() - o["loss"]) <= 1e-5 * abs(o["loss"]), (loss.item(), o["loss"]))
    check("  dvideo rows", rel(vt.grad.cpu().numpy(), o["dvideo"][lo2:hi2]) <= 2e-3, rel(vt.grad.cpu().numpy(), o["dvideo"][lo2:hi2]))
    check("  dtext rows", rel(tt.grad.cpu().numpy(), o["dtext"][lo2:hi2]) <= 2e-3, rel(tt.grad.cpu().numpy(), o["dtext"][lo2:hi2]))
    check("  dbias", abs(m2.bias.grad.item() - o["dbias"]) <= 2e-3 * max(abs(o["dbias"]), 1e-4), (m2.bias.grad.item(), o["dbias"]))

    # ---------------- retrieval: text database sharded by rows, exact-grid embeddings with planted ties ----------------
    Nv, M, D = 3000, 1237, 256
    vv = ro.exact_grid_embeddings(Nv, D, 3); tx = ro.exact_grid_embeddings(M, D, 4)
    tx[700] = tx[5]; tx[1236] = tx[640]                     # exact duplicates straddling the shards
    gt = np.random.default_rng(5).integers(0, M, size=Nv); gt[:4] = [5, 700, 640, 1236]
    keep = []
    r = compute_recall_at_k_streaming(torch.tensor(vv, device=dev), torch.tensor(tx, device=dev), torch.tensor(gt, device=dev),
                                      k_values=[1, 5, 10, 50], use_ddp=True, _counts_out=keep)
    sim = ro.similarity(vv, tx)
    ranks = ro.gt_ranks(sim, gt)
    check("retrieval rank counts bit-exact (text shards all-reduced)", bool((keep[0].cpu().numpy() + 1 == ranks).all()),
          int((keep[0].cpu().numpy() + 1 != ranks).sum()))
    check("retrieval recall == oracle", r == ro.recall_at_k_streaming(vv, tx, gt, [1, 5, 10, 50]), r)
    m = compute_metrics_streaming(torch.tensor(vv, device=dev), torch.tensor(tx, device=dev), torch.tensor(gt, device=dev),
                                  k_values=[1, 5, 10, 50], use_ddp=True)
    om = ro.metrics_streaming(vv, tx, gt, k_values=(1, 5, 10, 50))
    # normalised inputs are no longer exact: near-ties may swap one rank (fp32 numpy vs bf16x3 tensor core)
    check("retrieval MRR_V2T (normalised inputs)", abs(m["MRR_V2T"] - om["MRR_V2T"]) <= 1e-5, (m["MRR_V2T"], om["MRR_V2T"]))
    s, i = streaming_topk(torch.tensor(vv, device=dev), torch.tensor(tx, device=dev), 10, use_ddp=True)
    os_, oi = ro.topk_lowest_index(sim, 10)
    check("retrieval top-10 indices bit-exact", bool((i.cpu().numpy() == oi).all()), int((i.cpu().numpy() != oi).sum()))
    check("retrieval top-10 scores bit-exact", bool((s.cpu().numpy() == os_).all()), "")

    # ---------------- epoch-end embedding stores: ragged NCCL gather feeding the sharded sweep ----------------
    from pathlib import Path
    from deepcoro_clip_b200 import EmbeddingStore, epoch_end_retrieval_metrics
    gg = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "retrieval_gauss_300x200.npz")
    cut_v = [0] + [int(300 * (q + 1) / world) + (7 if q % 2 == 0 and q + 1 < world else 0) for q in range(world)]
    cut_v[-1] = 300
    cut_t = [int(200 * q / world) for q in range(world)] + [200]
    vs, ts = EmbeddingStore(64, capacity=32, device=dev), EmbeddingStore(64, capacity=16, device=dev)
    for a in range(cut_v[rank], cut_v[rank + 1], 37):                     # uneven batches, buffer growth on the device
        vs.append(torch.tensor(gg["video"][a:min(a + 37, cut_v[rank + 1])], device=dev))
    ts.append(torch.tensor(gg["text"][cut_t[rank]:cut_t[rank + 1]], device=dev))
    mets = epoch_end_retrieval_metrics(vs, ts, torch.tensor(gg["gt"][cut_v[rank]:cut_v[rank + 1]], device=dev),
                                       k_values=(1, 5, 10, 50))
    refm = dict(zip([str(k) for k in gg["keys"]], gg["values"]))
    check("embedding stores: ragged gather + sharded metrics == reference golden",
          all(mets[k] == refm[k] for k in ("Recall@1", "Recall@5", "Recall@10", "Recall@50")) and
          abs(mets["MRR_V2T"] - refm["MRR_V2T"]) < 1e-9, {k: mets[k] for k in ("Recall@1", "Recall@10", "MRR_V2T")})

    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print(f"dist check ok on {world} ranks", flush=True)


if __name__ == "__main__":
    main()
