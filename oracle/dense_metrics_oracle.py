"""ORACLE (test infrastructure, never imported by the product package).

CPU restatement of the reference's dense multi-label retrieval metrics (utils/retrieval_metrics.py:65-324) in numpy /
plain Python. Only tests/ may import this. Pinned against tests/golden/dense_metrics_*.npz, which oracle/gen_golden.py
produced by calling the UNMODIFIED reference functions on tie-free inputs, and against the known-answer cases of the
reference's own tests/test_retrieval_metrics.py:11-147.

Ranking rule: score descending, LOWEST INDEX FIRST on ties (stable sort) — the north_star tie rule; the reference's
torch.argsort(descending=True) leaves tie order unspecified, so parity with it is only defined on tie-free rows.

Follows: _normalize_ground_truth_sets :7-62, compute_recall_at_k :65-101, compute_mrr :104-161,
compute_ndcg_at_k :204-246, compute_median_rank :249-284, compute_map :287-324.
"""
from __future__ import annotations

import math

import numpy as np


def gt_sets(gt, n):
    out = []
    arr = gt
    if isinstance(arr, np.ndarray):
        arr = arr.tolist()
    for entry in arr:
        if isinstance(entry, (list, tuple, set)):
            out.append({int(v) for v in entry if v is not None and int(v) >= 0})
        elif entry is None:
            out.append(set())
        else:
            out.append({int(entry)} if int(entry) >= 0 else set())
    out += [set() for _ in range(n - len(out))]
    return out[:n]


def ranking(sim: np.ndarray) -> np.ndarray:
    """[N, M] column order per row: score descending, lowest index first."""
    return np.argsort(-sim, axis=1, kind="stable")


def recall_at_k(sim, gt, k_values):
    n, m = sim.shape
    sets = gt_sets(gt, n)
    order = ranking(sim)
    out = {}
    for k in k_values:
        ku = min(k, m)
        hits = [1.0 if (s and any(int(c) in s for c in order[i, :ku])) else 0.0 for i, s in enumerate(sets)]
        out[f"Recall@{k}"] = float(sum(hits) / len(hits)) if hits else 0.0
    return out


def _best_rank(order_row, s):
    best = None
    pos = {int(c): r + 1 for r, c in enumerate(order_row)}
    for g in s:
        if g in pos and (best is None or pos[g] < best):
            best = pos[g]
    return best


def mrr(sim, gt):
    n, m = sim.shape
    if m == 1:
        return {"MRR_V2T": 1.0}
    sim = np.nan_to_num(sim, nan=0.0, posinf=1e4, neginf=-1e4)
    sets = gt_sets(gt, n)
    order = ranking(sim)
    vals = []
    for i, s in enumerate(sets):
        b = _best_rank(order[i], s) if s else None
        vals.append(1.0 / b if b else 0.0)
    return {"MRR_V2T": sum(vals) / len(vals) if vals else 0.0}


def ndcg_at_k(sim, gt, k_values):
    n, m = sim.shape
    sets = gt_sets(gt, n)
    order = ranking(sim)
    out = {}
    for k in k_values:
        ke = min(k, m)
        vals = []
        for i, s in enumerate(sets):
            if not s:
                vals.append(0.0)
                continue
            dcg = sum(1.0 / math.log2(r + 2) for r in range(ke) if int(order[i, r]) in s)
            ideal = min(len(s), ke)
            idcg = sum(1.0 / math.log2(r + 2) for r in range(ideal))
            vals.append(dcg / idcg if ideal > 0 and idcg > 0 else 0.0)
        out[f"NDCG@{k}_V2T"] = float(np.asarray(vals, np.float32).mean())
    return out


def median_rank(sim, gt):
    n, m = sim.shape
    sets = gt_sets(gt, n)
    order = ranking(sim)
    ranks = []
    for i, s in enumerate(sets):
        b = _best_rank(order[i], s) if s else None
        ranks.append(b if b is not None else m)
    r = np.sort(np.asarray(ranks, np.float32))
    return int(r[(len(r) - 1) // 2])           # torch.median: the lower of the two middle values


def mean_ap(sim, gt):
    n, m = sim.shape
    sets = gt_sets(gt, n)
    order = ranking(sim)
    aps = []
    for i, s in enumerate(sets):
        hits, psum = 0, 0.0
        if s:
            for r, c in enumerate(order[i], start=1):
                if int(c) in s:
                    hits += 1
                    psum += hits / r
        aps.append(psum / hits if hits else 0.0)
    return float(np.asarray(aps, np.float32).mean())
