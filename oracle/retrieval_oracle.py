"""ORACLE (test infrastructure, never imported by the product package).

numpy restatement of /root/reference/utils/retrieval_metrics_streaming.py:
  compute_recall_at_k_streaming   :10-101   (chunked matmul + running top-k merge)
  compute_metrics_streaming       :104-197  (normalise, recall@k, MRR via argsort rank, alignment, norms)

Tie rule: the reference inherits torch.topk / argsort tie order, which is unspecified. BASELINE.json's north_star
defines ties as LOWEST INDEX FIRST; this oracle implements that rule with a stable descending sort
(rank_i = 1 + #{j: s_ij > s_ig} + #{j < g: s_ij == s_ig}). On tie-free inputs it is
bit-identical to the reference (pinned by tests/golden/retrieval_*.npz from oracle/gen_golden.py).
"""
from __future__ import annotations

import numpy as np


def similarity(video: np.ndarray, text: np.ndarray, dtype=np.float32) -> np.ndarray:
    return np.asarray(video, dtype=dtype) @ np.asarray(text, dtype=dtype).T


def topk_lowest_index(sim: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """Row-wise top-k by (score desc, index asc). Returns (scores [N,k], indices [N,k] int64)."""
    k = min(k, sim.shape[1])
    order = np.argsort(-sim, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(sim, order, axis=1), order.astype(np.int64)


def gt_ranks(sim: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """1-based rank of the ground-truth column under the lowest-index tie rule."""
    n = sim.shape[0]
    g = np.asarray(gt, dtype=np.int64)
    sg = sim[np.arange(n), g][:, None]
    greater = (sim > sg).sum(axis=1)
    cols = np.arange(sim.shape[1])[None, :]
    ties_before = ((sim == sg) & (cols < g[:, None])).sum(axis=1)
    return 1 + greater + ties_before


def recall_at_k_streaming(video, text, gt, k_values=(1, 5, 10, 50), dtype=np.float32, chunk: int = 4096) -> dict:
    """Recall@k in PERCENT (retrieval_metrics_streaming.py:99); does NOT normalise (:35-41)."""
    video = np.asarray(video, dtype=dtype)
    text = np.asarray(text, dtype=dtype)
    gt = np.asarray(gt, dtype=np.int64)
    n = video.shape[0]
    hits = {k: 0 for k in k_values}
    for s in range(0, n, chunk):
        sim = video[s:s + chunk] @ text.T
        ranks = gt_ranks(sim, gt[s:s + chunk])
        for k in k_values:
            hits[k] += int((ranks <= k).sum())
    out = {}
    k_max = max(k_values)
    width = min(k_max, text.shape[0])          # the reference's best_indices has min(k_max, M) columns (:64, :80)
    for k in k_values:
        if k <= width:                          # (:90) recall for k > width stays 0
            out[f"Recall@{k}"] = hits[k] / n * 100
        else:
            out[f"Recall@{k}"] = 0.0
    return out


def metrics_streaming(video, text, gt, k_values=(1, 5, 10, 50), dtype=np.float32, chunk: int = 4096) -> dict:
    """compute_metrics_streaming (:104-197): float32, F.normalize both sides, recall, MRR, alignment, norms."""
    v = np.asarray(video, dtype=dtype)
    t = np.asarray(text, dtype=dtype)
    gt = np.asarray(gt, dtype=np.int64)
    vn = np.maximum(np.sqrt((v * v).sum(1, keepdims=True)), 1e-12)
    tn = np.maximum(np.sqrt((t * t).sum(1, keepdims=True)), 1e-12)
    v = (v / vn).astype(dtype)
    t = (t / tn).astype(dtype)
    out = recall_at_k_streaming(v, t, gt, k_values, dtype, chunk)
    n = v.shape[0]
    mrr = 0.0
    for s in range(0, n, chunk):
        sim = v[s:s + chunk] @ t.T
        ranks = gt_ranks(sim, gt[s:s + chunk])
        for r in ranks:                          # host double accumulation in row order (:166)
            mrr += 1.0 / float(r)
    out["MRR_V2T"] = mrr / n
    out["alignment_score"] = float((v * t[gt]).sum(axis=1, dtype=np.float64).sum() / n)
    out["video_norm"] = float(np.sqrt((v * v).sum(1)).mean())
    out["text_norm"] = float(np.sqrt((t * t).sum(1)).mean())
    out["median_rank"] = 1                       # placeholder in the reference (:195)
    return out


def exact_grid_embeddings(n: int, d: int, seed: int) -> np.ndarray:
    """Entries k/128, k ~ U{-127..127}: every dot product over d <= 1024 terms is exact in fp32 in any order."""
    rng = np.random.default_rng(seed)
    return (rng.integers(-127, 128, size=(n, d)).astype(np.float32) / 128.0).astype(np.float32)
