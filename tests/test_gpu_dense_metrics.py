"""GPU parity of the dense multi-label retrieval metrics (deepcoro_clip_b200/retrieval_metrics.py)
against golden values produced by the UNMODIFIED reference (utils/retrieval_metrics.py), the numpy oracle (ties: lowest
index first) and the known-answer cases of the reference's own tests/test_retrieval_metrics.py:11-147.
Recall and median rank are exact; MRR is the same left-to-right double sum; MAP / NDCG are fp32 means (1e-6)."""
import numpy as np
import pytest
import torch

from oracle import dense_metrics_oracle as dmo
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gt_arg(g):
    return torch.tensor(g["gt"][:, 0]) if g["gt"].shape[1] == 1 else [[int(c) for c in row if c >= 0] for row in g["gt"]]


@pytest.mark.parametrize("name", ["dense_metrics_120x90_g1", "dense_metrics_200x300_g4", "dense_metrics_64x7_g3"])
def test_dense_metrics_match_reference_golden(name):
    from deepcoro_clip_b200 import retrieval_metrics as rm
    g = np.load(GOLDEN / f"{name}.npz")
    sim = torch.tensor(g["sim"], device=DEV)
    gt = _gt_arg(g)
    ks = [int(k) for k in g["k_values"]]
    rec = rm.compute_recall_at_k(sim, gt, ks)
    nd = rm.compute_ndcg_at_k(sim, gt, ks)
    for i, k in enumerate(ks):
        assert rec[f"Recall@{k}"] == float(g["recall"][i])
        assert abs(nd[f"NDCG@{k}_V2T"] - float(g["ndcg"][i])) <= 1e-6
    assert abs(rm.compute_mrr(sim, gt)["MRR_V2T"] - float(g["mrr"])) <= 1e-15
    assert abs(rm.compute_map(sim, gt) - float(g["map"])) <= 1e-6
    assert rm.compute_median_rank(sim, gt) == int(g["median_rank"])
    if g["gt"].shape[1] > 1:
        # the same ground truth as a -1 padded 2-D index tensor (device-side sort + duplicate removal) with a duplicate
        gt2 = torch.tensor(np.concatenate([g["gt"], g["gt"][:, :1]], axis=1))
        assert rm.compute_recall_at_k(sim, gt2, ks) == rec
        assert rm.compute_ndcg_at_k(sim, gt2, ks) == nd
        assert rm.compute_map(sim, gt2) == rm.compute_map(sim, gt)
    allm = rm.compute_all_dense_metrics(sim, gt, recall_k=ks, ndcg_k=ks)
    assert allm["Recall@5"] == rec["Recall@5"] and allm["MedianRank_V2T"] == int(g["median_rank"])
    # the materialised similarity matrix itself (normalize + matmul, bf16x3 operands ~ fp32)
    s2 = rm.compute_similarity_matrix(torch.tensor(g["video"], device=DEV), torch.tensor(g["text"], device=DEV))
    assert float((s2.cpu() - torch.tensor(g["sim"])).abs().max()) <= 5e-6      # bf16x3 operands: ~2^-17 per product


def test_known_answers_of_the_reference_tests():
    """tests/test_retrieval_metrics.py: identity (:11-43), anti-diagonal 5x5 (:82-102), k > M clamp (:138-147)."""
    from deepcoro_clip_b200 import retrieval_metrics as rm
    n = 5
    sim = torch.zeros(n, n, device=DEV)
    torch.diagonal(sim)[:] = 1.0
    gt = torch.arange(n)
    rec = rm.compute_recall_at_k(sim, gt, k_values=[1, 3, 5])
    assert rec == {"Recall@1": 1.0, "Recall@3": 1.0, "Recall@5": 1.0}
    assert rm.compute_mrr(sim, gt)["MRR_V2T"] == 1.0
    assert abs(rm.compute_map(sim, gt) - 1.0) < 1e-6
    assert rm.compute_median_rank(sim, gt) == 1
    nd = rm.compute_ndcg_at_k(sim, gt, k_values=[1, 3, 5])
    assert all(abs(v - 1.0) < 1e-6 for v in nd.values())
    # worst case: the true match has the LOWEST score of its row
    sim = torch.ones(n, n, device=DEV) - torch.eye(n, device=DEV)
    sim += torch.arange(n, device=DEV).float()[None, :] * 1e-3           # tie-free
    o = dmo.recall_at_k(sim.cpu().numpy(), gt.numpy(), [1])
    assert rm.compute_recall_at_k(sim, gt, [1]) == o
    assert rm.compute_mrr(sim, gt)["MRR_V2T"] < 0.5 and rm.compute_median_rank(sim, gt) > 2.5
    # k larger than the number of candidates clamps (and warns) instead of failing
    sim = torch.randn(4, 3, device=DEV)
    r = rm.compute_recall_at_k(sim, torch.tensor([0, 1, 2, 0]), k_values=[5])
    assert r["Recall@5"] == 1.0
    assert rm.compute_mrr(torch.randn(4, 1, device=DEV), torch.zeros(4, dtype=torch.long))["MRR_V2T"] == 1.0


@pytest.mark.parametrize("N,M,gmax,dtype", [(1000, 4097, 6, torch.float32), (513, 800, 16, torch.bfloat16), (300, 33, 1, torch.float16)])
def test_dense_metrics_vs_oracle_with_ties_and_edge_rows(N, M, gmax, dtype):
    """Quantised similarities (many exact ties -> lowest index first), empty ground-truth rows, out-of-range and
    negative indices, NaN / inf entries for the MRR sanitiser; all dtypes the kernel reads."""
    from deepcoro_clip_b200 import retrieval_metrics as rm
    rng = np.random.default_rng(N + M)
    sim = (np.round(rng.standard_normal((N, M)) * 4) / 4).astype(np.float32)      # ~40 distinct values: heavy ties
    gt = []
    for i in range(N):
        n_i = int(rng.integers(0, gmax + 1))
        row = [int(c) for c in rng.choice(M + 3, size=n_i, replace=False)]          # some indices >= M
        if i % 50 == 0:
            row.append(-1)
        gt.append(row)
    sim_t = torch.tensor(sim, device=DEV).to(dtype)
    sim_o = sim_t.float().cpu().numpy()
    ks = [1, 5, 10, 50]
    assert rm.compute_recall_at_k(sim_t, gt, ks) == dmo.recall_at_k(sim_o, gt, ks)
    assert rm.compute_median_rank(sim_t, gt) == dmo.median_rank(sim_o, gt)
    assert abs(rm.compute_map(sim_t, gt) - dmo.mean_ap(sim_o, gt)) <= 1e-6
    nd, ndo = rm.compute_ndcg_at_k(sim_t, gt, ks), dmo.ndcg_at_k(sim_o, gt, ks)
    assert all(abs(nd[k] - ndo[k]) <= 1e-6 for k in nd)
    sim_bad = sim.copy()
    sim_bad[::7, 3] = np.nan
    sim_bad[::11, 5] = np.inf
    sim_bad[::13, 2] = -np.inf
    got = rm.compute_mrr(torch.tensor(sim_bad, device=DEV), gt)["MRR_V2T"]
    assert abs(got - dmo.mrr(sim_bad, gt)["MRR_V2T"]) <= 1e-15
    # per-item ranks against a stable argsort
    t = rm._row_terms(sim_t, gt)
    order = dmo.ranking(sim_o)
    pos = np.empty_like(order)
    np.put_along_axis(pos, order, np.arange(M)[None, :].repeat(N, 0), axis=1)
    ranks = t["ranks"].cpu().numpy()
    for i in range(0, N, 37):
        want = sorted(int(pos[i, c]) + 1 for c in set(gt[i]) if 0 <= c < M)
        assert sorted(int(r) for r in ranks[i] if r > 0) == want


def test_alignment_and_norms():
    from deepcoro_clip_b200 import retrieval_metrics as rm
    rng = np.random.default_rng(3)
    v = rng.standard_normal((257, 96)).astype(np.float32) * 3
    t = rng.standard_normal((300, 96)).astype(np.float32) * 0.5
    gt = rng.integers(0, 300, size=257)
    vt, tt = torch.tensor(v, device=DEV), torch.tensor(t, device=DEV)
    vh = v / np.linalg.norm(v, axis=1, keepdims=True)
    th = t / np.linalg.norm(t, axis=1, keepdims=True)
    a = rm.compute_alignment_score(vt[:10], tt[:10], all_video_embeddings=vt, all_text_embeddings=tt,
                                   global_ground_truth_indices_tensor=torch.tensor(gt, device=DEV))
    assert abs(a - float((vh * th[gt]).sum(1).mean())) <= 2e-6
    a2 = rm.compute_alignment_score(vt, tt[:257])
    assert abs(a2 - float((vh * th[:257]).sum(1).mean())) <= 2e-6
    n = rm.compute_embedding_norms(vt, tt)
    assert abs(n["video_norm"] - float(np.linalg.norm(v, axis=1).mean())) <= 1e-4
    assert abs(n["text_norm"] - float(np.linalg.norm(t, axis=1).mean())) <= 1e-4
