"""GPU parity: streaming retrieval (rank counts, recall@k, MRR, top-k lists) vs the reference golden vectors and
the numpy oracle. Bar (BASELINE.json north_star): top-k indices and recall counts BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as ro
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.tensor(np.asarray(a), device=DEV)


def test_known_answers_identity_and_antidiagonal():
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_metrics_streaming
    g = np.load(GOLDEN / "retrieval_known.npz")
    eye = torch.eye(5, device=DEV)
    m = compute_metrics_streaming(eye, eye, torch.arange(5, device=DEV), k_values=[1, 3, 5])
    ref = dict(zip([str(k) for k in g["eye_keys"]], g["eye_values"]))
    for k in ("Recall@1", "Recall@3", "Recall@5", "MRR_V2T", "alignment_score", "video_norm", "text_norm"):
        assert abs(m[k] - ref[k]) < 1e-6, k
    assert m["median_rank"] == 1
    anti = torch.flip(torch.eye(5, device=DEV), dims=[1])
    m2 = compute_metrics_streaming(eye, anti, torch.arange(5, device=DEV), k_values=[1])
    assert m2["Recall@1"] == 20.0
    o = ro.metrics_streaming(np.eye(5, dtype=np.float32), np.flip(np.eye(5, dtype=np.float32), 1), np.arange(5), [1])
    assert abs(m2["MRR_V2T"] - o["MRR_V2T"]) < 1e-12     # all-ties rows: lowest-index rule == oracle


def test_gauss_golden_metrics():
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_metrics_streaming
    g = np.load(GOLDEN / "retrieval_gauss_300x200.npz")
    m = compute_metrics_streaming(_t(g["video"]), _t(g["text"]), _t(g["gt"]), k_values=[1, 5, 10, 50],
                                  video_chunk_size=128, text_chunk_size=64, device="cuda")
    ref = dict(zip([str(k) for k in g["keys"]], g["values"]))
    for k in ("Recall@1", "Recall@5", "Recall@10", "Recall@50"):
        assert m[k] == ref[k], (k, m[k], ref[k])
    assert abs(m["MRR_V2T"] - ref["MRR_V2T"]) < 1e-9
    assert abs(m["alignment_score"] - ref["alignment_score"]) < 1e-6
    assert set(m) == set(ref)


def test_exact_grid_golden_topk_and_recall_bit_exact():
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_recall_at_k_streaming, streaming_topk
    g = np.load(GOLDEN / "retrieval_grid_257x300.npz")
    v, t = _t(g["video"]), _t(g["text"])
    r = compute_recall_at_k_streaming(v, t, _t(g["gt"]), k_values=[1, 5, 10], device="cuda")
    ref = dict(zip([str(k) for k in g["keys"]], g["values"]))
    assert r == {k: ref[k] for k in r}
    s, i = streaming_topk(v, t, 10)
    sim = ro.similarity(g["video"], g["text"])
    ov, oi = ro.topk_lowest_index(sim, 10)
    assert (i.cpu().numpy() == oi).all()
    assert (s.cpu().numpy() == ov).all()                       # scores bit-exact too (exact arithmetic inputs)
    s11 = -np.sort(-sim, axis=1)[:, :11]
    tie_free = (np.diff(s11, axis=1) != 0).all(axis=1)
    assert (i.cpu().numpy()[tie_free] == g["topk_idx"][tie_free]).all()   # == torch.topk of the reference


@pytest.mark.parametrize("N,M,D,k", [(1000, 777, 512, 10), (130, 5000, 512, 50), (4096, 2048, 768, 5), (300, 40, 64, 50)])
def test_exact_grid_vs_oracle(N, M, D, k):
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_recall_at_k_streaming, streaming_topk
    v = ro.exact_grid_embeddings(N, D, 3)
    t = ro.exact_grid_embeddings(M, D, 4)
    gt = np.random.default_rng(5).integers(0, M, size=N)
    keep = []
    kv = [1, 5, 10, 50]
    r = compute_recall_at_k_streaming(_t(v), _t(t), _t(gt), k_values=kv, _counts_out=keep)
    sim = ro.similarity(v, t)
    ranks = ro.gt_ranks(sim, gt)
    assert (keep[0].cpu().numpy() + 1 == ranks).all()           # every rank count exact
    assert r == ro.recall_at_k_streaming(v, t, gt, kv)
    s, i = streaming_topk(_t(v), _t(t), k)
    ov, oi = ro.topk_lowest_index(sim, k)
    kk = min(k, M)
    assert (i.cpu().numpy()[:, :kk] == oi).all() and (s.cpu().numpy()[:, :kk] == ov).all()


def test_planted_ties_lowest_index():
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_recall_at_k_streaming, streaming_topk
    rng = np.random.default_rng(7)
    v = ro.exact_grid_embeddings(600, 128, 8)
    t = ro.exact_grid_embeddings(900, 128, 9)
    dup = rng.integers(0, 900, size=90)
    t[rng.integers(0, 900, size=90)] = t[dup]                  # exact duplicate text rows => exact score ties
    gt = rng.integers(0, 900, size=600)
    gt[:90] = dup                                               # many ground truths sit inside a tie group
    keep = []
    compute_recall_at_k_streaming(_t(v), _t(t), _t(gt), k_values=[1, 5], _counts_out=keep)
    sim = ro.similarity(v, t)
    assert (keep[0].cpu().numpy() + 1 == ro.gt_ranks(sim, gt)).all()
    s, i = streaming_topk(_t(v), _t(t), 16)
    _, oi = ro.topk_lowest_index(sim, 16)
    assert (i.cpu().numpy() == oi).all()


def test_gaussian_normalised_counts_match_oracle():
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_metrics_streaming
    rng = np.random.default_rng(4)
    v = rng.standard_normal((3000, 512)).astype(np.float32)
    t = rng.standard_normal((2000, 512)).astype(np.float32)
    gt = rng.integers(0, 2000, size=3000)
    m = compute_metrics_streaming(_t(v), _t(t), _t(gt), k_values=[1, 5, 10])
    o = ro.metrics_streaming(v, t, gt, [1, 5, 10])
    for k in ("Recall@1", "Recall@5", "Recall@10"):
        assert m[k] == o[k]
    assert abs(m["MRR_V2T"] - o["MRR_V2T"]) < 1e-7
    assert abs(m["alignment_score"] - o["alignment_score"]) < 1e-6


def test_duplicate_texts_tie_exactly_after_normalisation():
    """Inexact (Gaussian, L2-normalised) embeddings with duplicated text rows: the duplicates of the ground truth must
    tie EXACTLY with it (ground-truth similarity taken from the tensor core, not from a CUDA-core dot product), so the
    rank is 1 + the number of duplicates with a lower index, whatever the rounding."""
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_metrics_streaming, streaming_topk
    rng = np.random.default_rng(21)
    M, D = 500, 512
    t = rng.standard_normal((M, D)).astype(np.float32)
    t[100:150] = t[0:50]
    t[300:325] = t[0:25]
    gt = np.concatenate([np.arange(0, 50), np.arange(100, 150), np.arange(300, 325)])
    gt = np.tile(gt, 8)
    v = (t[gt] + 0.1 * rng.standard_normal((len(gt), D))).astype(np.float32)
    expect = np.where(gt >= 300, 3, np.where(gt >= 100, 2, 1))
    from deepcoro_clip_b200 import retrieval_metrics_streaming as rms
    keep = []
    vop, top, _, _, _ = rms._operands(_t(v), _t(t), True, "auto")
    rms._recall_from_operands(vop, top, _t(gt), [1, 2, 3], False, None, keep)
    assert (keep[0].cpu().numpy() + 1 == expect).all()
    m = compute_metrics_streaming(_t(v), _t(t), _t(gt), k_values=[1, 2, 3])
    assert m["Recall@1"] == 100.0 * (expect <= 1).mean() and m["Recall@3"] == 100.0
    assert abs(m["MRR_V2T"] - (1.0 / expect).mean()) < 1e-12
    s, i = streaming_topk(_t(v), _t(t), 3, normalize=True)
    i = i.cpu().numpy()
    base = gt % 100
    trip = base < 25
    assert (i[trip] == np.stack([base[trip], base[trip] + 100, base[trip] + 300], 1)).all()
    assert (i[~trip][:, :2] == np.stack([base[~trip], base[~trip] + 100], 1)).all()


def test_exact_in_bf16_probe():
    """precision="auto": the one-pass device probe agrees with the definition (x == bf16(x) everywhere) for fp32 / fp16 /
    bf16 inputs, strided rows, a single inexact element anywhere, and NaN (inexact)."""
    from deepcoro_clip_b200.retrieval_metrics_streaming import _exact_in_bf16
    from oracle import retrieval_oracle as ro
    g = torch.tensor(ro.exact_grid_embeddings(3000, 520, 9), device="cuda:0")           # entries k/128: exact in bf16
    assert _exact_in_bf16(g) and _exact_in_bf16(g, g[:7]) and _exact_in_bf16(g.bfloat16())
    assert _exact_in_bf16(g.half()) and _exact_in_bf16(g[:, :300]) and _exact_in_bf16(g.double())
    for pos in ((0, 0), (2999, 519), (1234, 77)):
        h = g.clone()
        h[pos] += 1e-4
        assert not _exact_in_bf16(h) and not _exact_in_bf16(g, h) and not _exact_in_bf16(h[:, :520:1])
    h = g.clone()
    h[5, 5] = float("nan")
    assert not _exact_in_bf16(h)
    assert not _exact_in_bf16(torch.randn(64, 64, device="cuda:0"))


def test_inference_topk_and_top5_bindings():
    """inference()'s torch.topk over the dense similarity matrix (runners/video_constrative_learning_runner.py:1757-1758)
    and save_retrieval_results' per-row top-5 (utils/wandb_logger.py:951-957) through streaming_topk: same indices on
    tie-free inputs, without the [N, M] matrix."""
    from deepcoro_clip_b200 import inference_topk_indices, top5_predictions
    g = torch.Generator().manual_seed(11)
    v = torch.nn.functional.normalize(torch.randn(700, 512, generator=g), dim=-1).to(DEV)
    t = torch.nn.functional.normalize(torch.randn(1500, 512, generator=g), dim=-1).to(DEV)
    sim = v @ t.T
    ref_s, ref_i = torch.topk(sim, k=6, dim=1)
    clear = (ref_s[:, :-1] - ref_s[:, 1:]).min(dim=1).values > 1e-5          # rows whose top-6 margins exceed the operand error
    idx = inference_topk_indices(v, t, 5)
    assert idx.shape == (700, 5) and idx.dtype == torch.int64 and clear.float().mean().item() > 0.9
    assert (idx[clear] == ref_i[clear][:, :5]).all()
    s5, i5 = top5_predictions(v, t)
    assert (i5 == idx).all() and (s5[clear] - ref_s[clear][:, :5]).abs().max().item() <= 2e-5
    assert inference_topk_indices(v, t[:3], 5).shape == (700, 3)             # k clamps to the number of texts (:951)


def test_embedding_store_on_device_single_process():
    """SURVEY 8f #3 on hardware: validation embeddings appended batch by batch into the device-resident store (buffer growth
    included), then the epoch-end metrics straight from the device buffers — the reference's golden values. The ragged
    multi-rank gather over NCCL is part of tools/gpu_check_dist.py (tests/test_gpu_multi.py)."""
    from deepcoro_clip_b200 import EmbeddingStore, epoch_end_retrieval_metrics
    g = np.load(GOLDEN / "retrieval_gauss_300x200.npz")
    vs, ts = EmbeddingStore(64, capacity=32, device=DEV), EmbeddingStore(64, capacity=16, device=DEV)
    for a in range(0, 300, 37):
        vs.append(_t(g["video"][a:a + 37]))
    for a in range(0, 200, 64):
        ts.append(torch.tensor(g["text"][a:a + 64]))                 # host batches are accepted as well
    assert vs.local().is_cuda and vs.local().shape == (300, 64) and ts.local().shape == (200, 64)
    m = epoch_end_retrieval_metrics(vs, ts, _t(g["gt"]), k_values=(1, 5, 10, 50))
    ref = dict(zip([str(k) for k in g["keys"]], g["values"]))
    for k in ("Recall@1", "Recall@5", "Recall@10", "Recall@50"):
        assert m[k] == ref[k], (k, m[k], ref[k])
    assert abs(m["MRR_V2T"] - ref["MRR_V2T"]) < 1e-9 and abs(m["alignment_score"] - ref["alignment_score"]) < 1e-6
    vs.reset()
    assert vs.local().shape[0] == 0
