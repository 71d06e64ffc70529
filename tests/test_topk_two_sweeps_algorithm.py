"""CPU: the ALGORITHM of the opt-in two-sweep top-k (csrc/retrieval.cu ColMaxEpi / kth_largest_kernel / CollectEpi /
topk_merge) restated in numpy with the kernels' own partitioning (128-column halves of 256-wide blocks, 32-column chunks,
one subset per (slot, position-in-chunk)), checked against a full stable sort: the k-th largest subset maximum is a lower
bound of the k-th best score, the candidate set contains every tie at the k-th score, and the merge applies the
lowest-index rule. The kernels themselves are covered by tests/test_gpu_z_topk_two_sweeps.py."""
import numpy as np
import pytest


def subset_maxima(S, segs):
    """part_max[row][slot][e]: slot = seg * 2 + wg, columns of 256-wide block b go to segment b * segs // n_blocks (the
    exact split does not matter for correctness, only disjointness does), wg = (col % 256) // 128, e = col % 32."""
    N, M = S.shape
    n_blocks = (M + 255) // 256
    pm = np.full((N, 2 * segs, 32), -np.inf, dtype=np.float32)
    col = np.arange(M)
    slot = (col // 256) * segs // n_blocks * 2 + (col % 256) // 128
    e = col % 32
    for j in range(M):
        pm[:, slot[j], e[j]] = np.maximum(pm[:, slot[j], e[j]], S[:, j])
    return pm


def kth_largest(vals, k):
    v = np.sort(vals.reshape(vals.shape[0], -1), axis=1)[:, ::-1]
    return v[:, k - 1] if v.shape[1] >= k else np.full(vals.shape[0], -np.inf, np.float32)


def two_sweeps(S, k, segs, cap):
    thr = kth_largest(subset_maxima(S, segs), k)
    out_s = np.full((S.shape[0], k), -np.inf, np.float32)
    out_i = np.full((S.shape[0], k), -1, np.int64)
    ncand = []
    for i in range(S.shape[0]):
        cand = np.nonzero(S[i] >= thr[i])[0]
        ncand.append(len(cand))
        if len(cand) > cap:
            return None, None, ncand
        order = sorted(cand, key=lambda j: (-S[i, j], j))[:k]        # topk_merge: score desc, index asc
        out_s[i, :len(order)] = S[i, order]
        out_i[i, :len(order)] = order
    return out_s, out_i, ncand


def exact(S, k):
    idx = np.argsort(-S, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(S, idx, 1), idx


@pytest.mark.parametrize("N,M,k,segs,ties", [(40, 1000, 10, 1, False), (16, 3000, 16, 2, False), (24, 700, 5, 3, True),
                                            (8, 257, 1, 1, True), (8, 40, 10, 1, False)])
def test_two_sweep_topk_equals_stable_sort(N, M, k, segs, ties):
    rng = np.random.default_rng(N * 1000 + M)
    S = rng.standard_normal((N, M)).astype(np.float32)
    if ties:
        S = np.round(S * 4) / 4          # heavy ties, including at the k-th score
    s, i, ncand = two_sweeps(S, k, segs, cap=max(64, 8 * k) if not ties else 10 ** 9)
    es, ei = exact(S, k)
    assert (i == ei).all() and (s == es).all()
    assert min(ncand) >= min(k, M)


def test_two_sweep_topk_fewer_columns_than_k_and_overflow():
    rng = np.random.default_rng(7)
    S = rng.standard_normal((4, 6)).astype(np.float32)
    s, i, _ = two_sweeps(S, 10, 1, cap=64)                       # k > M: threshold -inf, everything is a candidate
    assert (i[:, :6] == np.argsort(-S, axis=1, kind="stable")).all() and (i[:, 6:] == -1).all()
    S = np.zeros((2, 500), np.float32)                           # all equal: every column ties with the k-th score
    assert two_sweeps(S, 5, 1, cap=64)[0] is None                # -> overflow, the caller falls back to the register lists


def test_expected_candidates_per_row_is_small():
    """Capacity sizing: at the C4 shape (32,473 columns, k = 10, 64 subsets) a row collects ~11 candidates on average."""
    rng = np.random.default_rng(11)
    S = rng.standard_normal((64, 32473)).astype(np.float32)
    _, _, ncand = two_sweeps(S, 10, 1, cap=80)
    assert np.mean(ncand) < 16 and max(ncand) <= 80
