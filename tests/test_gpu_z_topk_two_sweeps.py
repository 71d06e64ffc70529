"""GPU parity of the OPT-IN two-sweep top-k (B200CLIP_TOPK2=1: retrieval_colmax -> kth_largest -> retrieval_collect ->
topk_merge) against the oracle, the golden vectors and the default register-list sweep: indices and scores BIT-EXACT.

Green on a B200 in round 2 (gpurun_out/r02b) and measured there (profiles/r02_topk.json): exact, but SLOWER than the register
lists at the C4 shape (17.1 ms against 10.4 ms for k = 10: the candidate append costs 13 ms), so the path stays opt-in and the
register lists stay the default. The algorithm is also checked on CPU in tests/test_topk_two_sweeps_algorithm.py."""
import os

import numpy as np
import pytest
import torch

from oracle import retrieval_oracle as ro
from tests.conftest import GOLDEN

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


def _both(v, t, k, monkeypatch):
    from deepcoro_clip_b200 import _lib
    from deepcoro_clip_b200.retrieval_metrics_streaming import streaming_topk
    monkeypatch.setenv("B200CLIP_TOPK2", "0")
    s0, i0 = streaming_topk(v, t, k, precision="bf16")
    monkeypatch.setenv("B200CLIP_TOPK2", "1")
    before = _lib.LAUNCHES
    s1, i1 = streaming_topk(v, t, k, precision="bf16")
    torch.cuda.synchronize()
    return (s0.cpu().numpy(), i0.cpu().numpy()), (s1.cpu().numpy(), i1.cpu().numpy()), _lib.LAUNCHES - before


def test_exact_grid_golden_bit_exact(monkeypatch):
    g = np.load(GOLDEN / "retrieval_grid_257x300.npz")
    v, t = torch.tensor(g["video"], device=DEV), torch.tensor(g["text"], device=DEV)
    (s0, i0), (s1, i1), launches = _both(v, t, 10, monkeypatch)
    ov, oi = ro.topk_lowest_index(ro.similarity(g["video"], g["text"]), 10)
    assert (i1 == oi).all() and (s1 == ov).all()
    assert (i1 == i0).all() and (s1 == s0).all()
    assert launches == 6          # 2 operand packs, colmax, kth_largest, collect, merge: the two-sweep path really ran


@pytest.mark.parametrize("N,M,D,k", [(1000, 5000, 128, 10), (300, 33000, 64, 16), (4100, 700, 256, 5), (130, 257, 64, 1)])
def test_matches_register_lists_on_exact_grids(N, M, D, k, monkeypatch):
    gen = torch.Generator().manual_seed(N + M)
    v = (torch.randint(-127, 128, (N, D), generator=gen).float() / 128).to(DEV)
    t = (torch.randint(-127, 128, (M, D), generator=gen).float() / 128).to(DEV)
    (s0, i0), (s1, i1), _ = _both(v, t, k, monkeypatch)
    assert (i1 == i0).all() and (s1 == s0).all()
    ov, oi = ro.topk_lowest_index(ro.similarity(v.cpu().numpy(), t.cpu().numpy()), k)
    assert (i1 == oi).all()


def test_planted_ties_lowest_index_first(monkeypatch):
    gen = torch.Generator().manual_seed(9)
    v = (torch.randint(-3, 4, (200, 64), generator=gen).float() / 4).to(DEV)       # coarse grid: many equal scores
    t = (torch.randint(-3, 4, (900, 64), generator=gen).float() / 4).to(DEV)
    t[500:520] = t[10:30]                                                          # exact duplicates of earlier texts
    (s0, i0), (s1, i1), _ = _both(v, t, 10, monkeypatch)
    ov, oi = ro.topk_lowest_index(ro.similarity(v.cpu().numpy(), t.cpu().numpy()), 10)
    assert (i1 == oi).all() and (s1 == ov).all() and (i0 == oi).all()


def test_overflow_falls_back_and_short_databases(monkeypatch):
    v = torch.ones(64, 64, device=DEV)
    t = torch.ones(3000, 64, device=DEV)                    # every score equal: all 3000 columns tie with the k-th
    (s0, i0), (s1, i1), launches = _both(v, t, 5, monkeypatch)
    assert (i1 == np.arange(5)[None, :]).all() and (i1 == i0).all()
    assert launches == 7                                    # ... + the register-list sweep after the overflow, merge
    gen = torch.Generator().manual_seed(3)
    v = (torch.randint(-127, 128, (40, 64), generator=gen).float() / 128).to(DEV)
    t = (torch.randint(-127, 128, (6, 64), generator=gen).float() / 128).to(DEV)
    (s0, i0), (s1, i1), _ = _both(v, t, 10, monkeypatch)   # k is clamped to the 6 texts
    assert i1.shape == (40, 6) and (i1 == i0).all() and (s1 == s0).all()
