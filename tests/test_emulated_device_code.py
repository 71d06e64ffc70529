"""CPU: the library's own plain-CUDA device code (csrc/retrieval_epi.cuh, csrc/alignment_diag.cuh — the very headers
compiled into libb200clip.so) built for the HOST under an emulation shim (tests/emul/cuda_emul.h: threads, warp shuffles,
ballots, barriers, atomics) and run against numpy. The tcgen05 tile engine cannot be emulated; tests/emul/retrieval_emul.cpp
replaces it by a driver that follows its epilogue contract (TileSeq item mode, TeCtx, chunk order, zero-filled padding).
This is how the kernels written after the round-1 GPU budget was spent (two-sweep top-k, alignment scalar tail) were
executed at all before their first run on hardware; the register-list / rank-count policies, which HAVE been validated
on the GPU, go through the same driver as a check of the emulation itself."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import GOLDEN

EMUL = Path(__file__).resolve().parent / "emul"
C_F, C_I, C_L = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong)


@pytest.fixture(scope="module")
def emul():
    so = EMUL / "libemul.so"
    srcs = [EMUL / "retrieval_emul.cpp", EMUL / "cuda_emul.h", EMUL.parents[1] / "deepcoro_clip_b200/csrc/retrieval_epi.cuh",
            EMUL.parents[1] / "deepcoro_clip_b200/csrc/alignment_diag.cuh", EMUL.parents[1] / "deepcoro_clip_b200/csrc/te_ctx.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", str(so), str(srcs[0])], check=True)
    return ctypes.CDLL(str(so))


def _p(a, t):
    return a.ctypes.data_as(t)


def exact_topk(S, k):
    idx = np.argsort(-S, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(S, idx, 1), idx


def two_sweeps(lib, S, k, segs, cap, col_offset=0):
    N, M = S.shape
    S = np.ascontiguousarray(S, np.float32)
    pm = np.full((N, 2 * segs, 32), -np.inf, np.float32)
    thr = np.zeros(N, np.float32)
    cnt = np.zeros(N, np.int32)
    bs = np.zeros((N, cap), np.float32)
    bi = np.full((N, cap), 0x7FFFFFFF, np.int32)
    os_, oi = np.zeros((N, k), np.float32), np.zeros((N, k), np.int64)
    ov = lib.emul_topk_two_sweeps(_p(S, C_F), N, M, k, segs, cap, col_offset, _p(pm, C_F), _p(thr, C_F), _p(cnt, C_I),
                                  _p(bs, C_F), _p(bi, C_I), _p(os_, C_F), _p(oi, C_L))
    return ov, os_, oi, thr, cnt


@pytest.mark.parametrize("N,M,k,segs,ties", [(130, 700, 10, 1, False), (70, 1300, 16, 3, False), (40, 257, 5, 2, True),
                                            (9, 520, 1, 2, True)])
def test_two_sweep_topk_device_code(emul, N, M, k, segs, ties):
    rng = np.random.default_rng(N + M)
    S = rng.standard_normal((N, M)).astype(np.float32)
    if ties:
        S = np.round(S * 2) / 2
    ov, s, i, thr, cnt = two_sweeps(emul, S, k, segs, cap=4096 if ties else max(64, 8 * k), col_offset=1000)
    assert ov == 0
    es, ei = exact_topk(S, k)
    assert (i == ei + 1000).all() and (s == es).all()
    assert (thr <= es[:, k - 1]).all()                      # the subset-maxima bound never exceeds the k-th best score
    assert (cnt == (S >= thr[:, None]).sum(1)).all()        # every candidate collected exactly once


def test_two_sweep_topk_device_code_overflow_and_short(emul):
    S = np.zeros((5, 600), np.float32)
    assert two_sweeps(emul, S, 5, 2, cap=64)[0] == 1        # all equal: overflow flag, caller falls back
    rng = np.random.default_rng(2)
    S = rng.standard_normal((4, 6)).astype(np.float32)      # fewer columns than k: -inf threshold, -1 / -inf padding
    ov, s, i, thr, _ = two_sweeps(emul, S, 10, 1, cap=64)
    assert ov == 0 and np.isneginf(thr).all()
    assert (i[:, :6] == np.argsort(-S, axis=1, kind="stable")).all() and (i[:, 6:] == -1).all()
    assert np.isneginf(s[:, 6:]).all()


def test_emulation_reproduces_gpu_validated_policies(emul):
    """RetrEpi<16> lists + topk_merge and RetrEpi<0> rank counts (measured on the GPU in round 1) through the same driver."""
    g = np.load(GOLDEN / "retrieval_grid_257x300.npz")
    S = np.ascontiguousarray(g["video"].astype(np.float64) @ g["text"].astype(np.float64).T, np.float32)   # exact grid
    N, M = S.shape
    k, segs = 10, 2
    ps = np.zeros((N, 2 * segs, k), np.float32)
    pi = np.zeros((N, 2 * segs, k), np.int32)
    os_, oi = np.zeros((N, k), np.float32), np.zeros((N, k), np.int64)
    emul.emul_topk_register_lists(_p(S, C_F), N, M, k, segs, 0, _p(ps, C_F), _p(pi, C_I), _p(os_, C_F), _p(oi, C_L))
    es, ei = exact_topk(S, k)
    assert (oi == ei).all() and (os_ == es).all()
    ov, s2, i2, _, _ = two_sweeps(emul, S, k, segs, cap=128)
    assert ov == 0 and (i2 == oi).all() and (s2 == os_).all()
    gt = g["gt"].astype(np.int64)
    sgt = S[np.arange(N), gt].copy()
    counts = np.zeros(N, np.int32)
    emul.emul_rank_counts(_p(S, C_F), N, M, segs, _p(sgt, C_F), _p(gt, C_L), 0, _p(counts, C_I))
    col = np.arange(M)[None, :]
    want = ((S > sgt[:, None]) | ((S == sgt[:, None]) & (col < gt[:, None]))).sum(1)
    assert (counts == want).all()
    ref = dict(zip([str(x) for x in g["keys"]], g["values"]))
    for kk in (1, 5, 10):
        assert (counts < kk).mean() * 100 == ref[f"Recall@{kk}"]


@pytest.mark.parametrize("name", ["align_b64_d512", "align_siglip_b130_d96", "align_b300_d200"])
def test_alignment_scalar_kernel_device_code(emul, name):
    g = np.load(GOLDEN / f"{name}.npz")
    gated = bool(g["use_siglip"])
    v = g["video"].astype(np.float64)
    t = g["text"].astype(np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    S = v @ t.T
    n = S.shape[0]
    tau = float(np.exp(g["log_temp"][0]))
    bound = 0.7310585786300049 if gated else 1.0
    scale2 = 1.4426950408889634 / tau
    shift2 = scale2 * bound - min(max(2.0 * bound * scale2 - 120.0, 0.0), 100.0)       # dyn_prep_kernel
    dyn = np.zeros(16, np.float32)
    dyn[0], dyn[1], dyn[2], dyn[3], dyn[6], dyn[7] = scale2, shift2, 1.0 / tau, tau, 0.6931471805599453 * shift2, 1.0
    f = S / (1.0 + np.exp(-S)) if gated else S
    P = np.exp2(f * float(dyn[0]) - float(dyn[1]))                                     # logits_lse_fwd contract
    sums = np.concatenate([P.sum(0), P.sum(1), np.diag(S)]).astype(np.float32)
    out = np.zeros(4, np.float32)
    emul.emul_alignment_diag(_p(sums, C_F), n, _p(dyn, C_F), int(gated), _p(out, C_F))
    for got, key in zip(out[:3], ("cosine_f64", "logprob_f64", "prob_f64")):
        ref = float(g[key])
        assert abs(float(got) - ref) <= 5e-6 * max(1.0, abs(ref)), (key, float(got), ref)


def test_two_sweep_topk_seeded_fuzz(emul):
    """Random (N, M, k, segments), every third case quantised to force ties at the k-th score: exact top-k, -1 padding."""
    rng = np.random.default_rng(0)
    for it in range(12):
        N = int(rng.integers(1, 300)); M = int(rng.integers(1, 1500)); k = int(rng.integers(1, 17))
        segs = int(rng.integers(1, (M + 255) // 256 + 1))
        S = rng.standard_normal((N, M)).astype(np.float32)
        if it % 3 == 0:
            S = np.round(S * 3) / 3
        ov, s, i, thr, cnt = two_sweeps(emul, S, k, segs, cap=4096)
        kk = min(k, M)
        es, ei = exact_topk(S, kk)
        assert ov == 0 and (i[:, :kk] == ei).all() and (s[:, :kk] == es).all() and (i[:, kk:] == -1).all(), (N, M, k, segs)
