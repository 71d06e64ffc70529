"""CPU: the attention-pool MMA kernels (csrc/attnpool_mma_kernels.cuh — the file compiled into libb200clip.so; its SASS was
checked to be unchanged by the split from attnpool_mma.cu) compiled for the HOST against an emulation of their hardware
primitives (tests/emul/pool_mma_prims_emul.h: mma.sync m16n8k16 and ldmatrix with the PTX fragment layouts, the swizzled
TMA tile, mbarriers, shuffles) and run against the closed forms of SURVEY Appendix A.4 in numpy. These kernels have been
validated on the GPU; the test pins the emulation to them, so that changes to the kernels can be executed on CPU before
they are timed on hardware."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

EMUL = Path(__file__).resolve().parent / "emul"
CSRC = EMUL.parents[1] / "deepcoro_clip_b200" / "csrc"
VP = ctypes.c_void_p


@pytest.fixture(scope="module")
def emul():
    so = EMUL / "libpoolemul.so"
    srcs = [EMUL / "pool_emul.cpp", EMUL / "pool_mma_prims_emul.h", EMUL / "cuda_emul.h", CSRC / "attnpool_mma_kernels.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", str(so), str(srcs[0])], check=True)
    return ctypes.CDLL(str(so))


def to_bf16(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return u.astype(np.uint16)


def bf16_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def ptr(a):
    return a.ctypes.data_as(VP) if a is not None else None


def make_inputs(B, N, D, H, seed, masked):
    rng = np.random.default_rng(seed)
    xb = to_bf16(rng.standard_normal((B, N, D)).astype(np.float32))
    x = bf16_to_f32(xb).astype(np.float64)
    qt = (rng.standard_normal((H, D)) * 0.2).astype(np.float32)
    mask = None
    if masked:
        mask = (rng.random((B, N)) < 0.2)
        mask[:, 0] = False
        mask = mask.astype(np.uint8)
    return xb, x, qt, mask


def softmax_stats(x, qt, mask):
    s = np.einsum("bnd,hd->bhn", x, qt.astype(np.float64))
    if mask is not None:
        s = np.where(mask[:, None, :] != 0, -np.inf, s)
    m = s.max(-1)
    e = np.exp(s - m[..., None])
    l = e.sum(-1)
    a = e / l[..., None]
    return s, m, l, a


@pytest.mark.parametrize("B,N,D,H,S,NW,masked", [(2, 70, 128, 8, 2, 8, False), (1, 45, 256, 4, 1, 16, True)])
def test_pool_forward_kernel_under_emulation(emul, B, N, D, H, S, NW, masked):
    xb, x, qt, mask = make_inputs(B, N, D, H, 1, masked)
    pm = np.zeros((B, S, H), np.float32)
    pl = np.zeros((B, S, H), np.float32)
    pa = np.zeros((B, S, H, D), np.float32)
    emul.emul_pool_fwd(ptr(xb), 1, ptr(mask), ptr(qt), None, B, N, D, H, S, NW, 2, ptr(pm), ptr(pl), ptr(pa))
    # merge of the split partials (attnpool_merge, attnpool.cu): common maximum, rescaled sums
    M = pm.max(1, keepdims=True)
    wgt = np.exp(pm - M)
    l = (pl * wgt).sum(1)
    xbar = (pa * wgt[..., None]).sum(1) / l[..., None]
    _, m_ref, l_ref, a = softmax_stats(x, qt, mask)
    ref = np.einsum("bhn,bnd->bhd", a, x)
    assert np.abs(xbar - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-6
    assert np.abs((M[:, 0] + np.log(l)) - (m_ref + np.log(l_ref))).max() <= 1e-5


def test_pool_forward_given_weights_under_emulation(emul):
    B, N, D, H, S = 2, 50, 128, 8, 2
    xb, x, _, _ = make_inputs(B, N, D, H, 2, False)
    w = np.random.default_rng(3).standard_normal((B, H, N)).astype(np.float32)
    pa = np.zeros((B, S, H, D), np.float32)
    emul.emul_pool_fwd(ptr(xb), 1, None, None, ptr(w), B, N, D, H, S, 8, 2, None, None, ptr(pa))
    ref = np.einsum("bhn,bnd->bhd", w.astype(np.float64), x)
    assert np.abs(pa.sum(1) - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-6


@pytest.mark.parametrize("fused_dq", [False, True])
@pytest.mark.parametrize("B,N,D,H,S,NW,masked", [(2, 70, 128, 8, 2, 8, False), (1, 45, 256, 4, 1, 16, True),
                                                (2, 100, 384, 8, 3, 8, True)])
def test_pool_backward_kernel_under_emulation(emul, B, N, D, H, S, NW, masked, fused_dq):
    xb, x, qt, mask = make_inputs(B, N, D, H, 4, masked)
    rng = np.random.default_rng(5)
    dxbar = rng.standard_normal((B, H, D)).astype(np.float32)
    _, m, l, a = softmax_stats(x, qt, mask)
    xbar = np.einsum("bhn,bnd->bhd", a, x)
    dx = np.zeros((B, N, D), np.uint16)
    ds = np.zeros((B, H, N), np.float32)
    part_dq = np.zeros((B, S, H, D), np.float32) if fused_dq else None
    emul.emul_pool_bwd(ptr(xb), 1, ptr(mask), ptr(qt), ptr(dxbar), ptr(xbar.astype(np.float32)), ptr(m.astype(np.float32)),
                       ptr(l.astype(np.float32)), B, N, D, H, S, NW, 2, ptr(dx), ptr(ds), None, ptr(part_dq))
    c = np.einsum("bhd,bhd->bh", dxbar.astype(np.float64), xbar)
    tdot = np.einsum("bhd,bnd->bhn", dxbar.astype(np.float64), x)
    ds_ref = a * (tdot - c[..., None])
    dx_ref = np.einsum("bhn,bhd->bnd", a, dxbar.astype(np.float64)) + np.einsum("bhn,hd->bnd", ds_ref, qt.astype(np.float64))
    assert np.abs(ds - ds_ref).max() <= 1e-4 * np.abs(ds_ref).max() + 1e-7
    got = bf16_to_f32(dx).astype(np.float64)
    # dx leaves as bf16 and its product uses the 16-bit hi parts of dxbar / qt only: bf16-level accuracy by design
    assert np.abs(got - dx_ref).max() <= 1e-2 * np.abs(dx_ref).max()
    assert np.linalg.norm(got - dx_ref) <= 4e-3 * np.linalg.norm(dx_ref)
    if fused_dq:
        # the opt-in kDq instantiation: dqt = sum_{b, n} ds_hn x_n from the same pass (written after the GPU budget was spent)
        dq_ref = np.einsum("bhn,bnd->hd", ds_ref, x)
        got_dq = part_dq.sum((0, 1))
        assert np.abs(got_dq - dq_ref).max() <= 2e-5 * np.abs(dq_ref).max() + 1e-6


def test_fused_dq_kernel_seeded_fuzz(emul):
    """Random shapes of the opt-in kDq backward (D in {128 .. 512}, 1..8 heads, token counts below / across the 32-token
    tile, more splits than tiles, masks, upstream dlse): ds and the fused query gradient against the closed forms."""
    rng = np.random.default_rng(0)
    for it in range(8):
        D = int(rng.choice([128, 256, 384, 512])); H = int(rng.integers(1, 9)); B = int(rng.integers(1, 4))
        N = int(rng.integers(1, 150)); S = int(rng.integers(1, 5)); masked = bool(rng.integers(0, 2))
        NW = 16 if D % 256 == 0 else 8
        xb, x, qt, mask = make_inputs(B, N, D, H, 100 + it, masked)
        dxbar = rng.standard_normal((B, H, D)).astype(np.float32)
        _, m, l, a = softmax_stats(x, qt, mask)
        xbar = np.einsum("bhn,bnd->bhd", a, x)
        dx = np.zeros((B, N, D), np.uint16); ds = np.zeros((B, H, N), np.float32); pdq = np.zeros((B, S, H, D), np.float32)
        dlse = rng.standard_normal((B, H)).astype(np.float32) if it % 2 else None
        emul.emul_pool_bwd(ptr(xb), 1, ptr(mask), ptr(qt), ptr(dxbar), ptr(xbar.astype(np.float32)), ptr(m.astype(np.float32)),
                           ptr(l.astype(np.float32)), B, N, D, H, S, NW, 2, ptr(dx), ptr(ds), ptr(dlse), ptr(pdq))
        c = np.einsum("bhd,bhd->bh", dxbar.astype(np.float64), xbar) - (dlse if dlse is not None else 0.0)
        ds_ref = a * (np.einsum("bhd,bnd->bhn", dxbar.astype(np.float64), x) - c[..., None])
        dq_ref = np.einsum("bhn,bnd->hd", ds_ref, x)
        case = (B, N, D, H, S, masked, dlse is not None)
        assert np.abs(ds - ds_ref).max() <= 2e-4 * np.abs(ds_ref).max() + 1e-9, case
        assert np.abs(pdq.sum((0, 1)) - dq_ref).max() <= 5e-5 * np.abs(dq_ref).max() + 1e-9, case
