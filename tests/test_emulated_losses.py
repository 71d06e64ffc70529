"""CPU: the package's own CLIP loss (deepcoro_clip_b200.loss: autograd function, arenas, label smoothing, diagonal and
normalise-backward corrections, the distributed plan) executed end to end on CPU — single process against the reference
goldens, and on TWO RANKS over gloo against the same full-batch goldens (every rank returns the full loss, the
local rows of the full-batch gradient and the identical log_temp gradient).

What runs underneath (tests/emul/loss_emul.cpp): the SHIPPED CUDA-core kernels (l2norm forward / backward, colsum, dyn_prep,
clip_finalize, clip_dlogtemp) under the host emulation, and MODELS of the two tcgen05 tile kernels that follow the contracts
of include/b200clip.h operation by operation (the tile kernels themselves are covered by the -m gpu tests only). No math of
the host code is replaced; the stand-ins are the device plumbing and the ctypes target."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import contrastive_oracle as co
from tests.conftest import GOLDEN

EMUL = Path(__file__).resolve().parent / "emul"
CSRC = EMUL.parents[1] / "deepcoro_clip_b200" / "csrc"


def build_emul():
    so = EMUL / "liblossemul.so"
    srcs = [EMUL / "loss_emul.cpp", EMUL / "pool_mma_prims_emul.h", EMUL / "cuda_emul.h", CSRC / "l2norm_kernels.cuh",
            CSRC / "scalars_kernels.cuh", CSRC / "siglip_kernels.cuh",
            CSRC / "retrieval_epi.cuh", CSRC / "alignment_diag.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", str(so), str(srcs[0])], check=True)
    return so


def patch_package(so, setattr_=setattr):
    """Points the package at the emulated library. Spawned gloo workers patch their own interpreter for good (plain
    setattr); the single-process tests pass pytest's monkeypatch.setattr so that the patch is undone after the test."""
    from deepcoro_clip_b200 import _lib, loss, ops
    emul = ctypes.CDLL(str(so))
    for name, (ret, types) in _lib._prototypes().items():
        fn = getattr(emul, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = ret, types

    def call(name, *args):
        rc = getattr(emul, "b200clip_" + name)(*[a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args])
        if rc != 0:
            raise _lib.B200ClipError(f"emulated b200clip_{name} failed with {rc}")

    setattr_(ops, "require_cuda", lambda *t: torch.device("cpu"))
    setattr_(ops, "call", call)
    setattr_(ops, "stream_ptr", lambda dev=None: 0)
    setattr_(_lib, "lib", lambda: emul)          # size queries (b200clip_rowlse_slots) go to the emulated library as well
    return loss


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


CASES = {"clip_c1_b64_d512": ("CLIPLoss", {}), "clip_ls_b48_d96": ("CLIPLoss", {"label_smoothing": 0.1}),
         "clip_b300_d200": ("CLIPLoss", {}),                 # D not a multiple of 64, B not a multiple of the tile height
         "contrastive_legacy_b32_d128": ("ContrastiveLoss", {}), "gated_siglip_legacy_b40_d128": ("SiglipLoss", {})}


@pytest.mark.parametrize("name", sorted(CASES))
def test_clip_loss_host_path_single_process(name, monkeypatch):
    loss_mod = patch_package(build_emul(), monkeypatch.setattr)
    g = np.load(GOLDEN / f"{name}.npz")
    cls, kw = CASES[name]
    v = torch.tensor(g["video"], dtype=torch.float32, requires_grad=True)
    t = torch.tensor(g["text"], dtype=torch.float32, requires_grad=True)
    lt = torch.tensor(g["log_temp"].astype(np.float32).reshape(1), requires_grad=True)
    out = getattr(loss_mod, cls)(**kw)(video_features=v, text_features=t, log_temp=lt)
    out.backward()
    loss, dv, dt, dlt = out.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item()
    ref = float(g["f32_loss"])
    assert abs(loss - ref) <= 1e-5 * abs(ref), (loss, ref)
    assert _rel(dv, g["f32_dvideo"]) <= 2e-3 and _rel(dt, g["f32_dtext"]) <= 2e-3
    rlt = float(g["f32_dlog_temp"].reshape(-1)[0])
    assert abs(dlt - rlt) <= 2e-3 * max(abs(rlt), 1e-3)


def _clip_body(rank, world, name):
    loss_mod = patch_package(build_emul())
    g = np.load(GOLDEN / f"{name}.npz")
    cls, kw = CASES[name]
    N = g["video"].shape[0]
    B = N // world
    lo, hi = rank * B, (rank + 1) * B
    v = torch.tensor(g["video"][lo:hi], dtype=torch.float32, requires_grad=True)
    t = torch.tensor(g["text"][lo:hi], dtype=torch.float32, requires_grad=True)
    lt = torch.tensor(g["log_temp"].astype(np.float32).reshape(1), requires_grad=True)
    loss = getattr(loss_mod, cls)(**kw)(video_features=v, text_features=t, log_temp=lt)
    loss.backward()
    return (loss.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item())


GLOO_CLIP = ["clip_c1_b64_d512", "clip_ls_b48_d96"]
GLOO_SIGLIP = ["siglip_mp_b32_t40_d64", "siglip_diag_b32_t32_d64", "siglip_entropy_b16_t32_d64"]
GLOO_LEGACY = [("contrastive", "contrastive_legacy_b32_d128"), ("siglip", "gated_siglip_legacy_b40_d128")]


def _all_rank(rank, world, port, out):
    """ONE two-rank gloo job runs every distributed scenario of this file (a spawn costs ~7 s of interpreter start-up)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        for name in GLOO_CLIP:
            res[("clip", name)] = _clip_body(rank, world, name)
        for name in GLOO_SIGLIP:
            res[("siglip", name)] = _siglip_body(rank, world, name)
        for cls, name in GLOO_LEGACY:
            res[("legacy", cls)] = _legacy_body(rank, world, cls, name)
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def gloo_results():
    build_emul()
    world = 2
    port = 29500 + (os.getpid() % 1500)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_all_rank, args=(world, port, out), nprocs=world, join=True)
    return {r: out[r] for r in range(world)}


@pytest.mark.parametrize("name", GLOO_CLIP)
def test_clip_loss_host_path_two_ranks_gloo(name, gloo_results):
    world = 2
    out = {r: gloo_results[r][("clip", name)] for r in range(world)}
    g = np.load(GOLDEN / f"{name}.npz")
    ref = float(g["f32_loss"])
    B = g["video"].shape[0] // world
    rlt = float(g["f32_dlog_temp"].reshape(-1)[0])
    for r in range(world):
        loss, dv, dt, dlt = out[r]
        assert abs(loss - ref) <= 1e-5 * abs(ref), (r, loss, ref)                     # the FULL loss on every rank
        assert _rel(dv, g["f32_dvideo"][r * B:(r + 1) * B]) <= 2e-3                    # local rows of the full gradient
        assert _rel(dt, g["f32_dtext"][r * B:(r + 1) * B]) <= 2e-3
        assert abs(dlt - rlt) <= 2e-3 * max(abs(rlt), 1e-3)                            # identical full value
    assert out[0][0] == out[1][0] and out[0][3] == out[1][3]


# ---------------------------------------------------------------------------------------------------------------------
# SigLIP multi-positive loss: shipped compaction / positive-correction / scalar-tail kernels under the
# emulation, contract models of the dense tile kernels (b200clip_logits_bwd mode 2, b200clip_siglip_dense_fwd)
# ---------------------------------------------------------------------------------------------------------------------
SIGLIP_CASES = {
    # the SigLIP classes kept next to the unified loss (no tau clamp / no logit clamp / label smoothing /
    # "pos_mask > 0" weight rule) are flags of the same kernels
    "pairwise_mp_b24_t40_d64": ("SiglipPairwiseFeatureLoss", dict(positive_weight=1.5, negative_weight=0.7)),
    "bce2_b40_d96": ("SigLIP2BCELoss", dict()),
    "bce2_ls_noclamp_b32_d64": ("SigLIP2BCELoss", dict(bias_init=-4.0, label_smoothing=0.1)),
    "mp2_ls_b20_t36_d64": ("SigLIP2MultiPositiveBCELoss", dict(bias_init=-3.0, positive_weight=2.0, negative_weight=0.5,
                                                               label_smoothing=0.2)),
    "mp2_diag_b18_t30_d64": ("SigLIP2MultiPositiveBCELoss", dict(bias_init=-5.0)),
    "siglip_diag_b32_t32_d64": ("SigLIPLoss", dict()),
    # the entropy regulariser (contrastive.py:19-68, 306-313): row statistics, device-side coefficient, gradient mode 3
    "siglip_entropy_b16_t32_d64": ("SigLIPLoss", dict(entropy_regularization=True, bias_init=-2.0, min_entropy_threshold=5.0)),
    "pairwise_entropy_auto_b16_t48_d64": ("SiglipPairwiseFeatureLoss", dict(auto_positive_weight=True, entropy_regularization=True,
                                                                           entropy_weight=0.2, min_entropy_threshold=6.0)),
    "siglip_mp_b32_t40_d64": ("SigLIPLoss", dict()),
    "siglip_mp_noweights_b24_t50_d96": ("SigLIPLoss", dict(positive_weight=2.0, negative_weight=0.5, use_severity_weights=False)),
    "siglip_autobalance_b16_t48_d64": ("SigLIPLoss", dict(auto_balance=True)),
}


def _siglip_inputs(g, lo=None, hi=None):
    sl = slice(lo, hi)
    v = torch.tensor(g["video"][sl], dtype=torch.float32, requires_grad=True)
    t = torch.tensor(g["text"], dtype=torch.float32, requires_grad=True)
    lt = torch.tensor(g["log_temp"].astype(np.float32).reshape(1), requires_grad=True)
    kw = {}
    if "in_pos_mask" in g.files:
        kw["pos_mask"] = torch.tensor(g["in_pos_mask"][sl], dtype=torch.float32)
    if "in_pos_weights" in g.files:
        kw["pos_weights"] = torch.tensor(g["in_pos_weights"][sl], dtype=torch.float32)
    return v, t, lt, kw


def _check_siglip(g, loss, dv, dt, dlt, db, rows=slice(None)):
    ref = float(g["f32_loss"])
    assert abs(loss - ref) <= 1e-5 * abs(ref), (loss, ref)
    assert _rel(dv, g["f32_dvideo"][rows]) <= 2e-3 and _rel(dt, g["f32_dtext"]) <= 2e-3
    rlt = float(g["f32_dlog_temp"].reshape(-1)[0])
    assert abs(dlt - rlt) <= 2e-3 * max(abs(rlt), 1e-4)
    if "f32_dbias" in g.files and db is not None:
        rb = float(g["f32_dbias"])
        assert abs(db - rb) <= 2e-3 * max(abs(rb), 1e-4)


@pytest.mark.parametrize("name", sorted(SIGLIP_CASES))
def test_siglip_loss_host_path_single_process(name, monkeypatch):
    loss_mod = patch_package(build_emul(), monkeypatch.setattr)
    g = np.load(GOLDEN / f"{name}.npz")
    cls, ckw = SIGLIP_CASES[name]
    mod = getattr(loss_mod, cls)(**ckw)
    v, t, lt, kw = _siglip_inputs(g)
    out = mod(video_features=v, text_features=t, log_temp=lt, **kw)
    out.backward()
    bias = getattr(mod, "bias", None)
    db = bias.grad.item() if isinstance(bias, torch.nn.Parameter) and bias.grad is not None else None
    _check_siglip(g, out.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item(), db)
    if ckw.get("entropy_regularization"):
        d = mod.get_entropy_diagnostics()                                     # the reference's seven keys, one host copy
        assert set(d) == {"entropy_mean", "entropy_min", "entropy_max", "entropy_normalized", "entropy_deficit",
                          "entropy_loss", "bce_loss"}
        assert all(np.isfinite(x) for x in d.values()) and d["entropy_min"] <= d["entropy_mean"] <= d["entropy_max"]
        assert abs(d["bce_loss"] + d["entropy_loss"] - out.item()) <= 1e-6 * abs(out.item())
    with torch.no_grad():                                                     # the no-grad (dense forward) route
        fwd_only = mod(video_features=v.detach(), text_features=t.detach(), log_temp=lt.detach(), **kw).item()
    assert abs(fwd_only - float(g["f32_loss"])) <= 1e-5 * abs(float(g["f32_loss"]))


def _siglip_body(rank, world, name):
    loss_mod = patch_package(build_emul())
    g = np.load(GOLDEN / f"{name}.npz")
    cls, ckw = SIGLIP_CASES[name]
    B = g["video"].shape[0] // world
    res = []
    # video rows / mask rows sharded, the same text on every rank: the sharded fast path (text_replicated=True: row slab per
    # rank, scalars and text gradient all-reduced) and the default reference-faithful path (video / masks gathered, the
    # [B_global, T] problem evaluated by every rank, nothing reduced) must both reproduce the full-batch golden
    for replicated in (True, False):
        mod = getattr(loss_mod, cls)(text_replicated=replicated, **ckw)
        v, t, lt, kw = _siglip_inputs(g, rank * B, (rank + 1) * B)
        loss = mod(video_features=v, text_features=t, log_temp=lt, **kw)
        loss.backward()
        res.append((loss.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item(), mod.bias.grad.item()))
    # per-rank DIFFERENT texts (what the reference's runner passes: each rank's own text batch): rank r holds half of the
    # texts and the matching half of the mask columns; reference semantics = the [B_global, T_r] problem on rank r
    T = g["text"].shape[0]
    Th = T // world
    cols = slice(rank * Th, (rank + 1) * Th)
    mod = getattr(loss_mod, cls)(**ckw)
    v, _, lt, kw = _siglip_inputs(g, rank * B, (rank + 1) * B)
    t = torch.tensor(g["text"][cols], dtype=torch.float32, requires_grad=True)
    kw = {k: x[:, cols].contiguous() for k, x in kw.items()}
    loss = mod(video_features=v, text_features=t, log_temp=lt, **kw)
    loss.backward()
    res.append((loss.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item(), mod.bias.grad.item()))
    # the fast path's contract check: different texts under text_replicated=True poison the loss (NaN), loudly
    mod = getattr(loss_mod, cls)(text_replicated=True, **ckw)
    with torch.no_grad():
        bad = mod(video_features=v.detach(), text_features=t.detach(), log_temp=lt.detach(), **kw).item()
    res.append(bad)
    return res


@pytest.mark.parametrize("name", GLOO_SIGLIP)
def test_siglip_loss_host_path_two_ranks_gloo(name, gloo_results):
    """video rows (and their mask / weight rows) sharded, text replicated; every rank returns the full loss, its
    own video-gradient rows, the FULL text gradient and the full scalar gradients."""
    world = 2
    out = {r: gloo_results[r][("siglip", name)] for r in range(world)}
    g = np.load(GOLDEN / f"{name}.npz")
    cls, ckw = SIGLIP_CASES[name]
    B = g["video"].shape[0] // world
    for mode in (0, 1):                                  # sharded fast path, reference-faithful gathered path
        for r in range(world):
            loss, dv, dt, dlt, db = out[r][mode]
            _check_siglip(g, loss, dv, dt, dlt, db, rows=slice(r * B, (r + 1) * B))
    assert out[0][0][0] == out[1][0][0]
    # per-rank texts: the oracle evaluates rank r's [B_global, T_r] problem (utils/loss/contrastive.py:252-263)
    T = g["text"].shape[0]
    Th = T // world
    okw = {k: v for k, v in ckw.items() if k not in ("entropy_regularization",)}
    # the gathered mask is the same on every rank: rank q's rows carry rank q's own column slice (its local texts)
    def gathered(key):
        if key not in g.files:
            return None
        return np.concatenate([g[key][q * B:(q + 1) * B, q * Th:(q + 1) * Th] for q in range(world)], axis=0)
    for r in range(world):
        cols = slice(r * Th, (r + 1) * Th)
        o = co.siglip_loss(g["video"], g["text"][cols], g["log_temp"], bias=ckw.get("bias_init", -10.0),
                           pos_mask=gathered("in_pos_mask"), pos_weights=gathered("in_pos_weights"),
                           entropy_regularization_on=bool(ckw.get("entropy_regularization", False)),
                           min_entropy_threshold=ckw.get("min_entropy_threshold", 2.0))
        loss, dv, dt, dlt, db = out[r][2]
        assert abs(loss - o["loss"]) <= 1e-5 * abs(o["loss"]), (r, loss, o["loss"])
        assert _rel(dv, o["dvideo"][r * B:(r + 1) * B]) <= 2e-3 and _rel(dt, o["dtext"]) <= 2e-3
        assert abs(dlt - o["dlog_temp"]) <= 2e-3 * max(abs(o["dlog_temp"]), 1e-4)
        assert abs(db - o["dbias"]) <= 2e-3 * max(abs(o["dbias"]), 1e-4)
        assert np.isnan(out[r][3])


@pytest.mark.parametrize("name", ["align_b64_d512", "align_siglip_b130_d96", "align_b300_d200"])
def test_alignment_diagnostics_end_to_end(name, monkeypatch):
    """deepcoro_clip_b200.alignment_diagnostics (logging half) through its real host code: shipped l2norm /
    dyn_prep / alignment scalar kernels under the emulation + the contract model of the forward tile kernel, against the
    transcribed runner lines (runners/video_constrative_learning_runner.py:1323-1335)."""
    patch_package(build_emul(), monkeypatch.setattr)
    from deepcoro_clip_b200.diagnostics import alignment_diagnostics
    g = np.load(GOLDEN / f"{name}.npz")
    r = alignment_diagnostics(torch.tensor(g["video"]), torch.tensor(g["text"]), torch.tensor(g["log_temp"].astype(np.float32)),
                              use_siglip=bool(g["use_siglip"]))
    for key, ref in (("alignment_cosine", "cosine_f64"), ("alignment_logprob", "logprob_f64"), ("alignment_prob", "prob_f64")):
        want = float(g[ref])
        assert abs(r[key].item() - want) <= 2e-5 * max(1.0, abs(want)), (key, r[key].item(), want)


def _legacy_body(rank, world, cls, name):
    loss_mod = patch_package(build_emul())
    g = np.load(GOLDEN / f"{name}.npz")
    B = g["video"].shape[0] // world
    v = torch.tensor(g["video"][rank * B:(rank + 1) * B], dtype=torch.float32, requires_grad=True)
    t = torch.tensor(g["text"][rank * B:(rank + 1) * B], dtype=torch.float32, requires_grad=True)
    lt = torch.tensor(g["log_temp"].astype(np.float32).reshape(1), requires_grad=True)
    mod = loss_mod.InfoNCELoss(use_ddp=True, loss_type=cls)               # the dispatcher picks the *DDP class
    loss = mod(v, t, lt)
    loss.backward()
    return (loss.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item())


@pytest.mark.parametrize("cls,name", GLOO_LEGACY)
def test_legacy_ddp_losses_two_ranks_gloo(cls, name, gloo_results):
    """utils/loss/losses.py:104-158, 213-276 through InfoNCELoss(use_ddp=True): ContrastiveLossDDP (no tau clamp) and the
    gated SiglipLossDDP, which rounds the features to fp16 before the gather (:243-244) — compared with the oracle on the
    fp16-rounded inputs."""
    from oracle import contrastive_oracle as co
    world = 2
    out = {r: gloo_results[r][("legacy", cls)] for r in range(world)}
    g = np.load(GOLDEN / f"{name}.npz")
    v, t = g["video"].astype(np.float32), g["text"].astype(np.float32)
    if cls == "siglip":
        v, t = v.astype(np.float16).astype(np.float32), t.astype(np.float16).astype(np.float32)
    o = co.clip_loss(v, t, float(g["log_temp"].reshape(-1)[0]), clamp_min=None, gated=(cls == "siglip"))
    B = v.shape[0] // world
    for r in range(world):
        loss, dv, dt, dlt = out[r]
        assert abs(loss - o["loss"]) <= 1e-5 * abs(o["loss"]), (r, loss, o["loss"])
        gtol = 2e-3 if cls == "contrastive" else 4e-3            # fp16 features: the gradient comes back rounded to fp16
        assert _rel(dv, o["dvideo"][r * B:(r + 1) * B]) <= gtol and _rel(dt, o["dtext"][r * B:(r + 1) * B]) <= gtol
        assert abs(dlt - o["dlog_temp"]) <= 2e-3 * max(abs(o["dlog_temp"]), 1e-3)
