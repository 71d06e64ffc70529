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

from tests.conftest import GOLDEN

EMUL = Path(__file__).resolve().parent / "emul"
CSRC = EMUL.parents[1] / "deepcoro_clip_b200" / "csrc"


def build_emul():
    so = EMUL / "liblossemul.so"
    srcs = [EMUL / "loss_emul.cpp", EMUL / "pool_mma_prims_emul.h", EMUL / "cuda_emul.h", CSRC / "l2norm_kernels.cuh",
            CSRC / "scalars_kernels.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-o", str(so), str(srcs[0])], check=True)
    return so


def patch_package(so):
    """Points the package at the emulated library (a plain function so that spawned gloo workers can call it too)."""
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

    ops.require_cuda = lambda *t: torch.device("cpu")
    ops.call = call
    ops.stream_ptr = lambda dev=None: 0
    return loss


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


CASES = {"clip_c1_b64_d512": ("CLIPLoss", {}), "clip_ls_b48_d96": ("CLIPLoss", {"label_smoothing": 0.1}),
         "contrastive_legacy_b32_d128": ("ContrastiveLoss", {}), "gated_siglip_legacy_b40_d128": ("SiglipLoss", {})}


def _single(name, q):
    loss_mod = patch_package(build_emul())
    g = np.load(GOLDEN / f"{name}.npz")
    cls, kw = CASES[name]
    v = torch.tensor(g["video"], dtype=torch.float32, requires_grad=True)
    t = torch.tensor(g["text"], dtype=torch.float32, requires_grad=True)
    lt = torch.tensor(g["log_temp"].astype(np.float32).reshape(1), requires_grad=True)
    out = getattr(loss_mod, cls)(**kw)(video_features=v, text_features=t, log_temp=lt)
    out.backward()
    q.put((out.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item()))


@pytest.mark.parametrize("name", sorted(CASES))
def test_clip_loss_host_path_single_process(name):
    """Run in a child process: the patch replaces module attributes of the package for the whole interpreter."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_single, args=(name, q))
    p.start()
    loss, dv, dt, dlt = q.get(timeout=600)
    p.join(60)
    g = np.load(GOLDEN / f"{name}.npz")
    ref = float(g["f32_loss"])
    assert abs(loss - ref) <= 1e-5 * abs(ref), (loss, ref)
    assert _rel(dv, g["f32_dvideo"]) <= 2e-3 and _rel(dt, g["f32_dtext"]) <= 2e-3
    rlt = float(g["f32_dlog_temp"].reshape(-1)[0])
    assert abs(dlt - rlt) <= 2e-3 * max(abs(rlt), 1e-3)


def _rank(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
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
        out[rank] = (loss.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["clip_c1_b64_d512", "clip_ls_b48_d96"])
def test_clip_loss_host_path_two_ranks_gloo(name):
    build_emul()
    world = 2
    port = 29500 + (os.getpid() % 1500)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank, args=(world, port, name, out), nprocs=world, join=True)
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
