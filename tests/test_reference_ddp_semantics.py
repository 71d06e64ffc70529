"""Build container only (imports the UNMODIFIED reference from /root/reference): pins the reading of the reference's DDP
semantics that the package implements, on two gloo ranks.

  * SigLIPLoss (utils/loss/contrastive.py:252-263): video / pos_mask / pos_weights are gathered, the text is NOT — every
    rank evaluates the [B_global, T_local] problem against ITS OWN texts; the video gradient is the own-row chunk of that
    problem's gradient (gather_with_gradient's backward, :87-91), the text gradient is the full one of the local texts.
    The oracle's single-process evaluation of that problem is what tests/test_emulated_losses.py checks the package's
    default (text_replicated=False) path against.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import contrastive_oracle as co
from tests.conftest import GOLDEN

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the build container")


def _rank(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.dont_write_bytecode = True
    sys.path.insert(0, str(REF))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from utils.loss.contrastive import SigLIPLoss
        g = np.load(GOLDEN / "siglip_mp_b32_t40_d64.npz")
        B, Th = g["video"].shape[0] // world, g["text"].shape[0] // world
        rows, cols = slice(rank * B, (rank + 1) * B), slice(rank * Th, (rank + 1) * Th)
        v = torch.tensor(g["video"][rows], dtype=torch.float32, requires_grad=True)
        t = torch.tensor(g["text"][cols], dtype=torch.float32, requires_grad=True)
        lt = torch.tensor(g["log_temp"].astype(np.float32).reshape(1), requires_grad=True)
        mod = SigLIPLoss()
        loss = mod(v, t, lt, pos_mask=torch.tensor(g["in_pos_mask"][rows, cols], dtype=torch.float32),
                   pos_weights=torch.tensor(g["in_pos_weights"][rows, cols], dtype=torch.float32))
        loss.backward()
        out[rank] = (loss.item(), v.grad.numpy(), t.grad.numpy(), lt.grad.item(), mod.bias.grad.item())
    finally:
        dist.destroy_process_group()


def test_reference_siglip_per_rank_texts_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank, args=(world, 31000 + os.getpid() % 1500, out), nprocs=world, join=True)
    g = np.load(GOLDEN / "siglip_mp_b32_t40_d64.npz")
    B, Th = g["video"].shape[0] // world, g["text"].shape[0] // world
    gathered = {k: np.concatenate([g[k][q * B:(q + 1) * B, q * Th:(q + 1) * Th] for q in range(world)], axis=0)
                for k in ("in_pos_mask", "in_pos_weights")}
    for r in range(world):
        o = co.siglip_loss(g["video"], g["text"][r * Th:(r + 1) * Th], g["log_temp"], pos_mask=gathered["in_pos_mask"],
                           pos_weights=gathered["in_pos_weights"])
        loss, dv, dt, dlt, db = out[r]
        assert abs(loss - o["loss"]) <= 2e-6 * abs(o["loss"])
        assert np.abs(dv - o["dvideo"][r * B:(r + 1) * B]).max() <= 1e-5 * np.abs(o["dvideo"]).max()
        assert np.abs(dt - o["dtext"]).max() <= 1e-5 * np.abs(o["dtext"]).max()
        assert abs(dlt - o["dlog_temp"]) <= 1e-5 * abs(o["dlog_temp"]) and abs(db - o["dbias"]) <= 1e-5 * abs(o["dbias"])
