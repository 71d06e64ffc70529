"""CPU, world_size 2, gloo: the host-side sharding plan (deepcoro_clip_b200/dist_plan.py) reproduces the reference's
DDP semantics when each rank's kernel outputs are emulated with the numpy oracle:
every rank obtains the FULL global loss; retrieval rank counts over text shards add up to the global ranks."""
import math
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import contrastive_oracle as co
from oracle import retrieval_oracle as ro


# the arithmetic of the plan, spelled out here (the package itself does it inside its kernels)
def row_slab(rank, rows_per_rank):
    return rank * rows_per_rank, (rank + 1) * rows_per_rank


def reduce_sum_(x, world_size, group=None):
    if world_size > 1:
        dist.all_reduce(x, group=group)
    return x


def clip_loss_from_sums(sum_row_lse, sum_col_lse, sum_tgt, n_global):
    """loss = 0.5 * [mean_i(r_i - tgt_i) + mean_j(c_j - tgt_j)], with sum_i tgt_i == sum_j tgt_j == sum_tgt."""
    return (0.5 / n_global) * (sum_row_lse + sum_col_lse) - sum_tgt / n_global


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepcoro_clip_b200 import dist_plan as dp
        assert dp.world() == (world, rank)
        B, D, tau = 24, 32, 0.07
        N = B * world
        rng = np.random.default_rng(0)                        # identical global data on every rank
        v = rng.standard_normal((N, D)); t = rng.standard_normal((N, D))
        lo, hi = row_slab(rank, B)
        vh, _ = co.l2_normalize(v); th, _ = co.l2_normalize(t)
        # --- what the rank's kernels would produce for its row slab (oracle stand-in) ---
        vall = dp.gather_rows(torch.tensor(vh[lo:hi]), world)
        tall = dp.gather_rows(torch.tensor(th[lo:hi]), world)
        assert np.allclose(vall.numpy(), vh) and np.allclose(tall.numpy(), th)      # rank-major concatenation
        L = vh[lo:hi] @ tall.numpy().T / tau                                         # slab logits [B, N]
        shift = 1.0 / tau
        rowsum = torch.tensor(np.exp(L - shift).sum(1))
        colsum = torch.tensor(np.exp(L - shift).sum(0))
        diag = torch.tensor([L[np.arange(B), lo + np.arange(B)].sum()])
        # --- the plan's collectives + assembly ---
        reduce_sum_(colsum, world)
        rowsum_all = dp.gather_rows(rowsum, world)
        reduce_sum_(diag, world)
        sum_r = (torch.log(rowsum_all) + shift).sum()
        sum_c = (torch.log(colsum) + shift).sum()
        loss = clip_loss_from_sums(sum_r, sum_c, diag[0], N)
        ref = co.clip_loss(v, t, math.log(tau), want_grads=False)["loss"]
        assert abs(float(loss) - ref) < 1e-10, (float(loss), ref)
        # --- retrieval: text shards, rank counts all-reduced ---
        M = 37
        tv = ro.exact_grid_embeddings(50, 16, 1); tt = ro.exact_grid_embeddings(M, 16, 2)
        tt[5] = tt[30]                                        # exact tie across the two shards
        gt = np.random.default_rng(3).integers(0, M, size=50); gt[0] = 30; gt[1] = 5
        s_lo, s_hi = dp.text_shard(M, world, rank)
        sim = ro.similarity(tv, tt)
        sg = sim[np.arange(50), gt][:, None]
        cols = np.arange(M)[None, :]
        better = (sim > sg) | ((sim == sg) & (cols < gt[:, None]))
        counts = torch.tensor(better[:, s_lo:s_hi].sum(1))
        reduce_sum_(counts, world)
        assert (counts.numpy() + 1 == ro.gt_ranks(sim, gt)).all()
        covered = dp.gather_rows(torch.tensor([[s_lo, s_hi]]), world).numpy()
        assert covered[0, 0] == 0 and covered[-1, 1] == M and (covered[1:, 0] == covered[:-1, 1]).all()
        out[rank] = float(loss)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plan():
    world = 2
    port = 29000 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world and abs(out[0] - out[1]) < 1e-12       # every rank returns the same full loss


def test_shard_helpers_single_process():
    from deepcoro_clip_b200 import dist_plan as dp
    assert dp.world() == (1, 0)
    assert row_slab(3, 128) == (384, 512)
    spans = [dp.text_shard(32473, 8, r) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 32473 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert dp.text_shard(3, 8, 7) == (3, 3)                # more ranks than rows: empty shard
    x = torch.arange(6).view(3, 2)
    assert dp.gather_rows(x, 1) is x


def _worker_siglip_pieces(rank, world, port, out):
    """Host pieces of the SigLIP variants under DDP: the text-row gather of SigLIP2BCELossDDP (forward = rank-major
    concatenation, backward = this rank's chunk without a reduce, utils/loss/siglip2_bce.py:194-224) and the assembly of
    the entropy regulariser's global statistics from per-rank {sum, min, max} triples."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepcoro_clip_b200.loss import _GatherTextRows
        B, D = 5, 4
        x = (torch.arange(B * D, dtype=torch.float32).view(B, D) + 100 * rank).requires_grad_(True)
        full = _GatherTextRows.apply(x, None)
        assert full.shape == (world * B, D)
        for r in range(world):
            assert torch.equal(full[r * B:(r + 1) * B], torch.arange(B * D, dtype=torch.float32).view(B, D) + 100 * r)
        w = torch.arange(world * B * D, dtype=torch.float32).view(world * B, D)
        (full * w).sum().backward()
        assert torch.equal(x.grad, w[rank * B:(rank + 1) * B])            # own chunk, no reduction, no 1/W
        # entropy statistics: every rank contributes (sum_i H_i, min_i H_i, max_i H_i) of its rows; the coefficient
        # kernel consumes the [W, 3] gather (loss.py: stats_all)
        rng = np.random.default_rng(7)
        H = rng.random(world * B) * 3.0                                    # global per-row entropies
        mine = H[rank * B:(rank + 1) * B]
        stats = torch.tensor([mine.sum(), mine.min(), mine.max()], dtype=torch.float64)
        stats_all = torch.empty(world * 3, dtype=torch.float64)
        dist.all_gather_into_tensor(stats_all, stats)
        stats_all = stats_all.view(world, 3)
        mean = float(stats_all[:, 0].sum() / (world * B))
        assert abs(mean - H.mean()) < 1e-12
        assert float(stats_all[:, 1].min()) == H.min() and float(stats_all[:, 2].max()) == H.max()
        out[rank] = mean
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_siglip_variant_pieces():
    world = 2
    port = 31000 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_siglip_pieces, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world and out[0] == out[1]


def _worker_ragged_gather(rank, world, port, out):
    """the ragged cross-rank gather of validation embeddings equals the reference's
    _gather_tensor_along_batch semantics (rank-major concatenation of the valid rows), 1-D and 2-D, empty ranks too."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepcoro_clip_b200.embedding_store import EmbeddingStore, gather_tensor_along_batch
        sizes = [5, 3]
        full = torch.arange(sum(sizes) * 4, dtype=torch.float32).view(sum(sizes), 4)
        lo = sum(sizes[:rank])
        mine = full[lo:lo + sizes[rank]]
        assert torch.equal(gather_tensor_along_batch(mine), full)
        assert torch.equal(gather_tensor_along_batch(mine[:, 0].contiguous()), full[:, 0])
        # equal sizes: no compaction step; an empty rank
        assert torch.equal(gather_tensor_along_batch(full[3 * rank:3 * rank + 3]), full[:6])
        e = gather_tensor_along_batch(full[:4] if rank == 0 else full[:0])
        assert torch.equal(e, full[:4])
        store = EmbeddingStore(dim=4, capacity=2, device="cpu")
        for i in range(0, sizes[rank], 2):                      # appended batch by batch, buffer grows
            store.append(mine[i:i + 2])
        assert store.n == sizes[rank] and torch.equal(store.local(), mine)
        assert torch.equal(store.gather(), full)
        store.reset()
        assert store.local().shape == (0, 4)
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_ragged_embedding_gather():
    world = 2
    port = 33000 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_ragged_gather, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
