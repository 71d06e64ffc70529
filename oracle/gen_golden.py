"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (imported from /root/reference,
torch CPU) on seeded synthetic inputs. Run in the build container only:

    python oracle/gen_golden.py

The GPU box has no /root/reference; it checks against these committed vectors instead.
"""
from __future__ import annotations

import math
import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

from utils.loss.contrastive import CLIPLoss, SigLIPLoss  # noqa: E402
from utils.loss.losses import ContrastiveLoss, SiglipLoss  # noqa: E402
from utils.loss.siglip_pairwise import SiglipPairwiseFeatureLoss  # noqa: E402
from utils.loss.siglip2_bce import SigLIP2BCELoss, SigLIP2MultiPositiveBCELoss  # noqa: E402
from utils.retrieval_metrics_streaming import compute_metrics_streaming, compute_recall_at_k_streaming  # noqa: E402
from models.rope_3d import Rope3D  # noqa: E402
from models.attention_pool import AttentionPool, AttentionPoolWithCLS  # noqa: E402
from models.video_aggregator import EnhancedVideoAggregator  # noqa: E402


def _np(t):
    return t.detach().cpu().double().numpy() if t.dtype != torch.bfloat16 else t.detach().float().numpy()


def run_loss(mod, v, t, lt, dtype, **kw):
    v = v.to(dtype).clone().requires_grad_(True)
    t = t.to(dtype).clone().requires_grad_(True)
    lt = lt.to(dtype).clone().requires_grad_(True)
    kw = {k: (x.to(dtype) if isinstance(x, torch.Tensor) else x) for k, x in kw.items()}
    if dtype == torch.float64:
        mod = mod.double()
    loss = mod(v, t, lt, **kw)
    loss.backward()
    out = dict(loss=_np(loss), dvideo=_np(v.grad), dtext=_np(t.grad), dlog_temp=_np(lt.grad))
    if hasattr(mod, "bias") and isinstance(mod.bias, torch.nn.Parameter) and mod.bias.grad is not None:
        out["dbias"] = _np(mod.bias.grad)
    return out


def save_loss_case(name, mod_fn, B, T, D, seed, log_temp, extra=None, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, D, generator=g) * scale
    t = torch.randn(T, D, generator=g) * scale
    lt = torch.tensor([log_temp])
    kw = extra(g) if extra else {}
    rec = dict(video=v.numpy(), text=t.numpy(), log_temp=np.array([log_temp]))
    for k, x in kw.items():
        rec["in_" + k] = x.numpy()
    # torch's CPU float64 autocast guard: the reference casts to .float() internally, so the "fp64" run is
    # only fp64 where the module does not force fp32; we therefore also record the fp32 run (the real reference).
    r32 = run_loss(mod_fn(), v, t, lt, torch.float32, **kw)
    for k, x in r32.items():
        rec["f32_" + k] = x
    np.savez_compressed(OUT / f"{name}.npz", **rec)
    print(name, "loss", float(r32["loss"]))


def losses():
    save_loss_case("clip_c1_b64_d512", lambda: CLIPLoss(), 64, 64, 512, 0, math.log(0.07))
    save_loss_case("clip_ls_b48_d96", lambda: CLIPLoss(label_smoothing=0.1), 48, 48, 96, 11, math.log(0.0588))
    save_loss_case("clip_clamp_b8_d64", lambda: CLIPLoss(), 8, 8, 64, 12, math.log(5e-5))
    save_loss_case("clip_b300_d200", lambda: CLIPLoss(), 300, 300, 200, 13, math.log(0.0588))
    save_loss_case("contrastive_legacy_b32_d128", lambda: ContrastiveLoss(), 32, 32, 128, 14, math.log(0.1))
    save_loss_case("gated_siglip_legacy_b40_d128", lambda: SiglipLoss(), 40, 40, 128, 15, math.log(0.1))

    def mp(B, T, npos, weights=True):
        def f(g):
            m = torch.zeros(B, T)
            idx = torch.arange(B) % T
            m[torch.arange(B), idx] = 1.0
            for _ in range(npos - 1):
                m[torch.arange(B), torch.randint(0, T, (B,), generator=g)] = 1.0
            out = {"pos_mask": m}
            if weights:
                choices = torch.tensor([1.0, 1.5, 2.5, 3.0])
                out["pos_weights"] = m * choices[torch.randint(0, 4, (B, T), generator=g)]
            return out
        return f

    save_loss_case("siglip_diag_b32_t32_d64", lambda: SigLIPLoss(), 32, 32, 64, 20, math.log(0.087))
    save_loss_case("siglip_mp_b32_t40_d64", lambda: SigLIPLoss(), 32, 40, 64, 21, math.log(0.087), mp(32, 40, 4))
    save_loss_case("siglip_mp_noweights_b24_t50_d96", lambda: SigLIPLoss(positive_weight=2.0, negative_weight=0.5,
                                                                         use_severity_weights=False),
                   24, 50, 96, 22, math.log(0.07), mp(24, 50, 3))
    save_loss_case("siglip_autobalance_b16_t48_d64", lambda: SigLIPLoss(auto_balance=True), 16, 48, 64, 23,
                   math.log(0.1), mp(16, 48, 3, weights=False))
    save_loss_case("siglip_entropy_b16_t32_d64", lambda: SigLIPLoss(entropy_regularization=True, bias_init=-2.0,
                                                                    min_entropy_threshold=5.0),
                   16, 32, 64, 24, math.log(0.05), mp(16, 32, 2))
    save_loss_case("siglip_bias0_b130_t260_d512", lambda: SigLIPLoss(bias_init=-1.0), 130, 260, 512, 25,
                   math.log(0.087), mp(130, 260, 4))


def siglip_variants():
    """Rows a6 of SURVEY §8: the SigLIP classes the reference keeps importable next to the unified SigLIPLoss."""
    def mp(B, T, npos, weights=True):
        def f(g):
            m = torch.zeros(B, T)
            m[torch.arange(B), torch.arange(B) % T] = 1.0
            for _ in range(npos - 1):
                m[torch.arange(B), torch.randint(0, T, (B,), generator=g)] = 1.0
            out = {"pos_mask": m}
            if weights:
                choices = torch.tensor([1.0, 1.5, 2.5, 3.0])
                out["pos_weights"] = m * choices[torch.randint(0, 4, (B, T), generator=g)]
            return out
        return f

    save_loss_case("pairwise_mp_b24_t40_d64", lambda: SiglipPairwiseFeatureLoss(positive_weight=1.5, negative_weight=0.7),
                   24, 40, 64, 30, math.log(0.087), mp(24, 40, 3))
    save_loss_case("pairwise_entropy_auto_b16_t48_d64",
                   lambda: SiglipPairwiseFeatureLoss(auto_positive_weight=True, entropy_regularization=True,
                                                     entropy_weight=0.2, min_entropy_threshold=6.0),
                   16, 48, 64, 31, math.log(0.06), mp(16, 48, 2, weights=False))
    save_loss_case("bce2_b40_d96", lambda: SigLIP2BCELoss(), 40, 40, 96, 32, math.log(0.1))
    save_loss_case("bce2_ls_noclamp_b32_d64", lambda: SigLIP2BCELoss(bias_init=-4.0, label_smoothing=0.1), 32, 32, 64, 33,
                   math.log(0.02))          # |S / tau + b| reaches ~50: the variant must NOT clamp at 30
    save_loss_case("mp2_ls_b20_t36_d64", lambda: SigLIP2MultiPositiveBCELoss(bias_init=-3.0, positive_weight=2.0,
                                                                            negative_weight=0.5, label_smoothing=0.2),
                   20, 36, 64, 34, math.log(0.08), mp(20, 36, 3))
    save_loss_case("mp2_diag_b18_t30_d64", lambda: SigLIP2MultiPositiveBCELoss(bias_init=-5.0), 18, 30, 64, 35,
                   math.log(0.04))          # clamp at +-30 active on part of the matrix


def multipos():
    """the logits-level multi-positive softmax losses (fp32 autograd through the reference classes)."""
    from utils.loss.multi_positive_infonce import MultiPositiveInfoNCELoss
    from utils.loss.weighted_siglip import WeightedSigLIPLoss
    for name, N, M, seed in (("multipos_48x64", 48, 64, 50), ("multipos_130x37", 130, 37, 51)):
        g = torch.Generator().manual_seed(seed)
        logits = torch.randn(N, M, generator=g) * 4.0
        mask = (torch.rand(N, M, generator=g) < 0.06).float()
        mask[torch.arange(min(N, M)), torch.arange(min(N, M))] = 1.0
        mask[N // 2] = 0.0                                    # a row without positives
        mask[:, M // 3] = 0.0                                 # a column without positives
        pw = torch.tensor([1.0, 1.5, 2.5, 3.0])[torch.randint(0, 4, (N, M), generator=g)]
        rec = dict(logits=logits.numpy(), mask=mask.numpy(), pos_weights=pw.numpy())
        for key, fn in (("wsl", lambda L: WeightedSigLIPLoss()(L, mask * pw - 0.2 * (1 - mask))),      # negatives clamp to 0
                        ("mpi_mean", lambda L: MultiPositiveInfoNCELoss()(L, mask, pw)),
                        ("mpi_sum_noweights", lambda L: MultiPositiveInfoNCELoss(reduction="sum")(L, mask)),
                        # use_importance_weighting (:57-93): rows / columns weighted by their summed raw pos_weights / mask
                        ("mpi_imp_mean", lambda L: MultiPositiveInfoNCELoss(use_importance_weighting=True)(L, mask, pw)),
                        ("mpi_imp_sum_noweights", lambda L: MultiPositiveInfoNCELoss(
                            reduction="sum", use_importance_weighting=True)(L, mask))):
            L = logits.clone().requires_grad_(True)
            loss = fn(L)
            loss.backward()
            rec[key + "_loss"] = _np(loss)
            rec[key + "_dlogits"] = _np(L.grad)
        np.savez_compressed(OUT / f"{name}.npz", **rec)
        print(name, {k: float(v) for k, v in rec.items() if k.endswith("_loss")})


def tokenmean():
    """the mean branch of VideoEncoder._pool_video_tokens (models/video_encoder.py:589-603), called as
    the UNBOUND reference method on a stub whose attention_pool is None (the encoder itself needs backbone weights)."""
    import types
    from models.video_encoder import VideoEncoder
    for name, B, N, L, D, seed in (("tokenmean_b2_n3_l50_d128", 2, 3, 50, 128, 70), ("tokenmean_b1_n2_l197_d256", 1, 2, 197, 256, 71)):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, N, L, D, generator=g, requires_grad=True)
        dout = torch.randn(B, N, D, generator=g)
        y = VideoEncoder._pool_video_tokens(types.SimpleNamespace(attention_pool=None), x)
        y.backward(dout)
        np.savez_compressed(OUT / f"{name}.npz", x=x.detach().numpy(), dout=dout.numpy(), out=_np(y), dx=_np(x.grad))
        print(name, float(y.abs().mean()))


def alignment():
    """SURVEY §8f #2 (logging half): the runner's per-step alignment diagnostics. The block is inline in the train step
    (runners/video_constrative_learning_runner.py:1323-1335) and the runner does not import outside the training
    stack, so the lines are transcribed here with the same torch calls, in fp32 (as the runner runs them) and fp64."""
    import torch.nn.functional as F
    for name, B, D, seed, lt, use_siglip in (("align_b64_d512", 64, 512, 60, math.log(0.07), False),
                                             ("align_siglip_b130_d96", 130, 96, 61, math.log(0.087), True),
                                             ("align_b300_d200", 300, 200, 62, math.log(0.05), False)):
        g = torch.Generator().manual_seed(seed)
        text = torch.randn(B, D, generator=g)
        video = 0.6 * text + torch.randn(B, D, generator=g)            # aligned pairs, as after some training
        rec = dict(video=video.numpy(), text=text.numpy(), log_temp=np.array([lt]), use_siglip=np.array(use_siglip))
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            video_emb, text_emb, log_temp = video.to(dt), text.to(dt), torch.tensor([lt], dtype=dt)
            video_norm = F.normalize(video_emb, dim=1)
            text_norm = F.normalize(text_emb, dim=1)
            similarity = torch.matmul(video_norm, text_norm.t())
            alignment_cosine_tensor = torch.diag(similarity).mean()
            logits_base = similarity
            if use_siglip:
                logits_base = similarity * torch.sigmoid(similarity)
            tau_value = torch.exp(log_temp.float() if dt == torch.float32 else log_temp)
            logits_matrix = logits_base / tau_value
            logprob_matrix = F.log_softmax(logits_matrix, dim=1)
            alignment_logprob_tensor = torch.diag(logprob_matrix).mean()
            alignment_prob_tensor = alignment_logprob_tensor.exp()
            rec["cosine_" + tag] = _np(alignment_cosine_tensor)
            rec["logprob_" + tag] = _np(alignment_logprob_tensor)
            rec["prob_" + tag] = _np(alignment_prob_tensor)
        np.savez_compressed(OUT / f"{name}.npz", **rec)
        print(name, float(rec["cosine_f64"]), float(rec["logprob_f64"]), float(rec["prob_f64"]))


def inline_multipos():
    """SURVEY §8f #2 (loss half): the runner's inline multi-positive branch. The block is inline in the train step
    (runners/video_constrative_learning_runner.py:1256-1322) and the runner does not import outside the training stack, so the
    lines are transcribed here with the same torch calls (the WeightedSigLIPLoss it calls IS the imported reference class), in
    fp32 (as the runner runs them) and fp64, with autograd for the gradients."""
    import torch.nn.functional as F
    from utils.loss.weighted_siglip import WeightedSigLIPLoss
    cases = (("inline_mp_weighted_b48_m64_d128", 48, 64, 128, 70, math.log(0.1), True, True, 0.0, 1.0),
             ("inline_mp_weighted_noweights_b33_m20_d96", 33, 20, 96, 71, math.log(0.07), True, False, 0.0, 1.0),
             ("inline_mp_weighted_margin_b40_m56_d512", 40, 56, 512, 72, math.log(0.087), True, True, 0.5, 1.0),
             ("inline_mp_bce_b48_m64_d128", 48, 64, 128, 73, math.log(0.1), False, True, 0.0, 0.5),
             ("inline_mp_bce_noweights_b20_m35_d200", 20, 35, 200, 74, math.log(0.2), False, False, 0.25, 1.0))
    for name, B, M, D, seed, lt, weighted, with_pw, margin, neg_w in cases:
        g = torch.Generator().manual_seed(seed)
        text = torch.randn(M, D, generator=g)
        owner = torch.randint(0, M, (B,), generator=g)
        video = 0.7 * text[owner] + torch.randn(B, D, generator=g)
        targets = torch.zeros(B, M)
        targets[torch.arange(B), owner] = 1.0
        for _ in range(2):
            targets[torch.arange(B), torch.randint(0, M, (B,), generator=g)] = 1.0
        targets[B - 1] = 0.0                                           # one row without any positive
        pw = None
        if with_pw:
            pw = targets * torch.tensor([1.0, 1.5, 2.5, 3.0])[torch.randint(0, 4, (B, M), generator=g)]
        abn = (torch.rand(M, generator=g) < 0.3).float() if margin > 0 else None
        rec = dict(video=video.numpy(), text=text.numpy(), log_temp=np.array([lt]), targets=targets.numpy(),
                   weighted=np.array(weighted), margin=np.array(margin), neg_weight=np.array(neg_w))
        if pw is not None:
            rec["pos_weights"] = pw.numpy()
        if abn is not None:
            rec["abnormal"] = abn.numpy()
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            video_emb = video.to(dt).clone().requires_grad_(True)
            text_emb = text.to(dt).clone().requires_grad_(True)
            log_temp = torch.tensor([lt], dtype=dt, requires_grad=True)
            positive_mask = targets.to(dt)
            positive_weights = pw.to(dt) if pw is not None else None
            # ---- runner lines :1256-1303 ----
            tg = positive_mask
            video_norm = F.normalize(video_emb, dim=1)
            text_norm = F.normalize(text_emb, dim=1)
            similarity = torch.matmul(video_norm, text_norm.t())
            gated = similarity * torch.sigmoid(similarity)
            temp_value = torch.exp(log_temp.float() if dt == torch.float32 else log_temp)
            logits_matrix = gated / temp_value
            if margin > 0 and abn is not None:
                abnormal_vector = abn.to(dt)
                if torch.count_nonzero(abnormal_vector) > 0:
                    logits_matrix = logits_matrix + abnormal_vector.unsqueeze(0) * margin
            if weighted:
                pos_weights_for_loss = tg
                if positive_weights is not None and torch.count_nonzero(positive_weights) > 0:
                    pos_weights_for_loss = tg * positive_weights
                clip_loss = WeightedSigLIPLoss()(logits=logits_matrix, positive_weights=pos_weights_for_loss)
            else:
                weight_matrix = torch.full_like(tg, neg_w)
                if positive_weights is not None:
                    weight_matrix = torch.where(tg > 0, positive_weights, weight_matrix)
                loss_sum = F.binary_cross_entropy_with_logits(logits_matrix, tg, weight=weight_matrix, reduction='sum')
                denom = max(1.0, float(tg.sum().item()))
                clip_loss = loss_sum / denom
            # ---- runner lines :1305-1322 ----
            with torch.no_grad():
                logprob_matrix = F.log_softmax(logits_matrix, dim=1)
                if positive_weights is not None and torch.count_nonzero(positive_weights) > 0:
                    pos_weight_mask = positive_weights * tg
                else:
                    pos_weight_mask = tg
                row_sums = pos_weight_mask.sum(dim=1)
                valid_rows = row_sums > 0
                normalized_weights = torch.zeros_like(pos_weight_mask)
                normalized_weights[valid_rows] = pos_weight_mask[valid_rows] / row_sums[valid_rows].unsqueeze(1)
                alignment_logprob_tensor = (logprob_matrix * normalized_weights).sum(dim=1)[valid_rows].mean()
                alignment_prob_tensor = alignment_logprob_tensor.exp()
                alignment_cosine_tensor = similarity[tg.bool()].mean()
            clip_loss.backward()
            rec["loss_" + tag] = _np(clip_loss)
            rec["dvideo_" + tag] = _np(video_emb.grad)
            rec["dtext_" + tag] = _np(text_emb.grad)
            rec["dlog_temp_" + tag] = _np(log_temp.grad)
            rec["logprob_" + tag] = _np(alignment_logprob_tensor)
            rec["prob_" + tag] = _np(alignment_prob_tensor)
            rec["cosine_" + tag] = _np(alignment_cosine_tensor)
        np.savez_compressed(OUT / f"{name}.npz", **rec)
        print(name, float(rec["loss_f64"]), float(rec["logprob_f64"]), float(rec["cosine_f64"]))


def dense_metrics():
    """utils/retrieval_metrics.py on tie-free Gaussian similarities with multi-label ground truth."""
    from utils.retrieval_metrics import (compute_map, compute_median_rank, compute_mrr, compute_ndcg_at_k,
                                         compute_recall_at_k, compute_similarity_matrix)
    for name, N, M, gmax, seed in (("dense_metrics_120x90_g1", 120, 90, 1, 40), ("dense_metrics_200x300_g4", 200, 300, 4, 41),
                                   ("dense_metrics_64x7_g3", 64, 7, 3, 42)):
        g = torch.Generator().manual_seed(seed)
        v = torch.randn(N, 48, generator=g)
        t = torch.randn(M, 48, generator=g)
        sim = compute_similarity_matrix(v, t)
        assert all(len(set(r.tolist())) == M for r in sim)                      # tie-free rows
        gt = torch.full((N, gmax), -1, dtype=torch.int64)
        for i in range(N):
            n_i = int(torch.randint(0 if gmax > 1 else 1, gmax + 1, (1,), generator=g))
            gt[i, :n_i] = torch.randperm(M, generator=g)[:n_i]
        gt_arg = gt[:, 0] if gmax == 1 else gt
        ks = [1, 5, 10]
        rec = compute_recall_at_k(sim, gt_arg, ks)
        nd = compute_ndcg_at_k(sim, gt_arg, ks)
        rec_v = np.array([rec[f"Recall@{k}"] for k in ks])
        nd_v = np.array([nd[f"NDCG@{k}_V2T"] for k in ks])
        np.savez_compressed(OUT / f"{name}.npz", video=v.numpy(), text=t.numpy(), sim=sim.numpy(), gt=gt.numpy(),
                            k_values=np.array(ks), recall=rec_v, ndcg=nd_v, mrr=np.array(compute_mrr(sim, gt_arg)["MRR_V2T"]),
                            map=np.array(compute_map(sim, gt_arg)), median_rank=np.array(compute_median_rank(sim, gt_arg)))
        print(name, rec, nd)


def retrieval():
    g = torch.Generator().manual_seed(30)
    v = torch.randn(300, 64, generator=g)
    t = torch.randn(200, 64, generator=g)
    gt = torch.randint(0, 200, (300,), generator=g)
    m = compute_metrics_streaming(v, t, gt, k_values=[1, 5, 10, 50], video_chunk_size=128, text_chunk_size=64,
                                  device="cpu")
    np.savez_compressed(OUT / "retrieval_gauss_300x200.npz", video=v.numpy(), text=t.numpy(), gt=gt.numpy(),
                        keys=np.array(list(m.keys())), values=np.array([float(x) for x in m.values()]))
    # exact-grid inputs (tie free by construction check) through the recall-only API (no normalisation)
    rng = np.random.default_rng(31)
    ve = (rng.integers(-127, 128, size=(257, 64)) / 128.0).astype(np.float32)
    te = (rng.integers(-127, 128, size=(300, 64)) / 128.0).astype(np.float32)
    gte = rng.integers(0, 300, size=257)
    r = compute_recall_at_k_streaming(torch.from_numpy(ve), torch.from_numpy(te), torch.from_numpy(gte),
                                      k_values=[1, 5, 10], video_chunk_size=100, text_chunk_size=77, device="cpu")
    sim = torch.from_numpy(ve) @ torch.from_numpy(te).t()
    top = torch.topk(sim, 10, dim=1)
    np.savez_compressed(OUT / "retrieval_grid_257x300.npz", video=ve, text=te, gt=gte,
                        keys=np.array(list(r.keys())), values=np.array([float(x) for x in r.values()]),
                        topk_idx=top.indices.numpy(), topk_val=top.values.numpy())
    # known-answer cases of tests/test_retrieval_metrics.py re-expressed for the streaming API
    eye = torch.eye(5)
    m1 = compute_metrics_streaming(eye, eye, torch.arange(5), k_values=[1, 3, 5], device="cpu")
    anti = torch.flip(torch.eye(5), dims=[1])
    m2 = compute_metrics_streaming(eye, anti, torch.arange(5), k_values=[1], device="cpu")
    np.savez_compressed(OUT / "retrieval_known.npz", eye_keys=np.array(list(m1.keys())),
                        eye_values=np.array([float(x) for x in m1.values()]), anti_keys=np.array(list(m2.keys())),
                        anti_values=np.array([float(x) for x in m2.values()]))
    print("retrieval", m, r, m1, m2)


def rope():
    for name, (B, heads, T, H, W, cls, dtype) in {
        "rope_f32_t2h3w4_cls": (2, 4, 2, 3, 4, 1, torch.float32),
        "rope_f32_t3h2w2": (1, 8, 3, 2, 2, 0, torch.float32),
        "rope_bf16_t4h7w7_cls": (2, 4, 4, 7, 7, 1, torch.bfloat16),
    }.items():
        g = torch.Generator().manual_seed(40)
        N = T * H * W + cls
        q = torch.randn(B, heads, N, 96, generator=g).to(dtype).requires_grad_(True)
        k = torch.randn(B, heads, N, 96, generator=g).to(dtype).requires_grad_(True)
        mod = Rope3D(embed_dim=96 * heads, num_heads=heads).eval()
        qr, kr = mod(q, k, T, H, W)
        gq = torch.randn(B, heads, N, 96, generator=g).to(dtype)
        gk = torch.randn(B, heads, N, 96, generator=g).to(dtype)
        (qr * gq).sum().backward(retain_graph=True)
        (kr * gk).sum().backward()
        sin, cos = mod._get_cached_freqs(T, H, W, q.device, dtype, cls)
        np.savez_compressed(OUT / f"{name}.npz", q=_np(q), k=_np(k), gq=_np(gq), gk=_np(gk), q_rot=_np(qr),
                            k_rot=_np(kr), dq=_np(q.grad), dk=_np(k.grad), sin=_np(sin), cos=_np(cos),
                            meta=np.array([B, heads, T, H, W, cls]))
        print(name, tuple(qr.shape))
    # mismatch branch: N != THW (+1) -> unchanged
    mod = Rope3D(embed_dim=96 * 2, num_heads=2).eval()
    q = torch.randn(1, 2, 7, 96)
    qr, kr = mod(q, q, 2, 2, 2)
    assert qr is q and kr is q


def attnpool():
    for name, (B, N, D, heads, out_dim, use_mask) in {
        "attnpool_b3_n50_d128_h8": (3, 50, 128, 8, None, False),
        "attnpool_b4_n37_d256_h4_mask_proj": (4, 37, 256, 4, 32, True),
    }.items():
        torch.manual_seed(50)
        mod = AttentionPool(D, num_heads=heads, output_dim=out_dim, dropout=0.0).double()
        # default init leaves in_proj_bias / out_proj.bias at 0: randomise so every term is exercised
        with torch.no_grad():
            mod.attn.in_proj_bias.normal_(std=0.3)
            mod.attn.out_proj.bias.normal_(std=0.3)
            mod.norm.weight.normal_(mean=1.0, std=0.2)
            mod.norm.bias.normal_(std=0.2)
            mod.query.normal_(std=0.5)
        g = torch.Generator().manual_seed(51)
        x = torch.randn(B, N, D, generator=g).double().requires_grad_(True)
        mask = None
        if use_mask:
            mask = torch.rand(B, N, generator=g) < 0.2
            mask[:, 0] = False
        out = mod(x, mask)
        go = torch.randn(out.shape, generator=g).double()
        (out * go).sum().backward()
        rec = dict(x=_np(x), go=_np(go), out=_np(out), dx=_np(x.grad), heads=np.array(heads),
                   mask=np.zeros((B, N), bool) if mask is None else mask.numpy(), has_mask=np.array(use_mask))
        names = {"query": mod.query, "in_proj_weight": mod.attn.in_proj_weight, "in_proj_bias": mod.attn.in_proj_bias,
                 "out_proj_weight": mod.attn.out_proj.weight, "out_proj_bias": mod.attn.out_proj.bias,
                 "norm_weight": mod.norm.weight, "norm_bias": mod.norm.bias}
        if out_dim is not None:
            names["proj_weight"] = mod.proj.weight
            names["proj_bias"] = mod.proj.bias
        for k, p in names.items():
            rec["p_" + k] = _np(p)
            rec["g_" + k] = _np(p.grad)
        np.savez_compressed(OUT / f"{name}.npz", **rec)
        print(name, tuple(out.shape))


CLS_POOL_PARAMS = {"cls_token": "cls_token", "in_proj_weight": "transformer.layers.0.self_attn.in_proj_weight",
                   "in_proj_bias": "transformer.layers.0.self_attn.in_proj_bias",
                   "out_proj_weight": "transformer.layers.0.self_attn.out_proj.weight",
                   "out_proj_bias": "transformer.layers.0.self_attn.out_proj.bias",
                   "linear1_weight": "transformer.layers.0.linear1.weight", "linear1_bias": "transformer.layers.0.linear1.bias",
                   "linear2_weight": "transformer.layers.0.linear2.weight", "linear2_bias": "transformer.layers.0.linear2.bias",
                   "norm1_weight": "transformer.layers.0.norm1.weight", "norm1_bias": "transformer.layers.0.norm1.bias",
                   "norm2_weight": "transformer.layers.0.norm2.weight", "norm2_bias": "transformer.layers.0.norm2.bias",
                   "norm_weight": "norm.weight", "norm_bias": "norm.bias", "proj_weight": "proj.weight", "proj_bias": "proj.bias"}


def clspool():
    """AttentionPoolWithCLS (models/attention_pool.py:104-197), fp64, train() so that autograd takes the plain
    (non-fused) nn.TransformerEncoderLayer path; dropout = 0 makes train() and eval() the same function."""
    for name, (B, N, D, heads, out_dim, use_mask) in {
        "clspool_b3_n50_d128_h8": (3, 50, 128, 8, None, False),
        "clspool_b4_n37_d128_h4_mask_proj": (4, 37, 128, 4, 32, True),
    }.items():
        torch.manual_seed(70)
        mod = AttentionPoolWithCLS(D, num_heads=heads, output_dim=out_dim, dropout=0.0).double()
        with torch.no_grad():       # default init zeroes the biases: randomise so every term is exercised
            for k, p in mod.named_parameters():
                if k.endswith("bias"):
                    p.normal_(std=0.3)
                elif "norm" in k:
                    p.normal_(mean=1.0, std=0.2)
            mod.cls_token.normal_(std=0.5)
            for k, p in mod.named_parameters():     # the two [2048, D] matrices are stored as fp16: make that lossless
                if "linear" in k and k.endswith("weight"):
                    p.copy_(p.half().double())
        g = torch.Generator().manual_seed(71)
        x = torch.randn(B, N, D, generator=g).double().requires_grad_(True)
        mask = None
        if use_mask:
            mask = torch.rand(B, N, generator=g) < 0.2
            mask[:, 0] = False
            mask[1, :] = True          # one sample with every token masked: the CLS key attends to itself only
        out = mod(x, mask)
        go = torch.randn(out.shape, generator=g).double()
        (out * go).sum().backward()
        rec = dict(x=_np(x), go=_np(go), out=_np(out), dx=_np(x.grad), heads=np.array(heads),
                   mask=np.zeros((B, N), bool) if mask is None else mask.numpy(), has_mask=np.array(use_mask))
        named = dict(mod.named_parameters())
        for k, full in CLS_POOL_PARAMS.items():
            if full in named:
                rec["p_" + k] = _np(named[full])
                rec["g_" + k] = _np(named[full].grad)
        for k in ("linear1_weight", "linear2_weight"):          # keep the fixture small: fp16 values (exact, see above),
            rec["p_" + k] = rec["p_" + k].astype(np.float16)    # gradient rows / columns subsampled by 16 along the
        rec["g_linear1_weight"] = rec["g_linear1_weight"][::16]  # 2048-wide hidden dimension
        rec["g_linear2_weight"] = rec["g_linear2_weight"][:, ::16]
        np.savez_compressed(OUT / f"{name}.npz", **rec)
        print(name, tuple(out.shape))


def qpool():
    for name, (B, N, D, use_mask) in {"qpool_b5_n4_d64": (5, 4, 64, False), "qpool_b6_n5_d128_mask": (6, 5, 128, True)}.items():
        torch.manual_seed(60)
        mod = EnhancedVideoAggregator(D, num_heads=4, dropout=0.0, aggregator_depth=0, max_segments=16).double()
        with torch.no_grad():
            mod.final_ln.weight.normal_(mean=1.0, std=0.2)
            mod.final_ln.bias.normal_(std=0.2)
            mod.attn_query.normal_(std=0.5)
            mod.pos_encoding.normal_(std=0.3)
        g = torch.Generator().manual_seed(61)
        x = torch.randn(B, N, D, generator=g).double().requires_grad_(True)
        mask = None
        if use_mask:
            mask = torch.rand(B, N, generator=g) < 0.3
            mask[0] = False
            mask[1] = True          # all-masked study: uniform-over-valid fallback => zeros
        out = mod(x, mask)
        go = torch.randn(out.shape, generator=g).double()
        (out * go).sum().backward()
        np.savez_compressed(OUT / f"{name}.npz", x=_np(x), go=_np(go), out=_np(out), dx=_np(x.grad),
                            mask=np.zeros((B, N), bool) if mask is None else mask.numpy(), has_mask=np.array(use_mask),
                            pos=_np(mod.pos_encoding), ln_w=_np(mod.final_ln.weight), ln_b=_np(mod.final_ln.bias),
                            attn_query=_np(mod.attn_query), g_pos=_np(mod.pos_encoding.grad),
                            g_ln_w=_np(mod.final_ln.weight.grad), g_ln_b=_np(mod.final_ln.bias.grad),
                            g_attn_query=_np(mod.attn_query.grad))
        print(name, tuple(out.shape))


def milpool():
    """MultiInstanceLinearProbing(pooling_mode="attention")._pool_instances on [B, N, D] and [B, N, L, D] inputs."""
    from models.multi_instance_linear_probing import MultiInstanceLinearProbing
    cases = {"milpool_b5_n6_d64_h32_mask": ((5, 6, 64), 32, True), "milpool_b3_n4_d128_h128": ((3, 4, 128), 128, False),
             "milpool_b2_n3_l20_d64_h24_mask": ((2, 3, 20, 64), 24, True)}
    for name, (shape, hd, use_mask) in cases.items():
        torch.manual_seed(70)
        mod = MultiInstanceLinearProbing(shape[-1], {"head": 3}, pooling_mode="attention", attention_hidden=hd).double()
        with torch.no_grad():
            for lin in (mod.attention_V, mod.attention_U, mod.attention_w):
                lin.bias.normal_(std=0.3)
            mod.attention_w.weight.mul_(3.0)
        g = torch.Generator().manual_seed(71)
        x = torch.randn(*shape, generator=g).double().requires_grad_(True)
        B, N = shape[:2]
        mask = None
        if use_mask:
            mask = torch.rand(B, N, generator=g) > 0.35
            mask[:, 0] = True
        out = mod._pool_instances(x, mask)
        go = torch.randn(out.shape, generator=g).double()
        (out * go).sum().backward()
        np.savez_compressed(OUT / f"{name}.npz", x=_np(x), go=_np(go), out=_np(out), dx=_np(x.grad),
                            mask=np.ones((B, N), bool) if mask is None else mask.numpy(), has_mask=np.array(use_mask),
                            V=_np(mod.attention_V.weight), bV=_np(mod.attention_V.bias), U=_np(mod.attention_U.weight),
                            bU=_np(mod.attention_U.bias), w=_np(mod.attention_w.weight), bw=_np(mod.attention_w.bias),
                            g_V=_np(mod.attention_V.weight.grad), g_bV=_np(mod.attention_V.bias.grad),
                            g_U=_np(mod.attention_U.weight.grad), g_bU=_np(mod.attention_U.bias.grad),
                            g_w=_np(mod.attention_w.weight.grad), g_bw=_np(mod.attention_w.bias.grad))
        print(name, tuple(out.shape))


if __name__ == "__main__":
    which = sys.argv[1:] or ["losses", "siglip_variants", "multipos", "inline_multipos", "tokenmean", "alignment", "dense_metrics", "retrieval", "rope", "attnpool", "clspool", "qpool", "milpool"]
    for name in which:
        globals()[name]()
