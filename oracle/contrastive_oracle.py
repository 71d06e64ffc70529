"""ORACLE (test infrastructure, never imported by the product package).

CPU restatement in numpy of the reference's contrastive losses, with closed-form gradients.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Pinned: tests/test_oracle_golden.py checks every function here against tests/golden/*.npz, which
oracle/gen_golden.py produced by importing the UNMODIFIED reference modules from /root/reference
(torch CPU, fp32 and fp64 autograd) in the build container.

Follows, line by line:
  CLIPLoss.forward            /root/reference/utils/loss/contrastive.py:140-164
  SigLIPLoss.forward          /root/reference/utils/loss/contrastive.py:250-315
  compute_entropy_regularization                          contrastive.py:19-68
  SiglipPairwiseFeatureLoss   /root/reference/utils/loss/siglip_pairwise.py:300-371
  SigLIP2BCELoss / DDP        /root/reference/utils/loss/siglip2_bce.py:79-106, 160-184
  SigLIP2MultiPositiveBCELoss /root/reference/utils/loss/siglip2_bce.py:284-331
  WeightedSigLIPLoss          /root/reference/utils/loss/weighted_siglip.py:33-51
  MultiPositiveInfoNCELoss    /root/reference/utils/loss/multi_positive_infonce.py:30-100
  ContrastiveLoss / DDP       /root/reference/utils/loss/losses.py:44-64, 132-158   (no tau clamp)
  SiglipLoss / DDP (gated)    /root/reference/utils/loss/losses.py:190-211, 241-276
  per-step alignment diagnostics  /root/reference/runners/video_constrative_learning_runner.py:1323-1335
      (an inline block of the runner's train step, not callable outside the training stack: its golden vectors come from
       a torch transcription of those lines in oracle/gen_golden.py, not from an imported function)
"""
from __future__ import annotations

import numpy as np


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> tuple[np.ndarray, np.ndarray]:
    """F.normalize(x, dim=-1): x / max(||x||, eps). Returns (xhat, norm_clamped)."""
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True))
    n = np.maximum(n, eps)
    return x / n, n


def _normalize_backward(dxhat: np.ndarray, xhat: np.ndarray, n: np.ndarray) -> np.ndarray:
    # d/dx of x/||x||  (rows with ||x|| < eps are clamped: derivative is dxhat/eps there; never hit in tests)
    return (dxhat - (dxhat * xhat).sum(axis=-1, keepdims=True) * xhat) / n


def _logsumexp(a: np.ndarray, axis: int) -> np.ndarray:
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def clip_loss(video, text, log_temp, *, label_smoothing: float = 0.0, clamp_min: float | None = 1e-4,
              gated: bool = False, dtype=np.float64, want_grads: bool = True) -> dict:
    """0.5*(CE(L, arange) + CE(L^T, arange)), L = f(S)/tau, S = vhat that^T.

    clamp_min=1e-4 -> CLIPLoss (contrastive.py:153); None -> legacy ContrastiveLoss (losses.py:53).
    gated=True -> f(s) = s*sigmoid(s) (losses.py:198), else identity.
    """
    v = np.asarray(video, dtype=dtype)
    t = np.asarray(text, dtype=dtype)
    lt = dtype(np.asarray(log_temp, dtype=np.float64).reshape(-1)[0])
    vh, vn = l2_normalize(v)
    th, tn = l2_normalize(t)
    S = vh @ th.T
    tau = np.exp(lt)
    clamped = False
    if clamp_min is not None and tau < clamp_min:
        tau, clamped = dtype(clamp_min), True
    if gated:
        sg = _sigmoid(S)
        F = S * sg
        Fp = sg * (1.0 + S * (1.0 - sg))
    else:
        F, Fp = S, None
    L = F / tau
    N = L.shape[0]
    eps = label_smoothing
    r = _logsumexp(L, 1)
    c = _logsumexp(L, 0)
    Y = (1.0 - eps) * np.eye(N, dtype=dtype) + eps / N
    tgt = (Y * L).sum()
    loss = 0.5 * ((r.sum() - tgt) / N + (c.sum() - tgt) / N)
    out = {"loss": float(loss), "row_lse": r, "col_lse": c, "tau": float(tau)}
    if not want_grads:
        return out
    G = ((np.exp(L - r[:, None]) - Y) + (np.exp(L - c[None, :]) - Y)) / (2.0 * N)      # dloss/dL
    dS = G / tau if Fp is None else G * Fp / tau
    dvh = dS @ th
    dth = dS.T @ vh
    out["dvideo"] = _normalize_backward(dvh, vh, vn)
    out["dtext"] = _normalize_backward(dth, th, tn)
    out["dlog_temp"] = 0.0 if clamped else float(-(G * L).sum())
    return out


def entropy_regularization(logits: np.ndarray, min_entropy_threshold: float = 2.0):
    """contrastive.py:19-68 — relu(thr - mean_i H(softmax_j logits_i.)) with log(p + 1e-10)."""
    z = logits - logits.max(axis=1, keepdims=True)
    p = np.exp(z)
    p = p / p.sum(axis=1, keepdims=True)
    ent = -(p * np.log(p + 1e-10)).sum(axis=1)
    mean_ent = ent.mean()
    deficit = max(min_entropy_threshold - mean_ent, 0.0)
    diag = {"entropy_mean": float(mean_ent), "entropy_min": float(ent.min()), "entropy_max": float(ent.max()),
            "entropy_normalized": float(mean_ent / np.log(logits.shape[1])), "entropy_deficit": float(deficit)}
    return deficit, diag, p, ent


def siglip_loss(video, text, log_temp, *, bias: float | None = -10.0, pos_mask=None, pos_weights=None,
                positive_weight: float = 1.0, negative_weight: float = 1.0, use_severity_weights: bool = True,
                auto_balance: bool = False, entropy_regularization_on: bool = False, entropy_weight: float = 0.1,
                min_entropy_threshold: float = 2.0, dtype=np.float64, want_grads: bool = True,
                variant: str = "unified", label_smoothing: float = 0.0) -> dict:
    """Sigmoid-BCE family (single process: the gather is the identity). ``variant``:
      "unified"        SigLIPLoss.forward, contrastive.py:250-315
      "pairwise"       SiglipPairwiseFeatureLoss.forward, siglip_pairwise.py:320-371 (no bias, tau unclamped,
                       weights where pos_mask > 0, negative_weight >= 0)
      "bce2"           SigLIP2BCELoss.forward, siglip2_bce.py:79-106 (identity labels, no clamps, label smoothing)
      "multipositive2" SigLIP2MultiPositiveBCELoss.forward, siglip2_bce.py:284-331 (logit clamp, raw mask labels,
                       label smoothing, weights by smoothed label > 0.5, pos_weights multiply)"""
    v = np.asarray(video, dtype=dtype)
    t = np.asarray(text, dtype=dtype)
    lt = dtype(np.asarray(log_temp, dtype=np.float64).reshape(-1)[0])
    if variant == "unified":
        positive_weight = max(float(positive_weight), 1e-6)
        negative_weight = max(float(negative_weight), 1e-6)
    elif variant == "pairwise":
        positive_weight = max(float(positive_weight), 1e-6)
        negative_weight = max(float(negative_weight), 0.0)
        bias = None
    vh, vn = l2_normalize(v)
    th, tn = l2_normalize(t)
    S = vh @ th.T
    tau = np.exp(lt)
    clamped = variant == "unified" and tau < 1e-4
    if clamped:
        tau = dtype(1e-4)
    b = dtype(0.0 if bias is None else bias)
    R = S / tau + b
    lc = np.inf if variant == "bce2" else 30.0
    L = np.clip(R, -lc, lc)
    B, T = L.shape
    if pos_mask is None or variant == "bce2":
        y = np.zeros((B, T), dtype=dtype)
        m = min(B, T)
        y[:m, :m] = np.eye(m, dtype=dtype)
        raw_mask = y
    else:
        raw_mask = np.asarray(pos_mask, dtype=dtype)
        y = raw_mask if variant == "multipositive2" else np.clip(raw_mask, 0.0, 1.0)
    if label_smoothing > 0:
        y = y * (1.0 - label_smoothing) + label_smoothing / 2.0
    if variant in ("unified", "pairwise"):
        w = np.full((B, T), negative_weight, dtype=dtype)
        if use_severity_weights and pos_weights is not None:
            pc = np.asarray(pos_weights, dtype=dtype) * positive_weight
        else:
            pc = np.full((B, T), positive_weight, dtype=dtype)
        if auto_balance:
            cnt_src = raw_mask if variant == "pairwise" else y
            pos_counts = np.maximum(cnt_src.sum(axis=1, keepdims=True), 1.0)
            ratio = np.maximum((T - pos_counts) / pos_counts, 1.0)
            pc = np.broadcast_to(ratio, (B, T))
        w = np.where((raw_mask > 0) if variant == "pairwise" else (y > 0.5), pc, w)
    elif variant == "bce2":
        w = np.ones((B, T), dtype=dtype)
    else:
        w = np.where(y > 0.5, dtype(positive_weight), dtype(negative_weight))
        if pos_weights is not None:
            w = np.where(y > 0.5, w * np.asarray(pos_weights, dtype=dtype), w)
    bce = np.maximum(L, 0.0) - L * y + np.log1p(np.exp(-np.abs(L)))
    loss = (w * bce).mean()
    out = {"bce_loss": float(loss), "tau": float(tau)}
    dL = w * (_sigmoid(L) - y) / (B * T)
    if entropy_regularization_on:
        deficit, diag, p, ent = entropy_regularization(L, min_entropy_threshold)
        out["entropy_diagnostics"] = diag
        loss = loss + entropy_weight * deficit
        if deficit > 0.0:
            # d(-mean H)/dL_ij = -(1/B) * dH_i/dL_ij ; dH_i/dz_j = -p_j*(log(p_j+e) + p_j/(p_j+e)) + p_j*sum_k p_k(...)
            e = 1e-10
            gterm = np.log(p + e) + p / (p + e)
            dH = -p * gterm + p * (p * gterm).sum(axis=1, keepdims=True)
            dL = dL + entropy_weight * (-dH / B)
    out["loss"] = float(loss)
    if not want_grads:
        return out
    dR = dL * ((R >= -lc) & (R <= lc))
    dS = dR / tau
    out["dvideo"] = _normalize_backward(dS @ th, vh, vn)
    out["dtext"] = _normalize_backward(dS.T @ vh, th, tn)
    out["dbias"] = float(dR.sum())
    out["dlog_temp"] = 0.0 if clamped else float(-(dR * (R - b)).sum())
    return out


def multipos_softmax_loss(logits, weights, *, mask=None, mode="weighted_siglip", eps=1e-6, reduction="mean",
                          dtype=np.float64, want_grads=True) -> dict:
    """WeightedSigLIPLoss.forward (utils/loss/weighted_siglip.py:33-51; mode "weighted_siglip": weights = positive_weights)
    and MultiPositiveInfoNCELoss.forward without importance weighting (utils/loss/multi_positive_infonce.py:30-100; mode
    "infonce": weights = pos_mask [* pos_weights], ``mask`` = pos_mask decides which rows / columns count), with the
    closed-form gradient with respect to the logits."""
    L = np.asarray(logits, dtype=dtype)
    w = np.maximum(np.asarray(weights, dtype=dtype), 0.0)
    N, M = L.shape
    lr = L - _logsumexp(L, 1)[:, None]
    lc = L - _logsumexp(L, 0)[None, :]
    P, Q = w.sum(1), w.sum(0)
    if mode == "weighted_siglip":
        dr, dc = np.maximum(P, eps), np.maximum(Q, eps)
        loss = 0.5 * ((-(w * lr).sum(1) / dr).mean() + (-(w * lc).sum(0) / dc).mean())
        gr, gc = 0.5 / (N * dr), 0.5 / (M * dc)
    else:
        mk = np.asarray(mask, dtype=dtype)
        rsel, csel = mk.sum(1) > 0, mk.sum(0) > 0
        dr, dc = np.maximum(P, 1.0), np.maximum(Q, 1.0)
        lrow = -(w * lr).sum(1) / dr
        lcol = -(w * lc).sum(0) / dc
        stacked = np.concatenate([lrow[rsel], lcol[csel]])
        if stacked.size == 0:
            return {"loss": 0.0, "dlogits": np.zeros_like(L)}
        scale = 1.0 / stacked.size if reduction == "mean" else 1.0
        loss = stacked.sum() * scale
        gr, gc = np.where(rsel, scale / dr, 0.0), np.where(csel, scale / dc, 0.0)
    out = {"loss": float(loss)}
    if want_grads:
        out["dlogits"] = gr[:, None] * (np.exp(lr) * P[:, None] - w) + gc[None, :] * (np.exp(lc) * Q[None, :] - w)
    return out


def alignment_diagnostics(video, text, log_temp, *, use_siglip: bool = False, dtype=np.float64) -> dict:
    """runners/video_constrative_learning_runner.py:1323-1335: mean diagonal cosine, mean diagonal row-log-softmax of
    the (gated when the loss name contains "siglip") logits at tau = exp(log_temp) (no clamp), and its exponential."""
    v = np.asarray(video, dtype=dtype)
    t = np.asarray(text, dtype=dtype)
    vh, _ = l2_normalize(v)
    th, _ = l2_normalize(t)
    S = vh @ th.T                                                       # :1324-1326
    cosine = np.diag(S).mean()                                          # :1327
    base = S * _sigmoid(S) if use_siglip else S                         # :1328-1331
    tau = np.exp(dtype(np.asarray(log_temp, dtype=np.float64).reshape(-1)[0]))   # :1332
    L = base / tau
    logprob = (np.diag(L) - _logsumexp(L, 1)).mean()                    # :1333-1334
    return {"alignment_cosine": float(cosine), "alignment_logprob": float(logprob),
            "alignment_prob": float(np.exp(logprob))}                   # :1335


def inline_multipositive(video, text, log_temp, targets, pos_weights=None, *, abnormal=None, margin: float = 0.0,
                         weighted: bool = True, eps: float = 1e-6, neg_weight: float = 1.0, dtype=np.float64) -> dict:
    """runners/video_constrative_learning_runner.py:1256-1322, the inline branch for batches with a positive mask: gated
    logits s*sigmoid(s)/tau (:1258-1263) + margin on abnormal text columns (:1265-1273), then WeightedSigLIPLoss
    (utils/loss/weighted_siglip.py:38-51; :1275-1283) or BCE-sum / max(1, sum targets) (:1284-1296); closed-form gradients
    w.r.t. the raw features and log_temp; alignment scalars of :1298-1311."""
    v = np.asarray(video, dtype=dtype); t = np.asarray(text, dtype=dtype)
    tg = np.asarray(targets, dtype=dtype)
    pw = None if pos_weights is None else np.asarray(pos_weights, dtype=dtype)
    vh, vn = l2_normalize(v); th, tn = l2_normalize(t)
    S = vh @ th.T
    sg = _sigmoid(S)
    G = S * sg
    tau = np.exp(dtype(np.asarray(log_temp, dtype=np.float64).reshape(-1)[0]))
    L = G / tau
    if abnormal is not None and margin > 0 and np.count_nonzero(abnormal) > 0:
        L = L + np.asarray(abnormal, dtype=dtype)[None, :] * margin
    B, M = L.shape
    use_pw = pw is not None and np.count_nonzero(pw) > 0
    lse_r = _logsumexp(L, 1); lse_c = _logsumexp(L, 0)
    if weighted:
        pos = np.maximum(tg * pw if use_pw else tg, 0.0)
        R = pos.sum(1); C = pos.sum(0)
        dr = np.maximum(R, eps); dc = np.maximum(C, eps)
        l_r = (lse_r * R - (pos * L).sum(1)) / dr
        l_c = (lse_c * C - (pos * L).sum(0)) / dc
        loss = 0.5 * (l_r.mean() + l_c.mean())
        dL = 0.5 / B * (R[:, None] * np.exp(L - lse_r[:, None]) - pos) / dr[:, None] \
            + 0.5 / M * (C[None, :] * np.exp(L - lse_c[None, :]) - pos) / dc[None, :]
    else:
        w = np.where(tg > 0, pw, neg_weight) if pw is not None else np.full_like(tg, neg_weight)
        bce = np.maximum(L, 0) - L * tg + np.log1p(np.exp(-np.abs(L)))
        denom = max(1.0, float(tg.sum()))
        loss = (w * bce).sum() / denom
        dL = w * (_sigmoid(L) - tg) / denom
    dS = dL / tau * sg * (1 + S * (1 - sg))
    dvh = dS @ th; dth = dS.T @ vh
    pwm = pw * tg if use_pw else tg
    rs = pwm.sum(1); valid = rs > 0
    logp = L - lse_r[:, None]
    alp = ((logp * pwm).sum(1)[valid] / rs[valid]).mean() if valid.any() else float("nan")
    cos = S[tg != 0].mean() if (tg != 0).any() else float("nan")
    return {"loss": float(loss), "dvideo": _normalize_backward(dvh, vh, vn), "dtext": _normalize_backward(dth, th, tn),
            "dlog_temp": float(-(dL * G / tau).sum()), "alignment_logprob": float(alp), "alignment_prob": float(np.exp(alp)),
            "alignment_cosine": float(cos)}


# ---- fp32 timing port used by bench.py (cpu_baseline / --impl reference); same algorithm, all host threads ----
def clip_fwd_bwd_f32(video: np.ndarray, text: np.ndarray, log_temp: float):
    """The reference step (normalize, matmul, 2x cross-entropy, backward) in fp32 numpy (BLAS threads)."""
    r = clip_loss(video, text, log_temp, dtype=np.float32)
    return r["loss"], r["dvideo"], r["dtext"], r["dlog_temp"]
