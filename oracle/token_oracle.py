"""ORACLE (test infrastructure, never imported by the product package).

numpy restatements of the token-side modules, with closed-form backward passes:
  rope3d_*            /root/reference/models/rope_3d.py:13-44 (tables), :130-184 (3D grid), :208-252 (apply)
  attention_pool_*    /root/reference/models/attention_pool.py:73-101 (nn.MultiheadAttention with ONE query
                      token + LayerNorm + optional Linear), written in the folded form of SURVEY Appendix A.4
  cls_pool_*          /root/reference/models/attention_pool.py:104-197 (AttentionPoolWithCLS: CLS token prepended,
                      ONE post-LN nn.TransformerEncoderLayer [self-attention, residual + norm1, relu feed-forward,
                      residual + norm2], CLS row -> norm -> optional Linear), written literally: K and V of all
                      N + 1 rows, no folding -- the kernels' lse-merge form is checked against it
  query_pool_*        /root/reference/models/video_aggregator.py:119-123, 128-158 (pos-enc, final LN, masked
                      softmax pooling with uniform fallback)
Pinned against the imported reference by tests/golden/{rope,attnpool,clspool,qpool}_*.npz (oracle/gen_golden.py).
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------
# RoPE 3D
# ------------------------------------------------------------------------------------------------
def rope3d_dims(head_dim: int) -> tuple[int, int, int]:
    if head_dim % 6 != 0:
        raise ValueError("head_dim must be divisible by 6")
    t = head_dim // 3
    return t, t, head_dim - 2 * t


def rope3d_angles(head_dim, T, H, W, n_special=0, temporal_base=10000.0, spatial_base=10000.0, temporal_scale=1.0,
                  dtype=np.float64) -> np.ndarray:
    """theta[n, c] for token n (n_special leading rows are 0) and channel c; pair-duplicated (rope_3d.py:44)."""
    td, hd, wd = rope3d_dims(head_dim)

    def freqs(dim, length, base):
        inv = 1.0 / (dtype(base) ** (np.arange(0, dim, 2, dtype=dtype) / dtype(dim)))
        f = np.arange(length, dtype=dtype)[:, None] * inv[None, :]
        return np.repeat(f, 2, axis=1)                         # (f0,f0,f1,f1,...)

    tf = freqs(td, T, temporal_base * temporal_scale)
    hf = freqs(hd, H, spatial_base)
    wf = freqs(wd, W, spatial_base)
    th = np.zeros((T, H, W, head_dim), dtype=dtype)
    th[..., :td] = tf[:, None, None, :]
    th[..., td:td + hd] = hf[None, :, None, :]
    th[..., td + hd:] = wf[None, None, :, :]
    th = th.reshape(T * H * W, head_dim)
    if n_special:
        th = np.concatenate([np.zeros((n_special, head_dim), dtype=dtype), th], axis=0)
    return th


def _rotate_half(x):
    out = np.empty_like(x)
    out[..., 0::2] = -x[..., 1::2]
    out[..., 1::2] = x[..., 0::2]
    return out


def rope3d_apply(x: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """x [B, heads, N, Dh]; y = x*cos + rotate_half(x)*sin (rope_3d.py:240-246)."""
    return x * cos[None, None] + _rotate_half(x) * sin[None, None]


def rope3d_apply_backward(dy: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """Inverse rotation: dx = dy*cos - rotate_half(dy)*sin  (tables are pair-duplicated)."""
    return dy * cos[None, None] - _rotate_half(dy) * sin[None, None]


def rope3d_forward(q, k, T, H, W, n_special=0, **kw):
    """Rope3D.forward incl. CLS auto-detect and the 'return unchanged on mismatch' branch (:208-221)."""
    N, Dh = q.shape[2], q.shape[3]
    if n_special == 0 and N == T * H * W + 1:
        n_special = 1
    if N != n_special + T * H * W:
        return q, k
    th = rope3d_angles(Dh, T, H, W, n_special, dtype=np.float64, **kw)
    c, s = np.cos(th), np.sin(th)
    return rope3d_apply(q, c, s), rope3d_apply(k, c, s)


# ------------------------------------------------------------------------------------------------
# AttentionPool (one learnable query, nn.MultiheadAttention algebra folded; eval / dropout = 0)
# ------------------------------------------------------------------------------------------------
def _layernorm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    xh = (x - mu) / np.sqrt(var + eps)
    return xh * w + b, xh, 1.0 / np.sqrt(var + eps)


def _layernorm_backward(dy, xh, rstd, w):
    dxh = dy * w
    D = xh.shape[-1]
    dx = (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True)) * rstd
    return dx, (dy * xh).reshape(-1, D).sum(0), dy.reshape(-1, D).sum(0)


def attention_pool_forward(x, params: dict, num_heads: int, mask=None, want_cache=False):
    """x [B,N,D]; params: query(1,1,D), in_proj_weight(3D,D), in_proj_bias(3D), out_proj_weight(D,D),
    out_proj_bias(D), norm_weight(D), norm_bias(D), optional proj_weight(O,D), proj_bias(O)."""
    x = np.asarray(x, dtype=np.float64)
    B, N, D = x.shape
    Hn = num_heads
    Dh = D // Hn
    Wq, Wk, Wv = (np.asarray(params["in_proj_weight"], dtype=np.float64)[i * D:(i + 1) * D] for i in range(3))
    bq, bk, bv = (np.asarray(params["in_proj_bias"], dtype=np.float64)[i * D:(i + 1) * D] for i in range(3))
    query = np.asarray(params["query"], dtype=np.float64).reshape(D)
    q0 = Wq @ query + bq                                              # [D]
    scale = 1.0 / np.sqrt(Dh)
    qt = np.stack([Wk[h * Dh:(h + 1) * Dh].T @ q0[h * Dh:(h + 1) * Dh] * scale for h in range(Hn)])   # [Hn, D]
    beta = np.array([q0[h * Dh:(h + 1) * Dh] @ bk[h * Dh:(h + 1) * Dh] * scale for h in range(Hn)])
    s = np.einsum("bnd,hd->bhn", x, qt) + beta[None, :, None]         # [B,Hn,N]
    if mask is not None:
        s = np.where(np.asarray(mask, dtype=bool)[:, None, :], -np.inf, s)
    m = s.max(-1, keepdims=True)
    e = np.exp(s - m)
    a = e / e.sum(-1, keepdims=True)                                  # [B,Hn,N]
    xbar = np.einsum("bhn,bnd->bhd", a, x)                            # [B,Hn,D]
    o = np.stack([xbar[:, h] @ Wv[h * Dh:(h + 1) * Dh].T + bv[h * Dh:(h + 1) * Dh] for h in range(Hn)], 1)   # [B,Hn,Dh]
    oc = o.reshape(B, D)
    Wo = np.asarray(params["out_proj_weight"], dtype=np.float64)
    bo = np.asarray(params["out_proj_bias"], dtype=np.float64)
    y = oc @ Wo.T + bo
    ln, xh, rstd = _layernorm(y, np.asarray(params["norm_weight"], np.float64), np.asarray(params["norm_bias"], np.float64))
    out = ln
    if "proj_weight" in params and params["proj_weight"] is not None:
        out = ln @ np.asarray(params["proj_weight"], np.float64).T + np.asarray(params["proj_bias"], np.float64)
    if want_cache:
        return out, dict(x=x, a=a, xbar=xbar, qt=qt, q0=q0, oc=oc, xh=xh, rstd=rstd, ln=ln, query=query,
                         Wq=Wq, Wk=Wk, Wv=Wv, Wo=Wo, scale=scale, Hn=Hn, Dh=Dh)
    return out


def attention_pool_backward(dout, cache: dict, params: dict) -> dict:
    """Gradients w.r.t. x and every parameter."""
    c = cache
    x, a, xbar, qt, q0 = c["x"], c["a"], c["xbar"], c["qt"], c["q0"]
    Hn, Dh, scale = c["Hn"], c["Dh"], c["scale"]
    B, N, D = x.shape
    g = {}
    dln = np.asarray(dout, np.float64)
    if "proj_weight" in params and params["proj_weight"] is not None:
        Wp = np.asarray(params["proj_weight"], np.float64)
        g["proj_weight"] = dln.T @ c["ln"]
        g["proj_bias"] = dln.sum(0)
        dln = dln @ Wp
    dy, g["norm_weight"], g["norm_bias"] = _layernorm_backward(dln, c["xh"], c["rstd"], np.asarray(params["norm_weight"], np.float64))
    g["out_proj_weight"] = dy.T @ c["oc"]
    g["out_proj_bias"] = dy.sum(0)
    do = (dy @ c["Wo"]).reshape(B, Hn, Dh)
    dWv = np.zeros((D, D)); dbv = np.zeros(D); dWk = np.zeros((D, D)); dq0 = np.zeros(D)
    dx = np.zeros_like(x)
    for h in range(Hn):
        sl = slice(h * Dh, (h + 1) * Dh)
        dxbar = do[:, h] @ c["Wv"][sl]                                 # [B,D]
        dWv[sl] = do[:, h].T @ xbar[:, h]
        dbv[sl] = do[:, h].sum(0)
        da = np.einsum("bd,bnd->bn", dxbar, x)
        ds = a[:, h] * (da - (dxbar * xbar[:, h]).sum(-1, keepdims=True))
        dx += a[:, h][:, :, None] * dxbar[:, None, :] + ds[:, :, None] * qt[h][None, None, :]
        dqt = np.einsum("bn,bnd->d", ds, x)                            # [D]
        dWk[sl] = np.outer(q0[sl], dqt) * scale
        dq0[sl] = c["Wk"][sl] @ dqt * scale          # (+ bk term: ds sums to zero per row, so d beta contributes 0)
    dWq = np.outer(dq0, c["query"])
    g["in_proj_weight"] = np.concatenate([dWq, dWk, dWv], 0)
    g["in_proj_bias"] = np.concatenate([dq0, np.zeros(D), dbv], 0)
    g["query"] = (c["Wq"].T @ dq0).reshape(1, 1, D)
    g["x"] = dx
    return g


# ------------------------------------------------------------------------------------------------
# AttentionPoolWithCLS (eval / dropout = 0); only row 0 of the layer output is used by the reference (:187-195)
# ------------------------------------------------------------------------------------------------
def cls_pool_forward(x, params: dict, num_heads: int, mask=None, want_cache=False):
    """x [B,N,D]; params: cls_token(1,1,D), in_proj_weight(3D,D), in_proj_bias(3D), out_proj_weight, out_proj_bias,
    linear1_weight(F,D), linear1_bias(F), linear2_weight(D,F), linear2_bias(D), norm1_*, norm2_* (layer), norm_*
    (final, attention_pool.py:152), optional proj_weight(O,D), proj_bias(O)."""
    f8 = lambda k: np.asarray(params[k], dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    B, N, D = x.shape
    Hn, Dh = num_heads, D // num_heads
    X = np.concatenate([np.broadcast_to(f8("cls_token").reshape(1, 1, D), (B, 1, D)), x], 1)      # :177-178
    mk = np.zeros((B, N + 1), bool)
    if mask is not None:
        mk[:, 1:] = np.asarray(mask, dtype=bool)                                                   # :181-184
    Wq, Wk, Wv = (f8("in_proj_weight")[i * D:(i + 1) * D] for i in range(3))
    bq, bk, bv = (f8("in_proj_bias")[i * D:(i + 1) * D] for i in range(3))
    scale = 1.0 / np.sqrt(Dh)
    q = X[:, 0] @ Wq.T + bq                                           # [B,D]   (query row 0 only)
    K = X @ Wk.T + bk                                                 # [B,N+1,D]
    V = X @ Wv.T + bv
    qh, Kh, Vh = q.reshape(B, Hn, Dh), K.reshape(B, N + 1, Hn, Dh), V.reshape(B, N + 1, Hn, Dh)
    s = np.einsum("bhk,bnhk->bhn", qh, Kh) * scale
    s = np.where(mk[:, None, :], -np.inf, s)
    e = np.exp(s - s.max(-1, keepdims=True))
    a = e / e.sum(-1, keepdims=True)                                  # [B,Hn,N+1]
    oc = np.einsum("bhn,bnhk->bhk", a, Vh).reshape(B, D)
    y0 = oc @ f8("out_proj_weight").T + f8("out_proj_bias")
    z1, xh1, rstd1 = _layernorm(X[:, 0] + y0, f8("norm1_weight"), f8("norm1_bias"))
    f1 = z1 @ f8("linear1_weight").T + f8("linear1_bias")
    hid = np.maximum(f1, 0.0)
    f2 = hid @ f8("linear2_weight").T + f8("linear2_bias")
    z2, xh2, rstd2 = _layernorm(z1 + f2, f8("norm2_weight"), f8("norm2_bias"))
    z3, xh3, rstd3 = _layernorm(z2, f8("norm_weight"), f8("norm_bias"))
    out = z3
    if params.get("proj_weight") is not None:
        out = z3 @ f8("proj_weight").T + f8("proj_bias")
    if want_cache:
        return out, dict(X=X, a=a, qh=qh, Kh=Kh, Vh=Vh, oc=oc, z1=z1, xh1=xh1, rstd1=rstd1, f1=f1, hid=hid, xh2=xh2,
                         rstd2=rstd2, z3=z3, xh3=xh3, rstd3=rstd3, Wq=Wq, Wk=Wk, Wv=Wv, scale=scale, Hn=Hn, Dh=Dh)
    return out


def cls_pool_backward(dout, cache: dict, params: dict) -> dict:
    """Gradients w.r.t. x, cls_token and every layer parameter (chain rule through the literal forward above)."""
    f8 = lambda k: np.asarray(params[k], dtype=np.float64)
    c = cache
    X, a, qh, Kh, Vh = c["X"], c["a"], c["qh"], c["Kh"], c["Vh"]
    B, N1, D = X.shape
    Hn, Dh, scale = c["Hn"], c["Dh"], c["scale"]
    g = {}
    dz3 = np.asarray(dout, np.float64)
    if params.get("proj_weight") is not None:
        g["proj_weight"] = dz3.T @ c["z3"]
        g["proj_bias"] = dz3.sum(0)
        dz3 = dz3 @ f8("proj_weight")
    dz2, g["norm_weight"], g["norm_bias"] = _layernorm_backward(dz3, c["xh3"], c["rstd3"], f8("norm_weight"))
    dr2, g["norm2_weight"], g["norm2_bias"] = _layernorm_backward(dz2, c["xh2"], c["rstd2"], f8("norm2_weight"))
    g["linear2_weight"] = dr2.T @ c["hid"]
    g["linear2_bias"] = dr2.sum(0)
    df1 = (dr2 @ f8("linear2_weight")) * (c["f1"] > 0)
    g["linear1_weight"] = df1.T @ c["z1"]
    g["linear1_bias"] = df1.sum(0)
    dz1 = dr2 + df1 @ f8("linear1_weight")
    dr1, g["norm1_weight"], g["norm1_bias"] = _layernorm_backward(dz1, c["xh1"], c["rstd1"], f8("norm1_weight"))
    g["out_proj_weight"] = dr1.T @ c["oc"]
    g["out_proj_bias"] = dr1.sum(0)
    do = (dr1 @ f8("out_proj_weight")).reshape(B, Hn, Dh)
    dVh = np.einsum("bhn,bhk->bnhk", a, do)
    da = np.einsum("bhk,bnhk->bhn", do, Vh)
    ds = a * (da - (a * da).sum(-1, keepdims=True))
    dq = (np.einsum("bhn,bnhk->bhk", ds, Kh) * scale).reshape(B, D)
    dK = (np.einsum("bhn,bhk->bnhk", ds, qh) * scale).reshape(B, N1, D)
    dV = dVh.reshape(B, N1, D)
    dX = dK @ c["Wk"] + dV @ c["Wv"]
    dX[:, 0] += dr1 + dq @ c["Wq"]
    g["in_proj_weight"] = np.concatenate([dq.T @ X[:, 0], np.einsum("bnd,bne->de", dK, X), np.einsum("bnd,bne->de", dV, X)], 0)
    g["in_proj_bias"] = np.concatenate([dq.sum(0), dK.sum((0, 1)), dV.sum((0, 1))], 0)
    g["cls_token"] = dX[:, 0].sum(0).reshape(1, 1, D)
    g["x"] = dX[:, 1:]
    return g


# ------------------------------------------------------------------------------------------------
# Multi-view query pool (EnhancedVideoAggregator without its transformer blocks)
# ------------------------------------------------------------------------------------------------
def query_pool_forward(x, pos_encoding, ln_w, ln_b, attn_query, mask=None, want_cache=False):
    """x [B,N,D] -> [B,D]: x + pos[:N]; LayerNorm; masked rows -> 0; softmax_n(q . x_n) (no 1/sqrt(D));
    all-masked rows fall back to uniform over valid (=> zero output)."""
    x = np.asarray(x, np.float64)
    B, N, D = x.shape
    if pos_encoding is not None:
        x = x + np.asarray(pos_encoding, np.float64).reshape(-1, D)[None, :N]
    ln, xh, rstd = _layernorm(x, np.asarray(ln_w, np.float64), np.asarray(ln_b, np.float64))
    q = np.asarray(attn_query, np.float64).reshape(D)
    if mask is not None:
        mk = np.asarray(mask, bool)
        ln = np.where(mk[..., None], 0.0, ln)
    s = ln @ q                                                          # [B,N]
    if mask is not None:
        s = np.where(mk, -np.inf, s)
        with np.errstate(invalid="ignore"):
            m = s.max(-1, keepdims=True)
            e = np.exp(s - m)
            w = e / e.sum(-1, keepdims=True)
        w = np.nan_to_num(w, nan=0.0, posinf=0.0, neginf=0.0)
        bad = w.sum(-1, keepdims=True) <= 0
        valid = (~mk).astype(np.float64)
        fb = valid / np.maximum(valid.sum(-1, keepdims=True), 1e-6)
        w = np.where(bad, fb, w)
    else:
        m = s.max(-1, keepdims=True)
        e = np.exp(s - m)
        w = e / e.sum(-1, keepdims=True)
    out = np.einsum("bn,bnd->bd", w, ln)
    if want_cache:
        return out, dict(ln=ln, xh=xh, rstd=rstd, w=w, q=q, mask=None if mask is None else mk)
    return out


def query_pool_backward(dout, cache, ln_w):
    c = cache
    ln, w, q = c["ln"], c["w"], c["q"]
    dout = np.asarray(dout, np.float64)
    dw = np.einsum("bd,bnd->bn", dout, ln)
    ds = w * (dw - (dw * w).sum(-1, keepdims=True))
    if c["mask"] is not None:
        allmasked = c["mask"].all(-1, keepdims=True)
        ds = np.where(allmasked, 0.0, ds)
    dln = w[..., None] * dout[:, None, :] + ds[..., None] * q[None, None, :]
    dq = np.einsum("bn,bnd->d", ds, ln)
    if c["mask"] is not None:
        dln = np.where(c["mask"][..., None], 0.0, dln)
    dx, dlw, dlb = _layernorm_backward(dln, c["xh"], c["rstd"], np.asarray(ln_w, np.float64))
    return dict(x=dx, pos=dx.sum(0), ln_w=dlw, ln_b=dlb, attn_query=dq.reshape(1, 1, -1))


# ------------------------------------------------------------------------------------------------
# Gated-attention MIL pooling (models/multi_instance_linear_probing.py:493-536)
# ------------------------------------------------------------------------------------------------
def mil_gated_pool_forward(x, mask, V, bV, U, bU, w, bw, want_cache=False):
    """One level (:499-507): x [S,L,D]; mask [S,L] True = valid (or None). A = softmax_l(w.(tanh(Vx+bV)*sigmoid(Ux+bU))+bw
    with invalid -> -inf); out = sum_l A_l x_l. A sequence without a valid instance gives NaN, as torch's softmax does."""
    x = np.asarray(x, np.float64)
    V, U = np.asarray(V, np.float64), np.asarray(U, np.float64)
    w = np.asarray(w, np.float64).reshape(-1)
    t = np.tanh(x @ V.T + np.asarray(bV, np.float64))
    g = 1.0 / (1.0 + np.exp(-(x @ U.T + np.asarray(bU, np.float64))))
    a = (t * g) @ w + float(np.asarray(bw).reshape(-1)[0])                 # [S,L]
    if mask is not None:
        a = np.where(np.asarray(mask, bool), a, -np.inf)
    with np.errstate(invalid="ignore"):
        e = np.exp(a - a.max(-1, keepdims=True))
        A = e / e.sum(-1, keepdims=True)
    out = np.einsum("sl,sld->sd", A, x)
    if want_cache:
        return out, dict(x=x, t=t, g=g, A=A, V=V, U=U, w=w)
    return out


def mil_gated_pool_backward(dout, c):
    x, t, g, A, V, U, w = c["x"], c["t"], c["g"], c["A"], c["V"], c["U"], c["w"]
    dout = np.asarray(dout, np.float64)
    dA = np.einsum("sd,sld->sl", dout, x)
    ds = A * (dA - (A * dA).sum(-1, keepdims=True))
    dpv = ds[..., None] * w * g * (1.0 - t * t)
    dpu = ds[..., None] * w * t * g * (1.0 - g)
    dx = A[..., None] * dout[:, None, :] + dpv @ V + dpu @ U
    return dict(x=dx, V=np.einsum("slh,sld->hd", dpv, x), bV=dpv.sum((0, 1)), U=np.einsum("slh,sld->hd", dpu, x),
                bU=dpu.sum((0, 1)), w=np.einsum("sl,slh->h", ds, t * g).reshape(1, -1), bw=np.array([ds.sum()]))


def mil_hierarchical_pool_forward(x, mask, V, bV, U, bU, w, bw, want_cache=False):
    """Two levels (:509-536): patch level over L per (b, n) without a mask, then the video level with the mask."""
    x = np.asarray(x, np.float64)
    B, N, L, D = x.shape
    emb, c1 = mil_gated_pool_forward(x.reshape(B * N, L, D), None, V, bV, U, bU, w, bw, want_cache=True)
    out, c2 = mil_gated_pool_forward(emb.reshape(B, N, D), mask, V, bV, U, bU, w, bw, want_cache=True)
    return (out, (c1, c2)) if want_cache else out


def mil_hierarchical_pool_backward(dout, caches):
    c1, c2 = caches
    g2 = mil_gated_pool_backward(dout, c2)
    S, L, D = c1["x"].shape
    g1 = mil_gated_pool_backward(g2["x"].reshape(S, D), c1)
    out = {k: g1[k] + g2[k] for k in ("V", "bV", "U", "bU", "w", "bw")}
    out["x"] = g1["x"].reshape(c2["x"].shape[0], c2["x"].shape[1], L, D)
    return out
