"""CPU: host-side logic that needs no GPU — the C ABI library loads and exports every symbol of include/b200clip.h,
RoPE tables equal the reference's bit for bit, module state-dicts keep the reference's parameter names, the registry
hook overwrites the right keys, and the product path refuses to run without CUDA (no CPU fallback)."""
import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN


def test_library_loads_and_exports_every_header_symbol():
    from deepcoro_clip_b200 import _lib
    lib = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.b200clip_abi_version() == 1
    assert b"invalid" in lib.b200clip_strerror(-22)


def test_header_is_plain_c_and_a_c_host_can_bind_the_library(tmp_path):
    """include/b200clip.h is the boundary: it must compile as C99 (no C++ / torch types) and a plain C program linked against
    libb200clip.so must be able to call it (the GPU-free entry points only — no compute without a device)."""
    import os
    import shutil
    import subprocess
    from deepcoro_clip_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    _lib.lib()
    root = _lib._PKG.parent
    src = tmp_path / "host.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "b200clip.h"\n'
                   'int main(void) {\n'
                   '  if (b200clip_abi_version() != B200CLIP_ABI_VERSION) return 1;\n'
                   '  if (strstr(b200clip_strerror(-22), "invalid") == NULL) return 2;\n'
                   '  if (b200clip_l2norm_fwd(NULL, 0, 0, 4, 8, NULL, 64, 64, -1, NULL, NULL, 0, 1, NULL) != -22) return 3;\n'
                   '  printf("abi %d\\n", b200clip_abi_version());\n  return 0;\n}\n')
    exe = tmp_path / "host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{root / 'include'}", str(src), "-o", str(exe),
                    f"-L{_lib._PKG}", "-lb200clip", f"-Wl,-rpath,{_lib._PKG}", "-Wl,--allow-shlib-undefined"], check=True)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = f"/usr/local/cuda/lib64:{env.get('LD_LIBRARY_PATH', '')}"
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "abi 1", (r.returncode, r.stdout, r.stderr)


@pytest.mark.parametrize("name,dtype", [("rope_f32_t2h3w4_cls", torch.float32), ("rope_bf16_t4h7w7_cls", torch.bfloat16)])
def test_rope_tables_bit_identical_to_reference(name, dtype):
    from deepcoro_clip_b200.rope_3d import Rope3D
    g = np.load(GOLDEN / f"{name}.npz")
    B, heads, T, H, W, cls = [int(x) for x in g["meta"]]
    mod = Rope3D(96 * heads, heads).eval()
    sin, cos = mod._get_cached_freqs(T, H, W, torch.device("cpu"), dtype, cls)
    assert (sin.float().numpy() == g["sin"].astype(np.float32)).all()
    assert (cos.float().numpy() == g["cos"].astype(np.float32)).all()
    assert (T, H, W, torch.device("cpu"), dtype, cls) in mod._cache
    assert mod.t_dim == mod.h_dim == mod.w_dim == 32 and mod.head_dim == 96


def test_rope_ctor_errors_and_mismatch_passthrough():
    from deepcoro_clip_b200.rope_3d import Rope3D
    with pytest.raises(ValueError):
        Rope3D(embed_dim=64, num_heads=8)          # head_dim 8 not divisible by 6
    mod = Rope3D(192, 2).eval()
    q = torch.randn(1, 2, 7, 96)
    qr, kr = mod(q, q, 2, 2, 2)                    # N != THW (+1): inputs returned unchanged, no kernel involved
    assert qr is q and kr is q
    assert len(mod.state_dict()) == 0


def test_state_dict_names_match_reference():
    from deepcoro_clip_b200.attention_pool import AttentionPool
    from deepcoro_clip_b200.video_aggregator import EnhancedVideoAggregator
    keys = set(AttentionPool(512, 8, output_dim=256).state_dict())
    assert keys == {"query", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
                    "norm.weight", "norm.bias", "proj.weight", "proj.bias"}
    from deepcoro_clip_b200.attention_pool import AttentionPoolWithCLS
    cls_keys = set(AttentionPoolWithCLS(128, 8, output_dim=64).state_dict())
    layer = "transformer.layers.0."
    assert cls_keys == {"cls_token", "norm.weight", "norm.bias", "proj.weight", "proj.bias"} | {
        layer + k for k in ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                            "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                            "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")}
    with pytest.raises(ValueError):
        AttentionPoolWithCLS(128, 8, num_layers=2)
    agg = set(EnhancedVideoAggregator(64, num_heads=4, aggregator_depth=1, max_segments=8).state_dict())
    assert {"pos_encoding", "attn_query", "final_ln.weight", "final_ln.bias", "blocks.0.norm1.weight",
            "blocks.0.attn.in_proj_weight", "blocks.0.mlp.0.weight", "blocks.0.mlp.3.bias"} <= agg


CLS_POOL_PARAMS = {"cls_token": "cls_token", "in_proj_weight": "self_attn.in_proj_weight",
                   "in_proj_bias": "self_attn.in_proj_bias", "out_proj_weight": "self_attn.out_proj.weight",
                   "out_proj_bias": "self_attn.out_proj.bias", "linear1_weight": "linear1.weight",
                   "linear1_bias": "linear1.bias", "linear2_weight": "linear2.weight", "linear2_bias": "linear2.bias",
                   "norm1_weight": "norm1.weight", "norm1_bias": "norm1.bias", "norm2_weight": "norm2.weight",
                   "norm2_bias": "norm2.bias", "norm_weight": "norm.weight", "norm_bias": "norm.bias",
                   "proj_weight": "proj.weight", "proj_bias": "proj.bias"}


def load_cls_pool(mod, g):
    """Copies the golden fixture's parameters into a (product) AttentionPoolWithCLS; returns {fixture key: Parameter}."""
    named = dict(mod.named_parameters())
    out = {}
    with torch.no_grad():
        for k, tail in CLS_POOL_PARAMS.items():
            if "p_" + k not in g.files:
                continue
            full = tail if tail.split(".")[0] in ("cls_token", "norm", "proj") else "transformer.layers.0." + tail
            named[full].copy_(torch.as_tensor(g["p_" + k].astype(np.float64)).to(named[full].dtype))
            out[k] = named[full]
    return out


@pytest.mark.parametrize("name", ["clspool_b3_n50_d128_h8", "clspool_b4_n37_d128_h4_mask_proj"])
def test_cls_pool_merge_algebra_with_torch_stand_in(name, monkeypatch):
    """Host-side algebra of AttentionPoolWithCLS (lse merge of the CLS key, folded scores, feed-forward tail) against
    the reference's fp64 output and gradients, with the streaming kernel replaced by a dense autograd stand-in that
    has the kernel's contract (xbar, sa, lse). The kernel itself is covered by the -m gpu tests."""
    from deepcoro_clip_b200 import attention_pool as ap
    g = np.load(GOLDEN / f"{name}.npz")

    def stand_in(x, qt, mask, drop_p, drop_seed, want_lse):
        assert want_lse and drop_p == 0.0
        s = torch.einsum("bnd,hd->bhn", x, qt.to(x.dtype))
        if mask is not None:
            s = s.masked_fill(mask[:, None, :], float("-inf"))
        lse = torch.logsumexp(s, dim=-1)
        a = torch.nan_to_num(torch.exp(s - lse.unsqueeze(-1)), nan=0.0)
        return torch.einsum("bhn,bnd->bhd", a, x), torch.ones_like(lse), lse

    monkeypatch.setattr(ap._StreamPool, "apply", staticmethod(stand_in))
    monkeypatch.setattr(torch.Tensor, "float", lambda t: t)          # keep the fp64 fixture in fp64 end to end
    out_dim = g["p_proj_weight"].shape[0] if "p_proj_weight" in g.files else None
    mod = ap.AttentionPoolWithCLS(g["x"].shape[2], int(g["heads"]), output_dim=out_dim).double().eval()
    params = load_cls_pool(mod, g)
    x = torch.tensor(g["x"], requires_grad=True)
    mask = torch.tensor(g["mask"]) if bool(g["has_mask"]) else None
    out = mod(x, mask)
    (out * torch.tensor(g["go"])).sum().backward()
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(x.grad.numpy(), g["dx"], rtol=1e-8, atol=1e-11)
    for k, prm in params.items():
        got = prm.grad.numpy()
        got = got[::16] if k == "linear1_weight" else got[:, ::16] if k == "linear2_weight" else got
        np.testing.assert_allclose(got, g["g_" + k], rtol=1e-8, atol=1e-10, err_msg=k)


def test_cls_pool_dropout_algebra_with_torch_stand_in(monkeypatch):
    """Training-mode algebra of AttentionPoolWithCLS (dropout over the N + 1 attention weights, value-bias weight
    w_x * sa + w_c) against the float64 replica of the reference math with the same masks; the streaming kernel is
    replaced by a dense stand-in that applies the kernel's counter-based mask."""
    from deepcoro_clip_b200 import attention_pool as ap
    from tests.test_gpu_tokens import _cls_pool_replica, _keep_mask
    B, N, D, H, p = 3, 29, 128, 4, 0.25

    def stand_in(x, qt, mask, drop_p, drop_seed, want_lse):
        assert want_lse and drop_p == p
        s = torch.einsum("bnd,hd->bhn", x, qt.to(x.dtype))
        if mask is not None:
            s = s.masked_fill(mask[:, None, :], float("-inf"))
        lse = torch.logsumexp(s, dim=-1)
        a = torch.exp(s - lse.unsqueeze(-1)) * _keep_mask(drop_seed, B, H, N, p, "cpu").to(x.dtype) / (1 - p)
        return torch.einsum("bhn,bnd->bhd", a, x), a.sum(-1), lse

    monkeypatch.setattr(ap._StreamPool, "apply", staticmethod(stand_in))
    monkeypatch.setattr(torch.Tensor, "float", lambda t: t)
    monkeypatch.setattr(ap.F, "dropout", lambda t, p_=0.5, training=True, inplace=False: t)
    torch.manual_seed(5)
    mod = ap.AttentionPoolWithCLS(D, H, output_dim=48, dropout=p).double().train()
    with torch.no_grad():
        mod.cls_token.normal_(std=0.5)
        for k, prm in mod.named_parameters():
            if k.endswith("bias"):
                prm.normal_(std=0.3)
    x = torch.randn(B, N, D, dtype=torch.float64, requires_grad=True)
    mask = torch.rand(B, N) < 0.2
    torch.manual_seed(99)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    keep_c = torch.rand(B, H) >= p
    torch.manual_seed(99)
    out = mod(x, mask)
    go = torch.randn_like(out)
    out.backward(go)
    y, xd, P = _cls_pool_replica(mod, x, mask, _keep_mask(seed, B, H, N, p, "cpu"), keep_c, p)
    y.backward(go)
    np.testing.assert_allclose(out.detach().numpy(), y.detach().numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(x.grad.numpy(), xd.grad.numpy(), rtol=1e-8, atol=1e-11)
    for k, prm in mod.named_parameters():
        np.testing.assert_allclose(prm.grad.numpy(), P[k].grad.numpy(), rtol=1e-8, atol=1e-10, err_msg=k)


@pytest.mark.parametrize("name", ["align_b64_d512", "align_siglip_b130_d96"])
def test_alignment_diagnostics_wiring_with_torch_stand_in(name, monkeypatch):
    """Host wiring of diagnostics.alignment_diagnostics (arena layout, argument order, dyn slots, the formula the scalar
    kernel implements) against the golden vectors, with the C-ABI entry points replaced by torch stand-ins that follow
    the contracts written in include/b200clip.h. The kernels themselves are covered by the -m gpu tests."""
    from deepcoro_clip_b200 import diagnostics as dg, ops
    g = np.load(GOLDEN / f"{name}.npz")
    LOG2E, LN2 = 1.4426950408889634, 0.6931471805599453

    def l2norm_operand(x, role=-1, normalize=True):
        x = x.double()
        return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12), None, x.shape[1]

    def dyn_prep(log_temp, bias, clamp_min, bound):
        assert bias is None and clamp_min == 0.0
        tau = float(torch.exp(log_temp.double().reshape(-1)[0]))
        scale2 = LOG2E / tau
        shift2 = scale2 * bound - min(max(2.0 * bound * scale2 - 120.0, 0.0), 100.0)
        d = torch.zeros(16, dtype=torch.float64)
        d[0], d[1], d[2], d[3], d[6], d[7] = scale2, shift2, 1.0 / tau, tau, LN2 * shift2, 1.0
        return d

    def call(fn, *a):
        if fn == "logits_lse_fwd":
            A, Bm, Ma, Nb, K, lda, ldb, s2, sh2, gated, dyn, skip_if_stable, rowsum, colsum, diag, diag_off, st = a
            assert skip_if_stable == 1 and dyn[11] == 0
            S = A @ Bm.T
            P = torch.exp2((S * torch.sigmoid(S) if gated else S) * dyn[0] - dyn[1])
            rowsum += P.sum(1)
            colsum += P.sum(0)
            diag.copy_(torch.diagonal(S, diag_off))
        elif fn == "logits_rowlse":
            assert a[9] == 1 and a[8][11] == 0            # only_if_stable; dyn[11] == 0: the launch gate stays closed
        elif fn == "alignment_diag":
            sums, n, dyn, gated, out, st = a
            d = sums[2 * n:3 * n]
            f = d * torch.sigmoid(d) if gated else d
            lp = (f * dyn[2] - (torch.log(sums[n:2 * n]) + dyn[6])).mean()
            out[0], out[1], out[2] = d.mean(), lp, torch.exp(lp)
        else:
            raise AssertionError(fn)

    monkeypatch.setattr(ops, "require_cuda", lambda *t: torch.device("cpu"))
    monkeypatch.setattr(ops, "stream_ptr", lambda dev=None: 0)
    monkeypatch.setattr(ops, "l2norm_operand", l2norm_operand)
    monkeypatch.setattr(ops, "dyn_prep", dyn_prep)
    monkeypatch.setattr(ops, "call", call)
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{**k, "dtype": torch.float64}))
    r = dg.alignment_diagnostics(torch.tensor(g["video"]), torch.tensor(g["text"]), torch.tensor(g["log_temp"]),
                                 use_siglip=bool(g["use_siglip"]))
    for key, ref in (("alignment_cosine", "cosine_f64"), ("alignment_logprob", "logprob_f64"),
                     ("alignment_prob", "prob_f64")):
        assert abs(float(r[key]) - float(g[ref])) <= 1e-10 * max(1.0, abs(float(g[ref]))), key
    with pytest.raises(ValueError):
        dg.alignment_diagnostics(torch.zeros(4, 8), torch.zeros(5, 8), 0.0)


def test_no_cpu_fallback():
    from deepcoro_clip_b200._lib import B200ClipError
    from deepcoro_clip_b200.loss import CLIPLoss, SigLIPLoss
    from deepcoro_clip_b200.retrieval_metrics_streaming import compute_metrics_streaming
    v, t = torch.randn(4, 8), torch.randn(4, 8)
    with pytest.raises(B200ClipError):
        CLIPLoss()(v, t, torch.tensor(0.0))
    with pytest.raises(B200ClipError):
        SigLIPLoss()(v, t, torch.tensor(0.0))
    with pytest.raises(B200ClipError):
        compute_metrics_streaming(v, t, torch.arange(4))
    from deepcoro_clip_b200.attention_pool import AttentionPool, AttentionPoolWithCLS
    for cls in (AttentionPool, AttentionPoolWithCLS):
        with pytest.raises(B200ClipError):
            cls(128, 8)(torch.randn(2, 5, 128))


def test_loss_module_signatures():
    import inspect
    from deepcoro_clip_b200 import loss as L
    assert list(inspect.signature(L.CLIPLoss.forward).parameters)[1:] == ["video_features", "text_features", "log_temp"]
    assert list(inspect.signature(L.SigLIPLoss.forward).parameters)[1:] == ["video_features", "text_features", "log_temp",
                                                                             "pos_mask", "pos_weights"]
    s = L.SigLIPLoss()
    assert isinstance(s.bias, torch.nn.Parameter) and float(s.bias) == -10.0 and s.bias.ndim == 0
    assert not isinstance(L.SigLIPLoss(learnable_bias=False).bias, torch.nn.Parameter)
    assert s.get_entropy_diagnostics() == {}
    assert L.InfoNCELoss(temperature=0.07).log_temp.requires_grad
    with pytest.raises(ValueError):
        L.InfoNCELoss(loss_type="bogus")(torch.randn(2, 4), torch.randn(2, 4))


def test_install_into_reference_registry():
    """Runs only where the reference checkout exists (the build container); the GPU box has no /root/reference."""
    import os
    import sys
    if not os.path.isdir("/root/reference/utils"):
        pytest.skip("/root/reference not present")
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        from utils.registry import LossRegistry, register_submodules
        import utils.loss.typing  # noqa: F401  (scripts/main.py import order: losses.py first)
        register_submodules("utils.loss")
        import deepcoro_clip_b200 as pkg
        rep = pkg.install("/root/reference")
        assert set(rep["losses"]) >= {"clip", "contrastive", "contrastive_ddp", "siglip", "siglip_ddp", "InfoNCE"}
        assert LossRegistry.get("contrastive") is pkg.CLIPLoss
        assert LossRegistry.get("siglip") is pkg.SigLIPLoss
        assert LossRegistry.get("siglip_ddp") is pkg.SiglipLossDDP
        assert LossRegistry.get("InfoNCE") is pkg.InfoNCELoss
        from utils.enums import LossType
        assert LossRegistry.get(LossType.CLIP) is pkg.CLIPLoss          # str-enum keys hash like their values
        inst = LossRegistry.get("contrastive")()                          # zero-arg constructor (project wiring)
        from utils.loss.typing import Loss
        assert callable(Loss(loss_type=inst).run)
        pkg.install("/root/reference", semantics="cold")
        assert LossRegistry.get("siglip") is pkg.SiglipLoss and LossRegistry.get("contrastive") is pkg.ContrastiveLoss
        with pytest.raises(ValueError):
            LossRegistry.get("INFONCE_LOSS_DDP")                           # dead config key stays an error
        import utils.retrieval_metrics_streaming as rms
        assert rms.compute_metrics_streaming is pkg.compute_metrics_streaming
        # SURVEY §8 row a6 classes are rebound by name on their home modules, §8f #1 functions on utils.retrieval_metrics
        import utils.loss.siglip2_bce as s2
        import utils.loss.siglip_pairwise as sp
        import utils.retrieval_metrics as urm
        assert s2.SigLIP2BCELoss is pkg.SigLIP2BCELoss and s2.SigLIP2MultiPositiveBCELoss is pkg.SigLIP2MultiPositiveBCELoss
        assert s2.SigLIP2BCELossDDP is pkg.SigLIP2BCELossDDP and sp.SiglipPairwiseFeatureLoss is pkg.SiglipPairwiseFeatureLoss
        import utils.loss.weighted_siglip as uws
        assert uws.WeightedSigLIPLoss is pkg.WeightedSigLIPLoss
        assert LossRegistry.get("multi_positive_infonce") is pkg.MultiPositiveInfoNCELoss
        import models.attention_pool as map_
        assert map_.AttentionPoolWithCLS is pkg.AttentionPoolWithCLS and map_.AttentionPool is pkg.AttentionPool
        if "models.video_encoder" in sys.modules:
            assert sys.modules["models.video_encoder"].AttentionPoolWithCLS is pkg.AttentionPoolWithCLS
        # the gated-attention pooling methods of the probing head are rebound; CPU inputs keep the reference's
        import models.multi_instance_linear_probing as mil
        assert hasattr(mil.MultiInstanceLinearProbing._attention_pooling, "reference")
        assert hasattr(mil.MultiInstanceLinearProbing._hierarchical_attention_pooling, "reference")
        probe = mil.MultiInstanceLinearProbing(32, {"h": 2}, pooling_mode="attention", attention_hidden=8)
        xs = torch.randn(2, 3, 32)
        assert torch.equal(probe._pool_instances(xs), type(probe)._attention_pooling.reference(probe, xs, torch.ones(2, 3, dtype=torch.bool)))
        assert urm.compute_mrr is pkg.retrieval_metrics.compute_mrr and urm.compute_map is pkg.retrieval_metrics.compute_map
        assert "utils.retrieval_metrics" in rep["metrics"] or "utils.retrieval_metrics" in pkg.install("/root/reference")["metrics"]
    finally:
        sys.path.remove("/root/reference")


def test_ground_truth_set_normalisation_matches_reference():
    """Host logic of the dense metrics: ground-truth specifications -> one index set per query, as the reference's
    _normalize_ground_truth_sets (utils/retrieval_metrics.py:7-62) builds them (compared with the imported reference where
    it is available, with the documented rules otherwise)."""
    import importlib.util
    import os
    from deepcoro_clip_b200.retrieval_metrics import _normalize_ground_truth_sets as ours
    cases = [
        (torch.tensor([3, 0, 7]), 3),
        (torch.tensor([[1, 2, -1], [4, -1, -1], [-1, -1, -1]]), 3),
        ([[1, 2], [3], [], None, 5], 5),
        ([1, 2, 3], 5),                      # shorter than the number of queries: padded with empty sets
        ([[1], [2], [3], [4]], 2),           # longer: truncated
        ((0, (1, 1, 2), {4, 5}), 3),
    ]
    expect = [
        [{3}, {0}, {7}], [{1, 2}, {4}, set()], [{1, 2}, {3}, set(), set(), {5}], [{1}, {2}, {3}, set(), set()],
        [{1}, {2}], [{0}, {1, 2}, {4, 5}],
    ]
    ref = None
    path = "/root/reference/utils/retrieval_metrics.py"
    if os.path.exists(path):
        spec = importlib.util.spec_from_file_location("_ref_retrieval_metrics", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref = mod._normalize_ground_truth_sets
    for (gt, n), want in zip(cases, expect):
        assert ours(gt, n) == want
        if ref is not None:
            assert ref(gt, n) == want
    with pytest.raises(ValueError):
        ours(torch.zeros(2, 2, 2, dtype=torch.long), 2)
    with pytest.raises(TypeError):
        ours(3.5, 1)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): exactly one JSON line on stdout with the
    keys of the contract, no GPU needed."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("contrastive loss fwd+bwd samples/s")
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "clip32k"


def test_plan_entry_points_agree_with_the_host_formulas():
    """Host-only entry points of the C ABI (no launch, no device needed): buffer sizes the Python side computes on its own must
    equal what the library expects (include/b200clip.h K3b, K11, K11b, the aggregator step)."""
    import ctypes
    from deepcoro_clip_b200 import _lib, ops
    lib = _lib.lib()
    out = ctypes.c_int64(0)
    for nx, ny in ((1, 1), (128, 256), (129, 257), (1500, 1500), (32768, 32768), (1024, 8192)):
        assert lib.b200clip_gstore_elems(nx, ny, ctypes.byref(out)) == 0
        assert out.value == ops.gstore_elems(nx, ny)
        assert out.value % 4096 == 0 and out.value >= nx * ny
    assert lib.b200clip_gstore_elems(0, 5, ctypes.byref(out)) != 0
    sizes = (ctypes.c_int64 * 3)()
    B, N, D, H, F = 8, 4, 512, 4, 2048
    assert lib.b200clip_aggregator_sizes(B, N, D, H, F, sizes) == 0
    al = lambda n: (n + 63) // 64 * 64
    R = B * N
    assert sizes[0] == 6 * al(R * D) + 2 * al(R) + al(3 * R * D) + al(B * H * N * N) + 2 * al(R * F)
    assert sizes[1] == 4 * al(R * D) + al(R * F) + al(3 * R * D)
    assert sizes[2] == 4 * D * D + 2 * F * D + 9 * D + F
    assert lib.b200clip_aggregator_sizes(B, 17, D, H, F, sizes) != 0          # more than 16 views: not a fused shape
    plan = (ctypes.c_int * 4)()
    assert lib.b200clip_milpool_plan(32, 1568, 512, 128, plan) == 0
    P, Z, chunks, nut = tuple(plan)
    assert 1 <= P <= 16 and Z >= 1 and chunks == (32 * 1568 + 63) // 64 and nut == 2
    assert lib.b200clip_milpool_plan(4, 8, 60, 128, plan) != 0                # D % 16 != 0
    assert lib.b200clip_milpool_ok(8, 512, 128) == 1 and lib.b200clip_milpool_ok(8, 512, 12) == 0
