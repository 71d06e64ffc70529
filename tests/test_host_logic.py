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
    agg = set(EnhancedVideoAggregator(64, num_heads=4, aggregator_depth=1, max_segments=8).state_dict())
    assert {"pos_encoding", "attn_query", "final_ln.weight", "final_ln.bias", "blocks.0.norm1.weight",
            "blocks.0.attn.in_proj_weight", "blocks.0.mlp.0.weight", "blocks.0.mlp.3.bias"} <= agg


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
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "clip32k"
