"""Plugs the B200 implementations into the reference's own extension points:

  * ``LossRegistry`` (utils/registry.py:18-24, plain dict overwrite — last registration wins) under the same
    ``LossType`` keys the reference registers;
  * the module-level names ``models.video_encoder.Rope3D`` / ``AttentionPool`` / ``AttentionPoolWithCLS`` /
    ``EnhancedVideoAggregator`` that ``VideoEncoder.__init__`` instantiates (models/video_encoder.py:12-13, 115-140,
    207-219); the gated-attention pooling methods of ``MultiInstanceLinearProbing``
    (models/multi_instance_linear_probing.py:493-536) for CUDA inputs;
  * the function names ``compute_metrics_streaming`` / ``compute_recall_at_k_streaming`` imported by
    runners/multitask_runner.py:35, and the dense multi-label metrics of utils/retrieval_metrics.py.

Call ``install()`` AFTER ``register_submodules("utils.loss")`` (scripts/main.py:26-30) so that these entries are
the last ones written."""
from __future__ import annotations

import importlib
import sys
from typing import Dict

from . import (attention_pool, loss, mil_pooling, multipos_loss, retrieval_metrics, retrieval_metrics_streaming, rope_3d,
               video_aggregator)

# key -> class, for the two import orders the reference can end up with (SURVEY §8b "registration order hazard")
_MAIN = {          # scripts/main.py order: utils/loss/contrastive.py registers last
    "clip": loss.CLIPLoss, "contrastive": loss.CLIPLoss, "contrastive_ddp": loss.CLIPLoss,
    "siglip": loss.SigLIPLoss, "siglip_pairwise": loss.SigLIPLoss, "siglip2_bce": loss.SigLIPLoss,
    "siglip2_bce_ddp": loss.SigLIPLoss, "siglip2_multi_positive": loss.SigLIPLoss,
    "siglip_ddp": loss.SiglipLossDDP, "InfoNCE": loss.InfoNCELoss,
    "multi_positive_infonce": multipos_loss.MultiPositiveInfoNCELoss,     # logits-level loss
}
_COLD = dict(_MAIN)   # cold import (alphabetical walk): utils/loss/losses.py overwrites three keys with legacy classes
_COLD.update({"contrastive": loss.ContrastiveLoss, "contrastive_ddp": loss.ContrastiveLossDDP, "siglip": loss.SiglipLoss})


def loss_table(semantics: str = "main") -> Dict[str, type]:
    if semantics not in ("main", "cold"):
        raise ValueError("semantics must be 'main' (scripts/main.py import order) or 'cold'")
    return dict(_MAIN if semantics == "main" else _COLD)


# classes the reference keeps importable by name without registering them (utils/loss/siglip_pairwise.py:36-40,
# utils/loss/siglip2_bce.py): rebound on their modules so `from utils.loss.siglip2_bce import SigLIP2BCELoss`
# (utils/loss/locca_loss.py:429) picks up the B200 implementation
_BY_NAME = {
    "utils.loss.siglip_pairwise": {"SiglipPairwiseFeatureLoss": loss.SiglipPairwiseFeatureLoss},
    "utils.loss.weighted_siglip": {"WeightedSigLIPLoss": multipos_loss.WeightedSigLIPLoss},
    "utils.loss.multi_positive_infonce": {"MultiPositiveInfoNCELoss": multipos_loss.MultiPositiveInfoNCELoss},
    "utils.loss.siglip2_bce": {"SigLIP2BCELoss": loss.SigLIP2BCELoss, "SigLIP2BCELossDDP": loss.SigLIP2BCELossDDP,
                               "SigLIP2MultiPositiveBCELoss": loss.SigLIP2MultiPositiveBCELoss},
}


def install(reference_root: str | None = None, semantics: str = "main", losses: bool = True, modules: bool = True,
            metrics: bool = True) -> dict:
    """Registers / rebinds everything; returns a report {what: [names]} of what was replaced."""
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    report = {"losses": [], "modules": [], "metrics": []}
    if losses:
        registry = importlib.import_module("utils.registry").LossRegistry
        for key, cls in loss_table(semantics).items():
            registry.register(key)(cls)
            report["losses"].append(key)
        for modname, names in _BY_NAME.items():
            try:
                mod = importlib.import_module(modname)
            except Exception:
                continue
            for n, cls in names.items():
                setattr(mod, n, cls)
                report["losses"].append(f"{modname}.{n}")
        # the runner binds WeightedSigLIPLoss by name at import (runners/video_constrative_learning_runner.py:35)
        runner_mod = sys.modules.get("runners.video_constrative_learning_runner")
        if runner_mod is not None and hasattr(runner_mod, "WeightedSigLIPLoss"):
            runner_mod.WeightedSigLIPLoss = multipos_loss.WeightedSigLIPLoss
            report["losses"].append("runners.video_constrative_learning_runner.WeightedSigLIPLoss")
    if modules:
        for modname in ("models.video_encoder",):
            mod = sys.modules.get(modname)
            if mod is None:
                try:
                    mod = importlib.import_module(modname)
                except Exception:      # backbone deps (torchvision / timm) may be absent: nothing to rebind then
                    mod = None
            if mod is not None:
                mod.Rope3D = rope_3d.Rope3D
                mod.AttentionPool = attention_pool.AttentionPool
                if hasattr(mod, "AttentionPoolWithCLS"):
                    mod.AttentionPoolWithCLS = attention_pool.AttentionPoolWithCLS
                if hasattr(mod, "EnhancedVideoAggregator"):
                    mod.EnhancedVideoAggregator = video_aggregator.EnhancedVideoAggregator
                # token pooling over the views (models/video_encoder.py:589-603): one batched pass instead of a Python loop
                # over the views; the mean branch (:603) becomes the pool kernel's uniform-weights mode
                enc = getattr(mod, "VideoEncoder", None)
                if enc is not None and hasattr(enc, "_pool_video_tokens") and not hasattr(enc._pool_video_tokens, "reference"):
                    def _pool(self, token_feats):
                        if token_feats.is_cuda and token_feats.dim() == 4:
                            return attention_pool.pool_video_tokens(self, token_feats)
                        return _pool.reference(self, token_feats)
                    _pool.reference = enc._pool_video_tokens
                    enc._pool_video_tokens = _pool
                    report["modules"].append(modname + ".VideoEncoder._pool_video_tokens")
                report["modules"].append(modname)
        for modname, names in (("models.rope_3d", ("Rope3D", "apply_rope_qk")), ("models.attention_pool", ("AttentionPool", "AttentionPoolWithCLS")),
                               ("models.video_aggregator", ("EnhancedVideoAggregator",))):
            mod = sys.modules.get(modname)
            if mod is not None:
                src = {"models.rope_3d": rope_3d, "models.attention_pool": attention_pool,
                       "models.video_aggregator": video_aggregator}[modname]
                for n in names:
                    setattr(mod, n, getattr(src, n))
                report["modules"].append(modname)
        # gated-attention MIL pooling (models/multi_instance_linear_probing.py:493-536): the two methods of the probing head
        mod = sys.modules.get("models.multi_instance_linear_probing")
        if mod is None:
            try:
                mod = importlib.import_module("models.multi_instance_linear_probing")
            except Exception:
                mod = None
        mil = getattr(mod, "MultiInstanceLinearProbing", None) if mod is not None else None
        if mil is not None and not hasattr(mil._attention_pooling, "reference"):
            def _att(self, x, mask):
                return mil_pooling.attention_pooling(self, x, mask) if x.is_cuda else _att.reference(self, x, mask)

            def _hier(self, x, mask):
                return mil_pooling.hierarchical_attention_pooling(self, x, mask) if x.is_cuda else _hier.reference(self, x, mask)
            _att.reference, _hier.reference = mil._attention_pooling, mil._hierarchical_attention_pooling
            mil._attention_pooling, mil._hierarchical_attention_pooling = _att, _hier
            report["modules"].append("models.multi_instance_linear_probing.MultiInstanceLinearProbing._attention_pooling")
    if metrics:
        for modname in ("utils.retrieval_metrics_streaming", "runners.multitask_runner"):
            mod = sys.modules.get(modname)
            if mod is None and modname == "utils.retrieval_metrics_streaming":
                try:
                    mod = importlib.import_module(modname)
                except Exception:
                    mod = None
            if mod is not None:
                mod.compute_metrics_streaming = retrieval_metrics_streaming.compute_metrics_streaming
                if hasattr(mod, "compute_recall_at_k_streaming"):
                    mod.compute_recall_at_k_streaming = retrieval_metrics_streaming.compute_recall_at_k_streaming
                report["metrics"].append(modname)
        # dense multi-label metrics (utils/retrieval_metrics.py): rebind the functions on their home module and on
        # every runner module that imported them by name (runners/multitask_runner.py:26-34, ..._runner_simple.py:19-27)
        dense = ("compute_recall_at_k", "compute_mrr", "compute_ndcg_at_k", "compute_median_rank", "compute_map",
                 "compute_similarity_matrix", "compute_embedding_norms", "compute_alignment_score")
        for modname in ("utils.retrieval_metrics", "runners.multitask_runner", "runners.video_constrative_learning_runner",
                        "runners.video_constrative_learning_runner_simple"):
            mod = sys.modules.get(modname)
            if mod is None and modname == "utils.retrieval_metrics":
                try:
                    mod = importlib.import_module(modname)
                except Exception:
                    mod = None
            if mod is not None:
                for n in dense:
                    if hasattr(mod, n):
                        setattr(mod, n, getattr(retrieval_metrics, n))
                report["metrics"].append(modname)
        # epoch-end ragged gather: the runner method keeps its signature (self, local_tensor, world_size)
        runner_mod = sys.modules.get("runners.video_constrative_learning_runner")
        cls = getattr(runner_mod, "VideoContrastiveLearningRunner", None) if runner_mod is not None else None
        if cls is not None and hasattr(cls, "_gather_tensor_along_batch"):
            from .embedding_store import gather_tensor_along_batch

            def _gather(self, local_tensor, world_size):
                return gather_tensor_along_batch(local_tensor, world_size)
            cls._gather_tensor_along_batch = _gather
            report["metrics"].append("VideoContrastiveLearningRunner._gather_tensor_along_batch")
    return report
