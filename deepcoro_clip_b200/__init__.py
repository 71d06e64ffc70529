"""deepcoro_clip_b200 — B200-native (sm_100a) contrastive head for DeepCORO_CLIP: CLIP / SigLIP losses, streaming
retrieval metrics, Rope3D, AttentionPool and the multi-view query pool behind the reference's own interfaces.
Hand-written CUDA (tcgen05 / TMEM / TMA) behind a C ABI (include/b200clip.h); no CPU or PyTorch fallback."""
from .attention_pool import AttentionPool, AttentionPoolWithCLS, pool_video_tokens, token_mean_pool
from .diagnostics import alignment_diagnostics
from .embedding_store import EmbeddingStore, epoch_end_retrieval_metrics, gather_tensor_along_batch
from .host_pipeline import GraphedLossStep, HostBatchPrefetcher
from .install import install, loss_table
from .loss import (CLIPLoss, ContrastiveLoss, ContrastiveLossDDP, InfoNCELoss, SigLIP2BCELoss, SigLIP2BCELossDDP,
                   SigLIP2MultiPositiveBCELoss, SigLIPLoss, SiglipLoss, SiglipLossDDP, SiglipPairwiseFeatureLoss,
                   clip_loss)
from . import retrieval_metrics
from .mil_pooling import GatedAttentionPooling, gated_attention_pool
from .multipos_loss import MultiPositiveInfoNCELoss, WeightedSigLIPLoss, inline_multipositive_loss
from .retrieval_metrics_streaming import (compute_metrics_streaming, compute_recall_at_k_streaming, inference_topk_indices,
                                          streaming_topk, top5_predictions)
from .rope_3d import Rope3D, apply_rope_qk
from .video_aggregator import EnhancedVideoAggregator, query_pool

__all__ = ["AttentionPool", "GatedAttentionPooling", "gated_attention_pool", "AttentionPoolWithCLS", "CLIPLoss", "GraphedLossStep", "HostBatchPrefetcher", "ContrastiveLoss", "ContrastiveLossDDP", "EmbeddingStore", "EnhancedVideoAggregator",
           "InfoNCELoss", "MultiPositiveInfoNCELoss", "inline_multipositive_loss", "Rope3D", "SigLIP2BCELoss", "SigLIP2BCELossDDP", "SigLIP2MultiPositiveBCELoss", "SigLIPLoss",
           "SiglipLoss", "SiglipLossDDP", "SiglipPairwiseFeatureLoss", "WeightedSigLIPLoss", "alignment_diagnostics", "apply_rope_qk", "clip_loss",
           "compute_metrics_streaming", "compute_recall_at_k_streaming", "epoch_end_retrieval_metrics", "gather_tensor_along_batch", "install", "loss_table", "query_pool",
           "streaming_topk", "inference_topk_indices", "top5_predictions", "token_mean_pool", "pool_video_tokens"]
