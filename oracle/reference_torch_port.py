"""ORACLE (test infrastructure, never imported by the product package).

The reference's CPU path for the headline step, op for op: the reference IS PyTorch, so its CPU implementation of
`CLIPLoss.forward` + `loss.backward()` is this sequence of ATen CPU kernels (multi-threaded through
torch.set_num_threads). /root/reference does not exist on the GPU box, so bench.py's `cpu_baseline` and
`--impl reference` legs time this transcription there (kind "port"); it is the faster and more faithful baseline than
the numpy restatement in contrastive_oracle.py (0.48 s against 0.70 s per step at N = 4096, D = 512 on 8 cores).
Only tests/ and bench.py's cpu_baseline / --impl reference legs may import this.

Pinned: tests/test_oracle_golden.py::test_torch_port_matches_reference checks it against the golden vectors that
oracle/gen_golden.py produced from the UNMODIFIED imported reference class (bit-identical loss is expected: same ops,
same order, same library).

Follows /root/reference/utils/loss/contrastive.py:140-164 (single process: gather_with_gradient is the identity, :94-101).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def clip_loss_step(video, text, log_temp: float, label_smoothing: float = 0.0):
    """One forward + backward of the reference CLIPLoss on CPU tensors (fp32). Returns (loss, dvideo, dtext, dlog_temp)."""
    video_features = torch.as_tensor(video).detach().clone().requires_grad_(True)
    text_features = torch.as_tensor(text).detach().clone().requires_grad_(True)
    lt = torch.tensor([float(log_temp)], dtype=torch.float32, requires_grad=True)
    v = F.normalize(video_features.float(), dim=-1)                                   # :146
    t = F.normalize(text_features.float(), dim=-1)                                    # :147
    similarity = torch.matmul(v, t.t())                                               # :150
    temp = torch.exp(lt.float()).clamp(min=1e-4)                                      # :153
    logits = similarity / temp                                                        # :154
    targets = torch.arange(logits.size(0), device=logits.device)                      # :157-158
    loss_v2t = F.cross_entropy(logits, targets, label_smoothing=label_smoothing)      # :161
    loss_t2v = F.cross_entropy(logits.t(), targets, label_smoothing=label_smoothing)  # :162
    loss = 0.5 * (loss_v2t + loss_t2v)                                                # :164
    loss.backward()
    return loss.detach(), video_features.grad, text_features.grad, lt.grad


def siglip_loss_step(video, text, log_temp: float, pos_mask, pos_weights, bias: float = -10.0):
    """One forward + backward of the reference SigLIPLoss (default constructor) on CPU tensors, op for op
    (/root/reference/utils/loss/contrastive.py:259-303, single process). Returns (loss, dvideo, dtext)."""
    video_features = torch.as_tensor(video).detach().clone().requires_grad_(True)
    text_features = torch.as_tensor(text).detach().clone().requires_grad_(True)
    lt = torch.tensor([float(log_temp)], dtype=torch.float32, requires_grad=True)
    b = torch.tensor(float(bias), requires_grad=True)
    v = F.normalize(video_features.float(), dim=-1)                                   # :259
    t = F.normalize(text_features.float(), dim=-1)                                    # :260
    similarity = torch.matmul(v, t.t())                                               # :263
    temp = torch.exp(lt.float()).clamp(min=1e-4)                                      # :266
    logits = (similarity / temp + b).clamp(-30, 30)                                   # :267-270
    targets = torch.as_tensor(pos_mask).float().clamp(0, 1)                           # :274
    weight_matrix = torch.ones_like(targets)                                          # :283 (negative_weight = 1)
    positive_contrib = torch.as_tensor(pos_weights).float() * 1.0                     # :287-289 (positive_weight = 1)
    weight_matrix = torch.where(targets > 0.5, positive_contrib, weight_matrix)       # :298
    loss = F.binary_cross_entropy_with_logits(logits, targets, weight=weight_matrix, reduction="mean")   # :301-303
    loss.backward()
    return loss.detach(), video_features.grad, text_features.grad


@torch.no_grad()
def retrieval_metrics_step(video, text, gt, k_values=(1, 5, 10), video_chunk_size: int = 2048,
                           text_chunk_size: int = 8192):
    """The reference's CPU path for recall@k + MRR over a [n_videos] x [n_texts] sweep, op for op
    (/root/reference/utils/retrieval_metrics_streaming.py:35-101 with device='cpu', then the MRR loop :141-172): chunked
    matmul / topk / cat / topk / gather for the recalls, a second chunked matmul + full argsort + per-row Python loop for
    the reciprocal ranks. Inputs are used as given (the caller normalises, as compute_metrics_streaming :128-131 does)."""
    video_features = torch.as_tensor(video).float()
    text_features = torch.as_tensor(text).detach().float()
    ground_truth_indices = torch.as_tensor(gt)
    n_videos, n_texts = video_features.size(0), text_features.size(0)
    recalls = {k: 0 for k in k_values}
    k_max = max(k_values)
    for v_start in range(0, n_videos, video_chunk_size):                                   # :47
        video_chunk = video_features[v_start:min(v_start + video_chunk_size, n_videos)]
        best_scores = best_indices = None
        for t_start in range(0, n_texts, text_chunk_size):                                 # :56
            text_chunk = text_features[t_start:min(t_start + text_chunk_size, n_texts)]
            similarity = torch.matmul(video_chunk, text_chunk.t())                         # :61
            chunk_scores, chunk_indices = torch.topk(similarity, k=min(k_max, similarity.size(1)), dim=1)   # :64-65
            chunk_indices = chunk_indices + t_start                                        # :68
            if best_scores is None:
                best_scores, best_indices = chunk_scores, chunk_indices
            else:
                all_scores = torch.cat([best_scores, chunk_scores], dim=1)                 # :76-77
                all_indices = torch.cat([best_indices, chunk_indices], dim=1)
                best_scores, top_idx = torch.topk(all_scores, k=min(k_max, all_scores.size(1)), dim=1)      # :80-81
                best_indices = torch.gather(all_indices, 1, top_idx)                       # :82
        chunk_gt = ground_truth_indices[v_start:v_start + video_chunk.size(0)]
        for k in k_values:                                                                 # :89-93
            if k <= best_indices.size(1):
                recalls[k] += (best_indices[:, :k] == chunk_gt.unsqueeze(1)).any(dim=1).sum().item()
    metrics = {f"Recall@{k}": (recalls[k] / n_videos) * 100 for k in k_values}            # :99
    mrr_sum = 0.0
    for v_start in range(0, n_videos, video_chunk_size):                                   # :143
        video_chunk = video_features[v_start:min(v_start + video_chunk_size, n_videos)]
        chunk_gt = ground_truth_indices[v_start:v_start + video_chunk.size(0)]
        all_scores = [torch.matmul(video_chunk, text_features[t:min(t + text_chunk_size, n_texts)].t())
                      for t in range(0, n_texts, text_chunk_size)]                         # :151-155
        sorted_indices = torch.argsort(torch.cat(all_scores, dim=1), dim=1, descending=True)   # :158-161
        for i in range(chunk_gt.size(0)):                                                  # :162-166
            rank = (sorted_indices[i] == chunk_gt[i].item()).nonzero(as_tuple=True)[0]
            if rank.numel() > 0:
                mrr_sum += 1.0 / (rank[0].item() + 1)
    metrics["MRR_V2T"] = mrr_sum / n_videos                                                # :172
    return metrics
