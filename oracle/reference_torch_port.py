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
