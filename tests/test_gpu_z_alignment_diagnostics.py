"""GPU parity: per-step alignment diagnostics (logging half) from one sweep of the fused logits forward
against the transcribed runner lines (runners/video_constrative_learning_runner.py:1323-1335; goldens align_*.npz) and
the numpy oracle. Tolerance: the runner computes these in fp32 — 2e-5 relative-or-absolute on each scalar (bf16x3
operands at these sizes); plain bf16 operands are compared with the oracle fed the same bf16-rounded operands."""
import numpy as np
import pytest
import torch

from oracle import contrastive_oracle as co
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = (("alignment_cosine", "cosine_f64"), ("alignment_logprob", "logprob_f64"), ("alignment_prob", "prob_f64"))


@pytest.mark.parametrize("name", ["align_b64_d512", "align_siglip_b130_d96", "align_b300_d200"])
def test_alignment_diagnostics_match_runner_golden(name):
    from deepcoro_clip_b200 import _lib
    from deepcoro_clip_b200.diagnostics import alignment_diagnostics
    g = np.load(GOLDEN / f"{name}.npz")
    v = torch.tensor(g["video"], dtype=torch.float32, device=DEV, requires_grad=True)
    t = torch.tensor(g["text"], dtype=torch.float32, device=DEV)
    lt = torch.tensor(g["log_temp"].astype(np.float32), device=DEV, requires_grad=True)
    before = _lib.LAUNCHES
    r = alignment_diagnostics(v, t, lt, use_siglip=bool(g["use_siglip"]))
    assert _lib.LAUNCHES - before == 6            # 2 normalise, dyn_prep, forward sweep (+ its gated stable twin), scalar tail
    for key, ref in KEYS:
        x = r[key]
        assert x.ndim == 0 and x.device.type == "cuda" and not x.requires_grad
        ref = float(g[ref])
        assert abs(x.item() - ref) <= 2e-5 * max(1.0, abs(ref)), (key, x.item(), ref)


def test_alignment_diagnostics_bf16_operands_large_batch():
    """Plain bf16 operands (the precision the training step itself uses at large batch), B = 2048, python-float log_temp:
    against the oracle on the bf16-rounded normalised operands."""
    from deepcoro_clip_b200.diagnostics import alignment_diagnostics
    gen = torch.Generator().manual_seed(63)
    B, D = 2048, 512
    text = torch.randn(B, D, generator=gen)
    video = 0.5 * text + torch.randn(B, D, generator=gen)
    lt = float(np.log(0.07))
    r = alignment_diagnostics(video.to(DEV), text.to(DEV), lt, precision="bf16")
    vh = torch.nn.functional.normalize(video, dim=1).bfloat16().double().numpy()
    th = torch.nn.functional.normalize(text, dim=1).bfloat16().double().numpy()
    o = co.alignment_diagnostics(vh, th, lt)      # re-normalising the rounded rows moves them by <= 2^-9 relative
    assert abs(r["alignment_cosine"].item() - o["alignment_cosine"]) <= 2e-3
    assert abs(r["alignment_logprob"].item() - o["alignment_logprob"]) <= 2e-2 * max(1.0, abs(o["alignment_logprob"]))


def test_alignment_diagnostics_errors():
    from deepcoro_clip_b200._lib import B200ClipError
    from deepcoro_clip_b200.diagnostics import alignment_diagnostics
    with pytest.raises(ValueError):
        alignment_diagnostics(torch.zeros(4, 64, device=DEV), torch.zeros(5, 64, device=DEV), 0.0)
    with pytest.raises(B200ClipError):
        alignment_diagnostics(torch.zeros(4, 64), torch.zeros(4, 64), 0.0)
