"""GPU parity of the logits-level multi-positive softmax losses (deepcoro_clip_b200/multipos_loss.py)
against goldens produced by the UNMODIFIED reference classes (fp32 autograd) and the numpy oracle.
Tolerances: loss 1e-5 relative (fp32 accumulation), dlogits 1e-4 relative Frobenius (fp32 elementwise, __expf)."""
import numpy as np
import pytest
import torch

from oracle import contrastive_oracle as co
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _run(fn, logits):
    L = torch.tensor(logits, dtype=torch.float32, device=DEV, requires_grad=True)
    loss = fn(L)
    assert loss.ndim == 0 and loss.requires_grad
    (loss * 1.5).backward()                       # upstream scale reaches the kernel through grad_out
    torch.cuda.synchronize()
    return loss.item(), L.grad.cpu().numpy() / 1.5


@pytest.mark.parametrize("name", ["multipos_48x64", "multipos_130x37"])
def test_multipos_losses_match_reference_golden(name):
    from deepcoro_clip_b200.multipos_loss import MultiPositiveInfoNCELoss, WeightedSigLIPLoss
    g = np.load(GOLDEN / f"{name}.npz")
    mk = torch.tensor(g["mask"], device=DEV); pw = torch.tensor(g["pos_weights"], device=DEV)
    cases = {
        "wsl": lambda L: WeightedSigLIPLoss()(L, mk * pw - 0.2 * (1 - mk)),
        "mpi_mean": lambda L: MultiPositiveInfoNCELoss()(L, mk, pw),
        "mpi_sum_noweights": lambda L: MultiPositiveInfoNCELoss(reduction="sum")(L, mk),
        "mpi_imp_mean": lambda L: MultiPositiveInfoNCELoss(use_importance_weighting=True)(L, mk, pw),
        "mpi_imp_sum_noweights": lambda L: MultiPositiveInfoNCELoss(reduction="sum", use_importance_weighting=True)(L, mk),
    }
    for key, fn in cases.items():
        loss, dl = _run(fn, g["logits"])
        ref = float(g[key + "_loss"])
        assert abs(loss - ref) <= 1e-5 * abs(ref), (key, loss, ref)
        assert _rel(dl, g[key + "_dlogits"]) <= 1e-4, key


@pytest.mark.parametrize("N,M", [(1000, 1300), (4096, 515), (33, 8192)])
def test_multipos_vs_oracle_and_edges(N, M):
    """Larger / ragged shapes, strided (non-contiguous row pitch) inputs, rows and columns without positives, and the
    all-negative batch (MultiPositiveInfoNCELoss returns 0 with a zero gradient)."""
    from deepcoro_clip_b200.multipos_loss import MultiPositiveInfoNCELoss, WeightedSigLIPLoss
    rng = np.random.default_rng(N + M)
    logits = (rng.standard_normal((N, M)) * 5).astype(np.float32)
    mask = (rng.random((N, M)) < 0.01).astype(np.float32)
    mask[::7] = 0.0
    mask[:, ::5] = 0.0
    pw = rng.choice([1.0, 1.5, 2.5, 3.0], size=(N, M)).astype(np.float32)
    mk_t = torch.tensor(mask, device=DEV); pw_t = torch.tensor(pw, device=DEV)
    loss, dl = _run(lambda L: MultiPositiveInfoNCELoss()(L, mk_t, pw_t), logits)
    o = co.multipos_softmax_loss(logits, mask * pw, mask=mask, mode="infonce")
    assert abs(loss - o["loss"]) <= 1e-5 * abs(o["loss"]) and _rel(dl, o["dlogits"]) <= 1e-4
    # strided views: logits / weights as column slices of wider buffers
    wide = torch.zeros(N, M + 13, device=DEV); wide[:, 5:5 + M] = torch.tensor(logits, device=DEV)
    wl = wide[:, 5:5 + M].detach().requires_grad_(True)
    wpos = torch.zeros(N, M + 3, device=DEV); wpos[:, :M] = mk_t * pw_t
    l2 = WeightedSigLIPLoss()(wl, wpos[:, :M])
    l2.backward()
    o2 = co.multipos_softmax_loss(logits, mask * pw, mode="weighted_siglip")
    assert abs(l2.item() - o2["loss"]) <= 1e-5 * abs(o2["loss"])
    assert _rel(wl.grad.cpu().numpy(), o2["dlogits"]) <= 1e-4
    # no positives anywhere
    z, dz = _run(lambda L: MultiPositiveInfoNCELoss()(L, torch.zeros_like(mk_t)), logits)
    assert z == 0.0 and float(np.abs(dz).max()) == 0.0
    with pytest.raises(ValueError):
        WeightedSigLIPLoss()(torch.zeros(3, 4, device=DEV), torch.zeros(3, 5, device=DEV))
    with pytest.raises(ValueError):
        MultiPositiveInfoNCELoss(reduction="median")
