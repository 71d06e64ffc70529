"""CPU, the literal drop-in scenario (only where the reference checkout exists: the build container): the UNMODIFIED
reference's own call path — `LossRegistry.get(name)()`, `Loss(loss_type=...)`, `loss.run(video_features=, text_features=,
log_temp=)` (projects/contrastive_pretraining_project.py:204-206, runners/video_constrative_learning_runner.py:1317-1321) and
`utils.retrieval_metrics_streaming.compute_metrics_streaming` — evaluated with the stock classes first, then again after
`deepcoro_clip_b200.install()` on the same inputs, with the package running on the emulated library (shipped CUDA-core
kernels, modelled tile kernels: tests/emul/loss_emul.cpp). Run in a child process: it imports the reference tree and patches
the package for the whole interpreter."""
import os

import pytest
import torch.multiprocessing as mp

REF = "/root/reference"


def _child(q):
    import sys
    import numpy as np
    import torch
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    from tests.test_emulated_losses import build_emul, patch_package
    from tests.test_emulated_retrieval import patch_retrieval
    import utils.loss.typing as ult                      # pulls utils.loss.losses first, as scripts/main.py does
    from utils.registry import LossRegistry, register_submodules
    from utils.enums import SubmoduleType
    register_submodules(SubmoduleType.LOSS)
    import utils.retrieval_metrics_streaming as urms

    g = torch.Generator().manual_seed(0)
    v = torch.randn(64, 512, generator=g)
    t = 0.4 * v + torch.randn(64, 512, generator=g)
    lt0 = float(np.log(0.07))
    results = {}

    def run(tag):
        out = {}
        for key in ("clip", "contrastive", "siglip"):
            vv, tt = v.clone().requires_grad_(True), t.clone().requires_grad_(True)
            lt = torch.tensor([lt0], requires_grad=True)
            mod = LossRegistry.get(key)()                # zero-argument constructor, as the project does
            loss = ult.Loss(loss_type=mod).run(video_features=vv, text_features=tt, log_temp=lt)
            loss.backward()
            out[key] = (type(mod).__module__, loss.item(), vv.grad.numpy().copy(), tt.grad.numpy().copy(), lt.grad.item())
        m = urms.compute_metrics_streaming(v, t, torch.arange(64), k_values=[1, 5, 10], video_chunk_size=32,
                                           text_chunk_size=16, device="cpu")
        out["metrics"] = {k: float(x) for k, x in m.items()}
        results[tag] = out

    run("stock")
    import deepcoro_clip_b200
    so = build_emul()
    patch_package(so)
    patch_retrieval(so)
    deepcoro_clip_b200.install(REF)
    run("b200")
    q.put(results)


@pytest.mark.skipif(not os.path.isdir(REF + "/utils"), reason="/root/reference not present (GPU box)")
def test_reference_call_path_before_and_after_install():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_child, args=(q,))
    p.start()
    res = q.get(timeout=900)
    p.join(60)
    stock, ours = res["stock"], res["b200"]
    import numpy as np
    for key in ("clip", "contrastive", "siglip"):
        m0, l0, dv0, dt0, dlt0 = stock[key]
        m1, l1, dv1, dt1, dlt1 = ours[key]
        assert m0.startswith("utils.loss") and m1.startswith("deepcoro_clip_b200"), (key, m0, m1)   # the swap happened
        assert abs(l1 - l0) <= 1e-5 * abs(l0), (key, l1, l0)
        rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
        assert rel(dv1, dv0) <= 2e-3 and rel(dt1, dt0) <= 2e-3, key
        assert abs(dlt1 - dlt0) <= 2e-3 * max(abs(dlt0), 1e-4), key
    for k, x in stock["metrics"].items():
        y = ours["metrics"][k]
        if k.startswith("Recall@") or k == "median_rank":
            assert y == x, (k, y, x)
        else:
            assert abs(y - x) <= 1e-6 * max(1.0, abs(x)), (k, y, x)
    assert set(stock["metrics"]) == set(ours["metrics"])
