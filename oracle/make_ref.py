"""ORACLE (test infrastructure). Recipe for oracle/_ref/: the UNMODIFIED reference files of the hot path, copied from
/root/reference where they lie (build container only) so that they travel to the GPU box with the gpurun snapshot.

    python oracle/make_ref.py            # no-op when /root/reference is absent (GPU box: uses what the snapshot carried)

oracle/_ref/ is git-ignored (never part of the history, never a product source) and NOT gpurun-ignored. With it,
``bench.py --impl reference`` and the ``cpu_baseline`` legs time the reference's own classes / functions on the host cores
(kind "reference"); without it they fall back to the op-for-op transcription oracle/reference_torch_port.py (kind "port").
The reference is pure Python on PyTorch: there is nothing to compile, the files are imported as they are.
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
# namespace packages (the reference ships no utils/__init__.py): only the files the path needs, nothing else of the tree
FILES = [
    "utils/enums.py", "utils/registry.py",                       # LossRegistry / LossType (imported by the loss modules)
    "utils/loss/contrastive.py",                                 # CLIPLoss, SigLIPLoss
    "utils/retrieval_metrics_streaming.py",                      # compute_recall_at_k_streaming, compute_metrics_streaming
    "models/rope_3d.py", "models/attention_pool.py", "models/video_aggregator.py",
]


def make_ref() -> bool:
    if not REF.exists():
        return OUT.exists()
    for rel in FILES:
        dst = OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / rel, dst)
    return True


def import_ref():
    """Puts oracle/_ref on sys.path (in front) and returns True when the unmodified reference files are available."""
    if not (OUT / "utils" / "loss" / "contrastive.py").exists():
        return False
    if str(OUT) not in sys.path:
        sys.path.insert(0, str(OUT))
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if make_ref() else "no /root/reference and no oracle/_ref: the torch port is used instead")
