"""Multi-GPU parity (NCCL, one rank per GPU): row-slab CLIP, row-parallel SigLIP and text-sharded retrieval against the
single-process oracle. Needs >= 2 visible GPUs; skipped otherwise (the 2-rank host logic is covered on CPU with gloo in
test_dist_plan_gloo.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_two_rank_nccl_parity():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / "gpu_check_dist.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist check ok" in r.stdout
