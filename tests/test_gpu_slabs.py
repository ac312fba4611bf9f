"""Multi-GPU slab decomposition (needs >= 2 GPUs; skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run under `gpurun --gpus 2`)")
def test_two_slabs_neighbour_counts_and_migration():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "slab_worker.py")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert "SLAB_OK" in proc.stdout and "mismatches=0" in proc.stdout, proc.stdout[-2000:]
