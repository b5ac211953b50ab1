"""Utterance-sharded driver on real GPUs (needs >= 2 visible devices; skipped on a one-GPU box)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_gather_over_nvlink_matches_local_compute(built_lib):
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"MULTI_GPU_OK {n}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
