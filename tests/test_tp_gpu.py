"""Tensor parallelism over 2 GPUs of one box (skipped when fewer are visible): head-sharded attention + row-sharded MLP
with NCCL all-reduce after o_proj and down_proj, checked against the oracle and for cross-rank agreement."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_tp2_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "tp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "TP2 tiny ok" in r.stdout and "TP2 mid ok" in r.stdout and "TP2 vl ok" in r.stdout, r.stdout[-2000:]
