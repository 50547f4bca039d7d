"""Real multi-GPU data-parallel equivalence: two processes, two GPUs, NCCL.  Skipped on boxes with one GPU
(run it with ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu``); the CPU suite covers the same host logic with gloo."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("overlap", ["0", "1"])
def test_two_rank_nccl_training_equals_single_process(overlap):
    """overlap = 1: the all-reduce runs unit by unit under the deferred weight-gradient GEMMs inside ONE CUDA graph (parallel.GradSync);
    0 (default): graph, one all-reduce of the flat buffer, graph."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "dp_worker.py")]
    env = dict(os.environ, PIVP_DP_OVERLAP=overlap)
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-4000:]
