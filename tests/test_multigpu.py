"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped otherwise): NCCL view-sharded step == single-GPU
accumulated step, for both exchange modes (NCCL all-reduce, fused P2P reduce-scatter+Adam+all-gather)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["allreduce", "fused_p2p"])
def test_two_gpu_step_equals_single_gpu_accumulation(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, GSB_TEST_MODE=mode)
    port = 29700 + os.getpid() % 100 + (0 if mode == "allreduce" else 1)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "_mgpu_worker.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
