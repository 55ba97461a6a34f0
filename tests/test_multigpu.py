"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped otherwise): a view-sharded step equals, bit for bit, a
single-GPU Adam step on the per-rank gradients summed in rank order -- for the NCCL all-reduce baseline, the fused
peer-memory kernel with flag barriers (default), the same kernel with NCCL brackets, and the flag-barrier variant
with the whole iteration (exchange included) replayed from per-rank CUDA graphs."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["allreduce", "fused_p2p", "fused_p2p_nccl", "fused_p2p+graph"])
def test_two_gpu_step_equals_single_gpu_accumulation(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, GSB_TEST_MODE=mode)
    port = 29700 + os.getpid() % 100 + {"allreduce": 0, "fused_p2p": 1, "fused_p2p_nccl": 2, "fused_p2p+graph": 3}[mode]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "_mgpu_worker.py")], env=env, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
