"""Multi-GPU parity on real devices (skipped on boxes with fewer than 2 GPUs; the CPU-side contract is covered by
tests/test_parallel_gloo.py): 2 ranks x E envs over NCCL must reproduce ONE process with 2E envs — replicas bit-identical
across ranks, logged losses / KL equal to 1e-3 relative, weights within the chaotic-trajectory gate (scripts/mgpu_parity.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
def test_two_ranks_reproduce_one_process_with_twice_the_envs():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "mgpu_parity.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=540)
    lines = [l for l in out.stdout.splitlines() if l.startswith("[")]
    assert out.returncode == 0 and len(lines) == 3 and all(l.rstrip().endswith("OK") for l in lines), out.stdout[-3000:] + out.stderr[-2000:]
