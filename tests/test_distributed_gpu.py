"""N > 1 on real GPUs: da4ml_b200.distributed.solve_sharded under torch.distributed.run with the NCCL backend (one process
per GPU), every problem checked against the CPU checker.  Needs at least two CUDA devices (skipped on a one-GPU box; the
CPU suite covers the same sharding logic with gloo, tests/test_distributed_cpu.py)."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
def test_solve_sharded_two_ranks_nccl(cuda_binary):
    n_dev = cuda_binary.device_info()['cuda_devices']
    if n_dev < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    world = 2 if n_dev < 4 else 4
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1', '--master-port', str(port),
                          str(ROOT / 'tests' / 'dist_gpu_worker.py')], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert f'DIST_OK world={world}' in out.stdout, (out.stdout + out.stderr)[-3000:]
