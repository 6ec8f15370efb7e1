"""Worker of tests/test_distributed_gpu.py (one process per GPU under torch.distributed.run, NCCL): the BASELINE config 5
stand-in stack through the product's multi-GPU entry point (da4ml_b200.distributed.solve_sharded, default CUDA solver), every
layer compared with the CPU checker on rank 0."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

import oracle  # noqa: E402
from conftest import assert_stage_equal, int_matrix  # noqa: E402

import da4ml_b200._binary as B  # noqa: E402
from da4ml_b200.distributed import shard_assignment, solve_sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    B.set_stream(torch.cuda.current_stream().cuda_stream)
    shapes = [(16, 64), (64, 64), (64, 32), (32, 32), (32, 5), (24, 48), (48, 24)]
    layers = [int_matrix(a, b, 6, 40 + i) for i, (a, b) in enumerate(shapes)]
    rng = np.random.default_rng(3)
    qs = [[(-float(2 ** rng.integers(1, 7)), float(2 ** rng.integers(1, 7)) - 1.0, 1.0) for _ in range(a)] for a, _ in shapes]
    ls = [[float(v) for v in rng.integers(0, 2, a)] for a, _ in shapes]
    got = solve_sharded(layers, hard_dc=2, qintervals=qs, latencies=ls)  # gathered on every rank
    assert len(got) == len(layers)
    mine = shard_assignment([1.0] * len(layers), world)  # (every rank got work: 7 problems over `world` ranks)
    assert all(len(m) > 0 for m in mine)
    if rank == 0:
        mod, kind = oracle.best()
        for W, q, l, r in zip(layers, qs, ls, got):
            want = mod.solve(W, hard_dc=2, qintervals=q, latencies=l)
            for i, (a, b) in enumerate(zip(r.stages, want, strict=True)):
                assert_stage_equal(a, b, f'{W.shape} stage{i} ')
        print(f'DIST_OK world={world} checker={kind} adders={[r.n_adders for r in got]}', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
