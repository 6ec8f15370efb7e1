"""N>1 path on CPU: world_size-2 gloo run of the sharding + gather logic, with the CPU checker standing in for
the CUDA batch solver (tests may do that; the product default is the CUDA solver)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import assert_stage_equal, int_matrix

from da4ml_b200.distributed import shard_assignment, solve_sharded


def test_shard_assignment_balances_and_is_deterministic():
    w = [100, 1, 1, 50, 49, 1, 1, 1]
    a = shard_assignment(w, 2)
    assert a == shard_assignment(w, 2)
    assert sorted(a[0] + a[1]) == list(range(8))
    loads = [sum(w[i] for i in s) for s in a]
    assert max(loads) - min(loads) <= 2
    assert shard_assignment([3, 2, 1], 8)[3:] == [[], [], [], [], []]


def _cpu_solver(kernels, **opts):
    from oracle import port

    return [port.solve(k, **opts) for k in kernels]


def _worker(rank, world, port_no, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        kernels = [int_matrix(6 + i % 3, 5 + i % 4, 4, i) for i in range(7)]
        res = solve_sharded(kernels, solver=_cpu_solver, search_all_decompose_dc=False, decompose_dc=-1)
        q.put((rank, [[{k: np.asarray(v) for k, v in st.items() if k != 'counters'} for st in r] for r in res]))
    finally:
        dist.destroy_process_group()


def test_solve_sharded_world2_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port_no = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    kernels = [int_matrix(6 + i % 3, 5 + i % 4, 4, i) for i in range(7)]
    want = _cpu_solver(kernels, search_all_decompose_dc=False, decompose_dc=-1)
    for rank in (0, 1):
        assert len(out[rank]) == len(want)
        for got, ref_ in zip(out[rank], want):
            for a, b in zip(got, ref_, strict=True):
                assert_stage_equal(a, b)
