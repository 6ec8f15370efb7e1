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


def _hetero(i, n_in):
    rng = np.random.default_rng(100 + i)
    q = [(-float(2 ** rng.integers(0, 6)), float(2 ** rng.integers(0, 6)) - 0.5, 0.5) for _ in range(n_in)]
    lat = [float(v) for v in rng.integers(0, 3, n_in)]
    return q, lat


def _cpu_solver_per_problem(kernels, qintervals=None, latencies=None, **opts):
    from oracle import port

    assert len(qintervals) == len(kernels) == len(latencies)
    return [port.solve(k, qintervals=q, latencies=l, **opts) for k, q, l in zip(kernels, qintervals, latencies)]


def _worker(rank, world, port_no, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port_no)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        kernels = [int_matrix(6 + i % 3, 5 + i % 4, 4, i) for i in range(7)]
        res = solve_sharded(kernels, solver=_cpu_solver, search_all_decompose_dc=False, decompose_dc=-1)
        # per-problem intervals / latencies must follow their problems to whichever rank solves them
        qs, ls = zip(*[_hetero(i, k.shape[0]) for i, k in enumerate(kernels)])
        res2 = solve_sharded(kernels, solver=_cpu_solver_per_problem, qintervals=list(qs), latencies=list(ls), search_all_decompose_dc=False, decompose_dc=-1)
        pack = lambda rs: [[{k: np.asarray(v) for k, v in st.items() if k != 'counters'} for st in r] for r in rs]  # noqa: E731
        q.put((rank, pack(res), pack(res2)))
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
    got_all = [q.get(timeout=120) for _ in procs]
    out = {r: a for r, a, _ in got_all}
    out2 = {r: b for r, _, b in got_all}
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
    qs, ls = zip(*[_hetero(i, k.shape[0]) for i, k in enumerate(kernels)])
    want2 = _cpu_solver_per_problem(kernels, qintervals=list(qs), latencies=list(ls), search_all_decompose_dc=False, decompose_dc=-1)
    for rank in (0, 1):
        for got, ref_ in zip(out2[rank], want2, strict=True):
            for a, b in zip(got, ref_, strict=True):
                assert_stage_equal(a, b)
