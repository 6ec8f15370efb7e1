"""Batched front door: dedupe + grouping logic on CPU (checker as solver), CUDA path under -m gpu."""
import numpy as np
import pytest
from conftest import assert_stage_equal, int_matrix

from da4ml_b200.batching import CmvmCall, solve_calls


def _cpu_solver(kernels, qintervals=None, latencies=None, **opts):
    from oracle import port

    _cpu_solver.n_problems += len(kernels)
    out = []
    for i, k in enumerate(kernels):
        q = None if qintervals is None else qintervals[i]
        l = None if latencies is None else latencies[i]
        out.append(port.solve(k, qintervals=q, latencies=l, **opts))
    return out


def _calls():
    W, V = int_matrix(6, 5, 4, 1), int_matrix(5, 7, 6, 2)
    q1 = [(-8.0, 7.0, 1.0)] * 6
    q2 = [(-4.0, 3.0, 0.5)] * 6
    return [
        CmvmCall(W, q1, [0.0] * 6),
        CmvmCall(W, q2, [0.0] * 6),
        CmvmCall(W.copy(), q1, [0.0] * 6),  # duplicate of call 0
        CmvmCall(V, None, None, dict(hard_dc=2, adder_size=1)),
        CmvmCall(V, None, None),  # same matrix, other options -> another group
        CmvmCall(W, q1, [0.0] * 6),  # duplicate again
    ]


def test_solve_calls_dedupes_and_groups():
    from oracle import port

    _cpu_solver.n_problems = 0
    calls = _calls()
    res = solve_calls(calls, solver=_cpu_solver)
    assert _cpu_solver.n_problems == 4  # six calls, four distinct problems
    assert res[0] is res[2] is res[5]
    for c, r in zip(calls, res):
        want = port.solve(c.kernel, qintervals=c.qintervals, latencies=c.latencies, **(c.options or {}))
        for a, b in zip(r, want, strict=True):
            assert_stage_equal(a, b)
    with pytest.raises(TypeError):
        solve_calls([CmvmCall(calls[0].kernel, options=dict(offload_fn=None))], solver=_cpu_solver)


@pytest.mark.gpu
def test_solve_calls_cuda(cuda_binary):
    calls = _calls()
    res = solve_calls(calls)  # default: CUDA batch solver, returns Pipelines
    for c, r in zip(calls, res):
        want = cuda_binary.solve(np.ascontiguousarray(c.kernel), qintervals=c.qintervals, latencies=c.latencies, **(c.options or {}))
        assert r == want
        keep = np.ones(c.kernel.shape[0], bool)
        assert np.array_equal(r.kernel[keep], c.kernel[keep])
