"""Batched front door for traced matmuls.

The reference issues one ``solve`` per left-hand row of a traced ``(N, F) @ W`` (``trace/fixed_variable_array.py:368-371``
calling ``cmvm()`` ``:58-82``): the same constant matrix, possibly different input intervals / latencies per row, all
independent.  ``solve_calls`` takes such a list of calls, removes exact duplicates (same matrix bytes, intervals,
latencies and options solve to the same graph -- the solver is deterministic), groups the rest by option set and sends
each group to the GPU as one batch.
"""

from __future__ import annotations

from collections.abc import Callable, Sequence
from typing import NamedTuple

import numpy as np

_OPT_KEYS = ('method0', 'method1', 'hard_dc', 'decompose_dc', 'adder_size', 'carry_size', 'search_all_decompose_dc')


class CmvmCall(NamedTuple):
    """Arguments of one ``da4ml.cmvm.solve`` call (``bindings.cc:235-248``)."""

    kernel: np.ndarray
    qintervals: Sequence[tuple[float, float, float]] | None = None
    latencies: Sequence[float] | None = None
    options: dict | None = None  # solver_options_t minus offload_fn


def _key(call: CmvmCall):
    k = np.ascontiguousarray(call.kernel, dtype=np.float32)
    q = None if call.qintervals is None else tuple(tuple(float(v) for v in qi) for qi in call.qintervals)
    l = None if call.latencies is None else tuple(float(v) for v in call.latencies)
    o = tuple(sorted((call.options or {}).items()))
    return (k.shape, k.tobytes(), q, l, o)


def solve_calls(calls: Sequence[CmvmCall], solver: Callable | None = None) -> list:
    """Solve many ``solve`` calls at once; returns one result per call, in order.

    ``solver(kernels, qintervals=..., latencies=..., **options) -> list`` defaults to the CUDA batch solver
    (``da4ml_b200._binary.solve_batch``); identical calls share one solve (and one result object).
    """
    if solver is None:
        from ._binary import solve_batch as solver  # noqa: PLC0415
    uniq: dict = {}
    order: list = []
    for c in calls:
        bad = set(c.options or {}) - set(_OPT_KEYS)
        if bad:
            raise TypeError(f'unsupported solver options for batching: {sorted(bad)}')
        k = _key(c)
        if k not in uniq:
            uniq[k] = c
        order.append(k)
    groups: dict = {}
    for k, c in uniq.items():
        groups.setdefault(k[4], []).append((k, c))
    solved: dict = {}
    for opt_items, members in groups.items():
        kernels = [np.ascontiguousarray(c.kernel, dtype=np.float32) for _, c in members]
        qs = [c.qintervals for _, c in members]
        ls = [c.latencies for _, c in members]
        kw = dict(opt_items)
        res = solver(kernels, qintervals=qs if any(q is not None for q in qs) else None, latencies=ls if any(l is not None for l in ls) else None, **kw)
        for (k, _), r in zip(members, res, strict=True):
            solved[k] = r
    return [solved[k] for k in order]


class _Placeholder:
    """Stands in for a solution while calls are being collected: maps inputs to ``n_out`` valid outputs of the caller's own
    kind (the first input, repeated), so that tracing code runs through; whatever it builds from them is thrown away."""

    def __init__(self, n_out: int):
        self.n_out = n_out

    def __call__(self, inp, *a, **k):
        first = np.asarray(inp, dtype=object).ravel()[0]
        out = np.empty(self.n_out, dtype=object)
        out[:] = [first] * self.n_out
        return out


class SolveBatcher:
    """Turns the one-``solve``-per-row pattern of traced matmuls (reference ``trace/fixed_variable_array.py:361-373``: a loop
    over the left-hand rows, each calling ``cmvm()`` -> ``solve()``) into ONE batched GPU solve without touching the
    calling code: hand ``batcher.solve`` to the caller in place of ``da4ml.cmvm.solve`` and run the traced function through
    ``batcher.run(fn)``.  ``run`` executes ``fn`` twice: the first pass only records the calls (placeholder solutions), then
    all recorded calls are solved at once (``solve_calls``: identical calls once, one GPU batch per option set), and the
    second pass gets the real solutions in call order."""

    def __init__(self, types_module=None, solver: Callable | None = None):
        self.types_module = types_module
        self.solver = solver
        self.calls: list[CmvmCall] = []
        self.raw: list = []  # RawPipeline of each call (after run)
        self._mode = 'idle'
        self._next = 0

    def solve(self, kernel, method0='wmc', method1='auto', hard_dc=-1, decompose_dc=-2, qintervals=None, latencies=None,
              adder_size=-1, carry_size=-1, search_all_decompose_dc=True):
        """Same signature as ``da4ml.cmvm.solve`` (bindings.cc:235-248)."""
        opts = dict(method0=method0, method1=method1, hard_dc=hard_dc, decompose_dc=decompose_dc, adder_size=adder_size, carry_size=carry_size,
                    search_all_decompose_dc=search_all_decompose_dc)
        if self._mode == 'record':
            k = np.ascontiguousarray(kernel, dtype=np.float32)
            self.calls.append(CmvmCall(k.copy(), None if qintervals is None else [tuple(map(float, q)) for q in qintervals],
                                       None if latencies is None else [float(v) for v in latencies], opts))
            return _Placeholder(k.shape[1])
        if self._mode == 'replay':
            i = self._next
            self._next += 1
            want = self.calls[i]
            if i >= len(self.raw) or np.asarray(kernel).shape != want.kernel.shape or not np.array_equal(np.asarray(kernel, dtype=np.float32), want.kernel):
                raise RuntimeError('SolveBatcher: the second pass issued different solve() calls than the first')
            return self.raw[i].to_pipeline(self.types_module)
        from ._binary import solve_raw  # noqa: PLC0415

        return solve_raw(np.ascontiguousarray(kernel, dtype=np.float32), qintervals=qintervals, latencies=latencies, **opts).to_pipeline(self.types_module)

    def run(self, fn: Callable):
        self.calls, self.raw, self._mode = [], [], 'record'
        try:
            fn()
            solver = self.solver
            if solver is None:
                from ._binary import solve_batch_raw as solver  # noqa: PLC0415
            self.raw = solve_calls(self.calls, solver=solver)
            self._mode, self._next = 'replay', 0
            return fn()
        finally:
            self._mode = 'idle'
