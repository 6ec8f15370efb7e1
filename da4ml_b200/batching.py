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
