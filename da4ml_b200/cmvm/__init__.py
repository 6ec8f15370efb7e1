"""Public facade of the CMVM path -- same names as the reference's ``da4ml.cmvm``
(``src/da4ml/cmvm/__init__.py:7-29``): ``solve``, ``kernel_decompose``, ``QInterval``, ``Op``, ``CombLogic``,
``solver_options_t``.  ``solve_batch`` is the B200-side addition for many independent matrices."""

from collections.abc import Callable
from typing import TypedDict

import numpy as np

from .._binary import kernel_decompose, solve, solve_batch
from ..types import CombLogic, Op, Pipeline, QInterval


class solver_options_t(TypedDict, total=False):
    method0: str
    method1: str
    hard_dc: int
    decompose_dc: int
    adder_size: int
    carry_size: int
    search_all_decompose_dc: bool
    offload_fn: None | Callable[[np.ndarray, object], np.ndarray]


__all__ = ['solve', 'solve_batch', 'QInterval', 'Op', 'CombLogic', 'Pipeline', 'kernel_decompose', 'solver_options_t']
