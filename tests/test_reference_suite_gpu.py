"""The properties the reference's own CMVM suite asserts (reference tests/test_cmvm.py:23-55: CSD digits rebuild the
matrix, m0 @ m1 rebuilds it for every delay constraint, every solver option combination yields a graph whose kernel is
the input) checked on this package over the same grid of sizes, bit widths and options.  Matrices are seeded.  -m gpu."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import da4ml_b200._binary as native  # noqa: E402

SIZES = (2, 4, 8)
WIDTHS = (2, 4, 8)
GRID = list(itertools.product(SIZES, WIDTHS))


def rounded_uniform(n: int, width: int) -> np.ndarray:
    """Same distribution as the reference fixture (round((u - 0.5) * 2**(width + 1))), reproducible."""
    u = np.random.default_rng(1000 * n + width).random((n, n))
    return np.round((u - 0.5) * 2.0 ** (width + 1)).astype(np.float32)


@pytest.mark.parametrize('n,width', GRID)
def test_digits_rebuild_matrix(n, width):
    W = rounded_uniform(n, width)
    digits, row_shift, col_shift = native.csd_decompose(W)
    weights = 2.0 ** np.arange(digits.shape[-1])
    rebuilt = (digits * weights).sum(-1) * 2.0 ** row_shift[:, None] * 2.0 ** col_shift[None, :]
    assert (rebuilt == W).all()


@pytest.mark.parametrize('n,width', GRID)
def test_two_factor_decomposition_rebuilds_matrix(n, width):
    W = rounded_uniform(n, width)
    for dc in (-2, -1, 0, 1, 2):
        left, right = native.kernel_decompose(W, dc=dc)
        assert (left @ right == W).all(), dc


OPTION_GRID = list(itertools.product((0, 2, -1), ('mc', 'wmc'), ('mc', 'wmc'), (0, -1, -2), (False, True)))


@pytest.mark.parametrize('n,width', GRID)
def test_every_option_combination_reproduces_kernel(n, width):
    W = rounded_uniform(n, width)
    for hard_dc, m0, m1, decompose_dc, search_all in OPTION_GRID:
        graph = native.solve(W, method0=m0, method1=m1, hard_dc=hard_dc, decompose_dc=decompose_dc,
                             search_all_decompose_dc=search_all, adder_size=1, carry_size=-1)
        assert (graph.kernel == W).all(), (hard_dc, m0, m1, decompose_dc, search_all)
