"""The reference's own tests/test_cmvm.py, pointed at this package (same parametrisation and assertions;
the kernel fixture is seeded).  -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from da4ml_b200._binary import csd_decompose, kernel_decompose, solve  # noqa: E402


@pytest.fixture(params=[2, 4, 8])
def n_dim(request) -> int:
    return request.param


@pytest.fixture(params=[2, 4, 8])
def bits(request) -> int:
    return request.param


@pytest.fixture
def kernel(n_dim, bits):
    rng = np.random.default_rng(n_dim * 100 + bits)
    return np.round((rng.random((n_dim, n_dim)) - 0.5) * 2 ** (bits + 1)).astype(np.float32)


def test_decompose(kernel):
    csd, shift0, shift1 = csd_decompose(kernel.astype(np.float32))
    shift2 = np.arange(csd.shape[-1])
    recon = csd * (2.0 ** shift0[:, None, None]) * (2.0 ** shift1[None, :, None]) * (2.0 ** shift2[None, None, :])
    assert np.all(np.sum(recon, axis=-1) == kernel)


@pytest.mark.parametrize('dc', [-2, -1, 0, 1, 2])
def test_kernel_decompose(kernel, dc: int):
    m0, m1 = kernel_decompose(kernel.astype(np.float32), dc=dc)
    assert np.all(m0 @ m1 == kernel)


@pytest.mark.parametrize('hard_dc', [0, 2, -1])
@pytest.mark.parametrize('method0', ['mc', 'wmc'])
@pytest.mark.parametrize('method1', ['mc', 'wmc'])
@pytest.mark.parametrize('decompose_dc', [0, -1, -2])
@pytest.mark.parametrize('search_all_decompose_dc', [False, True])
def test_solve(kernel, method0, method1, hard_dc, decompose_dc, search_all_decompose_dc):
    sol = solve(
        kernel,
        hard_dc=hard_dc,
        method0=method0,
        method1=method1,
        decompose_dc=decompose_dc,
        search_all_decompose_dc=search_all_decompose_dc,
        adder_size=1,
        carry_size=-1,
    )
    assert np.all(sol.kernel == kernel)
