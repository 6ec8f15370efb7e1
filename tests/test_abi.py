"""The C-ABI shared library loads and exports every symbol include/da4ml_b200_cmvm.h declares.  No GPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / 'include' / 'da4ml_b200_cmvm.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(da4ml_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import da4ml_b200._binary as B

    lib = ctypes.CDLL(str(B.lib_path()))
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in the header but not exported'
    assert set(syms) == set(B.EXPORTED_SYMBOLS)


def test_argument_validation_without_gpu():
    import da4ml_b200._binary as B

    with pytest.raises(TypeError):
        B.solve(np.ones((4, 4), np.float64))  # the reference binds kernel as noconvert float32
    with pytest.raises(RuntimeError):
        B.solve(np.ones((2, 2, 2), np.float32))
    with pytest.raises(ValueError):
        B.solve(np.ones((4, 4), np.float32), qintervals=[(-1.0, 1.0, 1.0)] * 3)


def test_no_cpu_fallback():
    import da4ml_b200._binary as B

    if B.device_info()['cuda_devices'] > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        B.solve(np.ones((4, 4), np.float32))
    with pytest.raises(RuntimeError, match='no CUDA device'):
        B.csd_decompose(np.ones((4, 4), np.float32))


def test_product_does_not_import_the_oracle():
    for path in (ROOT / 'da4ml_b200').rglob('*'):
        if path.suffix in ('.py', '.cu', '.cuh', '.h'):
            src = path.read_text()
            assert 'oracle' not in src.replace('no CPU fallback', ''), f'{path} mentions the oracle'


def test_product_library_contains_no_simulation_code():
    """The kernel sources carry `#ifdef DA_CPU_SIM` branches for the CPU kernel simulation of the tests; the shipped
    library is built without that switch and exports nothing of the shim."""
    import subprocess

    import __graft_entry__ as G
    import da4ml_b200._binary as B

    assert not any('DA_CPU_SIM' in f for f in G.NVCC_FLAGS)
    syms = subprocess.run(['nm', '-D', '--defined-only', str(B.lib_path())], capture_output=True, text=True, check=True).stdout
    assert 'simt' not in syms and 'sim_solve' not in syms
    for path in (ROOT / 'da4ml_b200').rglob('*.py'):
        assert 'simt' not in path.read_text(), f'{path} refers to the test shim'
