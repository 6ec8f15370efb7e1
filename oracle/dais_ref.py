"""ctypes window onto oracle/_ref/libdais_ref.so: the reference's DAIS interpreter compiled in place.
TEST INFRASTRUCTURE ONLY (checker for the GPU replay of adder graphs)."""
import ctypes as C
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / '_ref' / 'libdais_ref.so'
_lib = None


def available() -> bool:
    return _LIB_PATH.exists()


def run(program: np.ndarray, data: np.ndarray) -> np.ndarray:
    """reference `run_interp` (dais/bindings.cc): float64 [n_samples, n_in] -> float64 [n_samples, n_out]."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.ref_dais_last_error.restype = C.c_char_p
        _lib.ref_dais_run.argtypes = [C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double)]
    prog = np.ascontiguousarray(program, dtype=np.int32)
    n_in, n_out = int(prog[2]), int(prog[3])
    x = np.ascontiguousarray(np.asarray(data, dtype=np.float64).reshape(-1, n_in))
    out = np.zeros((x.shape[0], n_out), np.float64)
    if _lib.ref_dais_run(prog.ctypes.data_as(C.POINTER(C.c_int32)), prog.size, x.ctypes.data_as(C.POINTER(C.c_double)), x.shape[0], out.ctypes.data_as(C.POINTER(C.c_double))):
        raise RuntimeError(_lib.ref_dais_last_error().decode())
    return out
