"""ctypes window onto oracle/libcmvm_oracle.so -- the from-scratch CPU restatement (cmvm_oracle.cc).

TEST INFRASTRUCTURE ONLY.  Same calling surface as oracle/ref.py so tests can swap one for the other."""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / 'libcmvm_oracle.so'
_lib = None

_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_i8p = C.POINTER(C.c_int8)


def available() -> bool:
    return _LIB_PATH.exists()


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise FileNotFoundError(f'{_LIB_PATH} missing: run `make -C oracle`')
        L = C.CDLL(str(_LIB_PATH))
        L.orc_last_error.restype = C.c_char_p
        L.orc_solve.restype = C.c_void_p
        L.orc_solve.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_char_p, C.c_char_p, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.orc_solve_single.restype = C.c_void_p
        L.orc_solve_single.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_char_p, _f32p, _f32p, C.c_int, C.c_int]
        L.orc_partial.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_char_p, C.c_int64, C.c_double, C.POINTER(C.c_double)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_n_stages.restype = C.c_int64
        L.orc_n_stages.argtypes = [C.c_void_p]
        L.orc_stage_n_ops.restype = C.c_int64
        L.orc_stage_n_ops.argtypes = [C.c_void_p, C.c_int64]
        L.orc_stage_meta.argtypes = [C.c_void_p, C.c_int64, _i64p]
        L.orc_stage_copy.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p, _i64p, _i64p, _i64p, _f32p]
        L.orc_csd_decompose.restype = C.c_int64
        L.orc_csd_decompose.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int, _i8p, _i8p, _i8p]
        L.orc_int_arr_to_csd.restype = C.c_int64
        L.orc_int_arr_to_csd.argtypes = [C.POINTER(C.c_int32), C.c_int64, _i8p]
        L.orc_kernel_decompose.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int, _f32p, _f32p]
        L.orc_get_lsb_loc.argtypes = [C.c_float]
        L.orc_iceil_log2.argtypes = [C.c_float]
        L.orc_cost_add.argtypes = [_f32p, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_log2f.restype = C.c_float
        L.orc_log2f.argtypes = [C.c_float]
        L.orc_csd_weight.argtypes = [C.c_int32]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_i64p)


def _prep(kernel, qintervals, latencies):
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    assert k.ndim == 2
    q = None if qintervals is None else np.ascontiguousarray(np.asarray(qintervals, dtype=np.float32).reshape(-1, 3))
    l = None if latencies is None else np.ascontiguousarray(np.asarray(latencies, dtype=np.float32).reshape(-1))
    return k, q, l


def _stages(h):
    L = lib()
    out = []
    for s in range(L.orc_n_stages(h)):
        meta = np.zeros(11, np.int64)
        L.orc_stage_meta(h, s, _ip(meta))
        n_in, n_out = int(meta[0]), int(meta[1])
        n_ops = L.orc_stage_n_ops(h, s)
        st = dict(
            shape=(n_in, n_out),
            inp_shifts=np.zeros(n_in, np.int64),
            out_idxs=np.zeros(n_out, np.int64),
            out_shifts=np.zeros(n_out, np.int64),
            out_negs=np.zeros(n_out, np.int64),
            ops_i=np.zeros((n_ops, 4), np.int64),
            ops_f=np.zeros((n_ops, 5), np.float32),
            carry_size=int(meta[2]),
            adder_size=int(meta[3]),
            counters=dict(T=int(meta[4]), sum_F=int(meta[5]), sum_R=int(meta[6]), F0=int(meta[7]), R0=int(meta[8]), D0=int(meta[9]), D_final=int(meta[10])),
        )
        L.orc_stage_copy(h, s, _ip(st['inp_shifts']), _ip(st['out_idxs']), _ip(st['out_shifts']), _ip(st['out_negs']), _ip(st['ops_i']), _fp(st['ops_f']))
        out.append(st)
    return out


def solve(kernel, method0='wmc', method1='auto', hard_dc=-1, decompose_dc=-2, qintervals=None, latencies=None,
          adder_size=-1, carry_size=-1, search_all_decompose_dc=True):
    L = lib()
    k, q, l = _prep(kernel, qintervals, latencies)
    h = L.orc_solve(_fp(k), k.shape[0], k.shape[1], method0.encode(), method1.encode(), hard_dc, decompose_dc, _fp(q), _fp(l), adder_size, carry_size, int(search_all_decompose_dc))
    if not h:
        raise RuntimeError(L.orc_last_error().decode())
    try:
        return _stages(h)
    finally:
        L.orc_free(h)


def solve_single(kernel, method='wmc', qintervals=None, latencies=None, adder_size=-1, carry_size=-1):
    L = lib()
    k, q, l = _prep(kernel, qintervals, latencies)
    h = L.orc_solve_single(_fp(k), k.shape[0], k.shape[1], method.encode(), _fp(q), _fp(l), adder_size, carry_size)
    if not h:
        raise RuntimeError(L.orc_last_error().decode())
    try:
        return _stages(h)[0]
    finally:
        L.orc_free(h)


def partial(kernel, method='wmc', max_iters=-1, time_limit_s=0.0):
    """Bounded run of the greedy loop (default qint/lat): dict(iters, seconds, sum_F, sum_R, F0, R0, D0)."""
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    out = (C.c_double * 7)()
    if lib().orc_partial(_fp(k), k.shape[0], k.shape[1], method.encode(), max_iters, float(time_limit_s), out):
        raise RuntimeError(lib().orc_last_error().decode())
    return dict(iters=int(out[0]), seconds=float(out[1]), sum_F=int(out[2]), sum_R=int(out[3]), F0=int(out[4]), R0=int(out[5]), D0=int(out[6]))


def csd_decompose(kernel, center=True):
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    csd = np.zeros(k.size * 33, np.int8)
    s0 = np.zeros(k.shape[0], np.int8)
    s1 = np.zeros(k.shape[1], np.int8)
    N = lib().orc_csd_decompose(_fp(k), k.shape[0], k.shape[1], int(center), csd.ctypes.data_as(_i8p), s0.ctypes.data_as(_i8p), s1.ctypes.data_as(_i8p))
    return csd[: k.size * N].reshape(k.shape[0], k.shape[1], N).copy(), s0, s1


def int_arr_to_csd(x):
    a = np.ascontiguousarray(x, dtype=np.int32)
    out = np.zeros(a.size * 33, np.int8)
    N = lib().orc_int_arr_to_csd(a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, out.ctypes.data_as(_i8p))
    return out[: a.size * N].reshape(*a.shape, N).copy()


def kernel_decompose(kernel, dc=-2):
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    m0 = np.zeros(k.shape, np.float32)
    m1 = np.zeros((k.shape[1], k.shape[1]), np.float32)
    lib().orc_kernel_decompose(_fp(k), k.shape[0], k.shape[1], dc, _fp(m0), _fp(m1))
    return m0, m1


def get_lsb_loc(x):
    return lib().orc_get_lsb_loc(float(x))


def iceil_log2(x):
    return lib().orc_iceil_log2(float(x))


def log2f(x):
    return lib().orc_log2f(float(x))


def csd_weight(x):
    return lib().orc_csd_weight(int(x))


def cost_add(q0, q1, shift, sub, adder_size, carry_size):
    a = np.asarray(q0, np.float32)
    b = np.asarray(q1, np.float32)
    out = np.zeros(2, np.float32)
    lib().orc_cost_add(_fp(a), _fp(b), shift, int(sub), adder_size, carry_size, _fp(out))
    return float(out[0]), float(out[1])
