"""ctypes window onto oracle/_ref/libcmvm_ref.so -- the reference's own CMVM translation units
(api.cc, cmvm_core.cc, state_opr.cc, indexers.cc) compiled in place, see oracle/Makefile.

TEST INFRASTRUCTURE ONLY.  Results are returned as flat numpy arrays ("stage dicts") so that parity
tests compare them array-for-array with the CUDA path.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / '_ref' / 'libcmvm_ref.so'
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
            raise FileNotFoundError(f'{_LIB_PATH} missing: run `make -C oracle` where /root/reference exists')
        L = C.CDLL(str(_LIB_PATH))
        L.ref_last_error.restype = C.c_char_p
        L.ref_solve.restype = C.c_void_p
        L.ref_solve.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_char_p, C.c_char_p, C.c_int, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_solve_single.restype = C.c_void_p
        L.ref_solve_single.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_char_p, _f32p, _f32p, C.c_int, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_n_stages.restype = C.c_int64
        L.ref_n_stages.argtypes = [C.c_void_p]
        L.ref_stage_n_ops.restype = C.c_int64
        L.ref_stage_n_ops.argtypes = [C.c_void_p, C.c_int64]
        L.ref_stage_meta.argtypes = [C.c_void_p, C.c_int64, _i64p]
        L.ref_stage_copy.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p, _i64p, _i64p, _i64p, _f32p]
        L.ref_get_lsb_loc.argtypes = [C.c_float]
        L.ref_iceil_log2.argtypes = [C.c_float]
        L.ref_cost_add.argtypes = [_f32p, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int, _f32p]
        L.ref_qint_add.argtypes = [_f32p, _f32p, C.c_int64, C.c_int, C.c_int, _f32p]
        L.ref_overlap_and_accum.argtypes = [_f32p, _f32p, C.POINTER(C.c_int)]
        L.ref_csd_decompose.restype = C.c_int64
        L.ref_csd_decompose.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int, _i8p, _i8p, _i8p]
        L.ref_int_arr_to_csd.restype = C.c_int64
        L.ref_int_arr_to_csd.argtypes = [C.POINTER(C.c_int32), C.c_int64, _i8p]
        L.ref_kernel_decompose.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int, _f32p, _f32p]
        L.ref_log2f.restype = C.c_float
        L.ref_log2f.argtypes = [C.c_float]
        L.ref_trace.restype = C.c_void_p
        L.ref_trace.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_char_p, _f32p, _f32p, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_int]
        L.ref_trace_len.restype = C.c_int64
        L.ref_trace_len.argtypes = [C.c_void_p]
        L.ref_trace_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.ref_trace_copy.argtypes = [C.c_void_p, _i64p, _i64p, _i64p]
        L.ref_trace_free.argtypes = [C.c_void_p]
        L.ref_freq_init.restype = C.c_int64
        L.ref_freq_init.argtypes = [_f32p, C.c_int64, C.c_int64, _f32p, _i64p, C.c_int64]
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
    for s in range(L.ref_n_stages(h)):
        meta = np.zeros(4, np.int64)
        L.ref_stage_meta(h, s, _ip(meta))
        n_in, n_out = int(meta[0]), int(meta[1])
        n_ops = L.ref_stage_n_ops(h, s)
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
        )
        L.ref_stage_copy(h, s, _ip(st['inp_shifts']), _ip(st['out_idxs']), _ip(st['out_shifts']), _ip(st['out_negs']), _ip(st['ops_i']), _fp(st['ops_f']))
        out.append(st)
    return out


def solve(kernel, method0='wmc', method1='auto', hard_dc=-1, decompose_dc=-2, qintervals=None, latencies=None,
          adder_size=-1, carry_size=-1, search_all_decompose_dc=True):
    """reference api.cc:147 `solve` -> list of stage dicts."""
    L = lib()
    k, q, l = _prep(kernel, qintervals, latencies)
    h = L.ref_solve(_fp(k), k.shape[0], k.shape[1], method0.encode(), method1.encode(), hard_dc, decompose_dc, _fp(q), _fp(l), adder_size, carry_size, int(search_all_decompose_dc))
    if not h:
        raise RuntimeError(L.ref_last_error().decode())
    try:
        return _stages(h)
    finally:
        L.ref_free(h)


def solve_single(kernel, method='wmc', qintervals=None, latencies=None, adder_size=-1, carry_size=-1):
    """reference cmvm_core.cc:227 `solve_single` -> one stage dict."""
    L = lib()
    k, q, l = _prep(kernel, qintervals, latencies)
    h = L.ref_solve_single(_fp(k), k.shape[0], k.shape[1], method.encode(), _fp(q), _fp(l), adder_size, carry_size)
    if not h:
        raise RuntimeError(L.ref_last_error().decode())
    try:
        return _stages(h)[0]
    finally:
        L.ref_free(h)


def trace(kernel, method='wmc', qintervals=None, latencies=None, adder_size=-1, carry_size=-1, max_iters=-1,
          time_limit_s=0.0, counters=True):
    """Greedy-loop trace: chosen pairs per iteration + work counters (see ref_glue.cc:ref_trace)."""
    L = lib()
    k, q, l = _prep(kernel, qintervals, latencies)
    h = L.ref_trace(_fp(k), k.shape[0], k.shape[1], method.encode(), _fp(q), _fp(l), adder_size, carry_size, max_iters, float(time_limit_s), int(counters))
    if not h:
        raise RuntimeError(L.ref_last_error().decode())
    try:
        n = L.ref_trace_len(h)
        pairs = np.zeros((n, 4), np.int64)
        fs = np.zeros(n, np.int64)
        rs = np.zeros(n, np.int64)
        sc = np.zeros(6, np.float64)
        L.ref_trace_scalars(h, sc.ctypes.data_as(C.POINTER(C.c_double)))
        if n:
            L.ref_trace_copy(h, _ip(pairs), _ip(fs), _ip(rs) if counters else None)
        return dict(pairs=pairs, f_sizes=fs, r_sizes=rs, f0=int(sc[0]), r0=int(sc[1]), d0=int(sc[2]), d_final=int(sc[3]), seconds=sc[4], create_seconds=sc[5])
    finally:
        L.ref_trace_free(h)


def freq_init(kernel, qintervals=None):
    """Initial pair histogram (create_state, state_opr.cc:115-144) as [n,5] int64 (id0,id1,shift,sub,count)."""
    L = lib()
    k, q, _ = _prep(kernel, qintervals, None)
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 5), np.int64)
        n = L.ref_freq_init(_fp(k), k.shape[0], k.shape[1], _fp(q), _ip(out), cap)
        if n <= cap:
            return out[:n]
        cap = int(n)


def csd_decompose(kernel, center=True):
    L = lib()
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    csd = np.zeros(k.size * 33, np.int8)
    s0 = np.zeros(k.shape[0], np.int8)
    s1 = np.zeros(k.shape[1], np.int8)
    N = L.ref_csd_decompose(_fp(k), k.shape[0], k.shape[1], int(center), csd.ctypes.data_as(_i8p), s0.ctypes.data_as(_i8p), s1.ctypes.data_as(_i8p))
    return csd[: k.size * N].reshape(k.shape[0], k.shape[1], N).copy(), s0, s1


def int_arr_to_csd(x):
    L = lib()
    a = np.ascontiguousarray(x, dtype=np.int32)
    out = np.zeros(a.size * 33, np.int8)
    N = L.ref_int_arr_to_csd(a.ctypes.data_as(C.POINTER(C.c_int32)), a.size, out.ctypes.data_as(_i8p))
    return out[: a.size * N].reshape(*a.shape, N).copy()


def kernel_decompose(kernel, dc=-2):
    L = lib()
    k = np.ascontiguousarray(kernel, dtype=np.float32)
    m0 = np.zeros(k.shape, np.float32)
    m1 = np.zeros((k.shape[1], k.shape[1]), np.float32)
    L.ref_kernel_decompose(_fp(k), k.shape[0], k.shape[1], dc, _fp(m0), _fp(m1))
    return m0, m1


def get_lsb_loc(x: float) -> int:
    return lib().ref_get_lsb_loc(float(x))


def iceil_log2(x: float) -> int:
    return lib().ref_iceil_log2(float(x))


def log2f(x: float) -> float:
    return lib().ref_log2f(float(x))


def cost_add(q0, q1, shift, sub, adder_size, carry_size):
    a = np.asarray(q0, np.float32)
    b = np.asarray(q1, np.float32)
    out = np.zeros(2, np.float32)
    lib().ref_cost_add(_fp(a), _fp(b), shift, int(sub), adder_size, carry_size, _fp(out))
    return float(out[0]), float(out[1])


def qint_add(q0, q1, shift, sub0=False, sub1=False):
    a = np.asarray(q0, np.float32)
    b = np.asarray(q1, np.float32)
    out = np.zeros(3, np.float32)
    lib().ref_qint_add(_fp(a), _fp(b), shift, int(sub0), int(sub1), _fp(out))
    return tuple(float(v) for v in out)


def overlap_and_accum(q0, q1):
    a = np.asarray(q0, np.float32)
    b = np.asarray(q1, np.float32)
    out = (C.c_int * 2)()
    lib().ref_overlap_and_accum(_fp(a), _fp(b), out)
    return int(out[0]), int(out[1])
