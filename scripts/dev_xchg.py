import ctypes as C, sys
sys.path.insert(0, '.')
import da4ml_b200._binary as B
L = B._L
L.da4ml_cmvm_debug_xchg_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
for G in (2, 8, 32, 148):
    for work in (0, 1, 4):
        out = C.c_double(0)
        rc = L.da4ml_cmvm_debug_xchg_bench(G, 2000, work, C.byref(out))
        print(f'G={G} work={work}: rc={rc} {out.value:.2f} us/exchange', flush=True)
