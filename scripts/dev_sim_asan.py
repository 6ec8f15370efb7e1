"""Developer check (CPU): the simulated kernels under AddressSanitizer -- every simulated global / shared buffer is its own
heap allocation, so an out-of-bounds access of the kernels shows up as a heap-buffer-overflow.

  g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -ffp-contract=off -fPIC -shared -w tests/simt/sim_cmvm.cc -o /tmp/libsim_asan.so
  ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 LD_PRELOAD=$(g++ -print-file-name=libasan.so) python scripts/dev_sim_asan.py
"""
import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import simt
from pathlib import Path
simt.LIB = Path('/tmp/libsim_asan.so')
simt.build = lambda force=False: simt.LIB
from conftest import assert_stage_equal, int_matrix
from oracle import port
for variant in (0, 1):  # planner's layout; owner lists that spill to global memory + a small counter table
    if variant:
        simt.set_own_caps(hash_log=8, list_rows=4)
    for (n_in, n_out, bits, seed, G) in [(8, 8, 4, 0, 2), (14, 11, 6, 3, 3), (6, 70, 5, 31, 2), (1, 9, 8, 1, 2), (9, 1, 8, 2, 2)]:
        W = int_matrix(n_in, n_out, bits, seed)
        got, _ = simt.solve_single(W, 'wmc', ctas=G, cta_threads=64)
        assert_stage_equal(got, port.solve_single(W, 'wmc'))
        print('asan ok', variant, W.shape, flush=True)
    simt.set_segment_cap(1500)
    got, meta = simt.solve_single(int_matrix(16, 16, 6, 11), 'wmc', ctas=2, cta_threads=64)
    simt.set_segment_cap(0)
    print('asan ok compaction', variant, meta[9], flush=True)
    simt.set_own_caps()
m0, m1 = simt.kernel_decompose(int_matrix(12, 20, 8, 1), 1)
print('asan ok decompose')
