import sys, numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
from da4ml_b200.types import pipeline_from_arrays
n, bits = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)
for method in ('wmc', 'wmc-dc'):
    raw, _ = B.solve_single_raw(W, method)
    K = pipeline_from_arrays(raw.stages).solutions[0].kernel
    print('single', method, raw.n_adders, 'kernel_ok', bool(np.all(K == W)), raw.counters[0]['status'])
for dc in range(-1, int(np.ceil(np.log2(n))) + 1):
    m0, m1 = B.kernel_decompose(W, dc)
    raw = B.solve_raw(W, search_all_decompose_dc=False, decompose_dc=dc, hard_dc=1000)
    p = raw.to_pipeline()
    K0, K1 = p.solutions[0].kernel, p.solutions[1].kernel
    print(f'dc={dc}: adders={raw.n_adders} m0@m1==W {bool(np.all(m0.astype(np.float64) @ m1.astype(np.float64) == W))} K0==m0 {bool(np.all(K0 == m0))} K1==m1 {bool(np.all(K1 == m1))} '
          f'stage ops {[len(s["ops_i"]) for s in raw.stages]} nbits {[c["n_bits"] for c in raw.counters]} max|m0| {np.abs(m0).max()} max|m1| {np.abs(m1).max()} nnz1 {np.count_nonzero(m1)}', flush=True)
    if not np.all(K1 == m1):
        bad = np.argwhere(K1 != m1)
        print('   K1 bad', len(bad), bad[:5].tolist(), K1[tuple(bad[0])], m1[tuple(bad[0])])
    if not np.all(K0 == m0):
        bad = np.argwhere(K0 != m0)
        print('   K0 bad', len(bad), bad[:5].tolist(), K0[tuple(bad[0])], m0[tuple(bad[0])])
