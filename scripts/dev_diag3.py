import sys, numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
from da4ml_b200.types import pipeline_from_arrays
n, bits = int(sys.argv[1]), int(sys.argv[2])
methods = sys.argv[3].split(',')
G = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.default_rng(0)
W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)
B.set_group_size(G)
for m in methods:
    raw, tr = B.solve_single_raw(W, m)
    K = pipeline_from_arrays(raw.stages).solutions[0].kernel
    badc = np.unique(np.argwhere(K != W)[:, 1])
    print(m, 'adders', raw.n_adders, 'bad cols', badc.tolist(), raw.counters[0])
