import sys, numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
from da4ml_b200.types import pipeline_from_arrays
n_in, n_out, bits = 96, 16, 8
rng = np.random.default_rng(0)
W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n_in, n_out)).astype(np.float32)
raw, _ = B.solve_single_raw(W, 'dummy')
K = pipeline_from_arrays(raw.stages).solutions[0].kernel
print('bad cols', np.unique(np.argwhere(K != W)[:, 1]).tolist())
