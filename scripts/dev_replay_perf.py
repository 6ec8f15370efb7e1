"""Throughput of the DAIS replay kernels (N2): bytes = 24 B per (op, sample) [2 x int64 loads + 1 store]."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
rng = np.random.default_rng(0)
W = rng.integers(-128, 128, size=(256, 256)).astype(np.float32)
pipe = B.solve(W, search_all_decompose_dc=False, decompose_dc=-1)
sol = pipe.solutions[0]
prog = sol.to_binary()
n_ops = len(sol.ops)
for n in (256, 2048, 16384):
    x = rng.integers(-128, 128, size=(n, 256)).astype(np.float64)
    B.dais_interp_run(prog, x[:8])
    t0 = time.time(); y = B.dais_interp_run(prog, x); dt = time.time() - t0
    ok = np.array_equal(y, x @ W.astype(np.float64))
    print(f'{n_ops} ops x {n} samples: {dt*1e3:.1f} ms wall (incl. H2D/D2H), {24.0*n_ops*n/dt/1e9:.0f} GB/s algorithmic, exact={ok}', flush=True)
