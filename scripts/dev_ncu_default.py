"""Two default solves of the bench workload (256x256 int8 seed 0): the first warms up, ncu profiles a launch of the second
(launch order per solve: stage 0 of all candidates, stage 1 -> `-s 2 -c 1` is the second stage 0)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B

W = np.random.default_rng(0).integers(-128, 128, size=(256, 256)).astype(np.float32)
for _ in range(2):
    raw = B.solve_raw(W)
    print(raw.n_adders, raw.device_ms, raw.launches, flush=True)
