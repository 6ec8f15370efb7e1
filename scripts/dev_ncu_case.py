"""One solve kernel launch for ncu (under gpurun): python scripts/dev_ncu_case.py <owned|columns> [n] [bits] [G]
The first (tiny) launch is a warm-up: profile the second match (`-s 1 -c 1`)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B

kind = sys.argv[1] if len(sys.argv) > 1 else 'owned'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 8
G = int(sys.argv[4]) if len(sys.argv) > 4 else 14
rng = np.random.default_rng(0)
B.solve_single_raw(rng.integers(-8, 8, size=(8, 8)).astype(np.float32))
W = np.random.default_rng(0).integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)
B.set_group_size(G)
raw, _ = B.solve_single_raw(W, 'wmc')
print(kind, n, bits, G, raw.device_ms, raw.n_adders)
