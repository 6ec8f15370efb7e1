"""Small solves for compute-sanitizer runs (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
rng = np.random.default_rng(0)
for n, bits, kw in [(8, 4, {}), (24, 8, dict(hard_dc=2, adder_size=2, carry_size=4)), (40, 8, {})]:
    W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)
    raw = B.solve_raw(W, **kw)
    assert np.array_equal(raw.to_pipeline().kernel, W)
    print(n, bits, kw, raw.n_adders, flush=True)
B.set_group_size(8)
raw, _ = B.solve_single_raw(rng.integers(-128, 128, size=(32, 32)).astype(np.float32), 'wmc', trace_cap=4096)
print('single G=8', raw.n_adders)
