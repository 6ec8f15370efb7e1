import sys, time
import numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
def mat(n, bits, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)
B.solve_raw(mat(8, 4, 0))
for n, bits, nb in [(128, 6, 1), (128, 6, 8), (128, 6, 64), (64, 8, 64), (256, 8, 2)]:
    ks = [mat(n, bits, s) for s in range(nb)]
    B.solve_batch_raw(ks[:1])
    t0 = time.time(); res = B.solve_batch_raw(ks); dt = time.time() - t0
    print(f'batch {nb} x {n}x{n} int{bits}: wall {dt*1e3:.1f} ms -> {nb/dt:.2f} matrices/s, dev {res[0].device_ms:.1f} ms, G={res[0].counters[0]["group_ctas"]} lcap={res[0].counters[0]["smem_list_cap"]} adders[0]={res[0].n_adders}', flush=True)
