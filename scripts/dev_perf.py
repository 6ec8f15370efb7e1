"""Developer timing sweep (run under gpurun): single-stage and full solves at growing sizes."""
import sys, time, os
import numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B

def mat(n, bits, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)

cases = [(64, 8), (128, 6), (128, 8), (256, 8)]
if len(sys.argv) > 1:
    cases = [tuple(map(int, a.split('x'))) for a in sys.argv[1:]]
B.solve_single_raw(mat(8, 4, 0))  # warm-up (context, allocations)
for n, bits in cases:
    W = mat(n, bits, 0)
    for G in [int(g) for g in os.environ.get('GS', '0').split(',')]:
        B.set_group_size(G)
        t0 = time.time(); raw, _ = B.solve_single_raw(W, 'wmc'); t1 = time.time()
        c = raw.counters[0]
        print(f'single {n}x{n} b{bits} G={c["group_ctas"]}: wall {1e3*(t1-t0):.1f} ms dev {raw.device_ms:.1f} ms T={c["T"]} us/step={1e3*raw.device_ms/max(c["T"],1):.1f} adders={raw.n_adders} F0={c["F0"]} Fmax={c["F_max"]} sumF={c["sum_F"]:.3e} sumR={c["sum_R"]:.3e} D0={c["D0"]} compactions={c["compactions"]} rescanned={c["rescanned"]:.3e} list_max={c["list_max"]} lcap={c["smem_list_cap"]} phase_us/step={[round(v/1.9e3/max(c["T"],1),2) for v in c["phase_cycles"]]} peak_us/n>10us={[(round((v//1000000)/1.9e3,1), v%1000000) for v in c["phase_cycles_max"]]}', flush=True)
    B.set_group_size(0)
    if os.environ.get('FULL', '1') == '1':
        t0 = time.time(); raw = B.solve_raw(W); t1 = time.time()
        ok = bool(np.all(raw.to_pipeline().kernel == W))
        print(f'solve  {n}x{n} b{bits}: wall {1e3*(t1-t0):.1f} ms dev {raw.device_ms:.1f} ms adders={raw.n_adders} launches={raw.launches} kernel_ok={ok} stage0 T={raw.counters[0]["T"]} G={raw.counters[0]["group_ctas"]}', flush=True)
