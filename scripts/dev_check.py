"""Developer parity sweep: CUDA path vs oracle/_ref on seeded matrices (run under gpurun)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import ref
import da4ml_b200._binary as B

KEYS = ['inp_shifts', 'out_idxs', 'out_shifts', 'out_negs', 'ops_i', 'ops_f']

def cmp_stage(a, b, tag):
    ok = True
    for k in KEYS:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape or not np.array_equal(x, y, equal_nan=True):
            ok = False
            msg = f'shape {x.shape} vs {y.shape}'
            if x.shape == y.shape:
                bad = np.argwhere(x != y)
                msg = f'{len(bad)} diffs, first at {bad[0].tolist()}: {x[tuple(bad[0])]} vs {y[tuple(bad[0])]}'
            print(f'  MISMATCH {tag} {k}: {msg}')
    return ok

def main():
    print(B.device_info())
    rng = np.random.default_rng(0)
    n_bad = 0
    # 1. single stage with trace
    for (n_in, n_out, bits, method) in [(8, 8, 4, 'wmc'), (16, 16, 8, 'wmc'), (17, 5, 8, 'mc'), (32, 32, 8, 'wmc'), (32, 32, 8, 'wmc-dc'), (32, 24, 6, 'mc-dc'), (64, 64, 8, 'wmc')]:
        W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n_in, n_out)).astype(np.float32)
        t0 = time.time()
        raw, tr = B.solve_single_raw(W, method, trace_cap=1 << 16)
        t1 = time.time()
        rs = ref.solve_single(W, method)
        t2 = time.time()
        rt = ref.trace(W, method, counters=True)
        ok = cmp_stage(raw.stages[0], rs, f'single {n_in}x{n_out} b{bits} {method}')
        T = len(rt['pairs'])
        c = raw.counters[0]
        tr_ok = tr.shape[0] == T and np.array_equal(tr[:, :4], rt['pairs']) and np.array_equal(tr[:, 4], rt['f_sizes'])
        if not tr_ok:
            nb = min(tr.shape[0], T)
            d = np.argwhere((tr[:nb, :4] != rt['pairs'][:nb]).any(1) | (tr[:nb, 4] != rt['f_sizes'][:nb]))
            print(f'  TRACE MISMATCH T gpu={tr.shape[0]} ref={T} first diff step {d[0].tolist() if len(d) else None}')
            if len(d):
                s = int(d[0][0]); print('   gpu', tr[max(0,s-1):s+2].tolist(), '\n   ref', np.c_[rt['pairs'], rt['f_sizes']][max(0,s-1):s+2].tolist())
        cnt_ok = (c['F0'] == rt['f0'] and c['R0'] == rt['r0'] and c['sum_F'] == int(rt['f_sizes'].sum()) and c['sum_R'] == int(rt['r_sizes'].sum()) and c['D_final'] == rt['d_final'] and c['D0'] == rt['d0'])
        if not cnt_ok:
            print('  COUNTER MISMATCH', c, {k: rt[k] for k in ('f0', 'r0', 'd0', 'd_final')}, int(rt['f_sizes'].sum()), int(rt['r_sizes'].sum()))
        n_bad += (not ok) + (not tr_ok) + (not cnt_ok)
        print(f'single {n_in}x{n_out} b{bits} {method}: ok={ok} trace={tr_ok} counters={cnt_ok} T={c["T"]} adders={raw.n_adders} gpu {1e3*(t1-t0):.1f} ms (dev {raw.device_ms:.2f}) ref {1e3*(t2-t1):.1f} ms G={c["group_ctas"]}')
    # 2. decomposition helpers
    for (n_in, n_out, bits) in [(8, 8, 4), (16, 12, 8), (32, 32, 8)]:
        W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n_in, n_out)).astype(np.float32)
        a = B.csd_decompose(W); b = ref.csd_decompose(W)
        ok = all(np.array_equal(x, y) for x, y in zip(a, b))
        for dc in (-2, -1, 0, 1, 2):
            m = B.kernel_decompose(W, dc); r = ref.kernel_decompose(W, dc)
            ok = ok and np.array_equal(m[0], r[0]) and np.array_equal(m[1], r[1])
        n_bad += not ok
        print(f'decompose {n_in}x{n_out}: ok={ok}')
    # 3. full solve
    for (n, bits, kw) in [(8, 4, {}), (8, 4, dict(hard_dc=2)), (8, 4, dict(hard_dc=0, adder_size=1)), (16, 8, {}), (16, 8, dict(search_all_decompose_dc=False, decompose_dc=-1)), (32, 8, {}), (32, 8, dict(hard_dc=2, adder_size=1, carry_size=-1)), (64, 8, {})]:
        W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)
        t0 = time.time(); raw = B.solve_raw(W, **kw); t1 = time.time()
        rs = ref.solve(W, **kw); t2 = time.time()
        ok = len(raw.stages) == len(rs) and all(cmp_stage(a, b, f'solve {n} {kw} stage{i}') for i, (a, b) in enumerate(zip(raw.stages, rs)))
        kern_ok = bool(np.all(raw.to_pipeline().kernel == W))
        n_bad += (not ok) + (not kern_ok)
        print(f'solve {n}x{n} b{bits} {kw}: ok={ok} kernel={kern_ok} adders={raw.n_adders} gpu {1e3*(t1-t0):.1f} ms (dev {raw.device_ms:.2f}, {raw.launches} launches) ref {1e3*(t2-t1):.1f} ms')
    print('TOTAL BAD', n_bad)
    return n_bad

if __name__ == '__main__':
    sys.exit(1 if main() else 0)
