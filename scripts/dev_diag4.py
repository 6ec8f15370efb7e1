import sys, numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
from oracle import ref
from da4ml_b200.types import pipeline_from_arrays
for (n_in, n_out, bits) in [(64, 64, 8), (96, 16, 8), (128, 8, 8), (128, 128, 8)]:
    rng = np.random.default_rng(0)
    W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n_in, n_out)).astype(np.float32)
    raw, _ = B.solve_single_raw(W, 'dummy')
    st = raw.stages[0]
    rs = ref.solve_single(W, 'dummy')
    K = pipeline_from_arrays(raw.stages).solutions[0].kernel
    badc = np.unique(np.argwhere(K != W)[:, 1])
    same = np.array_equal(st['ops_i'], rs['ops_i']) and np.array_equal(st['ops_f'], rs['ops_f'])
    print(f'{n_in}x{n_out}: bad cols {badc.tolist()} ops equal {same} n_ops {len(st["ops_i"])} {len(rs["ops_i"])}')
    if not same:
        d = np.argwhere((st['ops_i'] != rs['ops_i']).any(1) | (st['ops_f'] != rs['ops_f']).any(1))[:, 0]
        print('  first diffs at', d[:10].tolist(), 'count', len(d))
        # column boundaries from the reference: out_idxs are last ops of each column
        oi = rs['out_idxs']
        for i in d[:3]:
            col = int(np.searchsorted(oi, i))
            start = n_in if col == 0 else oi[col - 1] + 1
            print(f'   op {i}: col {col} (ops {start}..{oi[col]}), offset in col {i - start}, K={oi[col] - start + 2}')
            print('     gpu', st['ops_i'][i].tolist(), st['ops_f'][i].tolist())
            print('     ref', rs['ops_i'][i].tolist(), rs['ops_f'][i].tolist())
