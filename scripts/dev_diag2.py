import sys, numpy as np
sys.path.insert(0, '.')
import da4ml_b200._binary as B
from da4ml_b200.types import pipeline_from_arrays
n, bits = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
W = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), size=(n, n)).astype(np.float32)
for G in (1, 16, 148):
    B.set_group_size(G)
    raw, tr = B.solve_single_raw(W, 'wmc', trace_cap=1 << 16)
    st = raw.stages[0]
    K = pipeline_from_arrays(raw.stages).solutions[0].kernel
    badc = np.unique(np.argwhere(K != W)[:, 1])
    oi, of = st['ops_i'], st['ops_f']
    zero = np.argwhere((oi[n:] == 0).all(1) & (of[n:] == 0).all(1))[:, 0] + n
    T = raw.counters[0]['T']
    print(f'G={G} n_ops={len(oi)} T={T} bad cols {badc.tolist()} zero-records {zero.tolist()[:20]} (count {len(zero)})')
    for c in badc[:4]:
        idx = st['out_idxs'][c]
        print('   col', c, 'out_idx', idx, 'op', oi[idx].tolist(), of[idx].tolist(), 'shift', st['out_shifts'][c], 'neg', st['out_negs'][c])
    bad_order = np.argwhere((oi[n:, 0] >= np.arange(n, len(oi))) | (oi[n:, 1] >= np.arange(n, len(oi))))
    print('   forward refs', len(bad_order))
    # which ops are referenced by nobody and are not outputs (dangling)
    used = np.zeros(len(oi), bool); used[oi[n:, 0]] = True; used[oi[n:, 1]] = True; used[st['out_idxs'][st['out_idxs'] >= 0]] = True
    print('   unused ops', np.argwhere(~used)[:, 0].tolist()[:20])
