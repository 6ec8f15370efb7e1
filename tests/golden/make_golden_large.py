"""Hours-long golden runs of the reference's own object code (oracle/_ref) for the configurations
bench.py measures: BASELINE config 3 (256x256 int8) and config 4 (128x128 int6, default solve).

    python tests/golden/make_golden_large.py c3_dc-1          # one candidate, 1 thread (~1 h)
    OMP_NUM_THREADS=5 python tests/golden/make_golden_large.py c3_default   # 10 candidates (~2 h; 5 and 8
                                                                             # threads both need two waves)
    python tests/golden/make_golden_large.py c4_default --seeds 0-7

Each run writes tests/golden/large/<name>.json (sha256 over the result arrays, adder count, stage op
counts, wall seconds, threads) and tests/golden/large/<name>.npz (the arrays, for diffing a mismatch);
`--merge` folds every json of that directory into tests/golden/index.json.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE))
from make_golden import KEYS, int_matrix, n_adders, stage_digest  # noqa: E402

OUT = HERE / 'large'


def run(name, spec, kw):
    from oracle import ref

    W = int_matrix(*spec[1:])
    threads = int(os.environ.get('OMP_NUM_THREADS', os.cpu_count()))
    t0 = time.time()
    stages = ref.solve(W, **kw)
    dt = time.time() - t0
    rec = dict(spec=list(spec), kwargs=kw, n_adders=n_adders(stages), sha256=stage_digest(stages), full=False,
               stage_ops=[int(len(st['ops_i'])) for st in stages], cpu_seconds=round(dt, 2), cpu_threads=threads,
               cpu_model=cpu_model(), stage_sha256=[stage_digest([st]) for st in stages])
    OUT.mkdir(exist_ok=True)
    (OUT / f'{name}.json').write_text(json.dumps(rec, indent=1, sort_keys=True))
    arrs = {}
    for i, st in enumerate(stages):
        for k in KEYS:
            arrs[f's{i}_{k}'] = st[k]
    np.savez_compressed(OUT / f'{name}.npz', **arrs)
    print(name, rec['n_adders'], f'{dt:.1f}s', flush=True)


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def seeds_of(s):
    out = []
    for part in s.split(','):
        if '-' in part:
            a, b = part.split('-')
            out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('what', nargs='?')
    ap.add_argument('--seeds', default='0')
    ap.add_argument('--merge', action='store_true')
    a = ap.parse_args()
    if a.merge:
        idx_path = HERE / 'index.json'
        index = json.loads(idx_path.read_text())
        for p in sorted(OUT.glob('*.json')):
            index[p.stem] = json.loads(p.read_text())
        idx_path.write_text(json.dumps(index, indent=1, sort_keys=True))
        return
    for seed in seeds_of(a.seeds):
        sfx = '' if seed == 0 else f'_s{seed}'
        if a.what == 'c3_dc-1':
            run(f'c3_256x256_int8_dc-1{sfx}', ('int', 256, 256, 8, seed), dict(search_all_decompose_dc=False, decompose_dc=-1))
        elif a.what == 'c3_default':
            run(f'c3_256x256_int8_default{sfx}', ('int', 256, 256, 8, seed), {})
        elif a.what == 'c4_default':
            run(f'c4_128x128_int6_default{sfx}', ('int', 128, 128, 6, seed), {})
        elif a.what == 'c4_dc-1':
            run(f'c4_128x128_int6_dc-1{sfx}', ('int', 128, 128, 6, seed), dict(search_all_decompose_dc=False, decompose_dc=-1))
        else:
            raise SystemExit(f'unknown job {a.what}')


if __name__ == '__main__':
    main()
