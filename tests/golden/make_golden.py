"""Generate golden vectors for the CMVM path from the reference's own code (oracle/_ref, i.e. the
unmodified reference TUs compiled in place).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py            # small cases (seconds)
    python tests/golden/make_golden.py --large    # adds 64x64 / 128x128 cases (minutes of CPU)

Small cases store the complete result arrays; large cases store adder counts, work counters and a
SHA-256 over the result arrays (the arrays themselves would be MBs).
"""
import argparse
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from oracle import ref  # noqa: E402

KEYS = ['inp_shifts', 'out_idxs', 'out_shifts', 'out_negs', 'ops_i', 'ops_f']


def int_matrix(n_in, n_out, bits, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-(2 ** (bits - 1)), 2 ** (bits - 1), size=(n_in, n_out)).astype(np.float32)


def pytest_matrix(n, bits, seed):
    # the generator of the reference's tests/test_cmvm.py:19, seeded
    rng = np.random.default_rng(seed)
    return np.round((rng.random((n, n)) - 0.5) * 2 ** (bits + 1)).astype(np.float32)


def stage_digest(stages):
    h = hashlib.sha256()
    for st in stages:
        for k in KEYS:
            a = np.ascontiguousarray(st[k])
            h.update(k.encode())
            h.update(str(a.shape).encode())
            h.update(a.tobytes())
    return h.hexdigest()


def n_adders(stages):
    return int(sum(int(np.count_nonzero(st['ops_i'][:, 2] >= 0)) for st in stages))


SMALL = [
    # name, matrix spec, solve kwargs
    ('c1_8x8_int4_default', ('int', 8, 8, 4, 0), {}),
    ('c1_8x8_int4_dc-1', ('int', 8, 8, 4, 0), dict(search_all_decompose_dc=False, decompose_dc=-1)),
    ('pytest_8_b4_harddc2_add1', ('pytest', 8, 4, 1), dict(hard_dc=2, adder_size=1, carry_size=-1)),
    ('pytest_8_b8_harddc0_mc', ('pytest', 8, 8, 2), dict(hard_dc=0, method0='mc', method1='mc', adder_size=1, carry_size=-1, search_all_decompose_dc=False)),
    ('pytest_4_b2_mc_wmc', ('pytest', 4, 2, 3), dict(method0='mc', method1='wmc', decompose_dc=0, hard_dc=2, adder_size=1, carry_size=-1)),
    ('int_16x16_int8_default', ('int', 16, 16, 8, 4), {}),
    ('int_17x5_int8_mcpdc', ('int', 17, 5, 8, 5), dict(method0='mc-pdc', method1='wmc-pdc', adder_size=4, carry_size=2)),
    ('int_12x20_int6_harddc1', ('int', 12, 20, 6, 6), dict(hard_dc=1, adder_size=2, carry_size=8)),
    ('int_32x32_int8_default', ('int', 32, 32, 8, 7), {}),
    ('int_48x40_int8_harddc2_wmcpdc', ('int', 48, 40, 8, 21), dict(hard_dc=2, method0='wmc-pdc', adder_size=4, carry_size=8)),
    ('int_40x56_int6_mcdc_dc1', ('int', 40, 56, 6, 22), dict(method0='mc-dc', method1='mc', decompose_dc=1, hard_dc=4, search_all_decompose_dc=False)),
    ('int_56x24_int7_harddc0', ('int', 56, 24, 7, 23), dict(hard_dc=0, adder_size=2, carry_size=-1)),
]
LARGE = [
    ('c2_64x64_int8_default', ('int', 64, 64, 8, 0), {}),
    ('c2_64x64_int8_dc-1', ('int', 64, 64, 8, 0), dict(search_all_decompose_dc=False, decompose_dc=-1)),
    ('c4_128x128_int6_dc-1', ('int', 128, 128, 6, 0), dict(search_all_decompose_dc=False, decompose_dc=-1)),
    ('128x128_int8_dc-1', ('int', 128, 128, 8, 0), dict(search_all_decompose_dc=False, decompose_dc=-1)),
]


def build(spec):
    if spec[0] == 'int':
        return int_matrix(*spec[1:])
    return pytest_matrix(*spec[1:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--large', action='store_true')
    args = ap.parse_args()
    index = {}
    idx_path = HERE / 'index.json'
    if idx_path.exists():
        index = json.loads(idx_path.read_text())
    for name, spec, kw in SMALL:
        W = build(spec)
        stages = ref.solve(W, **kw)
        arrs = {'kernel': W}
        for i, st in enumerate(stages):
            for k in KEYS:
                arrs[f's{i}_{k}'] = st[k]
        np.savez_compressed(HERE / f'{name}.npz', **arrs)
        index[name] = dict(spec=list(spec), kwargs=kw, n_adders=n_adders(stages), sha256=stage_digest(stages), full=True)
        print(name, index[name]['n_adders'], flush=True)
    # heterogeneous input intervals / latencies, single stage with trace
    rng = np.random.default_rng(11)
    W = int_matrix(16, 12, 6, 11)
    q = np.stack([-(2.0 ** rng.integers(0, 8, 16)), 2.0 ** rng.integers(0, 8, 16) - 2.0 ** -2, np.full(16, 0.25)], axis=1).astype(np.float32)
    q[3] = (0.0, 0.0, 1.0)
    lat = rng.integers(0, 4, 16).astype(np.float32)
    for method in ('wmc', 'wmc-dc', 'mc-dc', 'wmc-pdc'):
        st = ref.solve_single(W, method, q, lat, adder_size=3, carry_size=4)
        tr = ref.trace(W, method, q, lat, adder_size=3, carry_size=4)
        name = f'single_16x12_hetero_{method}'
        np.savez_compressed(HERE / f'{name}.npz', kernel=W, qint=q, lat=lat, pairs=tr['pairs'], f_sizes=tr['f_sizes'], **{f's0_{k}': st[k] for k in KEYS})
        index[name] = dict(spec=['hetero', 16, 12, 6, 11], kwargs=dict(method=method, adder_size=3, carry_size=4), n_adders=n_adders([st]), sha256=stage_digest([st]), full=True, single=True)
        print(name, index[name]['n_adders'], flush=True)
    if args.large:
        for name, spec, kw in LARGE:
            W = build(spec)
            stages = ref.solve(W, **kw)
            index[name] = dict(spec=list(spec), kwargs=kw, n_adders=n_adders(stages), sha256=stage_digest(stages), full=False,
                               stage_ops=[int(len(st['ops_i'])) for st in stages])
            print(name, index[name]['n_adders'], flush=True)
            idx_path.write_text(json.dumps(index, indent=1, sort_keys=True))
    idx_path.write_text(json.dumps(index, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()
