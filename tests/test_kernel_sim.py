"""The product's CUDA kernel sources (da4ml_b200/csrc/*.cuh, unmodified) executed on the CPU under the SIMT shim of
tests/simt -- every CUDA thread a fiber, barriers and warp collectives as rendezvous points -- and compared bit for bit
with the oracle.  This checks the kernels' *logic* (substitution, recount, lazy histogram, chunk-cached argmax, group
exchange, adder-tree finisher, planner-chosen layouts) without a GPU; the GPU suite remains the parity proof on the real
hardware.  Small sizes only: the simulation runs a few thousand fibers on one core."""
import numpy as np
import pytest
from conftest import assert_stage_equal, golden_cases, int_matrix, load_golden

import simt
from oracle import port

METHODS = ['mc', 'mc-dc', 'mc-pdc', 'wmc', 'wmc-dc', 'wmc-pdc', 'dummy']


def test_shim_selftest():
    assert simt.lib().sim_selftest() == 0, simt.lib().sim_last_error()


@pytest.mark.parametrize('method', METHODS)
def test_every_selector_two_ctas(method):
    W = int_matrix(8, 8, 4, 0)
    got, meta = simt.solve_single(W, method, ctas=2, cta_threads=64)
    assert_stage_equal(got, port.solve_single(W, method), f'{method} ')
    assert meta[0] == 0 and meta[12] == 2


@pytest.mark.parametrize('ctas,threads', [(1, 32), (1, 128), (3, 64), (5, 32), (2, 512)])
def test_group_geometry_does_not_change_the_graph(ctas, threads):
    W = int_matrix(14, 11, 6, 3)
    got, _ = simt.solve_single(W, 'wmc', ctas=ctas, cta_threads=threads)
    assert_stage_equal(got, port.solve_single(W, 'wmc'), f'{ctas}x{threads} ')


def test_heterogeneous_intervals_latencies_and_adder_cost():
    rng = np.random.default_rng(5)
    W = int_matrix(16, 12, 7, 9) * np.float32(0.25)
    q = np.stack([-(2.0 ** rng.integers(0, 8, 16)), 2.0 ** rng.integers(0, 8, 16) - 0.5, np.full(16, 0.5)], axis=1).astype(np.float32)
    q[3] = (0.0, 0.0, 1.0)  # a dead input (state_opr.cc:92-97)
    lat = rng.integers(0, 3, 16).astype(np.float32)
    for method in ('wmc-dc', 'mc-pdc'):
        kw = dict(qintervals=[tuple(map(float, r)) for r in q], latencies=[float(v) for v in lat], adder_size=2, carry_size=4)
        got, _ = simt.solve_single(W, method, ctas=3, cta_threads=64, **kw)
        assert_stage_equal(got, port.solve_single(W, method, **kw), f'{method} ')


def test_accounting_mode_counts_the_reference_work():
    W = int_matrix(16, 16, 6, 11)
    want = port.solve_single(W, 'wmc')
    cnt = port.partial(W, 'wmc')  # exact work counters of the reference algorithm
    for ctas, threads in ((2, 64), (1, 32), (3, 128)):
        got, meta = simt.solve_single(W, 'wmc', ctas=ctas, cta_threads=threads, accounting=True)
        assert_stage_equal(got, want, 'accounting ')
        assert (meta[2], meta[3], meta[4], meta[5], meta[6]) == (cnt['iters'], cnt['sum_F'], cnt['sum_R'], cnt['F0'], cnt['R0'])


@pytest.mark.parametrize('name', ['all_zero', 'one_by_one', 'single_output', 'single_input', 'zero_cols', 'repeated'])
def test_edge_matrices(name):
    W = {
        'all_zero': np.zeros((5, 6), np.float32),
        'one_by_one': np.array([[5.0]], np.float32),
        'single_output': int_matrix(9, 1, 8, 2),
        'single_input': int_matrix(1, 9, 8, 1),
        'zero_cols': np.pad(int_matrix(5, 4, 6, 3), ((1, 1), (2, 1))),
        'repeated': np.tile(int_matrix(10, 2, 8, 8), (1, 4)),
    }[name]
    got, _ = simt.solve_single(W, 'wmc', ctas=2, cta_threads=64)
    assert_stage_equal(got, port.solve_single(W, 'wmc'), name + ' ')


def test_golden_single_stage_cases():
    """The committed reference outputs (tests/golden) for the single-stage cases, reproduced by the simulated kernels."""
    seen = 0
    for name, meta in golden_cases().items():
        if not name.startswith('single_'):
            continue
        extra, stages = load_golden(name)
        kw = {k: v for k, v in meta['kwargs'].items() if k in ('adder_size', 'carry_size')}
        got, _ = simt.solve_single(extra['kernel'], meta['kwargs']['method'], qintervals=extra.get('qint'), latencies=extra.get('lat'), ctas=3, cta_threads=64, **kw)
        assert_stage_equal(got, stages[0], name + ' ')
        seen += 1
    assert seen >= 4


def test_larger_matrix_five_ctas():
    W = int_matrix(22, 20, 8, 21)
    got, meta = simt.solve_single(W, 'wmc', ctas=5, cta_threads=64)
    assert_stage_equal(got, port.solve_single(W, 'wmc'), '22x20 ')
    assert meta[9] >= 0 and meta[14] > 0


def test_wide_columns_wide_digits_spills_and_hash_passes():
    """More than 32 output columns (several bitmap words), more than 16 CSD bits, owner lists that spill to global memory,
    and a pair-counter hash table so small that the owned expressions have to be counted in several subsets."""
    W = int_matrix(6, 70, 5, 31)
    got, _ = simt.solve_single(W, 'wmc', ctas=3, cta_threads=64)
    assert_stage_equal(got, port.solve_single(W, 'wmc'), '6x70 ')
    W = int_matrix(5, 4, 20, 4)
    got, _ = simt.solve_single(W, 'wmc', ctas=2, cta_threads=32)
    assert_stage_equal(got, port.solve_single(W, 'wmc'), '20-bit ')
    W = int_matrix(5, 100, 8, 7)
    got, _ = simt.solve_single(W, 'wmc-dc', ctas=2, cta_threads=64)
    assert_stage_equal(got, port.solve_single(W, 'wmc-dc'), 'dense 5x100 ')
    W = int_matrix(6, 90, 6, 12) * (np.random.default_rng(1).random((6, 90)) < 0.4)
    got, _ = simt.solve_single(W, 'wmc', ctas=3, cta_threads=64)
    assert_stage_equal(got, port.solve_single(W, 'wmc'), 'sparse 6x90 ')
    W = int_matrix(18, 14, 8, 23)
    want = port.solve_single(W, 'wmc')
    try:
        for rows in (0, 4):  # every row / the rows beyond the fourth in global memory
            simt.set_own_caps(list_rows=rows)
            got, meta = simt.solve_single(W, 'wmc', ctas=2, cta_threads=64)
            assert meta[15] == rows
            assert_stage_equal(got, want, f'owner lists with {rows} shared rows ')
        simt.set_own_caps()
        simt.set_wide_rows(True)  # three words per list row (what problems with more than 16 CSD bits use)
        got, _ = simt.solve_single(W, 'wmc', ctas=2, cta_threads=64)
        assert_stage_equal(got, want, 'wide list rows ')
        simt.set_wide_rows(False)
        simt.set_own_caps(hash_log=8)  # 256 slots, 160 usable per pass: 9 owned inputs x 3 rows x 30 counters do not fit
        got, _ = simt.solve_single(W, 'wmc', ctas=2, cta_threads=64)
        assert_stage_equal(got, want, 'small hash table ')
        got, _ = simt.solve_single(W, 'wmc', ctas=1, cta_threads=32, accounting=True)
        assert_stage_equal(got, want, 'small hash table, one CTA ')
        simt.set_own_caps(spill_rows=0, list_rows=6)  # no room to spill: the lists overflow
        with pytest.raises(RuntimeError, match='capacity status 5'):
            simt.solve_single(W, 'wmc', ctas=2, cta_threads=64)
    finally:
        simt.set_own_caps()
        simt.set_wide_rows(False)


def test_batched_launch_reuses_group_workspaces():
    """Five jobs of different shapes on two groups of two CTAs: three jobs run back to back in one group's workspace
    (counter slab / cell pools / barrier counters carried over from the previous job)."""
    mats = [int_matrix(12, 10, 6, 40), int_matrix(6, 14, 5, 41), int_matrix(9, 9, 7, 42), np.zeros((4, 5), np.float32), int_matrix(10, 33, 4, 43)]
    got = simt.solve_many(mats, 'wmc', ctas=2, groups=2, cta_threads=64)
    for i, (W, st) in enumerate(zip(mats, got)):
        assert_stage_equal(st, port.solve_single(W, 'wmc'), f'job {i} ')


@pytest.mark.parametrize('mode', [1, 2])
def test_result_does_not_depend_on_the_thread_schedule(mode):
    """The simulator resumes runnable threads in descending / pseudo-random order instead of ascending: a kernel that
    only works under one order is missing a barrier."""
    W = int_matrix(12, 36, 6, 17)
    want = port.solve_single(W, 'wmc-dc')
    simt.set_schedule(mode)
    try:
        got, _ = simt.solve_single(W, 'wmc-dc', ctas=3, cta_threads=64)
        many = simt.solve_many([W, int_matrix(7, 9, 5, 18)], 'wmc-dc', ctas=2, groups=1, cta_threads=64)
    finally:
        simt.set_schedule(0)
    assert_stage_equal(got, want, f'schedule {mode} ')
    assert_stage_equal(many[0], want, f'schedule {mode}, batched ')


@pytest.mark.parametrize('seed', range(0, 60, 2))
def test_random_family_through_the_kernel(seed):
    """The random family of tests/test_oracle_cross.py (ragged shapes, sparse / fractional / 12-bit / constant / diagonal
    matrices, heterogeneous, unsigned and dead input intervals, latencies, adder / carry sizes, every selector) through the
    simulated kernel."""
    from test_oracle_cross import METHODS as FAMILY_METHODS
    from test_oracle_cross import random_intervals, random_latencies, random_matrix, shapes

    rng = np.random.default_rng(7000 + seed)
    n_in, n_out = shapes(rng)
    W, kind = random_matrix(rng, n_in, n_out)
    kw = dict(qintervals=random_intervals(rng, n_in), latencies=random_latencies(rng, n_in))
    method = str(rng.choice(FAMILY_METHODS + ['dummy']))
    kw.update(adder_size=int(rng.choice([-1, 1, 3, 8])), carry_size=int(rng.choice([-1, 1, 4])))
    want = port.solve_single(W, method, **kw)
    got, _ = simt.solve_single(W, method, ctas=1 + seed % 3, cta_threads=64, **kw)
    assert_stage_equal(got, want, f'{kind} {W.shape} {method} ')


@pytest.mark.parametrize('dc', [-2, -1, 0, 1, 2, 5])
def test_decomposition_kernels(dc):
    """centre / all-pairs CSD distance / Prim MST + M0, M1 construction (cmvm_decompose.cuh) against the oracle."""
    for n_in, n_out, bits, seed in [(8, 8, 4, 0), (12, 20, 8, 1), (5, 33, 6, 2), (9, 1, 8, 3)]:
        W = int_matrix(n_in, n_out, bits, seed)
        m0, m1 = simt.kernel_decompose(W, dc)
        w0, w1 = port.kernel_decompose(W, dc)
        assert np.array_equal(m0, w0) and np.array_equal(m1, w1), (n_in, n_out, dc)
        assert np.array_equal(m0.astype(np.float64) @ m1.astype(np.float64), W)


def test_small_capacities_compact_or_report_never_hang():
    """With deliberately small buffers the kernels either still produce the reference graph (after compacting the
    histogram segment) or stop with a capacity status that the host driver answers with a retry -- never a deadlock.
    (The first version of this test found a real one: threads already harvesting changed what slower threads of the same
    CTA decided about a mid-step compaction.)"""
    W = int_matrix(16, 16, 6, 11)
    want = port.solve_single(W, 'wmc')
    outcomes = {}
    try:
        for cap in (2000, 1200, 700):
            simt.set_segment_cap(cap)
            try:
                got, meta = simt.solve_single(W, 'wmc', ctas=2, cta_threads=64)
                assert_stage_equal(got, want, f'segment {cap} ')
                outcomes[cap] = int(meta[9])
            except RuntimeError as e:
                assert 'capacity status 2' in str(e)
                outcomes[cap] = 'overflow'
        simt.set_segment_cap(0)
        assert outcomes[2000] != 'overflow' and outcomes[2000] >= 1  # compacted and finished
        assert outcomes[700] == 'overflow'
        for knob, status in ([dict(e_cap=20), 1], [dict(pool=150), 5]):
            simt.set_caps(**knob)
            with pytest.raises(RuntimeError, match=f'capacity status {status}'):
                simt.solve_single(W, 'wmc', ctas=2, cta_threads=64)
            simt.set_caps()
    finally:
        simt.set_segment_cap(0)
        simt.set_caps()


def test_race_check(tmp_path):
    """tests/simt/tsan_main.cc: the simulated kernels built with ThreadSanitizer, every CUDA thread announced as a TSan
    fiber; only __syncthreads / __syncwarp and the acquire-release group counters order memory (shuffles and votes do
    not).  No unordered conflicting accesses may be reported.  (This check found the missing group barrier after the
    stamp zeroing at problem start, and it flags a removed __syncwarp / __syncthreads.)"""
    import shutil
    import subprocess

    gxx = '/usr/bin/g++' if shutil.which('/usr/bin/g++') else 'g++'
    if shutil.which('setarch') is None or not subprocess.run([gxx, '-print-file-name=libtsan.so'], capture_output=True, text=True).stdout.strip().startswith('/'):
        pytest.skip('needs libtsan and setarch')
    exe = tmp_path / 'sim_tsan'
    subprocess.run([gxx, '-O1', '-g', '-fsanitize=thread', '-DSIMT_TSAN', '-std=c++17', '-ffp-contract=off', '-w', str(simt.HERE / 'tsan_main.cc'), '-o', str(exe)], check=True)
    for schedule in ('2',):  # pseudo-random thread order (the ascending one is what every other test of this file runs)
        r = subprocess.run(['setarch', 'x86_64', '-R', str(exe), schedule], capture_output=True, text=True, timeout=900)
        assert 'WARNING: ThreadSanitizer' not in r.stderr + r.stdout, (r.stderr + r.stdout)[-3000:]
        assert r.returncode == 0, (r.stderr + r.stdout)[-2000:]
        assert r.stdout.count(' ops, ') == 8 and 'FAILED' not in r.stdout and 'batch: rc 0' in r.stdout


def test_uncleared_buffers_may_hold_garbage():
    """The host driver clears only the counter slab, the barrier / exchange words and result_meta before a launch; every
    other buffer is filled with 0xA5 here and the graphs must not change."""
    mats = [int_matrix(14, 11, 6, 3), np.zeros((4, 5), np.float32), int_matrix(6, 40, 5, 31)]
    simt.set_poison(True)
    try:
        got = simt.solve_many(mats, 'wmc-dc', ctas=2, groups=1, cta_threads=64)
        simt.set_own_caps(list_rows=2)
        single, _ = simt.solve_single(mats[0], 'wmc-dc', ctas=3, cta_threads=64)
    finally:
        simt.set_poison(False)
        simt.set_own_caps()
    for W, st in zip(mats, got):
        assert_stage_equal(st, port.solve_single(W, 'wmc-dc'), 'poisoned ')
    assert_stage_equal(single, port.solve_single(mats[0], 'wmc-dc'), 'poisoned, spilled lists ')
