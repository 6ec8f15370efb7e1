"""Scalar numerics of the product (host side of cmvm_num.cuh through the C ABI) against the CPU checker,
which calls libm exactly as the reference does.  No GPU needed: these entry points are plain functions."""
import numpy as np
import pytest

import da4ml_b200._binary as B
from oracle import port


def _pow2_neighbours():
    vals = []
    for e in range(-40, 41):
        base = np.float32(2.0**e)
        u = base.view(np.uint32)
        for d in range(-40, 41):
            vals.append(np.uint32(int(u) + d).view(np.float32))
    return [float(v) for v in vals]


def test_iceil_log2_and_lsb_loc():
    rng = np.random.default_rng(0)
    xs = list(rng.standard_normal(2000).astype(np.float32) * 1000) + _pow2_neighbours() + [0.0, 1.0, 0.5, 3.0, 2.0**-126, float(np.float32(np.inf))]
    for x in xs:
        x = float(np.float32(x))
        assert B.iceil_log2(x) == port.iceil_log2(x), x
        assert B.get_lsb_loc(x) == port.get_lsb_loc(x), x


def test_log2f_stand_in_matches_libm_where_it_matters():
    # trunc() and ceil() of log2f are what the reference uses (cmvm_core.cc:137,187; state_opr.cc:57-62);
    # (float)log2((double)x) must give the same integers as glibc's log2f next to every power of two
    for x in _pow2_neighbours():
        if x <= 0 or not np.isfinite(x):
            continue
        got = np.float32(np.log2(np.float64(x)))
        want = np.float32(port.log2f(x))
        assert np.trunc(got) == np.trunc(want) and np.ceil(got) == np.ceil(want), x


@pytest.mark.parametrize('seed', range(4))
def test_cost_add_matches_checker(seed):
    rng = np.random.default_rng(seed)
    for _ in range(300):
        steps = 2.0 ** rng.integers(-6, 4, 2)
        q0 = (-float(rng.integers(0, 2) * rng.integers(1, 1 << 12)) * steps[0], float(rng.integers(1, 1 << 12)) * steps[0], steps[0])
        q1 = (-float(rng.integers(0, 2) * rng.integers(1, 1 << 12)) * steps[1], float(rng.integers(1, 1 << 12)) * steps[1], steps[1])
        shift, sub = int(rng.integers(-8, 9)), bool(rng.integers(0, 2))
        for a, c in ((-1, -1), (1, -1), (-1, 4), (6, 3)):
            assert B.cost_add(q0, q1, shift, sub, a, c) == port.cost_add(q0, q1, shift, sub, a, c)


def test_cost_add_power_of_two_edges():
    # |max+step| landing exactly on / next to a power of two is where log2 rounding could bite
    for e in range(1, 24):
        for d in (-2, -1, 0, 1, 2):
            q0 = (-float(2**e + d), float(2**e + d - 1), 1.0)
            q1 = (-128.0, 127.0, 1.0)
            assert B.cost_add(q0, q1, 0, False, 1, 1) == port.cost_add(q0, q1, 0, False, 1, 1), (e, d)


def test_csd_weight_is_naf_weight():
    # the device computes CSD digit counts with the NAF bit trick; the checker runs the reference's threshold loop
    def naf_weight(x):
        u = abs(int(x))
        h = u >> 1
        return bin((u + h) ^ h).count('1')

    rng = np.random.default_rng(5)
    xs = list(range(-600, 601)) + [int(v) for v in rng.integers(-(2**24), 2**24, 3000)]
    for x in xs:
        assert port.csd_weight(x) == naf_weight(x), x
