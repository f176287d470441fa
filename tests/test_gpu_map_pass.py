"""Kernel-level GPU parity test of ONE max-log-MAP pass (the reference's log_map16, TD16:84-879) through the test hook
oai_turbo_debug_map16, against the oracle port's orc_log_map16 (pinned on the compiled reference's log_map16 by
tests/test_oracle_pin.py): a-posteriori LLRs of every trellis step, bit for bit.  Inputs are crafted for the corners of
the exact policy's unsigned representation: branch metrics of exactly -16384 ("hazard" steps: sat(s +- p) <= -32767),
sparse (per-segment body choice) and dense (whole pass in the hazard form), metrics saturated at -32768, tail metrics
that make the first beta step positive."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def port_map(P, y, K, term):
    """the reference's lane layout (position k of lane l at 8*k + l), tails as the decoder places them (TD16:1075-1130)"""
    W = K // 8
    pos = np.arange(K)
    st = (pos % W) * 8 + pos // W
    s = np.zeros(K + 16, np.int16)
    p = np.zeros(K + 16, np.int16)
    s[st] = y[0:3 * K:3]
    p[st] = y[(2 if term else 1):3 * K:3]
    t = y[3 * K:]
    for i in range(3):
        if term == 0:
            s[K + i] = t[2 * i]; p[K + i] = t[2 * i + 1]
        else:
            s[K + 8 + i] = t[6 + 2 * i]; p[K + i] = t[7 + 2 * i]
    e = np.zeros(K + 16, np.int16)
    P.orc_log_map16(s, p, e, K, term, None, None)
    return e[:K]


def craft(K, kind, rng):
    n = 3 * K + 12
    if kind == "sparse_hazard":                 # +-3000 noise (exact policy), ~1 % of the steps with sat(s+p) or sat(s-p) at the floor
        y = rng.integers(-3000, 3001, size=n).astype(np.int16)
        idx = rng.choice(K, size=max(2, K // 100), replace=False)
        for j, k in enumerate(idx):
            s, p1, p2 = [(-20000, -13000, 13000), (-16384, -16383, 16384), (-32768, -1, 1), (-16384, -16384, 16383)][j % 4]
            y[3 * k], y[3 * k + 1], y[3 * k + 2] = s, p1, p2
    elif kind == "dense_hazard":                # every other step at the floor
        y = rng.integers(-3000, 3001, size=n).astype(np.int16)
        y[0:3 * K:6] = -32768
        y[1:3 * K:6] = rng.choice(np.array([-32768, -1, 0, 1, 32767], dtype=np.int16), size=y[1:3 * K:6].size)
        y[2:3 * K:6] = rng.choice(np.array([-32768, -1, 0, 1, 32767], dtype=np.int16), size=y[2:3 * K:6].size)
    elif kind == "edges":                       # sums right at -32768 / -32767 / -32766 and +32766 / +32767
        y = rng.integers(-200, 201, size=n).astype(np.int16)
        for k in range(0, K, 3):
            s, p = [(-16384, -16384), (-16384, -16383), (-16383, -16383), (16383, 16383), (16384, 16383), (-32768, 32767),
                    (32767, -32768), (0, -32768), (-1, -32767)][(k // 3) % 9]
            y[3 * k], y[3 * k + 1], y[3 * k + 2] = s, p, -p if p != -32768 else 32767
    elif kind == "saturated":                   # strong consistent signal: path metrics pile up at -32768
        y = np.where(rng.random(n) < 0.5, 30000, -30000).astype(np.int16)
        y[::7] = rng.integers(-32768, 32768, size=y[::7].size)
    elif kind == "positive_tail":               # large positive termination values: the first beta step starts > 0
        y = rng.integers(-2500, 2501, size=n).astype(np.int16)
        y[3 * K:] = np.array([32767, 32767, -32768, 32767, 32767, -32768, 32767, 32767, 32767, -32768, 32767, 32767], dtype=np.int16)
    else:
        raise ValueError(kind)
    return y


@pytest.mark.parametrize("kind", ["sparse_hazard", "dense_hazard", "edges", "saturated", "positive_tail"])
def test_one_map_pass_exact_policy_corners(capi, port, kind):
    rng = np.random.default_rng(sum(map(ord, kind)))
    for K in (40, 48, 64, 104, 512, 1056, 6144):
        for rep in range(2):
            y = craft(K, kind, rng)
            for term in (0, 1):
                want = port_map(port, y, K, term)
                for policy in (2, 0):           # 2: exact policy forced, 0: the guard decides (must pick the exact one here)
                    got = capi.debug_map16(y, K, term, policy)
                    d = np.nonzero(got != want)[0]
                    assert d.size == 0, (kind, K, term, policy, d.size, [(int(i), int(got[i]), int(want[i])) for i in d[:5]])


def test_one_map_pass_fast_policy_matches_where_the_guard_allows(capi, port):
    """policy 1 forces the non-saturating fast path; inside the guard (|y| <= 1000 here) it must equal the reference."""
    rng = np.random.default_rng(3)
    for K in (40, 512, 6144):
        for amp in (16, 300, 1000):
            y = rng.integers(-amp, amp + 1, size=3 * K + 12).astype(np.int16)
            for term in (0, 1):
                want = port_map(port, y, K, term)
                for policy in (1, 2, 0):
                    assert np.array_equal(capi.debug_map16(y, K, term, policy), want), (K, amp, term, policy)


def test_one_map_pass_tracked_policy(capi, port):
    """policy 4 forces the tracked fast pass (fast arithmetic + a-posteriori range certificate, retry on the exact policy
    when it fails): beyond the a-priori guard it must equal the reference whether the certificate holds (moderate
    amplitudes, coded signals) or not (amplitudes at which the reference saturates), for every K class incl. K = 40 where
    the re-run covers the lane, with ordinary and extreme tail metrics; policy 0 (decide) must agree."""
    from bench import coded_inputs
    rng = np.random.default_rng(17)
    for K in (40, 48, 104, 512, 1056, 6144):
        cases = [rng.integers(-amp, amp + 1, size=3 * K + 12).astype(np.int16) for amp in (700, 1500, 3000, 5000, 8000, 12000, 30000)]
        y = rng.integers(-2500, 2501, size=3 * K + 12).astype(np.int16)
        y[3 * K:] = np.array([32767, 32767, -32768, 32767, 32767, -32768, 32767, 32767, 32767, -32768, 32767, 32767], dtype=np.int16)
        cases.append(y)                                                    # tail metrics far outside the body's range
        y = rng.integers(-300, 301, size=3 * K + 12).astype(np.int16)
        y[3 * (K // 2)] = 9000                                              # one outlier drives the maxima, the spreads stay small
        cases.append(y)
        if K in (512, 6144):
            for A in (128, 256, 512, 1500):
                cases.append(coded_inputs(K, 1, 1.08, 99 + A, A=A)[0][0])
        for y in cases:
            for term in (0, 1):
                want = port_map(port, y, K, term)
                for policy in (4, 0):
                    got = capi.debug_map16(y, K, term, policy)
                    d = np.nonzero(got != want)[0]
                    assert d.size == 0, (K, int(np.abs(y).max()), term, policy, d.size, [(int(i), int(got[i]), int(want[i])) for i in d[:5]])
