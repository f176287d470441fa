"""CPU-only: host-side integer logic of the C ABI (no kernels involved) against the oracle."""
import ctypes as C

import pytest


def test_segmentation_params_match_oracle(port):
    """oai_lte_segmentation_params follows lte_segmentation.c:52-134 (parameter part): compared with the oracle
    port (itself pinned to the compiled reference in test_oracle_pin.py) over small, boundary and large B."""
    from openair4g_b200 import capi
    Bs = list(range(24, 6300, 7)) + list(range(6100, 100000, 211)) + [40, 41, 512, 513, 1024, 1025, 2048, 2049, 6144, 6145,
                                                                     75376 + 24, 30576 + 24, 7736 + 24, 97920, 97921, 200000]
    for B in Bs:
        v = [C.c_uint32(0) for _ in range(6)]
        r2 = port.orc_lte_segmentation(B, *[C.byref(x) for x in v])
        want = tuple(x.value for x in v)
        r1, got = capi.lte_segmentation_params(B)
        assert r1 == r2, B
        if r1 == 0:
            assert tuple(got[k] for k in ("C", "Cplus", "Cminus", "Kplus", "Kminus", "F")) == want, (B, got, want)
    assert capi.lte_segmentation_params(75376 + 24)[1] == {"C": 13, "Cplus": 13, "Cminus": 0, "Kplus": 5824, "Kminus": 5760, "F": 0}


def test_harness_gold_sequence_matches_oracle(port):
    """The harness' own 36.211 7.2 sequence generator (openair4g_b200/sim/txchain.py) against the oracle port of
    lte_gold_generic (pinned to the compiled reference in test_oracle_pin.py)."""
    import numpy as np
    from openair4g_b200.sim import txchain
    for c_init in (0, (0x1234 << 14) + (7 << 9), (0xFFFF << 14) + (1 << 13) + (9 << 9) + 503, 0x7FFFFFFF):
        n = 32 * 40
        words = np.zeros(n // 32, dtype=np.uint32)
        port.orc_gold_words(c_init, words.ctypes.data, words.size)
        pos = np.arange(n)
        want = ((words[pos >> 5] >> (pos & 31).astype(np.uint32)) & 1).astype(np.uint8)
        assert np.array_equal(txchain.gold_sequence(c_init, n), want), hex(c_init)


def test_ulsch_control_sizes_match_the_port():
    """oai_ulsch_control_sizes (host integer rule of the C ABI, ulsch_decoding.c:381-468) against the pinned port"""
    import ctypes as C
    import numpy as np
    from openair4g_b200 import capi
    from oracle import loader
    P = loader.port()
    rng = np.random.default_rng(5)
    for _ in range(400):
        nb_rb = int(rng.integers(1, 101))
        Qm = int(rng.choice([2, 4, 6]))
        Nsymb = int(rng.choice([9, 10, 11, 12]))
        O_RI, O_ACK, Or1 = int(rng.integers(0, 2)), int(rng.integers(0, 3)), int(rng.choice([0, 4, 11, 12, 20, 64]))
        sumKr = int(rng.integers(40, nb_rb * 144 * Qm + 41))
        betas = [int(rng.integers(8, 161)) for _ in range(3)]
        z = loader.UlSizes()
        want = P.orc_ulsch_control_sizes(O_RI, O_ACK, Or1, 12 * nb_rb, Nsymb, betas[0], betas[1], betas[2], sumKr, nb_rb, Qm, Nsymb, C.byref(z))
        rc, got = capi.ulsch_control_sizes(O_RI, O_ACK, Or1, 12 * nb_rb, Nsymb, betas[0], betas[1], betas[2], sumKr, nb_rb, Qm, Nsymb)
        assert rc == want
        if rc == 0:
            assert got == {"Qprime_RI": z.Qprime_RI, "Qprime_ACK": z.Qprime_ACK, "Qprime_CQI": z.Qprime_CQI, "G": z.G,
                           "Hprime": z.Hprime, "Hpp": z.Hpp}
