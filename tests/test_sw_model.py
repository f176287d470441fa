"""CPU model of the optional sliding-window mode (oracle/port/td16_sw_port.c).  The mode is this repo's own variant, so
there is no reference output to pin it on; what is checked here: it delivers the transmitted bits wherever the bit-exact
port does with margin, it obeys the reference's return-value rules, and its outputs on fixed seeded inputs do not drift
(a digest frozen when the model was written: the GPU kernel is tested bit for bit against the model, tests/test_gpu_sw.py)."""
import hashlib

import numpy as np
import pytest

from oracle import loader, vectors


@pytest.mark.parametrize("K", [40, 200, 512, 1008, 1024, 2048, 3904, 6144])
def test_clean_blocks_decode_to_the_transmitted_bits(K):
    for i, A in enumerate((6, 30, 400, 9000)):
        y, bits = vectors.llr_block(K, 10 + i, "clean", A=A, sigma_over_A=0.55)
        out, ret = loader.port_decode16_sw(y, K, 6, 1)
        ref, rret = loader.port_decode16(y, K, 6, 1)
        assert ret <= 6 and rret <= 6
        assert np.array_equal(out, ref)


def test_return_values():
    K = 1024
    y, _ = vectors.llr_block(K, 3, "clean", A=20, sigma_over_A=0.4)
    assert loader.port_decode16_sw(y, K, 1, 1)[1] == 2                 # one iteration: no CRC test at all (TD16:1267)
    assert loader.port_decode16_sw(y, K, 6, 1)[1] == 2                 # earliest exit is iteration 2
    noise = np.random.default_rng(1).integers(-50, 51, size=3 * K + 12).astype(np.int16)
    assert loader.port_decode16_sw(noise, K, 5, 1)[1] == 6
    assert loader.port_decode16_sw(y, 1000, 6, 1)[1] == 255            # not a turbo block size
    assert loader.port_decode16_sw(y, K, 6, 4)[1] == 255


def test_scaling_shift():
    P = loader.port()
    for A, want in ((10, 0), (24, 0), (25, 1), (200, 4), (25000, 10)):
        y = np.full(600, A, dtype=np.int16)
        assert P.orc_sw_shift(y, 600) == want, A
    assert [P.orc_sw_windows(K) for K in (40, 504, 512, 1008, 1024, 2016, 2048, 6144)] == [8, 8, 16, 16, 32, 32, 64, 64]


def test_model_outputs_are_frozen():
    h = hashlib.sha256()
    rng = np.random.default_rng(77)
    for i, K in enumerate((40, 320, 512, 1024, 1984, 2048, 4096, 6144)):
        for A, sig in ((8, 1.0), (300, 1.1), (5000, 0.9)):
            y, _ = vectors.llr_block(K, 500 + i, "waterfall", A=A, sigma_over_A=sig)
            out, ret = loader.port_decode16_sw(y, K, 6, 1)
            h.update(out.tobytes() + bytes([ret]))
        y = rng.integers(-32768, 32768, size=3 * K + 12).astype(np.int16)
        out, ret = loader.port_decode16_sw(y, K, 3, 0, F=16 if K >= 512 else 0)
        h.update(out.tobytes() + bytes([ret]))
    assert h.hexdigest() == FROZEN, h.hexdigest()


FROZEN = "0da57dbd2cd75eccc8c0b5769b496fd786fce39d11298c837335ec7f843cbbdd"
