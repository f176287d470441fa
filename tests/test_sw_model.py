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


def test_reference_error_floor_on_punctured_blocks_comes_from_its_5_step_rerun(tmp_path):
    """on heavily punctured blocks (90 % of the parity soft bits zero, the MCS28 regime) the reference's decoder
    keeps losing blocks that are decodable.  Evidence for the cause: the same port with the lane-boundary re-run lengthened
    from 5 to 40 steps (TD16:171,189,232-259,541-549,585-587) loses none of them, and neither does the sliding-window model."""
    import ctypes as C
    import os
    import re
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "port")
    src = open(os.path.join(here, "td16_port.c")).read()
    long_src = re.sub(r"#define RERUN\s+\d+", "#define RERUN 40", src).replace("loopval = (n - 40) >> 3;", "loopval = (n - 320) >> 3;")
    assert long_src != src
    (tmp_path / "td16_long.c").write_text(long_src)
    others = [os.path.join(here, f) for f in os.listdir(here) if f.endswith(".c") and f != "td16_port.c"]
    lib = str(tmp_path / "liblong.so")
    subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-fPIC", "-shared", "-I" + here, "-o", lib, str(tmp_path / "td16_long.c")] + others + ["-lm", "-lpthread"])
    L = C.CDLL(lib)
    L.orc_turbo_decoder16.restype = C.c_uint8
    K, keep, lost = 5824, 10, [0, 0, 0]
    idx = np.arange(K)
    for i in range(60):
        y = vectors.llr_block(K, 300 + i, "waterfall", A=40, sigma_over_A=0.48)[0].copy()
        y[3 * idx[(idx % keep) != 0] + 1] = 0
        y[3 * idx[((idx + keep // 2) % keep) != 0] + 2] = 0
        out = np.zeros(K // 8 + 4, dtype=np.uint8)
        lost[0] += loader.port_decode16(y, K, 6, 1)[1] > 6
        lost[1] += L.orc_turbo_decoder16(y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_uint16(K), C.c_uint8(6), C.c_uint8(1), C.c_uint8(0)) > 6
        lost[2] += loader.port_decode16_sw(y, K, 6, 1)[1] > 6
    assert lost[0] >= 2 and lost[1] == 0 and lost[2] == 0, lost
