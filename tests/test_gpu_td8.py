"""GPU parity tests of the 8-bit turbo decoder (phy_threegpplte_turbo_decoder8) through the C ABI.
Domain: K >= 256, K % 16 == 0."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader, vectors  # noqa: E402
from test_golden import iter_td8  # noqa: E402
from test_oracle_pin import ALL_K  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td8()
    return c


def test_td8_golden_vectors(capi):
    blocks, want = [], []
    for y, out, K, max_it, crc, F, ret in iter_td8():
        blocks.append({"y": y, "K": K, "max_iterations": max_it, "crc_type": crc, "F": F, "llr8": 1})
        want.append((out, ret))
    outs, status = capi.decode_batch(blocks)
    bad = [(i, b["K"], st, w[1]) for i, (w, ob, st, b) in enumerate(zip(want, outs, status, blocks))
           if st != w[1] or not np.array_equal(ob, w[0])]
    assert not bad, bad[:10]
    y, out, K, max_it, crc, F, ret = next(iter_td8())
    r, b = capi.phy_threegpplte_turbo_decoder8(y, K, 0, 0, max_it, crc, F)
    assert r == ret and np.array_equal(b, out)


def test_td8_all_sizes_vs_oracle_mixed_with_16bit(capi):
    """All 145 sizes of the domain, several scaling brackets; 16-bit blocks in the same submit."""
    blocks, want = [], []
    for i, K in enumerate(k for k in ALL_K if k >= 256 and k % 16 == 0):
        for regime, A in (("clean", 8), ("waterfall", 40), ("noise", 8), ("waterfall", 300), ("full", 8)):
            y, _ = vectors.llr_block(K, 4000 + i, regime, A=A, crc_type=i & 1)
            blocks.append({"y": y, "K": K, "max_iterations": 6, "crc_type": i & 1, "llr8": 1})
            want.append(loader.port_decode8(y, K, 6, i & 1))
        if i % 10 == 0:
            y, _ = vectors.llr_block(K, 4000 + i, "waterfall", crc_type=1)
            blocks.append({"y": y, "K": K, "max_iterations": 6, "crc_type": 1, "llr8": 0})
            want.append(loader.port_decode16(y, K, 6, 1))
    outs, status = capi.decode_batch(blocks)
    bad = [(i, b["K"], b.get("llr8"), st, w[1]) for i, (w, ob, st, b) in enumerate(zip(want, outs, status, blocks))
           if st != w[1] or not np.array_equal(ob, w[0])]
    assert not bad, bad[:10]


def test_td8_limits_and_domain(capi):
    blocks, want = [], []
    for max_it, crc in ((0, 1), (1, 1), (2, 0), (3, 2), (4, 3), (8, 1)):
        for regime in ("clean", "noise"):
            y, _ = vectors.llr_block(512, max_it, regime, crc_type=min(crc, 1))
            blocks.append({"y": y, "K": 512, "max_iterations": max_it, "crc_type": crc, "llr8": 1})
            want.append(loader.port_decode8(y, 512, max_it, crc))
    outs, status = capi.decode_batch(blocks)
    for (wb, wr), ob, st, b in zip(want, outs, status, blocks):
        assert st == wr, (b["max_iterations"], b["crc_type"], st, wr)
        if b["max_iterations"] > 1:
            assert np.array_equal(ob, wb)
    y = np.zeros(3 * 6144 + 12, dtype=np.int16)
    assert capi.phy_threegpplte_turbo_decoder8(y, 40, 0, 0, 4, 1, 0)[0] == 255      # outside the domain
    assert capi.phy_threegpplte_turbo_decoder8(y, 264, 0, 0, 4, 1, 0)[0] == 255
    assert capi.phy_threegpplte_turbo_decoder8(y, 512, 0, 0, 4, 5, 0)[0] == 255     # illegal crc_type
