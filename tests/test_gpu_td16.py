"""GPU parity tests of the 16-bit turbo decoder, all through the C ABI
(openair4g_b200.capi -> liboai_turbo_b200.so).  Bit-exact bar: decoded bytes and the
return value (iterations used / max+1 / 255) equal the oracle's on identical int16 input."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader, vectors  # noqa: E402
from test_golden import iter_td16  # noqa: E402
from test_oracle_pin import ALL_K  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def _check(capi, blocks, want):
    outs, status = capi.decode_batch(blocks)
    bad = []
    for i, ((wb, wr), ob, st, b) in enumerate(zip(want, outs, status, blocks)):
        if st != wr or (b["max_iterations"] > 1 and not np.array_equal(ob, wb)):
            bad.append((i, b["K"], st, wr))
    assert not bad, bad[:10]


def test_golden_vectors_batch(capi):
    """The committed compiled-reference vectors, submitted as ONE mixed-K batch."""
    blocks, want = [], []
    for y, out, K, max_it, crc, F, ret in iter_td16():
        blocks.append({"y": y, "K": K, "max_iterations": max_it, "crc_type": crc, "F": F})
        want.append((out, ret))
    _check(capi, blocks, want)


def test_golden_vectors_single_call(capi):
    """Same vectors through the reference-signature entry point, one call per block."""
    n = 0
    for y, out, K, max_it, crc, F, ret in iter_td16():
        if n % 3 == 0:
            r, b = capi.phy_threegpplte_turbo_decoder16(y, K, 0, 0, max_it, crc, F)
            assert r == ret, (K, max_it, crc, F, r, ret)
            if max_it > 1:
                assert np.array_equal(b, out), (K, max_it, crc, F)
        n += 1


def test_all_188_block_sizes_vs_oracle(capi):
    blocks, want = [], []
    for i, K in enumerate(ALL_K):
        for regime in ("clean", "waterfall", "noise", "full"):
            y, _ = vectors.llr_block(K, 1000 + i, regime, crc_type=i & 1)
            blocks.append({"y": y, "K": K, "max_iterations": 6, "crc_type": i & 1})
            want.append(loader.port_decode16(y, K, 6, i & 1))
    _check(capi, blocks, want)
    assert {w[1] for w in want} >= {2, 3, 4, 7}


@pytest.mark.parametrize("A", [4, 32, 128, 600, 1024, 1500, 4000, 8192, 30000])
def test_amplitudes_across_the_fast_path_guard(capi, A):
    """Amplitudes below, around and far above the non-saturating fast-path guard: the guard
    must hand hot blocks to the exact saturating kernel."""
    blocks, want = [], []
    for K in (40, 200, 512, 1056, 2048, 6144):
        for regime in ("clean", "waterfall"):
            y, _ = vectors.llr_block(K, A, regime, A=A)
            blocks.append({"y": y, "K": K, "max_iterations": 6, "crc_type": 1})
            want.append(loader.port_decode16(y, K, 6, 1))
    _check(capi, blocks, want)


def test_iteration_limits_crc_types_filler_bad_args(capi):
    blocks, want = [], []
    for K in (40, 504, 1024, 3904):
        for max_it in (0, 1, 2, 3, 4, 8):
            for regime in ("clean", "waterfall", "noise"):
                y, _ = vectors.llr_block(K, max_it, regime)
                blocks.append({"y": y, "K": K, "max_iterations": max_it, "crc_type": 1})
                want.append(loader.port_decode16(y, K, max_it, 1))
        for crc in (0, 1, 2, 3):
            y, _ = vectors.llr_block(K, 9, "clean", crc_type=min(crc, 1))
            blocks.append({"y": y, "K": K, "max_iterations": 4, "crc_type": crc})
            want.append(loader.port_decode16(y, K, 4, crc))
        for F in (8, 16, 40):
            if F < K - 24:
                y, _ = vectors.llr_block(K, F, "clean", crc_type=0, F=F)
                blocks.append({"y": y, "K": K, "max_iterations": 6, "crc_type": 0, "F": F})
                want.append(loader.port_decode16(y, K, 6, 0, F))
    # max_iterations 0/1: the reference never writes decoded_bytes; compare status only
    outs, status = capi.decode_batch(blocks)
    for (wb, wr), ob, st, b in zip(want, outs, status, blocks):
        assert st == wr, (b["K"], b["max_iterations"], st, wr)
        if b["max_iterations"] > 1:
            assert np.array_equal(ob, wb)
    y = np.zeros(3 * 520 + 12, dtype=np.int16)
    assert capi.phy_threegpplte_turbo_decoder16(y, 520, 0, 0, 4, 1, 0)[0] == 255    # illegal K
    assert capi.phy_threegpplte_turbo_decoder16(y, 512, 0, 0, 4, 4, 0)[0] == 255    # illegal crc_type
    outs, status = capi.decode_batch([{"y": y, "K": 520, "max_iterations": 4, "crc_type": 1},
                                      {"y": y, "K": 512, "max_iterations": 4, "crc_type": 7}])
    assert status == [255, 255]


def test_dl_stop_after_first_failure(capi):
    """dlsch_decoding.c:400,417,448-451: blocks after the first failing one are not decoded."""
    K = 1024
    blocks = []
    for r in range(5):
        y, _ = vectors.llr_block(K, r, "noise" if r == 2 else "clean")
        blocks.append({"y": y, "K": K, "max_iterations": 4, "crc_type": 1, "tb_id": 7})
    for r in range(3):
        y, _ = vectors.llr_block(K, 10 + r, "clean")
        blocks.append({"y": y, "K": K, "max_iterations": 4, "crc_type": 1, "tb_id": 8})
    outs, status = capi.decode_batch(blocks, flags=capi.BATCH_DL_STOP_AFTER_FAILURE)
    assert status[:3] == [2, 2, 5] and status[3:5] == [0xFE, 0xFE] and status[5:] == [2, 2, 2]
    assert not outs[3].any() and not outs[4].any() and outs[5].any()


def test_device_resident_plan_large_batch(capi):
    """Throughput-mode entry point: inputs/outputs stay in HBM (torch only supplies memory)."""
    import torch
    K, n = 6144, 600
    ys, want = [], []
    for i in range(12):
        y, _ = vectors.llr_block(K, 50 + i, ("clean", "waterfall", "noise")[i % 3])
        ys.append(y)
        want.append(loader.port_decode16(y, K, 6, 1))
    stride = 3 * K + 12 + 4
    host = np.zeros((n, stride), dtype=np.int16)
    for i in range(n):
        host[i, :3 * K + 12] = ys[i % 12]
    y_dev = torch.from_numpy(host).cuda()
    out_dev = torch.zeros((n, K // 8), dtype=torch.uint8, device="cuda")
    st_dev = torch.zeros(n, dtype=torch.uint8, device="cuda")
    plan = capi.DevPlan(n, K, 6, 1)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):                      # second run reuses the workspace
        launches = plan.decode(y_dev.data_ptr(), stride, out_dev.data_ptr(), K // 8, st_dev.data_ptr(), s)
        torch.cuda.synchronize()
    assert launches > 0
    out, st = out_dev.cpu().numpy(), st_dev.cpu().numpy()
    for i in range(n):
        assert st[i] == want[i % 12][1], i
        assert np.array_equal(out[i], want[i % 12][0]), i
    plan.close()


def test_tracked_passes_in_whole_decodes(capi):
    """Whole decodes at demapper-like amplitudes, i.e. beyond the a-priori guard: the passes run tracked (fast arithmetic,
    a-posteriori range certificate), hand over to the exact policy when the certificate fails or gets close, and blocks of
    all three classes share one batch (and hence warps and compaction lists).  Random K out of all 188 sizes, coded signals
    at A = 100 ... 6000 in the clean / waterfall / hopeless regimes, uniform noise, 4 or 6 iterations; bytes and return
    values against the port."""
    from openair4g_b200.sim import txchain
    rng = np.random.default_rng(2026)
    Ks = txchain.k_list()
    blocks, want = [], []
    for i in range(480):
        K = int(Ks[rng.integers(0, len(Ks))]) if i % 4 else (6144, 5824, 3904, 40)[(i // 4) % 4]
        A = int(rng.choice([100, 180, 256, 400, 700, 1200, 2500, 6000]))
        kind = i % 5
        if kind == 4:
            y = rng.integers(-A, A + 1, size=3 * K + 12).astype(np.int16)
        else:
            y, _ = vectors.llr_block(K, 5000 + i, "clean" if kind < 2 else "waterfall", A=A, sigma_over_A=(0.5, 0.8, 1.08, 1.6)[kind])
        max_it = 4 if i % 3 == 0 else 6
        blocks.append({"y": y, "K": K, "max_iterations": max_it, "crc_type": 1})
        want.append(loader.port_decode16(y, K, max_it, 1))
    blocks += [{"y": vectors.llr_block(6144, 7, "noise")[0], "K": 6144, "max_iterations": 6, "crc_type": 1}]       # one block inside the guard
    want.append(loader.port_decode16(blocks[-1]["y"], 6144, 6, 1))
    _check(capi, blocks, want)
    assert len({w[1] for w in want}) >= 4                      # several different return values (2 ... max+1) were exercised
