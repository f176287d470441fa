"""Optional sliding-window mode (OAI_BATCH_SLIDING_WINDOW; NOT bit-exact with the reference by design): the kernel
(openair4g_b200/csrc/td16_sw.cuh) against its CPU model (oracle/port/td16_sw_port.c) bit for bit -- decoded bytes and
return values --, and against the bit-exact mode where both must agree (blocks far from the decoding threshold)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader, vectors  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def _check(capi, blocks, want):
    outs, status = capi.decode_batch(blocks, flags=capi.BATCH_SLIDING_WINDOW)
    bad = []
    for i, ((wb, wr), ob, st, b) in enumerate(zip(want, outs, status, blocks)):
        if st != wr or (b["max_iterations"] > 1 and not np.array_equal(ob, wb)):
            bad.append((i, b["K"], st, wr))
    assert not bad, (len(bad), bad[:10])


def _want(b):
    return loader.port_decode16_sw(b["y"], b["K"], b["max_iterations"], b["crc_type"], b.get("F", 0))


@pytest.mark.parametrize("K", [40, 48, 104, 504, 512, 528, 1008, 1024, 1056, 2016, 2048, 2112, 3904, 5824, 6144])
def test_one_size_against_the_model(capi, K):
    """each window class (8 / 16 / 32 / 64 windows) and both ends of its K range: clean, waterfall and hopeless blocks,
    small and large amplitudes (scaling shift 0...7), 1 ... 6 iterations"""
    rng = np.random.default_rng(K)
    blocks = []
    for i in range(24):
        A = (8, 40, 300, 2500, 20000)[i % 5]
        kind = i % 4
        if kind == 3:
            y = rng.integers(-A, A + 1, size=3 * K + 12).astype(np.int16)
        else:
            y, _ = vectors.llr_block(K, 900 + i, "clean" if kind == 0 else "waterfall", A=A, sigma_over_A=(0.5, 1.0, 1.3)[kind])
        blocks.append({"y": y, "K": K, "max_iterations": (6, 4, 2, 1, 6, 3)[i % 6], "crc_type": 1})
    _check(capi, blocks, [_want(b) for b in blocks])


def test_mixed_sizes_in_one_batch(capi):
    """all 188 sizes in one submit (groups of 1 / 2 / 4 / 8 blocks per warp, ragged groups at every change of K), CRC24A
    with filler bits, CRC16 / CRC8, blocks that are not to be decoded, full-range soft bits"""
    from openair4g_b200.sim import txchain
    rng = np.random.default_rng(7)
    blocks = []
    for n, K in enumerate(txchain.k_list()):
        for rep in range(1 + (n % 3)):
            i = len(blocks)
            if i % 7 == 6:
                y = rng.integers(-32768, 32768, size=3 * K + 12).astype(np.int16)
                blocks.append({"y": y, "K": K, "max_iterations": 3, "crc_type": 1})
                continue
            crc = (1, 0, 1, 2, 1, 3)[i % 6]
            F = (0, 8, 24, 40)[i % 4] if (crc == 0 and K >= 512) else 0
            y, _ = vectors.llr_block(K, 3000 + i, "clean" if i % 2 else "waterfall", A=(16, 200)[i % 2], crc_type=crc, F=F,
                                     sigma_over_A=(0.9, 0.5)[i % 2])
            blocks.append({"y": y, "K": K, "max_iterations": 5, "crc_type": crc, "F": F})
    want = [_want(b) for b in blocks]
    for i in (5, 100, 101, 300):
        blocks[i]["decode_enable"] = 0
        want[i] = (want[i][0], capi.STATUS_NOT_DECODED)
    outs, status = capi.decode_batch(blocks, flags=capi.BATCH_SLIDING_WINDOW)
    bad = [(i, b["K"], st, w[1]) for i, (w, ob, st, b) in enumerate(zip(want, outs, status, blocks))
           if st != w[1] or (st != capi.STATUS_NOT_DECODED and not np.array_equal(ob, w[0]))]
    assert not bad, (len(bad), bad[:10])
    assert sum(1 for w in want if w[1] <= 5) > len(want) // 2


def test_agrees_with_the_bit_exact_mode_away_from_the_threshold(capi):
    """clean blocks: both modes must deliver the transmitted bits (the return values may differ by an iteration)"""
    blocks, sent = [], []
    for i, K in enumerate((40, 512, 1024, 2048, 3904, 6144) * 4):
        y, bits = vectors.llr_block(K, 40 + i, "clean", A=(12, 150)[i % 2], sigma_over_A=0.6)
        blocks.append({"y": y, "K": K, "max_iterations": 6, "crc_type": 1})
        sent.append(bits)
    o1, s1 = capi.decode_batch(blocks)
    o2, s2 = capi.decode_batch(blocks, flags=capi.BATCH_SLIDING_WINDOW)
    for a, b, sa, sb, bits in zip(o1, o2, s1, s2, sent):
        assert sa <= 6 and sb <= 6 and np.array_equal(a, b)


@pytest.mark.parametrize("case", [(75376, 90000, 6, 4, (), 0), (30576, 57600, 4, 6, (2,), 1), (7736, 14400, 4, 6, (), 1),
                                  (3000, 4800, 2, 4, (), 0), (40000, 60000, 4, 4, (5,), 1)])
def test_front_end_blocks_and_transport_blocks(capi, case):
    """rate dematching + sub-block deinterleaving in front of the sliding-window kernel (y is materialised on the device),
    transport-block reassembly and return value behind it, against the same chain around the model"""
    from oracle import chain
    tbs, G, Qm, max_it, noise, uplink = case
    tb = chain.make_tb(tbs, G, Qm, seed=11, noise_blocks=noise)
    rx = chain.rx_tb(tb, max_it, downlink=not uplink, dec=loader.port_decode16_sw)
    Cn, F = tb["seg"][0], tb["seg"][5]
    blocks, off = [], 0
    for r, K in enumerate(tb["Ks"]):
        e = tb["e"][off:off + tb["E"][r]]
        off += tb["E"][r]
        blocks.append({"y": e, "K": K, "max_iterations": max_it, "crc_type": 0 if Cn == 1 else 1, "F": F if r == 0 else 0,
                       "dematch": {"G": tb["G"], "C": Cn, "r": r, "rvidx": tb["rv"], "clear": 1, "Qm": tb["Qm"], "w": None}})
    flags = capi.BATCH_SLIDING_WINDOW | (0 if uplink else capi.BATCH_DL_STOP_AFTER_FAILURE)
    outs, status, tbo = capi.decode_batch(blocks, flags=flags, tbs=[{"first_cb": 0, "C": Cn, "uplink": uplink}])
    ret, valid, b = tbo[0]
    assert ret == rx["ret"], (ret, rx["ret"])
    for r in range(Cn):
        if rx["status"][r] is None:
            assert status[r] == capi.STATUS_NOT_DECODED
        else:
            assert status[r] == rx["status"][r] and np.array_equal(outs[r], rx["c"][r]), r
    if rx["b"] is not None:
        assert np.array_equal(b[:valid], rx["b"][:valid])
        if not noise:
            assert np.array_equal(b[:valid], tb["b"][:valid])


def test_iteration_limits_and_mixed_decoders(capi):
    """max_iterations 0 (nothing runs, return value 1) ... 8; 8-bit blocks of the same submit are not touched by the flag"""
    rng = np.random.default_rng(3)
    blocks, want = [], []
    for i, mi in enumerate((0, 1, 2, 3, 8, 0, 8, 5)):
        K = (6144, 1024, 512, 40)[i % 4]
        y, _ = vectors.llr_block(K, 80 + i, "waterfall", A=10, sigma_over_A=(0.7, 1.2)[i % 2])
        blocks.append({"y": y, "K": K, "max_iterations": mi, "crc_type": 1})
        want.append(loader.port_decode16_sw(y, K, mi, 1))
    for i in range(4):
        y, _ = vectors.llr_block(2048, 90 + i, "clean", A=20, sigma_over_A=0.6)
        blocks.append({"y": y, "K": 2048, "max_iterations": 4, "crc_type": 1, "llr8": 1})
        want.append(loader.port_decode8(y, 2048, 4, 1))
    outs, status = capi.decode_batch(blocks, flags=capi.BATCH_SLIDING_WINDOW)
    for i, ((wb, wr), ob, st, b) in enumerate(zip(want, outs, status, blocks)):
        assert st == wr, (i, st, wr)
        if b["max_iterations"] > 1:
            assert np.array_equal(ob, wb), i


def test_device_resident_plan_mode_switch(capi):
    import torch
    K, B = 3904, 96
    ys = np.stack([vectors.llr_block(K, 300 + i, "waterfall", A=9, sigma_over_A=(0.6, 1.05, 1.3)[i % 3])[0][:3 * K + 12] for i in range(B)])
    y = torch.from_numpy(ys).cuda()
    out = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda")
    st = torch.zeros(B, dtype=torch.uint8, device="cuda")
    plan = capi.DevPlan(B, K, 6, 1)
    stream = torch.cuda.current_stream().cuda_stream
    for flags, dec in ((capi.BATCH_SLIDING_WINDOW, loader.port_decode16_sw), (0, loader.port_decode16), (capi.BATCH_SLIDING_WINDOW, loader.port_decode16_sw)):
        plan.set_mode(flags)
        out.zero_()
        plan.decode(y.data_ptr(), 3 * K + 12, out.data_ptr(), K // 8, st.data_ptr(), stream)
        torch.cuda.synchronize()
        for i in range(B):
            wb, wr = dec(ys[i], K, 6, 1)
            assert int(st[i]) == wr and np.array_equal(out[i].cpu().numpy(), wb), (flags, i)
    with pytest.raises(RuntimeError):
        plan.set_mode(8)
    plan.close()
    p8 = capi.DevPlan(4, 2048, 4, 1, llr8=1)
    with pytest.raises(RuntimeError):
        p8.set_mode(capi.BATCH_SLIDING_WINDOW)
    p8.close()
