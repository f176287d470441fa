"""GPU parity tests of the receive-side front end (NULL map, rate dematching, sub-block
deinterleaving) through the C ABI: the reference-signature calls and the fused batched path."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader, vectors  # noqa: E402
from test_golden import iter_rm  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def test_front_end_calls_against_golden(capi):
    """generate_dummy_w / lte_rate_matching_turbo_rx (3 HARQ rounds) / sub_block_deinterleaving_turbo
    against the vectors produced by the compiled reference."""
    for key, par, es, ws, ds, dummy in iter_rm():
        K, F, G, Cb, Qm, Nl, r, E, RTC = (int(v) for v in par)
        D = K + 4
        dw = np.zeros_like(dummy)
        assert capi.generate_dummy_w(D, dw, F if r == 0 else 0) == RTC
        assert np.array_equal(dw, dummy), key
        w = np.zeros(3 * 32 * RTC, dtype=np.int16)
        for rnd, rv in enumerate((0, 2, 1)):
            e = np.ascontiguousarray(es[rnd])
            rc, Eo = capi.lte_rate_matching_turbo_rx(RTC, G, w, dw, e, Cb, 1827072, 8, 1, rv, 1 if rnd == 0 else 0, Qm, Nl, r)
            assert rc == 0 and Eo == E
            assert np.array_equal(w, ws[rnd]), (key, rnd)
        d = np.zeros(96 + 3 * D + 16, dtype=np.int16)
        capi.sub_block_deinterleaving_turbo(D, d, 96, w)
        assert np.array_equal(d, ds[0]), key


@pytest.mark.parametrize("K,F", [(40, 0), (40, 8), (104, 16), (512, 0), (1056, 24), (6144, 56)])
def test_front_end_calls_random_vs_oracle(capi, K, F):
    """Full-range int16 buffers (wrap-around accumulation), stale NULL marks, odd parameters."""
    P = loader.port()
    rng = np.random.default_rng(K * 7 + F)
    D = K + 4
    RTC = (D + 31) // 32
    Kpi = 32 * RTC
    dw1 = np.zeros(3 * Kpi, dtype=np.uint8)
    dw1[rng.integers(0, 3 * Kpi, size=5)] = 2              # stale marks must survive (the call only sets)
    dw2 = dw1.copy()
    assert capi.generate_dummy_w(D, dw1, F) == P.orc_generate_dummy_w(D, dw2, F) == RTC
    assert np.array_equal(dw1, dw2)
    for (G, C_, Qm, Nl, r, rv, Mdl, Kmimo) in [(3 * K + 100, 1, 2, 1, 0, 0, 8, 1), (2 * K, 1, 4, 1, 0, 2, 8, 1),
                                                (7 * K, 2, 6, 1, 1, 3, 8, 1), (5 * K + 6, 3, 2, 2, 2, 1, 4, 2),
                                                (40 * K, 13, 6, 1, 12, 0, 8, 1), (60 * K, 2, 2, 1, 0, 1, 1, 2)]:
        w1 = rng.integers(-32768, 32768, size=3 * Kpi).astype(np.int16)
        w2 = w1.copy()
        E2 = C.c_uint32(0)
        for clear, rvx in ((1, rv), (0, (rv + 1) % 4), (0, rv)):
            e = rng.integers(-32768, 32768, size=60 * K // C_ + 64).astype(np.int16)
            rc, E1 = capi.lte_rate_matching_turbo_rx(RTC, G, w1, dw1, e, C_, 1827072, Mdl, Kmimo, rvx, clear, Qm, Nl, r)
            assert P.orc_lte_rate_matching_turbo_rx(RTC, G, w2, dw2, e, C_, 1827072, Mdl, Kmimo, rvx, clear, Qm, Nl, r, C.byref(E2)) == 0
            assert rc == 0 and E1 == E2.value
            assert np.array_equal(w1, w2), (G, C_, rvx, clear)
        d1 = np.full(96 + 3 * D + 16, 777, dtype=np.int16)
        d2 = d1.copy()
        capi.sub_block_deinterleaving_turbo(D, d1, 96, w1)
        P.orc_sub_block_deinterleaving_turbo(D, d2.ctypes.data + 96 * 2, w2)
        assert np.array_equal(d1, d2)
    assert capi.lte_rate_matching_turbo_rx(RTC, 100, w1, dw1, e, 0, 1827072, 8, 1, 0, 1, 2, 1, 0)[0] == -1
    assert capi.lte_rate_matching_turbo_rx(RTC, 100, w1, dw1, e, 1, 1827072, 8, 0, 0, 1, 2, 1, 0)[0] == -1


def _tx(K, blk, F, G, Cb, Qm, r, rv, A, sigma):
    """info -> encode -> sub-block interleave -> rate match -> BPSK LLRs e (oracle TX mirror)."""
    P = loader.port()
    info = vectors.info_block(K, blk, crc_type=0 if Cb == 1 else 1, F=F)
    bits = vectors.encode(info)
    if F:
        bits[0:3 * F:3] = 2
        bits[1:3 * F:3] = 2                                 # filler bits are NULL in streams 0 and 1
    D = K + 4
    RTC = (D + 31) // 32
    w = np.zeros(3 * 32 * RTC, dtype=np.uint8)
    P.orc_sub_block_interleaving_turbo(D, np.ascontiguousarray(bits), w)
    e = np.zeros(G // Cb + 64, dtype=np.uint8)
    E = P.orc_lte_rate_matching_turbo(RTC, G, w, e, Cb, 1827072, 8, 1, rv, Qm, 1, r)
    rng = np.random.default_rng([K, blk, rv])
    llr = A * (2 * e[:E].astype(np.int64) - 1) + np.rint(sigma * A * rng.standard_normal(E)).astype(np.int64)
    return info, llr.astype(np.int16), E, RTC


def _oracle_chain(K, F, G, Cb, Qm, r, rv, clear, e, w, max_it, crc):
    P = loader.port()
    D = K + 4
    RTC = (D + 31) // 32
    dw = np.zeros(3 * 32 * RTC, dtype=np.uint8)
    P.orc_generate_dummy_w(D, dw, F if r == 0 else 0)
    Eo = C.c_uint32(0)
    assert P.orc_lte_rate_matching_turbo_rx(RTC, G, w, dw, e, Cb, 1827072, 8, 1, rv, clear, Qm, 1, r, C.byref(Eo)) == 0
    d = np.zeros(96 + 3 * D + 16, dtype=np.int16)
    P.orc_sub_block_deinterleaving_turbo(D, d.ctypes.data + 96 * 2, w)
    return loader.port_decode16(d[96:], K, max_it, crc, F if r == 0 else 0)


def test_fused_front_end_batch_with_harq_rounds(capi):
    """dematch_enable=1: e in, decoded bytes out, HARQ buffer w combined and returned; two
    rounds (rv 0 at low SNR, then rv 2 combined) for the dlsim / ulsim shapes of BASELINE.json."""
    shapes = [(5824, 0, 90000, 13, 6, (0, 1, 12)), (3904, 0, 14400, 2, 4, (0, 1)), (6144, 0, 57600, 5, 4, (0, 4)),
              (104, 16, 600, 1, 2, (0,)), (40, 0, 132, 1, 2, (0,))]
    for sigma in (0.6, 1.4):
        blocks, want, w_gpu, w_ref, infos = [], [], [], [], []
        for (K, F, G, Cb, Qm, rs) in shapes:
            for r in rs:
                info, e, E, RTC = _tx(K, r, F if r == 0 else 0, G, Cb, Qm, r, 0, 8, sigma)
                wg = np.zeros(3 * 32 * RTC, dtype=np.int16)
                wr = wg.copy()
                crc = 0 if Cb == 1 else 1
                blocks.append({"y": e, "K": K, "max_iterations": 6, "crc_type": crc, "F": F if r == 0 else 0,
                               "dematch": {"G": G, "C": Cb, "r": r, "rvidx": 0, "clear": 1, "Qm": Qm, "w": wg}})
                want.append(_oracle_chain(K, F, G, Cb, Qm, r, 0, 1, e, wr, 6, crc))
                w_gpu.append(wg); w_ref.append(wr); infos.append((K, F, G, Cb, Qm, r, crc))
        outs, status = capi.decode_batch(blocks)
        for i, ((wb, wrr), ob, st) in enumerate(zip(want, outs, status)):
            assert st == wrr and np.array_equal(ob, wb), (sigma, infos[i], st, wrr)
            assert np.array_equal(w_gpu[i], w_ref[i]), (sigma, infos[i])
        # second round: rv 2, combined into the same w (clear = 0)
        blocks2, want2 = [], []
        for i, (K, F, G, Cb, Qm, r, crc) in enumerate(infos):
            info, e, E, RTC = _tx(K, r, F if r == 0 else 0, G, Cb, Qm, r, 2, 8, sigma)
            blocks2.append({"y": e, "K": K, "max_iterations": 6, "crc_type": crc, "F": F if r == 0 else 0,
                            "dematch": {"G": G, "C": Cb, "r": r, "rvidx": 2, "clear": 0, "Qm": Qm, "w": w_gpu[i]}})
            want2.append(_oracle_chain(K, F, G, Cb, Qm, r, 2, 0, e, w_ref[i], 6, crc))
        outs, status = capi.decode_batch(blocks2)
        for i, ((wb, wrr), ob, st) in enumerate(zip(want2, outs, status)):
            assert st == wrr and np.array_equal(ob, wb), ("round2", sigma, infos[i], st, wrr)
            assert np.array_equal(w_gpu[i], w_ref[i])
    # front end only (dlsch_decoding.c:417): w still combined, block reported as not decoded
    K, F, G, Cb, Qm, r = 1056, 0, 4000, 1, 2, 0
    info, e, E, RTC = _tx(K, 5, 0, G, Cb, Qm, r, 0, 8, 0.5)
    wg = np.zeros(3 * 32 * RTC, dtype=np.int16)
    wr = wg.copy()
    _oracle_chain(K, F, G, Cb, Qm, r, 0, 1, e, wr, 4, 0)
    outs, status = capi.decode_batch([{"y": e, "K": K, "max_iterations": 4, "crc_type": 0, "decode_enable": 0,
                                       "dematch": {"G": G, "C": Cb, "r": r, "rvidx": 0, "clear": 1, "Qm": Qm, "w": wg}}])
    assert status == [0xFE] and np.array_equal(wg, wr)


def test_device_resident_harq_pool(capi):
    """HARQ soft buffers kept in HBM (oai_turbo_harq_pool_*): three rounds (rv 0, 2, 3; the last one wraps around the
    circular buffer) through pool slots give the same decoded bytes, return values and buffer contents as the oracle
    chain with a host buffer, while no w crosses PCIe; `clear` re-initialises a slot; slots are independent."""
    shapes = [(5824, 0, 90000, 13, 6, (0, 12)), (3904, 0, 14400, 2, 4, (0, 1)), (6144, 0, 57600, 5, 4, (4,)), (104, 16, 600, 1, 2, (0,))]
    infos = [(K, F, G, Cb, Qm, r, 0 if Cb == 1 else 1) for (K, F, G, Cb, Qm, rs) in shapes for r in rs]
    pool = capi.HarqPool(len(infos) + 3, 6144)
    slots = [(5 * i + 2) % (len(infos) + 3) for i in range(len(infos))]
    assert len(set(slots)) == len(slots)
    w_ref = [None] * len(infos)
    for rnd, (rv, clear) in enumerate([(0, 1), (2, 0), (3, 0), (0, 1)]):          # the 4th round restarts the processes
        blocks, want = [], []
        for i, (K, F, G, Cb, Qm, r, crc) in enumerate(infos):
            info, e, E, RTC = _tx(K, 40 + r, F if r == 0 else 0, G, Cb, Qm, r, rv, 8, 1.5)
            if w_ref[i] is None:
                w_ref[i] = np.zeros(3 * 32 * RTC, dtype=np.int16)
            blocks.append({"y": e, "K": K, "max_iterations": 6, "crc_type": crc, "F": F if r == 0 else 0,
                           "dematch": {"G": G, "C": Cb, "r": r, "rvidx": rv, "clear": clear, "Qm": Qm, "w": None,
                                       "harq_pool": pool, "harq_slot": slots[i]}})
            want.append(_oracle_chain(K, F, G, Cb, Qm, r, rv, clear, e, w_ref[i], 6, crc))
        outs, status = capi.decode_batch(blocks)
        for i, ((wb, wrr), ob, st) in enumerate(zip(want, outs, status)):
            assert st == wrr and np.array_equal(ob, wb), (rnd, infos[i], st, wrr)
            K, Cb = infos[i][0], infos[i][3]
            Kpi = 32 * ((K + 4 + 31) // 32)
            Ncb = min(1827072 // 8 // Cb, 3 * Kpi)
            assert np.array_equal(pool.read(slots[i], Ncb), w_ref[i][:Ncb]), (rnd, infos[i])
    assert {w for _, w in want} != {7}
    pool.close()


def test_front_end_descrambling(capi):
    """scr_enable: the soft bits arrive still scrambled (downlink: dlsch_unscrambling runs right before dlsch_decoding in
    the UE); the GPU front end applies the Gold-sequence signs while dematching.  Oracle = port of dlsch_unscrambling
    (its sequence generator pinned to the compiled reference) followed by the oracle chain.  dlsim shape (13 blocks of one
    codeword, different r_offsets) + a second codeword with another c_init + full-range values incl. -32768."""
    P = loader.port()
    K, F, G, Cb, Qm = 5824, 0, 90000, 13, 6
    rng = np.random.default_rng(99)
    blocks, want, w_gpu, w_ref = [], [], [], []
    for cw, c_init in enumerate([(0x1234 << 14) + (0 << 13) + (7 << 9) + 101, (0xBEEF << 14) + (1 << 13) + (3 << 9) + 17]):
        es, offs, off = [], [], 0
        for r in range(Cb):
            info, e, E, RTC = _tx(K, 70 + 20 * cw + r, 0, G, Cb, Qm, r, 0, 8, 0.3)
            if cw == 1 and r == 5:
                e = rng.integers(-32768, 32768, size=E).astype(np.int16)      # saturating / wrapping values
                e[:4] = [-32768, 32767, -32768, 0]
            es.append(e); offs.append(off); off += E
        assert off == G
        clean = np.concatenate(es)
        n_loop = 32 * (1 + G // 32)
        scr = np.zeros(n_loop, dtype=np.int16)
        scr[:G] = clean
        P.orc_dlsch_unscrambling(c_init, scr, n_loop)        # the sign flip is its own inverse (-32768 maps to itself)
        scrambled = scr[:G].copy()
        back = np.zeros(n_loop, dtype=np.int16)
        back[:G] = scrambled
        P.orc_dlsch_unscrambling(c_init, back, n_loop)       # what the reference hands to dlsch_decoding
        for r in range(Cb):
            E = es[r].size
            e_descr = back[offs[r]:offs[r] + E].copy()
            wg = np.zeros(3 * 32 * ((K + 4 + 31) // 32), dtype=np.int16)
            wr = wg.copy()
            blocks.append({"y": scrambled[offs[r]:offs[r] + E].copy(), "K": K, "max_iterations": 6, "crc_type": 1, "F": 0, "tb_id": cw,
                           "dematch": {"G": G, "C": Cb, "r": r, "rvidx": 0, "clear": 1, "Qm": Qm, "w": wg,
                                       "scr_c_init": c_init, "scr_offset": offs[r]}})
            want.append(_oracle_chain(K, F, G, Cb, Qm, r, 0, 1, e_descr, wr, 6, 1))
            w_gpu.append(wg); w_ref.append(wr)
    outs, status = capi.decode_batch(blocks)
    for i, ((wb, wrr), ob, st) in enumerate(zip(want, outs, status)):
        assert st == wrr and np.array_equal(ob, wb), (i, st, wrr)
        assert np.array_equal(w_gpu[i], w_ref[i]), i
    assert 2 in [w for _, w in want]
