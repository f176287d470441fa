"""Pins the CPU restatement (oracle/port/*.c) to the reference itself: the reference
sources compiled in place (oracle/_ref/libref_oai.so) are run on
the same inputs and every output must be identical.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import loader, vectors

ALL_K = [40 + 8 * i for i in range(60)] + [528 + 16 * i for i in range(32)] + \
        [1056 + 32 * i for i in range(32)] + [2112 + 64 * i for i in range(64)]


def test_qpp_table_matches_reference_header(port):
    path = "/root/reference/openair1/PHY/CODING/lte_interleaver2.h"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    from oracle.gen_interleaver import parse_f1f2, k_list
    pairs = parse_f1f2(path)
    assert k_list() == ALL_K
    for i, K in enumerate(ALL_K):
        assert port.orc_qpp_index(K) == i and port.orc_qpp_K(i) == K
        assert (port.orc_qpp_f1(i), port.orc_qpp_f2(i)) == pairs[i]
    assert port.orc_qpp_index(6145) == -1 and port.orc_qpp_index(520) == -1 and port.orc_qpp_index(32) == -1


def test_crc_matches_reference(port, ref):
    rng = np.random.default_rng(1)
    msg = np.frombuffer(b"Thebigredfox", dtype=np.uint8).copy()   # crc_byte.c:219 self-test string
    cases = [(msg, 8 * msg.size)]
    for n in (1, 3, 5, 64, 765, 768):
        buf = rng.integers(0, 256, size=n + 1).astype(np.uint8)
        for bits in (8 * n, 8 * n - 3, 8 * n + 5, 8 * n - 24 if n > 3 else 8 * n):
            cases.append((buf, bits))
    for buf, bits in cases:
        for a, b in (("orc_crc24a", "ref_crc24a"), ("orc_crc24b", "ref_crc24b"),
                     ("orc_crc16", "ref_crc16"), ("orc_crc8", "ref_crc8")):
            assert getattr(port, a)(buf, bits) == getattr(ref, b)(buf, bits), (a, bits)


def test_segmentation_matches_reference(port, ref):
    def run(fn, B, pre):
        v = [C.c_uint32(0) for _ in range(6)]
        r = fn(*pre, B, *[C.byref(x) for x in v])
        return r, tuple(x.value for x in v)
    Bs = list(range(24, 6300, 7)) + list(range(6100, 100000, 211)) + [6144, 6145, 75376 + 24, 30576 + 24, 7736 + 24]
    for B in Bs:
        r1, v1 = run(ref.ref_lte_segmentation, B, (None, None))
        r2, v2 = run(port.orc_lte_segmentation, B, ())
        assert r1 == r2, B
        if r1 == 0:
            assert v1 == v2, (B, v1, v2)
    # the BASELINE shapes
    assert run(port.orc_lte_segmentation, 75376 + 24, ())[1] == (13, 13, 0, 5824, 5760, 0)
    assert run(port.orc_lte_segmentation, 30576 + 24, ())[1][0:4] == (5, 5, 0, 6144)
    assert run(port.orc_lte_segmentation, 7736 + 24, ())[1][0:4] == (2, 2, 0, 3904)


@pytest.mark.parametrize("K,F", [(40, 0), (40, 8), (104, 16), (512, 0), (1056, 24), (3904, 0), (5824, 0), (6144, 0), (6144, 56)])
def test_dummy_w_dematch_deinterleave_match_reference(port, ref, K, F):
    rng = np.random.default_rng(K + F)
    D = K + 4
    RTC = (D + 31) // 32
    Kpi = 32 * RTC
    dw1 = np.zeros(3 * Kpi + 64, dtype=np.uint8)
    dw2 = dw1.copy()
    assert ref.ref_generate_dummy_w(D, dw1, F) == port.orc_generate_dummy_w(D, dw2, F) == RTC
    assert np.array_equal(dw1, dw2)
    for (G, C_, Qm, Nl, r, rv, Mdl, Kmimo) in [(3 * K + 100, 1, 2, 1, 0, 0, 8, 1), (2 * K, 1, 4, 1, 0, 2, 8, 1),
                                                (7 * K, 2, 6, 1, 1, 3, 8, 1), (5 * K + 6, 3, 2, 2, 2, 1, 4, 2),
                                                (40 * K, 13, 6, 1, 12, 0, 8, 1)]:
        E1, E2 = C.c_uint32(0), C.c_uint32(0)
        w1 = rng.integers(-32768, 32768, size=3 * Kpi).astype(np.int16)
        w2 = w1.copy()
        for clear, rvx in ((1, rv), (0, (rv + 1) % 4), (0, rv)):      # HARQ rounds accumulate
            e = rng.integers(-32768, 32768, size=2 * G // C_ + 64).astype(np.int16)
            a = ref.ref_lte_rate_matching_turbo_rx(RTC, G, w1, dw1, e, C_, 1827072, Mdl, Kmimo, rvx, clear, Qm, Nl, r, C.byref(E1))
            b = port.orc_lte_rate_matching_turbo_rx(RTC, G, w2, dw2, e, C_, 1827072, Mdl, Kmimo, rvx, clear, Qm, Nl, r, C.byref(E2))
            assert a == b == 0 and E1.value == E2.value
            assert np.array_equal(w1, w2)
        d1 = np.full(96 + 3 * D + 16, 777, dtype=np.int16)
        d2 = d1.copy()
        ref.ref_sub_block_deinterleaving_turbo(D, d1.ctypes.data + 96 * 2, w1)
        port.orc_sub_block_deinterleaving_turbo(D, d2.ctypes.data + 96 * 2, w2)
        assert np.array_equal(d1, d2)
    E1 = C.c_uint32(0)
    assert port.orc_lte_rate_matching_turbo_rx(RTC, 100, w2, dw2, e, 0, 1827072, 8, 1, 0, 1, 2, 1, 0, C.byref(E1)) == -1
    assert ref.ref_lte_rate_matching_turbo_rx(RTC, 100, w1, dw1, e, 1, 1827072, 8, 0, 0, 1, 2, 1, 0, C.byref(E1)) == -1


def test_tx_chain_port_matches_reference(port, ref):
    """The TX mirror used for test vectors (encoder for even byte counts -- the reference's
    byte-pair interleaver, 3gpplte_sse.c:321, leaves the last byte of an odd-length block
    unwritten --, sub-block interleaver, rate matching)."""
    for K in (48, 512, 1056, 5824, 6144):
        info = vectors.info_block(K, 3)
        a = vectors.encode(info)
        out = loader.aligned(3 * K + 12 + 64, np.uint8)
        inp = loader.aligned(K // 8 + 64, np.uint8)
        inp[:K // 8] = info
        i = port.orc_qpp_index(K)
        ref.ref_threegpplte_turbo_encoder(inp, K // 8, out.ctypes.data, 0, port.orc_qpp_f1(i), port.orc_qpp_f2(i))
        assert np.array_equal(a, out[:3 * K + 12])
        D = K + 4
        RTC = (D + 31) // 32
        d_ref = np.full(96 + 3 * D + 16, 2, dtype=np.uint8)
        d_ref[96:96 + 3 * D] = a
        w1 = np.zeros(3 * 32 * RTC, dtype=np.uint8)
        w2 = np.zeros_like(w1)
        assert ref.ref_sub_block_interleaving_turbo(D, d_ref.ctypes.data + 96, w1) == RTC
        assert port.orc_sub_block_interleaving_turbo(D, np.ascontiguousarray(a), w2) == RTC
        assert np.array_equal(w1, w2)
        for G, Cc, Qm, r, rv in ((3 * K, 1, 2, 0, 0), (2 * K + 8, 1, 4, 0, 1), (12 * K, 5, 6, 4, 2), (4 * K, 1, 2, 0, 3)):
            e1 = np.zeros(4 * K + 64, dtype=np.uint8)
            e2 = np.zeros_like(e1)
            E1 = ref.ref_lte_rate_matching_turbo(RTC, G, w1, e1, Cc, 1827072, 8, 1, rv, Qm, 1, r, 25, 0)
            E2 = port.orc_lte_rate_matching_turbo(RTC, G, w2, e2, Cc, 1827072, 8, 1, rv, Qm, 1, r)
            assert E1 == E2 and np.array_equal(e1, e2)
        if K >= 5824:                                 # limited soft buffer (Nir/C < 3*Kpi): the TX side gives up, RM:508-511
            for Cc in (13, 14):
                E1 = ref.ref_lte_rate_matching_turbo(RTC, 20 * K, w1, e1, Cc, 1827072, 8, 1, 0, 2, 1, 0, 25, 0)
                E2 = port.orc_lte_rate_matching_turbo(RTC, 20 * K, w2, e2, Cc, 1827072, 8, 1, 0, 2, 1, 0)
                assert E1 == E2 and (E1 == 0) == (1827072 // 8 // Cc < 3 * 32 * RTC)


def _cmp16(K, blk, regime, crc, max_it, A=8, F=0):
    y, info = vectors.llr_block(K, blk, regime, A=A, crc_type=crc if crc < 2 else 1, F=F)
    b1, r1 = loader.ref_decode16(y, K, max_it, crc, F)
    b2, r2 = loader.port_decode16(y, K, max_it, crc, F)
    assert r1 == r2, (K, regime, crc, max_it, A, r1, r2)
    if max_it > 1:
        assert np.array_equal(b1, b2), (K, regime, crc, max_it, A)
    return r1, b1, info


def test_td16_all_block_sizes(ref):
    """All 188 K x {clean, waterfall, noise, full-range} x CRC24A/B, 6 iterations."""
    hist = {}
    for i, K in enumerate(ALL_K):
        for regime in ("clean", "waterfall", "noise", "full"):
            r, b, info = _cmp16(K, i, regime, i & 1, 6)
            hist[r] = hist.get(r, 0) + 1
            if regime == "clean" and K >= 512:
                assert r == 2 and np.array_equal(b, info)
            if regime == "noise" and K >= 512:
                assert r == 7
    assert set(hist) >= {2, 3, 4, 7}


@pytest.mark.parametrize("A", [4, 32, 128, 1024, 4000, 8192])
def test_td16_amplitudes_and_saturation(ref, A):
    for K in (40, 200, 512, 1056, 2048, 6144):
        for regime in ("clean", "waterfall"):
            _cmp16(K, A, regime, 1, 6, A=A)


def test_td16_iteration_counts_crc_types_filler(ref):
    for K in (40, 504, 1024, 3904):
        for max_it in (1, 2, 3, 4, 8):
            for regime in ("clean", "waterfall", "noise"):
                _cmp16(K, max_it, regime, 1, max_it)
        for crc in (0, 1, 2, 3):
            _cmp16(K, 9, "clean", crc, 4)
        for F in (8, 16, 40):
            if F < K - 24:
                r, b, info = _cmp16(K, F, "clean", 0, 6, F=F)
                if K >= 504:
                    assert r == 2 and np.array_equal(b, info)
    y = np.zeros(3 * 40 + 12, dtype=np.int16)
    assert loader.port_decode16(y, 40, 4, 4)[1] == 255 == loader.ref_decode16(y, 40, 4, 4)[1]
    y = np.zeros(3 * 520 + 12, dtype=np.int16)
    assert loader.port_decode16(y, 520, 4, 1)[1] == 255 == loader.ref_decode16(y, 520, 4, 1)[1]


def test_kat0_reference_selftest_scenario(port):
    """KAT #0: the reference's own embedded self-test scenario
    (3gpplte_turbo_decoder_sse.c:2600-2652): bytes 07 a5 11 92 fe + CRC24A, K=64,
    LLR = +-15, 6 iterations -> decoded == input."""
    info = np.zeros(8, dtype=np.uint8)
    info[:5] = [0x07, 0xA5, 0x11, 0x92, 0xFE]
    c = port.orc_crc24a(info, 40) >> 8
    info[5:8] = [(c >> 16) & 255, (c >> 8) & 255, c & 255]
    bits = vectors.encode(info).astype(np.int16)
    y = (15 * (2 * bits - 1)).astype(np.int16)
    b, r = loader.port_decode16(y, 64, 6, 0)
    assert r == 2 and np.array_equal(b, info)
    if loader.ref() is not None:
        b2, r2 = loader.ref_decode16(y, 64, 6, 0)
        assert r2 == r and np.array_equal(b2, b)


def test_log_map16_internals_match_reference(port, ref):
    """alpha/beta/ext dumps of one MAP pass agree element-wise with the reference's
    log_map16 (exported by the in-place build)."""
    rng = np.random.default_rng(5)
    for n in (40, 48, 512, 6144):
        for amp in (20, 3000, 32767):
            sys_ = loader.aligned(n + 64, np.int16)
            par = loader.aligned(n + 64, np.int16)
            sys_[:n + 16] = rng.integers(-amp, amp + 1, size=n + 16)
            par[:n + 16] = rng.integers(-amp, amp + 1, size=n + 16)
            ab = 8 * (n + 16)
            a1, b1 = loader.aligned(ab, np.int16), loader.aligned(ab, np.int16)
            m11, m10 = loader.aligned(n + 64, np.int16), loader.aligned(n + 64, np.int16)
            e1 = loader.aligned(n + 128, np.int16)
            for term in (0, 1):
                ref.log_map16(sys_.ctypes.data, par.ctypes.data, m11.ctypes.data, m10.ctypes.data,
                              a1.ctypes.data, b1.ctypes.data, e1.ctypes.data, n, term, 0, 0, None, None, None, None)
                a2, b2 = np.zeros(ab, np.int16), np.zeros(ab, np.int16)
                e2 = np.zeros(n + 16, np.int16)
                port.orc_log_map16(np.ascontiguousarray(sys_[:n + 16]), np.ascontiguousarray(par[:n + 16]), e2, n, term,
                                   a2.ctypes.data, b2.ctypes.data)
                W = n // 8
                assert np.array_equal(e1[:n], e2[:n])
                assert np.array_equal(a1[:64 * (W + 1)], a2[:64 * (W + 1)])
                assert np.array_equal(b1[:64 * (W + 1)], b2[:64 * (W + 1)])


def test_log_map16_corner_inputs_match_reference(port, ref):
    """The crafted corner inputs of tests/test_gpu_map_pass.py (branch metrics of exactly -16384, metrics saturated at
    -32768, positive tail metrics) through the port's and the compiled reference's log_map16: LLRs, alpha and beta."""
    from test_gpu_map_pass import craft
    for kind in ("sparse_hazard", "dense_hazard", "edges", "saturated", "positive_tail"):
        rng = np.random.default_rng(sum(map(ord, kind)))
        for n in (40, 48, 512, 6144):
            y = craft(n, kind, rng)
            W = n // 8
            pos = np.arange(n)
            st = (pos % W) * 8 + pos // W
            for term in (0, 1):
                sys_ = loader.aligned(n + 64, np.int16)
                par = loader.aligned(n + 64, np.int16)
                sys_[:] = 0
                par[:] = 0
                sys_[st] = y[0:3 * n:3]
                par[st] = y[(2 if term else 1):3 * n:3]
                t = y[3 * n:]
                for i in range(3):
                    if term == 0:
                        sys_[n + i] = t[2 * i]; par[n + i] = t[2 * i + 1]
                    else:
                        sys_[n + 8 + i] = t[6 + 2 * i]; par[n + i] = t[7 + 2 * i]
                ab = 8 * (n + 16)
                a1, b1 = loader.aligned(ab, np.int16), loader.aligned(ab, np.int16)
                m11, m10 = loader.aligned(n + 64, np.int16), loader.aligned(n + 64, np.int16)
                e1 = loader.aligned(n + 128, np.int16)
                ref.log_map16(sys_.ctypes.data, par.ctypes.data, m11.ctypes.data, m10.ctypes.data,
                              a1.ctypes.data, b1.ctypes.data, e1.ctypes.data, n, term, 0, 0, None, None, None, None)
                a2, b2 = np.zeros(ab, np.int16), np.zeros(ab, np.int16)
                e2 = np.zeros(n + 16, np.int16)
                port.orc_log_map16(np.ascontiguousarray(sys_[:n + 16]), np.ascontiguousarray(par[:n + 16]), e2, n, term,
                                   a2.ctypes.data, b2.ctypes.data)
                assert np.array_equal(e1[:n], e2[:n]), (kind, n, term)
                assert np.array_equal(a1[:64 * (W + 1)], a2[:64 * (W + 1)]), (kind, n, term)
                assert np.array_equal(b1[:64 * (W + 1)], b2[:64 * (W + 1)]), (kind, n, term)


def test_td8_all_sizes_of_its_domain(ref):
    """8-bit decoder: every K >= 256 with K % 16 == 0 (145 sizes), all input-scaling brackets
    (amplitudes 8..2000 and full-range int16), both hard-decision rules (K % 128 == 0 or not)."""
    hist = {}
    n = 0
    for i, K in enumerate(k for k in ALL_K if k >= 256 and k % 16 == 0):
        for regime, A in (("clean", 8), ("waterfall", 8), ("noise", 8), ("full", 8), ("waterfall", 40),
                          ("waterfall", 100), ("clean", 300), ("waterfall", 2000)):
            y, info = vectors.llr_block(K, 3000 + i, regime, A=A, crc_type=i & 1)
            b1, r1 = loader.ref_decode16(y, K, 6, i & 1, which=8)
            b2, r2 = loader.port_decode8(y, K, 6, i & 1)
            assert r1 == r2 and np.array_equal(b1, b2), (K, regime, A)
            hist[r1] = hist.get(r1, 0) + 1
            n += 1
    assert n == 145 * 8 and set(hist) >= {2, 3, 4, 7}
    for max_it, crc in ((1, 1), (2, 0), (3, 2), (4, 3)):
        y, _ = vectors.llr_block(512, 9, "clean", crc_type=min(crc, 1))
        assert loader.ref_decode16(y, 512, max_it, crc, which=8)[1] == loader.port_decode8(y, 512, max_it, crc)[1]
    y = np.zeros(3 * 40 + 12 + 64, dtype=np.int16)
    assert loader.port_decode8(y, 40, 4, 1)[1] == 254            # outside the parity domain (reference overruns)


def test_gold_sequence_matches_reference(port, ref):
    """lte_gold_generic (LTE_REFSIG/lte_gold.c:151-180): port vs the compiled reference, word by word, for DL c_init
    values (rnti<<14 + q<<13 + (Ns>>1)<<9 + Nid_cell) and arbitrary 31-bit ones; plus the 36.211 7.2 definition."""
    rng = np.random.default_rng(7)
    inits = [0, 1, (0x1234 << 14) + (0 << 13) + (7 << 9) + 101, (0xFFFF << 14) + (1 << 13) + (9 << 9) + 503] + \
        [int(v) for v in rng.integers(0, 1 << 31, size=12)]
    for c_init in inits:
        a1, a2, b1, b2 = C.c_uint32(0), C.c_uint32(c_init), C.c_uint32(0), C.c_uint32(c_init)
        for i in range(300):
            assert port.orc_lte_gold_generic(C.byref(a1), C.byref(a2), 1 if i == 0 else 0) == \
                ref.ref_lte_gold_generic(C.byref(b1), C.byref(b2), 1 if i == 0 else 0), (c_init, i)
    # the sequence is the standard's c(n) = x1(n+1600) xor x2(n+1600)
    c_init = inits[2]
    x1 = [1] + [0] * 30
    x2 = [(c_init >> i) & 1 for i in range(31)]
    for n in range(1600 + 96):
        x1.append(x1[n + 3] ^ x1[n])
        x2.append(x2[n + 3] ^ x2[n + 2] ^ x2[n + 1] ^ x2[n])
    words = np.zeros(3, dtype=np.uint32)
    port.orc_gold_words(c_init, words.ctypes.data, 3)
    for n in range(96):
        assert ((int(words[n >> 5]) >> (n & 31)) & 1) == (x1[n + 1600] ^ x2[n + 1600]), n
    # descrambling: sign 2c-1, int16 wrap of -32768
    llr = np.array([100, -100, -32768, 32767] * 16, dtype=np.int16)
    want = llr.copy()
    port.orc_dlsch_unscrambling(c_init, llr, llr.size)
    for k in range(llr.size):
        c = (int(words[k >> 5]) >> (k & 31)) & 1
        v = int(want[k]) if c else -int(want[k])
        assert int(llr[k]) == ((v + 32768) % 65536) - 32768, k


def test_ul_front_port_matches_compiled_ulsch_decoding(ref):
    """oracle/port/ulfront_port.c against the reference's own ulsch_decoding() (compiled in place behind shim4) on random
    soft bits: every Qm, both cyclic prefixes, 0/1/2 ACK bits with and without bundling, RI, CQI, int16 extremes."""
    import ctypes as C
    from oracle import ulgen
    P = loader.port()
    rng = np.random.default_rng(77)
    cases = [(25, 16, 7736, 12, 0, 0, 0, 0, 1, 0), (25, 16, 7736, 12, 2, 1, 20, 0, 1, 0), (100, 16, 30576, 12, 1, 1, 40, 0, 1, 0),
             (100, 26, 61664, 12, 2, 1, 40, 0, 1, 0), (10, 5, 872, 12, 1, 1, 11, 0, 1, 0), (10, 5, 872, 12, 2, 0, 0, 1, 2, 0),
             (10, 5, 872, 12, 1, 0, 0, 1, 3, 0), (25, 16, 6200, 10, 2, 1, 20, 0, 1, 1), (25, 16, 6200, 10, 1, 1, 20, 1, 4, 1),
             (1, 5, 72, 12, 2, 1, 0, 0, 1, 0), (4, 26, 1000, 12, 2, 1, 0, 0, 1, 0), (8, 16, 2216, 12, 1, 1, 0, 0, 1, 0),
             (50, 24, 30576, 12, 2, 1, 30, 1, 2, 0), (6, 10, 504, 12, 1, 1, 5, 0, 1, 0)]
    for case in cases:
        par = ulgen.params(*case, max_it=2)
        z, Qm = par["z"], par["Qm"]
        assert (z.Hpp * Qm) % 32 == 0
        llr = rng.integers(-3000, 3001, size=z.Hpp * Qm).astype(np.int16)
        llr[rng.integers(0, llr.size, 12)] = -32768
        llr[rng.integers(0, llr.size, 12)] = 32767
        r = loader.ref_ulsch_decoding(par["ref"], llr)
        e = np.zeros(14 * 1200 * 6, dtype=np.int16)
        qa, qr = np.zeros(18, dtype=np.int16), np.zeros(6, dtype=np.int16)
        qc = np.zeros(2560, dtype=np.int8)
        oa, orr = np.zeros(4, dtype=np.uint8), np.zeros(2, dtype=np.uint8)
        assert P.orc_ulsch_front(llr, par["c_init"], Qm, C.byref(z), par["Ncp"], par["O_ACK"], par["O_RI"], par["bundling"],
                                 par["Nbundled"], e, qa, qr, qc, oa, orr) == 0
        ne = (z.Hprime - z.Qprime_CQI) * Qm
        assert np.array_equal(e[:ne], r["e"][:ne]), case
        assert np.array_equal(qc[:z.Q_CQI], r["q_cqi"][:z.Q_CQI]), case
        if par["O_ACK"]:
            k = 3 if par["O_ACK"] == 2 else 1
            assert np.array_equal(qa[:k], r["q_ACK"][:k]) and np.array_equal(oa[:par["O_ACK"]], r["o_ACK"][:par["O_ACK"]]), case
        if par["O_RI"]:
            assert np.array_equal(qr[:Qm], r["q_RI"][:Qm]) and orr[0] == r["o_RI"][0], case
