"""GPU parity tests of the transmit-side mirror through the C ABI: turbo encoder, sub-block
interleaver, rate matching -- the reference-signature calls and the fused batched call -- against the oracle
(compiled reference where it travels, C port otherwise), and a TX -> RX round trip at full batch size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader, vectors  # noqa: E402
from openair4g_b200.sim import txchain  # noqa: E402

NSOFT = 1827072


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def _info(K, seed):
    return np.random.default_rng([K, seed]).integers(0, 256, size=K // 8).astype(np.uint8)


def test_encoder_all_block_sizes(capi, port):
    """threegpplte_turbo_encoder on all 188 K against the port (standard code, every length) and, for even byte counts
    (3gpplte_sse.c:321 leaves the last interleaved byte of odd lengths unwritten), against the compiled reference."""
    ref = loader.ref()
    n_ref = 0
    for K in txchain.k_list():
        info = _info(K, 1)
        got = capi.threegpplte_turbo_encoder(info)
        assert np.array_equal(got, vectors.encode(info)), K
        if ref is not None and (K // 8) % 2 == 0:
            out = loader.aligned(3 * K + 12 + 64, np.uint8)
            inp = loader.aligned(K // 8 + 64, np.uint8)
            inp[:K // 8] = info
            i = port.orc_qpp_index(K)
            ref.ref_threegpplte_turbo_encoder(inp, K // 8, out.ctypes.data, 0, port.orc_qpp_f1(i), port.orc_qpp_f2(i))
            assert np.array_equal(got, out[:3 * K + 12]), K
            n_ref += 1
    assert ref is None or n_ref > 100
    # all-zero and all-one inputs (termination from a non-trivial state), illegal length leaves the output alone
    for K in (40, 6144):
        for v in (0, 255):
            info = np.full(K // 8, v, dtype=np.uint8)
            assert np.array_equal(capi.threegpplte_turbo_encoder(info), vectors.encode(info))
    assert np.all(capi.threegpplte_turbo_encoder(np.zeros(4, dtype=np.uint8)) == 255)      # 32 bits: "Illegal frame length!"
    assert np.all(capi.threegpplte_turbo_encoder(np.zeros(65, dtype=np.uint8)) == 255)     # 520 bits is not a block size


@pytest.mark.parametrize("K", [40, 48, 104, 512, 1056, 3904, 5824, 6144])
def test_interleaver_and_rate_matching_calls(capi, port, K):
    """sub_block_interleaving_turbo (arbitrary bytes in front of d are read like the reference reads them, d[3D+2] is
    written) and lte_rate_matching_turbo (rv 0..3, wrap-around, repetition, several C / Qm / Nl / Mdlharq / Kmimo)."""
    ref = loader.ref()
    rng = np.random.default_rng(K)
    D = K + 4
    RTC = (D + 31) // 32
    Kpi = 32 * RTC
    d1 = rng.integers(0, 2, size=96 + 3 * D + 16).astype(np.uint8)
    d1[:96] = rng.integers(0, 4, size=96)                      # includes LTE_NULL (2) and a foreign value
    d2 = d1.copy()
    w1 = np.zeros(3 * Kpi, dtype=np.uint8)
    w2 = np.zeros_like(w1)
    assert capi.sub_block_interleaving_turbo(D, d1, 96, w1) == RTC
    if ref is not None:
        assert ref.ref_sub_block_interleaving_turbo(D, d2.ctypes.data + 96, w2) == RTC
        assert np.array_equal(w1, w2) and np.array_equal(d1, d2)
    # the callers' layout: LTE_NULL in front of d
    d1[:96] = 2
    a = np.ascontiguousarray(d1[96:96 + 3 * D])
    assert capi.sub_block_interleaving_turbo(D, d1, 96, w1) == RTC
    assert port.orc_sub_block_interleaving_turbo(D, a, w2) == RTC
    assert np.array_equal(w1, w2)
    for (G, C_, Qm, Nl, r, rv, Mdl, Kmimo) in [(3 * K + 100, 1, 2, 1, 0, 0, 8, 1), (2 * K, 1, 4, 1, 0, 2, 8, 1),
                                                (7 * K, 2, 6, 1, 1, 3, 8, 1), (5 * K + 6, 3, 2, 2, 2, 1, 4, 2),
                                                (40 * K, 13, 6, 1, 12, 0, 8, 1), (12 * K, 1, 2, 1, 0, 3, 8, 1)]:
        e1 = np.full(13 * K + 64, 9, dtype=np.uint8)
        e2 = e1.copy()
        E1 = capi.lte_rate_matching_turbo(RTC, G, w1, e1, C_, NSOFT, Mdl, Kmimo, rv, Qm, Nl, r, 25, 0)
        if ref is not None:
            E2 = ref.ref_lte_rate_matching_turbo(RTC, G, w2, e2, C_, NSOFT, Mdl, Kmimo, rv, Qm, Nl, r, 25, 0)
        else:
            E2 = port.orc_lte_rate_matching_turbo(RTC, G, w2, e2, C_, NSOFT, Mdl, Kmimo, rv, Qm, Nl, r)
        assert E1 == E2 and np.array_equal(e1, e2), (G, C_, Qm, Nl, r, rv)
    # soft buffer smaller than the circular buffer: the reference gives up and returns 0
    if K >= 512:
        e1 = np.full(64, 9, dtype=np.uint8)
        assert capi.lte_rate_matching_turbo(RTC, 3 * K, w1, e1, 64, 3 * Kpi * 8 * 32, 8, 1, 0, 2, 1, 0, 25, 0) == 0
        assert np.all(e1 == 9)


def _oracle_tx(port, info, K, F_null, G, Cb, Qm, Nl, r, rv, Mdl=8, Kmimo=1):
    bits = vectors.encode(info)
    if F_null:
        bits[0:3 * F_null:3] = 2
        bits[1:3 * F_null:3] = 2
    D = K + 4
    RTC = (D + 31) // 32
    w = np.zeros(3 * 32 * RTC, dtype=np.uint8)
    port.orc_sub_block_interleaving_turbo(D, np.ascontiguousarray(bits), w)
    e = np.full(G + 64, 9, dtype=np.uint8)
    E = port.orc_lte_rate_matching_turbo(RTC, G, w, e, Cb, NSOFT, Mdl, Kmimo, rv, Qm, Nl, r)
    return e[:E].copy()


def test_tx_batch_against_oracle_chain(capi, port):
    """oai_turbo_tx_batch on a mixed batch (block sizes from both ends of the table, filler bits in both conventions,
    all redundancy versions, segmented transport blocks) = encoder -> interleaver -> rate matching of the oracle."""
    rng = np.random.default_rng(5)
    blocks, want = [], []
    ks = txchain.k_list()
    for i in range(160):
        K = int(ks[rng.integers(0, len(ks))]) if i >= 8 else (40, 48, 56, 6144, 6080, 512, 1024, 2048)[i]
        Cb = int(rng.choice([1, 1, 2, 5, 13]))
        r = int(rng.integers(0, Cb))
        Qm = int(rng.choice([2, 4, 6]))
        Nl = int(rng.choice([1, 1, 2]))
        rv = int(rng.integers(0, 4))
        G = int(rng.integers(K // 2 + 20, 5 * K)) // (Qm * Nl) * (Qm * Nl) * Cb
        F = int(rng.choice([0, 0, 8, 16, 40, 56])) if (r == 0 and K >= 104) else 0
        fn = int(rng.integers(0, 2))
        info = _info(K, i)
        if F:
            info[:F // 8] = 0
        blocks.append({"c": info, "K": K, "F": F, "filler_null": fn, "G": G, "C": Cb, "r": r, "rvidx": rv, "Qm": Qm, "Nl": Nl})
        want.append(_oracle_tx(port, info, K, F if fn else 0, G, Cb, Qm, Nl, r, rv))
    got = capi.tx_batch(blocks)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.size == w.size and np.array_equal(g, w), (i, {k: v for k, v in blocks[i].items() if k != "c"})
    # the numpy TX chain of the harness (36.212 filler convention) gives the same bits
    K, F, G, Cb, Qm = 1056, 24, 9000, 3, 4
    info = _info(K, 99)
    info[:3] = 0
    cbits = np.unpackbits(info)[None, :]
    e_np, E = txchain.rate_match(txchain.turbo_encode(cbits), K, F, G, Cb, Qm, 1, 0, 2)
    (e_gpu,) = capi.tx_batch([{"c": info, "K": K, "F": F, "filler_null": 1, "G": G, "C": Cb, "r": 0, "rvidx": 2, "Qm": Qm}])
    assert E == e_gpu.size and np.array_equal(e_np[0], e_gpu)


def test_tx_rx_round_trip_full_size(capi):
    """Size-independent property at a full-size batch: what the GPU TX chain emits, the GPU RX chain (dematching +
    deinterleaving + 16-bit decoder) turns back into the transmitted bytes, for every block; rv 0 at rate ~0.53 and a
    second transmission (rv 2) combined into the same soft buffers."""
    K, G, Qm, n = 6144, 11520, 2, 2048
    infos = [vectors.info_block(K, 1000 + i, crc_type=0) for i in range(64)]
    tx = [{"c": infos[i % 64], "K": K, "G": G, "C": 1, "r": 0, "rvidx": 0, "Qm": Qm} for i in range(n)]
    es = capi.tx_batch(tx)
    assert all(e.size == G for e in es)
    assert all(np.array_equal(es[i], es[i % 64]) for i in range(64, n, 97))
    ws = [np.zeros(3 * 6176, dtype=np.int16) for _ in range(n)]
    blocks = [{"y": (16 * (2 * es[i].astype(np.int16) - 1)).astype(np.int16), "K": K, "max_iterations": 6, "crc_type": 0,
               "dematch": {"G": G, "C": 1, "r": 0, "rvidx": 0, "clear": 1, "Qm": Qm, "w": ws[i]}} for i in range(n)]
    outs, status = capi.decode_batch(blocks)
    assert all(s <= 2 for s in status)
    assert all(np.array_equal(outs[i][:K // 8], infos[i % 64]) for i in range(n))
    # retransmission with rv 2: new bits, combined
    tx2 = [dict(t, rvidx=2) for t in tx[:256]]
    es2 = capi.tx_batch(tx2)
    assert not np.array_equal(es2[0], es[0])
    blocks2 = [{"y": (16 * (2 * es2[i].astype(np.int16) - 1)).astype(np.int16), "K": K, "max_iterations": 6, "crc_type": 0,
                "dematch": {"G": G, "C": 1, "r": 0, "rvidx": 2, "clear": 0, "Qm": Qm, "w": ws[i]}} for i in range(256)]
    outs2, status2 = capi.decode_batch(blocks2)
    assert all(s <= 2 for s in status2)
    assert all(np.array_equal(outs2[i][:K // 8], infos[i % 64]) for i in range(256))
    for i in (0, 100, 255):                                   # both transmissions landed: |w| reaches 32 where they overlap
        assert np.abs(ws[i]).max() == 32 and np.count_nonzero(ws[i]) > G
