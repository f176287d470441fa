"""ulsim/dlsim-equivalent harness on the GPU decoder: BASELINE configs[0] and [1] shapes.  Every
block the harness decodes is re-decoded by the oracle chain from the same soft bits and must
agree bit for bit; the BLER curve must fall with SNR."""
import ctypes as C

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


class Recorder:
    """wraps capi.decode_batch to keep the submitted blocks and results for the oracle cross-check"""
    def __init__(self, capi):
        self.capi, self.calls = capi, []
        self.BATCH_DL_STOP_AFTER_FAILURE = capi.BATCH_DL_STOP_AFTER_FAILURE

    def decode_batch(self, blocks, flags=0):
        w_before = [b["dematch"]["w"].copy() for b in blocks]
        outs, status = self.capi.decode_batch(blocks, flags=flags)
        w_after = [b["dematch"]["w"].copy() for b in blocks]
        self.calls.append((blocks, w_before, w_after, outs, status, flags))
        return outs, status


def oracle_check(rec, llr8=0):
    P = loader.port()
    n = 0
    for blocks, w_before, w_after, outs, status, flags in rec.calls:
        failed_tb = set()
        for b, w0, w1, ob, st in zip(blocks, w_before, w_after, outs, status):
            dm, K, F = b["dematch"], b["K"], b["F"]
            D = K + 4
            RTC = (D + 31) // 32
            dw = np.zeros(3 * 32 * RTC, dtype=np.uint8)
            P.orc_generate_dummy_w(D, dw, F)
            E = C.c_uint32(0)
            w = w0.copy()
            yy = b["y"]
            if dm.get("scr_c_init") is not None:               # dlsch_unscrambling of this block's slice of the codeword
                off, nb = dm["scr_offset"], b["y"].size
                words = np.zeros((off + nb + 31) // 32, dtype=np.uint32)
                P.orc_gold_words(dm["scr_c_init"], words.ctypes.data, words.size)
                pos = off + np.arange(nb)
                cbit = (words[pos >> 5] >> (pos & 31).astype(np.uint32)) & 1
                yy = np.where(cbit == 1, b["y"].astype(np.int32), -b["y"].astype(np.int32)).astype(np.int16)
            assert P.orc_lte_rate_matching_turbo_rx(RTC, dm["G"], w, dw, yy, dm["C"], 1827072, dm["Mdlharq"], dm["Kmimo"],
                                                    dm["rvidx"], dm["clear"], dm["Qm"], dm["Nl"], dm["r"], C.byref(E)) == 0
            assert np.array_equal(w, w1), "HARQ buffer differs from the oracle"
            if flags and b["tb_id"] in failed_tb:
                assert st == 0xFE and not ob.any()
                continue
            d = np.zeros(96 + 3 * D + 16, dtype=np.int16)
            P.orc_sub_block_deinterleaving_turbo(D, d.ctypes.data + 96 * 2, w)
            dec = loader.port_decode8 if llr8 else loader.port_decode16
            wb, wr = dec(d[96:], K, b["max_iterations"], b["crc_type"], F)
            assert st == wr and np.array_equal(ob, wb), (K, dm["r"], st, wr)
            if wr > b["max_iterations"]:
                failed_tb.add(b["tb_id"])
            n += 1
    return n


def test_ulsim_25prb_mcs16_sweep(capi):
    from openair4g_b200.sim import linksim
    rec = Recorder(capi)
    sim = linksim.LinkSim(linksim.ULSIM_25PRB_MCS16, max_iterations=4, seed=7)
    assert (sim.C, sim.K, sim.F) == (2, 3904, 0)
    bler = []
    for snr in (5.5, 6.5, 7.5, 9.0):
        r = sim.run(snr, 24, max_rounds=2, capi=rec)
        bler.append(r["bler_round0"])
        assert r["mismatch_vs_tx"] == 0
        assert r["residual_bler"] <= r["bler_round0"]
    assert bler[0] > 0.5 and bler[-1] == 0.0 and all(a >= b for a, b in zip(bler, bler[1:])), bler
    assert oracle_check(rec) > 300


def test_dlsim_100prb_mcs28_subframes(capi):
    from openair4g_b200.sim import linksim
    rec = Recorder(capi)
    sim = linksim.LinkSim(linksim.DLSIM_100PRB_MCS28, max_iterations=4, seed=3)
    assert (sim.C, sim.K, sim.F) == (13, 5824, 0)
    hi = sim.run(25.0, 6, capi=rec)
    lo = sim.run(17.0, 6, capi=rec)
    assert hi["bler_round0"] == 0.0 and hi["mismatch_vs_tx"] == 0 and hi["avg_iterations"] <= 3.0
    assert lo["bler_round0"] == 1.0
    assert oracle_check(rec) >= 13 * 6


def test_ulsim_8bit_decoder_switch(capi):
    from openair4g_b200.sim import linksim
    rec = Recorder(capi)
    sim = linksim.LinkSim(linksim.UL_100PRB_MCS16, max_iterations=4, llr8=1, seed=5)
    assert (sim.C, sim.K) == (5, 6144)
    r = sim.run(8.0, 6, capi=rec)
    assert r["mismatch_vs_tx"] == 0
    assert oracle_check(rec, llr8=1) >= 5


def test_gpu_tx_switch_gives_identical_runs(capi):
    """gpu_tx=True (oai_turbo_tx_batch instead of the numpy TX chain): same seeds -> the same transmitted bits, hence the
    same soft bits, HARQ buffers, return values and decoded bytes, on a DL (scrambled, C=13) and an UL (2 rounds) shape."""
    from openair4g_b200.sim import linksim
    for cfg, snr, n, rounds in ((linksim.DLSIM_100PRB_MCS28, 21.0, 4, 1), (linksim.ULSIM_25PRB_MCS16, 6.5, 12, 2)):
        runs = []
        for gpu_tx in (False, True):
            rec = Recorder(capi)
            sim = linksim.LinkSim(cfg, max_iterations=4, seed=11, gpu_tx=gpu_tx)
            r = sim.run(snr, n, max_rounds=rounds, capi=rec)
            runs.append((r, rec))
        (ra, reca), (rb, recb) = runs
        assert ra["tb_err"] == rb["tb_err"] and ra["cb_err"] == rb["cb_err"] and np.array_equal(ra["iters"], rb["iters"])
        assert len(reca.calls) == len(recb.calls)
        for ca, cb in zip(reca.calls, recb.calls):
            assert ca[4] == cb[4]
            assert all(np.array_equal(x["y"], y["y"]) for x, y in zip(ca[0], cb[0]))
            assert all(np.array_equal(x, y) for x, y in zip(ca[2], cb[2]))
            assert all(np.array_equal(x, y) for x, y in zip(ca[3], cb[3]))


def test_ulsim_through_the_uplink_front_end(capi):
    """LinkSim(ul_front=...): every subframe reaches the GPU as the multiplexed, scrambled soft bits of the whole PUSCH
    allocation (HARQ-ACK + RI + CQI present) and comes back as an assembled transport block; every subframe is re-run
    through the oracle (port of ulsch_decoding's front + the uplink chain) from the same soft bits and must agree."""
    from openair4g_b200.sim import linksim
    from oracle import chain

    class Rec:
        def __init__(self):
            self.calls = []

        def decode_batch(self, blocks, flags=0, tbs=None):
            w0 = [b["dematch"]["w"].copy() for b in blocks]
            res = capi.decode_batch(blocks, flags=flags, tbs=tbs)
            self.calls.append((blocks, tbs, w0, [b["dematch"]["w"].copy() for b in blocks], res))
            return res
    rec = Rec()
    rec.BATCH_DL_STOP_AFTER_FAILURE = capi.BATCH_DL_STOP_AFTER_FAILURE
    sim = linksim.LinkSim(linksim.ULSIM_25PRB_MCS16, max_iterations=4, seed=9, ul_front={"O_ACK": 2, "O_RI": 1, "Or1": 20})
    z = sim.ul_sizes
    assert z["Qprime_ACK"] > 0 and z["Qprime_RI"] > 0 and z["Qprime_CQI"] > 0 and sim.G < linksim.ULSIM_25PRB_MCS16.G
    bler = []
    for snr in (5.5, 7.0, 9.5):
        r = sim.run(snr, 16, max_rounds=2, capi=rec)
        bler.append(r["bler_round0"])
        assert r["mismatch_vs_tx"] == 0
    assert bler[0] > 0.5 and bler[-1] == 0.0, bler
    P = loader.port()
    checked = 0
    for blocks, tbs, w0, w1, (outs, status, tbo) in rec.calls:
        for t, (ret, valid, b) in zip(tbs, tbo):
            uf = t["ul_front"]
            zz = loader.UlSizes()
            Qm, Cn = uf["Qm"], t["C"]
            zz.Qprime_RI, zz.Qprime_ACK, zz.Qprime_CQI, zz.Hprime = uf["Qprime_RI"], uf["Qprime_ACK"], uf["Qprime_CQI"], uf["Hprime"]
            zz.Hpp, zz.Cmux = uf["Hprime"] + uf["Qprime_RI"], uf["Cmux"]
            zz.Rmux_prime, zz.Q_CQI = zz.Hpp // zz.Cmux, Qm * uf["Qprime_CQI"]
            zz.G = (uf["Hprime"] - uf["Qprime_CQI"]) * Qm
            e = np.zeros(zz.G + 8, dtype=np.int16)
            qa, qr = np.zeros(18, dtype=np.int16), np.zeros(6, dtype=np.int16)
            qc = np.zeros(zz.Q_CQI + 8, dtype=np.int8)
            oa, orr = np.zeros(4, dtype=np.uint8), np.zeros(2, dtype=np.uint8)
            assert P.orc_ulsch_front(uf["llr"], uf["c_init"], Qm, C.byref(zz), 0, uf["O_ACK"], uf["O_RI"], 0, 1, e, qa, qr, qc, oa, orr) == 0
            o = uf["out"]
            assert np.array_equal(o["q_ACK"][:3], qa[:3]) and np.array_equal(o["q_RI"][:Qm], qr[:Qm])
            assert np.array_equal(o["q_cqi"][:zz.Q_CQI], qc[:zz.Q_CQI]) and np.array_equal(o["o_ACK"], oa[:2]) and o["o_RI"][0] == orr[0]
            bl = blocks[t["first_cb"]:t["first_cb"] + Cn]
            dm = bl[0]["dematch"]
            Gp = zz.G // Qm
            tb = {"seg": (Cn, Cn, 0, bl[0]["K"], 0, 0), "Ks": [x["K"] for x in bl], "G": zz.G, "Qm": Qm, "Nl": 1, "Mdlharq": 8, "Kmimo": 1,
                  "rv": dm["rvidx"], "e": e[:zz.G].copy(), "E": [Qm * (Gp // Cn + (1 if r >= Cn - Gp % Cn else 0)) for r in range(Cn)]}
            rx = chain.rx_tb(tb, bl[0]["max_iterations"], downlink=False, w=[x.copy() for x in w0[t["first_cb"]:t["first_cb"] + Cn]],
                             clear=dm["clear"])
            assert ret == rx["ret"] and valid == rx["b_valid"] and np.array_equal(b[:valid], rx["b"][:valid])
            for r in range(Cn):
                i = t["first_cb"] + r
                assert status[i] == rx["status"][r] and np.array_equal(outs[i], rx["c"][r])
                assert np.array_equal(w1[i], rx["w"][r])
            checked += 1
    assert checked >= 48
