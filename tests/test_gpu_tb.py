"""Transport-block reassembly and return value on the GPU (oai_turbo_submit_tbs) against the oracle
chain (oracle/chain.py: dlsch_decoding.c:417,448-512 and ulsch_decoding.c:1380-1409 restated on top of the pinned port)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import chain  # noqa: E402

STOP = 1          # OAI_BATCH_DL_STOP_AFTER_FAILURE


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def _blocks(tb, max_it, llr8=0, tb_id=0):
    Cn, F = tb["seg"][0], tb["seg"][5]
    out, off = [], 0
    for r, K in enumerate(tb["Ks"]):
        e = tb["e"][off:off + tb["E"][r]]
        off += tb["E"][r]
        out.append({"y": e, "K": K, "max_iterations": max_it, "crc_type": 0 if Cn == 1 else 1, "F": F if r == 0 else 0,
                    "llr8": llr8, "tb_id": tb_id,
                    "dematch": {"G": tb["G"], "C": Cn, "r": r, "rvidx": tb["rv"], "clear": 1, "Qm": tb["Qm"], "w": None}})
    return out


CASES = [  # tbs, G, Qm, max_it, noise blocks
    (75376, 90000, 6, 4, ()),            # dlsim: C = 13 x K = 5824
    (75376, 90000, 6, 4, (4,)),          # block 4 fails
    (75376, 90000, 6, 4, (0,)),          # block 0 fails
    (30576, 57600, 4, 6, ()),            # UL 100 PRB MCS16: C = 5 x K = 6144
    (30576, 57600, 4, 6, (2,)),
    (30576, 57600, 4, 6, (0, 4)),
    (7736, 14400, 4, 6, ()),             # ulsim 25 PRB MCS16: C = 2 x K = 3904
    (7736, 14400, 4, 6, (1,)),
    (3000, 4800, 2, 4, ()),              # C = 1, K = 3072 with F = 48 filler bits, CRC24A
    (3000, 4800, 2, 4, (0,)),
    (6192, 9000, 2, 4, ()),              # C = 2 x K = 3136 with F = 8
    (40000, 60000, 4, 4, (5,)),          # mixed sizes: C- = 2 x 5696, C+ = 5 x 5760
]


@pytest.mark.parametrize("uplink", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_tb_outputs_match_reference_chain(capi, case, uplink):
    tbs_bits, G, Qm, max_it, noise = case
    tb = chain.make_tb(tbs_bits, G, Qm, seed=5 + uplink, noise_blocks=noise, sigma_over_A=0.25)
    rx = chain.rx_tb(tb, max_it, downlink=not uplink)
    Cn = len(tb["Ks"])
    for cb_out in (True, False):
        outs, status, tbo = capi.decode_batch(_blocks(tb, max_it), flags=0 if uplink else STOP,
                                              tbs=[{"first_cb": 0, "C": Cn, "uplink": uplink}], cb_out=cb_out)
        ret, valid, b = tbo[0]
        assert ret == rx["ret"], (ret, rx["ret"])
        if uplink:
            assert valid == rx["b_valid"]
            assert np.array_equal(b[:valid], rx["b"][:valid])
            assert (b[valid:] == 0xA5).all()                      # nothing written behind the reference's final offset
            if not noise:
                assert np.array_equal(b[:valid], tb["b"])
        else:
            if rx["b"] is None:
                assert valid == 0 and (b == 0xA5).all()              # NACK: b untouched
            else:
                assert valid == rx["b"].size and np.array_equal(b[:valid], rx["b"]) and np.array_equal(rx["b"], tb["b"])
        for r in range(Cn):
            want = rx["status"][r]
            assert status[r] == (0xFE if want is None else want), (r, status[r], want)
            if cb_out:
                assert np.array_equal(outs[r], rx["c"][r]), r


def test_many_transport_blocks_in_one_submit(capi):
    """UL and DL transport blocks of different shapes in one batch, 8-bit decoder for one of them; illegal block inside a TB"""
    specs = [(7736, 14400, 4, 4, (), 1, 0), (75376, 90000, 6, 4, (7,), 0, 0), (30576, 57600, 4, 4, (), 1, 1),
             (3000, 4800, 2, 4, (), 0, 0), (7736, 14400, 4, 4, (0,), 1, 0)]
    blocks, tds, rxs, tbl = [], [], [], []
    for i, (tbs_bits, G, Qm, max_it, noise, uplink, llr8) in enumerate(specs):
        tb = chain.make_tb(tbs_bits, G, Qm, seed=40 + i, noise_blocks=noise, sigma_over_A=0.25)
        rxs.append(chain.rx_tb(tb, max_it, downlink=not uplink, llr8=llr8))
        tds.append({"first_cb": len(blocks), "C": len(tb["Ks"]), "uplink": uplink})
        blocks += _blocks(tb, max_it, llr8=llr8, tb_id=i)
        tbl.append(tb)
    outs, status, tbo = capi.decode_batch(blocks, flags=STOP, tbs=tds, cb_out=False)
    for i, ((ret, valid, b), rx) in enumerate(zip(tbo, rxs)):
        assert ret == rx["ret"], i
        if specs[i][5]:
            assert valid == rx["b_valid"] and np.array_equal(b[:valid], rx["b"][:valid]), i
        elif rx["b"] is None:
            assert valid == 0 and (b == 0xA5).all(), i
        else:
            assert np.array_equal(b[:valid], rx["b"]), i
    # a transport block whose first block has an illegal size: it never reaches the GPU and counts as failed
    bad = _blocks(tbl[0], 4)
    bad[0]["K"] = 3900
    _, status, tbo = capi.decode_batch(bad, tbs=[{"first_cb": 0, "C": 2, "uplink": 1}])
    assert status[0] == 255 and tbo[0][0] == 5
    assert tbo[0][1] == rxs[0]["c"][1].size - 3 and np.array_equal(tbo[0][2][:tbo[0][1]], rxs[0]["c"][1][:-3])
    with pytest.raises(RuntimeError):
        capi.decode_batch(_blocks(tbl[0], 4), tbs=[{"first_cb": 1, "C": 2, "uplink": 1}])      # runs past the array
