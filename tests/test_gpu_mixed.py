"""GPU integration test: ONE submit with a random mixture of everything the batched call supports -- plain 16-bit and
8-bit decoder blocks of different K / iteration limits / CRC types / filler bits, front-end blocks with host-authoritative
or pool-resident HARQ buffers, with and without descrambling, front-end-only blocks -- checked block by block against the
oracle.  Catches ordering / offset / grouping mistakes of the host side that single-feature tests cannot."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader, vectors  # noqa: E402
from test_gpu_rm import _tx, _oracle_chain  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_mixed_batch(capi, seed):
    rng = np.random.default_rng(seed)
    P = loader.port()
    pool = capi.HarqPool(64, 6144)
    blocks, checks = [], []
    fe_shapes = [(3904, 0, 14400, 2, 4), (5824, 0, 90000, 13, 6), (104, 16, 600, 1, 2), (6144, 0, 57600, 5, 4), (1056, 0, 4000, 1, 2)]
    next_slot = 0
    for i in range(48):
        kind = rng.choice(["y16", "y16", "y8", "fe_host", "fe_pool", "fe_scr", "fe_only"])
        max_it = int(rng.integers(1, 7))
        if kind in ("y16", "y8"):
            K = int(rng.choice([40, 104, 512, 1056, 3904, 6144] if kind == "y16" else [256, 512, 1056, 5824, 6144]))
            crc = int(rng.integers(0, 2))
            F = int(rng.choice([0, 8, 24])) if (crc == 0 and kind == "y16" and K >= 104) else 0
            regime = str(rng.choice(["clean", "waterfall", "noise", "full"]))
            y, _ = vectors.llr_block(K, 500 + 100 * seed + i, regime, crc_type=crc, F=F)
            if kind == "y8":
                ypad = np.zeros(3 * K + 12 + 36, dtype=np.int16)
                ypad[:3 * K + 12] = y
                y = ypad
                want = loader.port_decode8(y, K, max_it, crc, F)
            else:
                want = loader.port_decode16(y, K, max_it, crc, F)
            blocks.append({"y": y, "K": K, "max_iterations": max_it, "crc_type": crc, "F": F, "llr8": 1 if kind == "y8" else 0, "tb_id": 1000 + i})
            checks.append(("dec", want, None, None, None))
            continue
        K, F, G, Cb, Qm = fe_shapes[int(rng.integers(0, len(fe_shapes)))]
        r = int(rng.integers(0, Cb))
        Fr = F if r == 0 else 0
        crc = 0 if Cb == 1 else 1
        rv = int(rng.choice([0, 2]))
        info, e, E, RTC = _tx(K, 800 + 100 * seed + i, Fr, G, Cb, Qm, r, rv, 8, float(rng.choice([0.4, 1.2])))
        w_ref = rng.integers(-50, 51, size=3 * 32 * RTC).astype(np.int16)       # a previous round's content
        clear = int(rng.integers(0, 2))
        dm = {"G": G, "C": Cb, "r": r, "rvidx": rv, "clear": clear, "Qm": Qm}
        e_in, e_oracle = e, e
        if kind == "fe_scr":
            c_init = int(rng.integers(0, 1 << 30))
            off = int(rng.integers(0, 5000))
            words = np.zeros((off + E + 31) // 32, dtype=np.uint32)
            P.orc_gold_words(c_init, words.ctypes.data, words.size)
            pos = off + np.arange(E)
            cbit = (words[pos >> 5] >> (pos & 31).astype(np.uint32)) & 1
            e_in = np.where(cbit == 1, e.astype(np.int32), -e.astype(np.int32)).astype(np.int16)   # scrambled as received
            dm.update({"scr_c_init": c_init, "scr_offset": off})
        use_pool = kind == "fe_pool" or (kind == "fe_scr" and bool(rng.integers(0, 2)))
        if use_pool:
            clear = 1                                                           # a fresh slot: nothing to combine with
            dm["clear"] = 1
            dm.update({"w": None, "harq_pool": pool, "harq_slot": next_slot})
            slot = next_slot
            next_slot += 1
            w_gpu = None
        else:
            w_gpu = w_ref.copy()
            dm["w"] = w_gpu
            slot = None
        dec_en = 0 if kind == "fe_only" else 1
        blocks.append({"y": e_in, "K": K, "max_iterations": max_it, "crc_type": crc, "F": Fr, "decode_enable": dec_en,
                       "tb_id": 2000 + i, "dematch": dm})
        w_o = w_ref.copy()
        want = _oracle_chain(K, F, G, Cb, Qm, r, rv, clear, e_oracle, w_o, max_it, crc)
        Kpi = 32 * RTC
        Ncb = min(1827072 // 8 // Cb, 3 * Kpi)
        checks.append(("fe" if dec_en else "fe_only", want, w_gpu, (w_o, Ncb), slot))
    outs, status = capi.decode_batch(blocks)
    for i, (b, (kind, want, w_gpu, wexp, slot)) in enumerate(zip(blocks, checks)):
        if kind == "fe_only":
            assert status[i] == 0xFE, (i, status[i])
        else:
            assert status[i] == want[1], (i, b["K"], b.get("llr8"), kind, status[i], want[1])
            if b["max_iterations"] > 1:
                assert np.array_equal(outs[i][:b["K"] // 8], want[0]), (i, b["K"], kind)
        if wexp is not None:
            w_o, Ncb = wexp
            got = pool.read(slot, Ncb) if slot is not None else w_gpu[:Ncb]
            assert np.array_equal(got, w_o[:Ncb]), (i, kind, "HARQ buffer")
    pool.close()


def test_recycled_handle_after_front_end_batch_then_all_invalid_batch(capi):
    """handles are pooled; a batch whose every block is rejected returns early from submit() and must not leave
    the previous batch's front-end bookkeeping behind (wait() would copy a stale HARQ buffer into the caller's w)."""
    from oracle import chain
    tb = chain.make_tb(7736, 14400, 4, seed=2, sigma_over_A=0.5)
    rx = chain.rx_tb(tb, 4, downlink=False)
    blocks, off = [], 0
    ws = [np.zeros(3 * 3936, dtype=np.int16) for _ in range(2)]
    for r in range(2):
        e = tb["e"][off:off + tb["E"][r]]
        off += tb["E"][r]
        blocks.append({"y": e, "K": 3904, "max_iterations": 4, "crc_type": 1,
                       "dematch": {"G": 14400, "C": 2, "r": r, "rvidx": 0, "clear": 1, "Qm": 4, "w": ws[r]}})
    outs, status = capi.decode_batch(blocks)                      # front-end batch with host-owned w
    assert status == rx["status"] and all(np.array_equal(ws[r], rx["w"][r]) for r in range(2))
    canary = [np.full(3 * 3936, 1234, dtype=np.int16) for _ in range(2)]
    bad = [{"y": np.zeros(3 * 3904 + 12, dtype=np.int16), "K": 3900, "max_iterations": 4, "crc_type": 1,
            "dematch": {"G": 14400, "C": 2, "r": r, "rvidx": 0, "clear": 1, "Qm": 4, "w": canary[r]}} for r in range(2)]
    outs, status = capi.decode_batch(bad)                         # every block illegal: nothing reaches the GPU
    assert status == [255, 255]
    assert all((c == 1234).all() for c in canary), "a stale HARQ buffer was written back"
    outs, status = capi.decode_batch(blocks)                      # and the handle still works
    assert status == rx["status"]
