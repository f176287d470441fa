"""Parity off GPU 0 and across devices: the batched call with gpu=k != current device, a HARQ pool
on that GPU, and the property that no entry point moves the caller's current device.  Plus the narrow (int8) soft-bit
feed of the front end, which must give the same bytes / status / HARQ contents as the int16 feed and the oracle chain.

The multi-GPU cases skip on a one-GPU box (run them with `gpurun --gpus 2`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import chain  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available()
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def _tb_blocks(tb, max_it, pool=None, slot0=0, in_fmt=0, w=None):
    """descriptors (capi.decode_batch dicts) of one uplink transport block from its soft bits e"""
    Cn, F = tb["seg"][0], tb["seg"][5]
    blocks, off = [], 0
    for r, K in enumerate(tb["Ks"]):
        e = tb["e"][off:off + tb["E"][r]]
        off += tb["E"][r]
        dm = {"G": tb["G"], "C": Cn, "r": r, "rvidx": tb["rv"], "clear": 1, "Qm": tb["Qm"], "w": None if w is None else w[r]}
        if pool is not None:
            dm["harq_pool"], dm["harq_slot"] = pool, slot0 + r
        blocks.append({"y": e.astype(np.int8) if in_fmt else e, "K": K, "max_iterations": max_it, "crc_type": 0 if Cn == 1 else 1,
                       "F": F if r == 0 else 0, "dematch": dm, "tb_id": 1, "in_fmt": in_fmt})
    return blocks


def _check(tb, rx, outs, status):
    for r, K in enumerate(tb["Ks"]):
        assert status[r] == rx["status"][r], (r, status[r], rx["status"][r])
        assert np.array_equal(outs[r], rx["c"][r]), r


@pytest.mark.parametrize("tbs,G,Qm", [(30576, 57600, 4), (7736, 14400, 4), (3000, 4800, 2)])
def test_int8_soft_bit_feed_equals_int16_feed(capi, tbs, G, Qm):
    tb = chain.make_tb(tbs, G, Qm, seed=3, sigma_over_A=0.6)          # |e| <= 127 by construction (A = 8)
    assert np.abs(tb["e"]).max() <= 127
    rx = chain.rx_tb(tb, 6, downlink=False)
    n = len(tb["Ks"])
    for fmt in (0, 1):
        pool = capi.HarqPool(n, max(tb["Ks"]))
        outs, status = capi.decode_batch(_tb_blocks(tb, 6, pool=pool, in_fmt=fmt))
        _check(tb, rx, outs, status)
        for r in range(n):
            ncb = rx["w"][r].size
            assert np.array_equal(pool.read(r, ncb), rx["w"][r]), (fmt, r)
        pool.close()
    # host-authoritative w with the int8 feed, pageable memory (staged copies)
    w = [np.zeros_like(x) for x in rx["w"]]
    outs, status = capi.decode_batch(_tb_blocks(tb, 6, in_fmt=1, w=w))
    _check(tb, rx, outs, status)
    for r in range(n):
        assert np.array_equal(w[r], rx["w"][r])


def test_int8_feed_rejected_without_front_end(capi):
    y = np.zeros(3 * 40 + 12, dtype=np.int16)
    with pytest.raises(RuntimeError):
        capi.decode_batch([{"y": y, "K": 40, "max_iterations": 2, "crc_type": 1, "in_fmt": 1}])


def test_other_gpu_batch_pool_and_current_device(capi):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    tb = chain.make_tb(30576, 57600, 4, seed=11, sigma_over_A=0.9)    # 5 x K=6144 near the waterfall
    rx = chain.rx_tb(tb, 6, downlink=False)
    n = len(tb["Ks"])
    for gpu in range(1, torch.cuda.device_count()):
        pool = capi.HarqPool(n, 6144, gpu=gpu)
        assert torch.cuda.current_device() == 0
        outs, status = capi.decode_batch(_tb_blocks(tb, 6, pool=pool), gpu=gpu)
        assert torch.cuda.current_device() == 0, "the library moved the caller's current device"
        _check(tb, rx, outs, status)
        for r in range(n):
            assert np.array_equal(pool.read(r, rx["w"][r].size), rx["w"][r])
        assert torch.cuda.current_device() == 0
        # a pool on another GPU than the batch is refused
        with pytest.raises(RuntimeError):
            capi.decode_batch(_tb_blocks(tb, 6, pool=pool), gpu=0)
        pool.close()
        # the single-call entry points still run on GPU 0 afterwards, and the TX mirror follows its gpu argument
        ret, by = capi.phy_threegpplte_turbo_decoder16(rx["y"][0], 6144, 0, 0, 6, 1, 0)
        assert ret == rx["status"][0] and np.array_equal(by, rx["c"][0])
        info = tb["cb"][0]
        (e0,) = capi.tx_batch([{"c": info, "K": 6144, "G": 57600, "C": 5, "r": 0, "rvidx": 0, "Qm": 4}], gpu=0)
        (e1,) = capi.tx_batch([{"c": info, "K": 6144, "G": 57600, "C": 5, "r": 0, "rvidx": 0, "Qm": 4}], gpu=gpu)
        assert np.array_equal(e0, e1) and torch.cuda.current_device() == 0


def test_one_thread_drives_all_gpus_concurrently(capi):
    """one submit per GPU in flight at the same time from one thread, then the waits (the by-cell sharding of configs[3]
    in its single-process form)"""
    import ctypes as C
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs two GPUs")
    from openair4g_b200 import sharding
    tbs = [chain.make_tb(7736, 14400, 4, seed=20 + i, sigma_over_A=0.7) for i in range(2 * ng)]
    rxs = [chain.rx_tb(t, 4, downlink=False) for t in tbs]
    owner = sharding.assign_by_cell(list(range(len(tbs))), ng)           # cell i = transport block i
    pools = [capi.HarqPool(2 * len(tbs), 3904, gpu=g) for g in range(ng)]
    handles, keep = [], []
    for g in range(ng):
        mine = [i for i in range(len(tbs)) if owner[i] == g]
        blocks = []
        for i in mine:
            blocks += _tb_blocks(tbs[i], 4, pool=pools[g], slot0=2 * i)
        n = len(blocks)
        descs = (capi.CbDesc * n)()
        status = np.full(n, 255, dtype=np.uint8)
        outs = [np.zeros(b["K"] // 8, dtype=np.uint8) for b in blocks]
        for j, b in enumerate(blocks):
            d = descs[j]
            y = np.ascontiguousarray(b["y"], dtype=np.int16)
            keep.append(y)
            d.in_ = y.ctypes.data; d.decoded_bytes = outs[j].ctypes.data; d.status = status.ctypes.data + j
            d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable, d.dematch_enable = b["K"], 4, b["crc_type"], b["F"], 1, 1
            dm = b["dematch"]
            d.G, d.C, d.r, d.rvidx, d.clear, d.Qm, d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = dm["G"], dm["C"], dm["r"], 0, 1, dm["Qm"], 1, 8, 1, 1827072
            d.harq_pool = pools[g].handle; d.harq_slot = dm["harq_slot"]
        h = C.c_void_p()
        assert capi.lib.oai_turbo_submit_batch(descs, n, 0, g, C.byref(h)) == 0, capi.last_error()
        handles.append((h, mine, outs, status, descs))
    for h, mine, outs, status, _ in handles:
        assert capi.lib.oai_turbo_wait(h) == 0, capi.last_error()
        j = 0
        for i in mine:
            for r in range(len(tbs[i]["Ks"])):
                assert status[j] == rxs[i]["status"][r] and np.array_equal(outs[j], rxs[i]["c"][r]), (i, r)
                j += 1
    for p in pools:
        p.close()
