"""GPU tests at batch scale (BASELINE configs[2]: 1-65536 code blocks, K = 40...6144): batch-size edge cases of the
4-threads-per-block / 8-blocks-per-warp / 16-blocks-per-CTA mapping, the pipelined host-buffer path (parts, page-locked
buffers, direct output) and size-independent properties at a full wave of K=6144 blocks.  All through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import loader, vectors  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from openair4g_b200 import capi as c
    c.init_td16()
    return c


def _distinct(K, count, seed0):
    ys, want, infos = [], [], []
    for i in range(count):
        regime = ("clean", "waterfall", "noise")[i % 3]
        y, info = vectors.llr_block(K, seed0 + i, regime)
        ys.append(y)
        infos.append((regime, info))
        want.append(loader.port_decode16(y, K, 6, 1))
    return ys, want, infos


@pytest.mark.parametrize("K", [40, 512, 1024, 2048, 4096, 6144])
def test_batch_size_edges(capi, K):
    """Batch sizes around the warp (8 blocks) and CTA (16 blocks) boundaries of k_map16, one submit each."""
    ys, want, _ = _distinct(K, 6, 7000 + K)
    for n in (1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33, 100):
        blocks = [{"y": ys[i % 6], "K": K, "max_iterations": 6, "crc_type": 1} for i in range(n)]
        outs, status = capi.decode_batch(blocks)
        for i in range(n):
            assert status[i] == want[i % 6][1], (K, n, i)
            assert np.array_equal(outs[i][:K // 8], want[i % 6][0]), (K, n, i)


def test_pipelined_host_batch(capi):
    """6200 blocks of K=6144 through oai_turbo_submit_batch from page-locked memory: large enough for the pipelined
    form (2 parts, copies overlapping decodes, per-part D2H straight into the caller's page-locked output)."""
    K, n, nd = 6144, 6200, 12
    ys, want, infos = _distinct(K, nd, 9100)
    pin = capi.PinnedArray((n, 3 * K + 12), np.int16)
    for i in range(n):
        pin.array[i] = ys[(i * 5) % nd]
    call = capi.HostBatchCall(pin.array, K, 6, 1)
    for _ in range(2):                                  # second run reuses the pooled batch object
        out, st = call.run()
        for i in range(n):
            j = (i * 5) % nd
            assert st[i] == want[j][1], i
            assert np.array_equal(out[i], want[j][0]), i
    # round trip: every clean-regime block decodes to the bytes that were encoded, in 2 iterations
    for i in range(n):
        regime, info = infos[(i * 5) % nd]
        if regime == "clean":
            assert st[i] == 2 and np.array_equal(out[i], info)


def test_full_wave_properties(capi):
    """One full wave of the MAP kernel (14208 blocks, K=6144), device-resident: replicated inputs must give identical
    outputs wherever they sit in the batch (no cross-block interference), the status histogram is the oracle's, and the
    XOR of all decoded words (a checksum of checksums) matches the one computed from the oracle's results."""
    import torch
    K, n, nd = 6144, 14208, 9
    ys, want, _ = _distinct(K, nd, 9300)
    row = 3 * K + 12
    host = np.zeros((nd, row), dtype=np.int16)
    for j in range(nd):
        host[j] = ys[j]
    idx = torch.from_numpy((np.arange(n) * 7) % nd).cuda()
    y_dev = torch.from_numpy(host).cuda()[idx].contiguous()
    out_dev = torch.zeros((n, K // 8), dtype=torch.uint8, device="cuda")
    st_dev = torch.zeros(n, dtype=torch.uint8, device="cuda")
    plan = capi.DevPlan(n, K, 6, 1)
    plan.decode(y_dev.data_ptr(), row, out_dev.data_ptr(), K // 8, st_dev.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out, st = out_dev.cpu().numpy(), st_dev.cpu().numpy()
    src = (np.arange(n) * 7) % nd
    want_out = np.stack([w[0] for w in want])[src]
    want_st = np.array([w[1] for w in want], dtype=np.uint8)[src]
    assert np.array_equal(st, want_st)
    assert np.array_equal(np.bincount(st, minlength=8), np.bincount(want_st, minlength=8))
    assert np.array_equal(np.bitwise_xor.reduce(out.view(np.uint32), axis=0), np.bitwise_xor.reduce(want_out.view(np.uint32), axis=0))
    assert np.array_equal(out, want_out)
    plan.close()


def test_td8_large_batch_device_resident(capi):
    """8-bit decoder, device-resident plan, both hard-decision modes (K=6144: n % 128 == 0; K=5824: not)."""
    import torch
    for K in (6144, 5824):
        n, nd = 1000, 6
        ys, want = [], []
        for i in range(nd):
            y, _ = vectors.llr_block(K, 9500 + i, ("clean", "waterfall", "noise")[i % 3])
            ypad = np.zeros(3 * K + 12 + 36, dtype=np.int16)       # the reference reads up to 28 int16 past y
            ypad[:3 * K + 12] = y
            ys.append(ypad)
            want.append(loader.port_decode8(ypad, K, 6, 1))
        row = ys[0].size
        host = np.stack([ys[i % nd] for i in range(n)])
        y_dev = torch.from_numpy(host).cuda()
        out_dev = torch.zeros((n, K // 8), dtype=torch.uint8, device="cuda")
        st_dev = torch.zeros(n, dtype=torch.uint8, device="cuda")
        plan = capi.DevPlan(n, K, 6, 1, llr8=1)
        plan.decode(y_dev.data_ptr(), row, out_dev.data_ptr(), K // 8, st_dev.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        out, st = out_dev.cpu().numpy(), st_dev.cpu().numpy()
        for i in range(n):
            assert st[i] == want[i % nd][1], (K, i)
            assert np.array_equal(out[i], want[i % nd][0]), (K, i)
        plan.close()


def test_pipelined_host_batch_from_pageable_memory(capi):
    """Same as above from ordinary (pageable) numpy memory: inputs go through the library's pinned staging area part by
    part, outputs through the staged copy-back."""
    K, n, nd = 6144, 6100, 6
    ys, want, _ = _distinct(K, nd, 9700)
    y = np.stack([ys[(i * 5) % nd] for i in range(n)])
    blocks = [{"y": y[i], "K": K, "max_iterations": 6, "crc_type": 1} for i in range(n)]
    # decode_batch copies each y into its own array: build ONE contiguous pageable buffer instead
    import ctypes as C
    out = np.zeros((n, K // 8), dtype=np.uint8)
    status = np.zeros(n, dtype=np.uint8)
    descs = (capi.CbDesc * n)()
    for i in range(n):
        d = descs[i]
        d.in_ = y.ctypes.data + i * y.strides[0]
        d.decoded_bytes = out.ctypes.data + i * (K // 8)
        d.status = status.ctypes.data + i
        d.K, d.max_iterations, d.crc_type, d.decode_enable = K, 6, 1, 1
    h = C.c_void_p()
    assert capi.lib.oai_turbo_submit_batch(descs, n, 0, -1, C.byref(h)) == 0, capi.last_error()
    assert capi.lib.oai_turbo_wait(h) == 0
    for i in range(n):
        j = (i * 5) % nd
        assert status[i] == want[j][1], i
        assert np.array_equal(out[i], want[j][0]), i


def test_concurrent_callers(capi):
    """The reference entry point is called from up to 10 eNB RX threads at once (lte-softmodem.c:1197-1304): every
    thread owns its batch object and stream, the device context is shared."""
    import threading
    K = 2048
    ys, want, _ = _distinct(K, 6, 9800)
    errors = []

    def worker(tid):
        try:
            for rep in range(8):
                j = (tid + rep) % 6
                r, b = capi.phy_threegpplte_turbo_decoder16(ys[j], K, 0, 0, 6, 1, 0)
                if r != want[j][1] or not np.array_equal(b, want[j][0]):
                    errors.append((tid, rep, r, want[j][1]))
                outs, st = capi.decode_batch([{"y": ys[(j + k) % 6], "K": K, "max_iterations": 6, "crc_type": 1} for k in range(3)])
                for k in range(3):
                    if st[k] != want[(j + k) % 6][1] or not np.array_equal(outs[k][:K // 8], want[(j + k) % 6][0]):
                        errors.append((tid, rep, "batch", k))
        except Exception as ex:                           # noqa: BLE001
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(10)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_small_batch_graph_replay_follows_new_inputs(capi):
    """Batches of <= 1024 blocks replay a cached CUDA graph of their launch sequence (keyed by block count, iteration limit
    and buffers).  Successive batches of the same shape but different contents, block sizes, CRC types and early-exit
    behaviour must each match the oracle; a different iteration limit or block count must not reuse a stale graph."""
    from oracle import loader, vectors
    plans = [
        [(6144, "clean", 6, 1), (512, "noise", 6, 1), (40, "clean", 6, 0), (1056, "waterfall", 6, 1)],
        [(512, "noise", 6, 1), (6144, "waterfall", 6, 1), (6144, "noise", 6, 0), (48, "clean", 6, 1)],      # same count and limit
        [(6144, "clean", 6, 1), (512, "noise", 6, 1), (40, "clean", 6, 0), (1056, "waterfall", 6, 1)],      # first shape again
        [(6144, "waterfall", 3, 1), (512, "noise", 3, 1), (40, "clean", 3, 0), (1056, "clean", 3, 1)],      # other iteration limit
        [(6144, "noise", 6, 1), (512, "clean", 6, 1), (104, "full", 6, 1)],                                 # other block count
        [(6144, "full", 6, 1), (2048, "full", 6, 1), (40, "full", 6, 0), (1056, "full", 6, 1)],             # exact policy
    ]
    for rnd in range(2):
        for pi, plan in enumerate(plans):
            blocks, want = [], []
            for j, (K, regime, max_it, crc) in enumerate(plan):
                y, _ = vectors.llr_block(K, 100 * rnd + 10 * pi + j, regime, crc_type=crc)
                blocks.append({"y": y, "K": K, "max_iterations": max_it, "crc_type": crc})
                want.append(loader.port_decode16(y, K, max_it, crc))
            outs, status = capi.decode_batch(blocks)
            for (wb, wr), ob, st, b in zip(want, outs, status, blocks):
                assert st == wr, (rnd, pi, b["K"], st, wr)
                if b["max_iterations"] > 1:
                    assert np.array_equal(ob, wb), (rnd, pi, b["K"])


def test_narrow_feed_and_part_streams(capi, monkeypatch):
    """The pipelined host batch packs parts whose soft bits fit int8 on the host (narrow feed) and decodes its parts on
    several compute streams.  9000 blocks of K=6144 = 3 parts: one part holds a value beyond int8 and must go over as
    int16, the others as int8 -- results must equal the oracle's and the un-pipelined, un-packed call's, from page-locked
    and from pageable caller memory."""
    K, n, nd = 6144, 9000, 9
    ys, want, _ = _distinct(K, nd, 9400)
    assert all(np.abs(y).max() <= 127 for y in ys)
    big = ys[3].copy()
    big[777] = 300                                        # one soft bit beyond int8
    wbig = loader.port_decode16(big, K, 6, 1)
    pin = capi.PinnedArray((n, 3 * K + 12), np.int16)
    for i in range(n):
        pin.array[i] = ys[(i * 7) % nd]
    pin.array[4000] = big                                 # in the second part
    call = capi.HostBatchCall(pin.array, K, 6, 1)
    monkeypatch.setenv("OAI_TURBO_PACK_THREADS", "4")     # the narrow feed is opt-in
    monkeypatch.setenv("OAI_TURBO_PACK_MIN_GBS", "0")     # keep it on whatever this box's pack rate is
    out, st = call.run()
    out, st = out.copy(), st.copy()
    for i in range(n):
        wb, wr = wbig if i == 4000 else want[(i * 7) % nd]
        assert st[i] == wr and np.array_equal(out[i], wb), i
    monkeypatch.setenv("OAI_TURBO_NO_NARROW_FEED", "1")
    out2, st2 = call.run()
    assert np.array_equal(out2, out) and np.array_equal(st2, st)
    monkeypatch.setenv("OAI_TURBO_NO_PIPELINE", "1")
    out3, st3 = call.run()
    assert np.array_equal(out3, out) and np.array_equal(st3, st)
    monkeypatch.delenv("OAI_TURBO_NO_PIPELINE")
    monkeypatch.delenv("OAI_TURBO_NO_NARROW_FEED")
    pageable = np.array(pin.array)                        # pageable caller memory: packed straight from it
    call2 = capi.HostBatchCall(pageable, K, 6, 1)
    out4, st4 = call2.run()
    assert np.array_equal(out4, out) and np.array_equal(st4, st)
