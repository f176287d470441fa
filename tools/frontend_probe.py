"""Throughput of the fused front end + decoder from host soft bits e (multi-cell uplink shape, BASELINE configs[3]):
100 PRB MCS16 -> 5 code blocks of K=6144 per UE, G=57600, Qm=4, E=11520 per block."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
from openair4g_b200 import capi
capi.init_td16()
K, G, Cb, Qm = 6144, 57600, 5, 4
E = G // Cb
def run(n_ue, use_pool, rounds=3):
    n = n_ue * Cb
    pin = capi.PinnedArray((n_ue, G), np.int16)
    rng = np.random.default_rng(1)
    pin.array[...] = rng.integers(-16, 17, size=(n_ue, G)).astype(np.int16)
    out = capi.PinnedArray((n, K // 8), np.uint8)
    status = np.zeros(n, dtype=np.uint8)
    pool = capi.HarqPool(n, K) if use_pool else None
    wbuf = None if use_pool else np.zeros((n, 3 * 6176), dtype=np.int16)
    descs = (capi.CbDesc * n)()
    for u in range(n_ue):
        for r in range(Cb):
            i = u * Cb + r
            d = descs[i]
            d.in_ = pin.array.ctypes.data + (u * G + r * E) * 2
            d.decoded_bytes = out.array.ctypes.data + i * (K // 8)
            d.status = status.ctypes.data + i
            d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable, d.dematch_enable = K, 6, 1, 0, 1, 1
            d.G, d.C, d.r, d.rvidx, d.clear, d.Qm, d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = G, Cb, r, 0, 1, Qm, 1, 8, 1, 1827072
            d.tb_id = u
            if use_pool:
                d.harq_pool = pool.handle; d.harq_slot = i
            else:
                d.w = wbuf.ctypes.data + i * 3 * 6176 * 2
    def call():
        h = C.c_void_p()
        rc = capi.lib.oai_turbo_submit_batch(descs, n, 0, -1, C.byref(h))
        assert rc == 0, capi.last_error()
        assert capi.lib.oai_turbo_wait(h) == 0
    call()
    t0 = time.perf_counter()
    for _ in range(rounds): call()
    dt = (time.perf_counter() - t0) / rounds
    print("%5d UEs x 5 blocks (K=6144, E=11520) %s: %.2f ms -> %.0f Mbit/s info, status %s" % (
        n_ue, "HARQ pool in HBM" if use_pool else "host-authoritative w", dt * 1e3, n * K / dt / 1e6, sorted(set(status))))
    if pool: pool.close()
for n_ue in (256, 2048, 8192):
    run(n_ue, True)
    run(n_ue, False)
