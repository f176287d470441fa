"""e feed (configs[3] shape), two batches in flight, with host timestamps of submit / wait: python tools/inflight_probe.py [int8]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openair4g_b200 import capi
capi.init_td16()
K, G, Cb, Qm = 6144, 57600, 5, 4
E = G // Cb
n_ue = 8192; n = n_ue * Cb
fmt = 1 if (len(sys.argv) > 1 and sys.argv[1] == "int8") else 0
dt_np = np.int8 if fmt else np.int16
pin = capi.PinnedArray((n_ue, G), dt_np)
pin.array[...] = np.random.default_rng(1).integers(-16, 17, size=(n_ue, G)).astype(dt_np)
lanes = []
for _ in range(2):
    out = capi.PinnedArray((n, K // 8), np.uint8); status = np.zeros(n, dtype=np.uint8); pool = capi.HarqPool(n, K)
    descs = (capi.CbDesc * n)()
    for u in range(n_ue):
        for r in range(Cb):
            i = u * Cb + r; d = descs[i]
            d.in_ = pin.array.ctypes.data + (u * G + r * E) * pin.array.itemsize
            d.decoded_bytes = out.array.ctypes.data + i * (K // 8); d.status = status.ctypes.data + i
            d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable, d.dematch_enable = K, 6, 1, 0, 1, 1
            d.G, d.C, d.r, d.rvidx, d.clear, d.Qm, d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = G, Cb, r, 0, 1, Qm, 1, 8, 1, 1827072
            d.tb_id = u; d.harq_pool = pool.handle; d.harq_slot = i; d.in_fmt = fmt
    lanes.append((descs, out, status, pool))
def submit(l):
    h = C.c_void_p(); assert capi.lib.oai_turbo_submit_batch(lanes[l][0], n, 0, -1, C.byref(h)) == 0; return h
def loop(steps, log=False):
    t0 = time.perf_counter(); pending = None
    for i in range(steps):
        a = time.perf_counter(); h = submit(i & 1); b = time.perf_counter()
        if pending is not None: capi.lib.oai_turbo_wait(pending)
        c = time.perf_counter()
        if log: print("step %d: submit %.2f..%.2f  wait(prev) returns %.2f" % (i, 1e3 * (a - t0), 1e3 * (b - t0), 1e3 * (c - t0)))
        pending = h
    capi.lib.oai_turbo_wait(pending)
    return time.perf_counter() - t0
loop(4)
dt = loop(8, log=True)
print("%s: %.2f ms per step -> %.0f Mbit/s" % (dt_np.__name__, 1e3 * dt / 8, 8 * n * K / dt / 1e6))
