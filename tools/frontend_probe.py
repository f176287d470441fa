"""Throughput of the fused front end + decoder from host soft bits e (multi-cell uplink shape, BASELINE configs[3]):
100 PRB MCS16 -> 5 code blocks of K=6144 per UE, G=57600, Qm=4, E=11520 per block.  One submit at a time, HARQ pool in HBM.
  python tools/frontend_probe.py [n_ue] [int8]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openair4g_b200 import capi
capi.init_td16()
K, G, Cb, Qm = 6144, 57600, 5, 4
E = G // Cb
n_ue = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
fmt = 1 if (len(sys.argv) > 2 and sys.argv[2] == "int8") else 0
n = n_ue * Cb
dt_np = np.int8 if fmt else np.int16
pin = capi.PinnedArray((n_ue, G), dt_np)
rng = np.random.default_rng(1)
pin.array[...] = rng.integers(-16, 17, size=(n_ue, G)).astype(dt_np)
out = capi.PinnedArray((n, K // 8), np.uint8)
status = np.zeros(n, dtype=np.uint8)
pool = capi.HarqPool(n, K)
descs = (capi.CbDesc * n)()
for u in range(n_ue):
    for r in range(Cb):
        i = u * Cb + r
        d = descs[i]
        d.in_ = pin.array.ctypes.data + (u * G + r * E) * pin.array.itemsize
        d.decoded_bytes = out.array.ctypes.data + i * (K // 8)
        d.status = status.ctypes.data + i
        d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable, d.dematch_enable = K, 6, 1, 0, 1, 1
        d.G, d.C, d.r, d.rvidx, d.clear, d.Qm, d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = G, Cb, r, 0, 1, Qm, 1, 8, 1, 1827072
        d.tb_id = u; d.harq_pool = pool.handle; d.harq_slot = i; d.in_fmt = fmt
def call():
    h = C.c_void_p()
    assert capi.lib.oai_turbo_submit_batch(descs, n, 0, -1, C.byref(h)) == 0, capi.last_error()
    assert capi.lib.oai_turbo_wait(h) == 0
call()
t0 = time.perf_counter()
for _ in range(3): call()
dt = (time.perf_counter() - t0) / 3
print("%5d UEs x 5 blocks (K=6144, E=11520, %s soft bits): %.2f ms -> %.0f Mbit/s info, status %s" % (n_ue, dt_np.__name__, dt * 1e3, n * K / dt / 1e6, sorted(set(status.tolist()))))
