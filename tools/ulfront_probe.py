"""configs[3] shape through the uplink front end + transport-block assembly (oai_turbo_submit_tbs with oai_ul_front_t):
time spent inside submit() (host bookkeeping + enqueue) and inside wait(), one call at a time.
  python tools/ulfront_probe.py [n_ue] [int8]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openair4g_b200 import capi
capi.init_td16()
K, G, Cb, Qm = 6144, 57600, 5, 4
n_ue = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
fmt = 1 if (len(sys.argv) > 2 and sys.argv[2] == "int8") else 0
n = n_ue * Cb
dt_np = np.int8 if fmt else np.int16
pin = capi.PinnedArray((n_ue, G), dt_np)
pin.array[...] = np.random.default_rng(1).integers(-16, 17, size=(n_ue, G)).astype(dt_np)
rc, z = capi.ulsch_control_sizes(0, 1, 0, 1200, 12, 40, 40, 16, Cb * K, 100, Qm, 12)
status = np.zeros(n, dtype=np.uint8)
pool = capi.HarqPool(n, K)
tb_bytes = Cb * (K // 8) - 3 * Cb
bbuf = capi.PinnedArray((n_ue, tb_bytes), np.uint8)
ret = np.zeros(n_ue, dtype=np.uint8); oack = np.zeros((n_ue, 2), dtype=np.uint8)
descs = (capi.CbDesc * n)(); ufs = (capi.UlFront * n_ue)(); tbs = (capi.TbDesc * n_ue)()
for u in range(n_ue):
    f = ufs[u]
    f.llr, f.llr_fmt, f.c_init = pin.array.ctypes.data + u * G * pin.array.itemsize, fmt, (0x1234 << 14) + (3 << 9) + (u % 504)
    f.Qm, f.Ncp, f.O_ACK, f.O_RI, f.bundling, f.Nbundled, f.Cmux = Qm, 0, 1, 0, 0, 1, 12
    f.Qprime_RI, f.Qprime_ACK, f.Qprime_CQI, f.Hprime = 0, z["Qprime_ACK"], 0, z["Hprime"]
    f.o_ACK = oack.ctypes.data + 2 * u
    t = tbs[u]
    t.first_cb, t.C, t.uplink, t.b, t.b_capacity, t.ret = u * Cb, Cb, 1, bbuf.array.ctypes.data + u * tb_bytes, tb_bytes, ret.ctypes.data + u
    t.ul_front = C.pointer(f)
    for r in range(Cb):
        d = descs[u * Cb + r]
        d.status = status.ctypes.data + u * Cb + r
        d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable, d.dematch_enable = K, 6, 1, 0, 1, 1
        d.G, d.C, d.r, d.rvidx, d.clear, d.Qm, d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = G, Cb, r, 0, 1, Qm, 1, 8, 1, 1827072
        d.tb_id = u; d.harq_pool = pool.handle; d.harq_slot = u * Cb + r
def call():
    h = C.c_void_p()
    t0 = time.perf_counter()
    assert capi.lib.oai_turbo_submit_tbs(descs, n, tbs, n_ue, 0, -1, C.byref(h)) == 0, capi.last_error()
    t1 = time.perf_counter()
    assert capi.lib.oai_turbo_wait(h) == 0
    return t1 - t0, time.perf_counter() - t1
call(); call()
r = [call() for _ in range(4)]
print("%d allocations (%d blocks), %s llr: submit %.2f ms, wait %.2f ms -> %.0f Mbit/s; ret %s" % (
    n_ue, n, dt_np.__name__, 1e3 * np.median([a for a, _ in r]), 1e3 * np.median([b for _, b in r]),
    n * K / np.median([a + b for a, b in r]) / 1e6, sorted(set(ret.tolist()))))
