"""Throughput of the GPU TX mirror (oai_turbo_tx_batch): K=6144, E=11520 (the multi-cell UL shape), host pointers
(pinned staging inside the call) and device pointers.  Prints info Mbit/s through the encoder + rate matching."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from openair4g_b200 import capi  # noqa: E402

capi.init_td16()
K, G, n = 6144, 11520, 42624
rng = np.random.default_rng(1)
info = rng.integers(0, 256, size=(n, K // 8)).astype(np.uint8)
e = np.zeros((n, G), dtype=np.uint8)
descs = (capi.TxDesc * n)()
for i in range(n):
    d = descs[i]
    d.c = info[i].ctypes.data; d.e = e[i].ctypes.data
    d.K = K; d.G = G; d.Nsoft = 1827072; d.C = 1; d.Mdlharq = 8; d.Kmimo = 1; d.rvidx = 0; d.Qm = 2; d.Nl = 1; d.r = 0
for rep in range(3):
    t0 = time.perf_counter()
    rc = capi.lib.oai_turbo_tx_batch(descs, n, 0, -1)
    dt = time.perf_counter() - t0
    assert rc == 0, capi.last_error()
    print("host pointers  : %7.2f ms -> %8.0f Mbit/s info (%d blocks, %d launches so far)" % (dt * 1e3, n * K / dt / 1e6, n, capi.launch_count()))
# device pointers
c_dev = torch.from_numpy(info).cuda()
e_dev = torch.zeros((n, G), dtype=torch.uint8, device="cuda")
for i in range(n):
    descs[i].c = c_dev.data_ptr() + i * (K // 8)
    descs[i].e = e_dev.data_ptr() + i * G
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    rc = capi.lib.oai_turbo_tx_batch(descs, n, capi.TX_DEVICE_POINTERS, -1)
    dt = time.perf_counter() - t0
    assert rc == 0, capi.last_error()
    print("device pointers: %7.2f ms -> %8.0f Mbit/s info" % (dt * 1e3, n * K / dt / 1e6))
assert np.array_equal(e_dev.cpu().numpy(), e), "device-pointer and host-pointer results differ"
print("results identical")
