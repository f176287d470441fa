import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3*K+12; Be = 4096
ynp = torch.randint(-16, 17, (Be, row), dtype=torch.int16).pin_memory().numpy()
calls = [capi.HostBatchCall(ynp, K, 6, 1) for _ in range(2)]
for c in calls: c.run(); c.run()
pending = None; log = []
T0 = time.perf_counter()
for i in range(6):
    t0 = time.perf_counter(); h = calls[i & 1].submit(); t1 = time.perf_counter()
    if pending is not None: pending[0].wait(pending[1])
    t2 = time.perf_counter()
    log.append((i, (t0-T0)*1e3, (t1-t0)*1e3, (t2-t1)*1e3))
    pending = (calls[i & 1], h)
pending[0].wait(pending[1])
for l in log: print("step %d at %.2f ms: submit %.2f ms, wait(prev) %.2f ms" % l)
print("total %.2f ms" % ((time.perf_counter()-T0)*1e3))
