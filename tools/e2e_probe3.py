import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3*K+12; Be = 23680
y = torch.randint(-16, 17, (Be, row), dtype=torch.int16).pin_memory()
call = capi.HostBatchCall(y.numpy(), K, 6, 1)
for _ in range(2): call.run()
for i in range(4):
    t0 = time.perf_counter(); h = call.submit(); t1 = time.perf_counter(); call.wait(h); t2 = time.perf_counter()
    print("step %d: submit %.2f ms, wait %.2f ms, total %.2f ms -> %.0f Mbit/s" % (i, (t1-t0)*1e3, (t2-t1)*1e3, (t2-t0)*1e3, Be*K/(t2-t0)/1e6))
