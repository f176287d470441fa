import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3*K+12; Be = 4096
y_pin = torch.randint(-16, 17, (Be, row), dtype=torch.int16).pin_memory()
d = torch.empty((Be, row), dtype=torch.int16, device="cuda")
for _ in range(3): d.copy_(y_pin, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(y_pin, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter()-t0)/5
print("H2D pinned %.1f MB in %.2f ms = %.1f GB/s" % (y_pin.numel()*2/1e6, dt*1e3, y_pin.numel()*2/dt/1e9))
ynp = y_pin.numpy()
for nchunk in (1, 2, 4, 8):
    calls = [capi.HostBatchCall(ynp[i*Be//nchunk:(i+1)*Be//nchunk], K, 6, 1) for i in range(nchunk)]
    import ctypes as C
    def run_all():
        hs = []
        for c in calls:
            h = C.c_void_p()
            rc = capi.lib.oai_turbo_submit_batch(c.descs, c.n, 0, -1, C.byref(h)); assert rc == 0
            hs.append(h)
        for h in hs:
            assert capi.lib.oai_turbo_wait(h) == 0
    for _ in range(2): run_all()
    t0 = time.perf_counter()
    for _ in range(5): run_all()
    dt = (time.perf_counter()-t0)/5
    print("chunks", nchunk, "e2e %.2f ms/step -> %.0f Mbit/s" % (dt*1e3, Be*K/dt/1e6))
# host-side cost only: time submit (returns after enqueue) vs wait
c = capi.HostBatchCall(ynp, K, 6, 1)
import ctypes as C
for _ in range(2): c.run()
h = C.c_void_p(); t0 = time.perf_counter(); capi.lib.oai_turbo_submit_batch(c.descs, c.n, 0, -1, C.byref(h)); t1 = time.perf_counter(); capi.lib.oai_turbo_wait(h); t2 = time.perf_counter()
print("submit %.2f ms, wait %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3))
# pipelined, two in flight
calls = [capi.HostBatchCall(ynp, K, 6, 1) for _ in range(2)]
for c in calls: c.run(); c.run()
for depth_name, fn in (("sequential", None), ("pipelined", 1)):
    t0 = time.perf_counter(); steps = 10; pending = None
    for i in range(steps):
        if fn is None:
            calls[i & 1].run()
        else:
            h = calls[i & 1].submit()
            if pending is not None: pending[0].wait(pending[1])
            pending = (calls[i & 1], h)
    if pending is not None: pending[0].wait(pending[1])
    dt = (time.perf_counter() - t0) / steps
    print(depth_name, "%.2f ms/step -> %.0f Mbit/s" % (dt * 1e3, Be * K / dt / 1e6))
