"""Host-buffer path (oai_turbo_submit_batch + oai_turbo_wait, page-locked buffers) with the narrow input feed at different
pack-thread counts; run each configuration in a fresh process:  python tools/e2e_pack_probe.py <threads|off> [blocks]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1] if len(sys.argv) > 1 else "12"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 42624
if mode == "off":
    os.environ["OAI_TURBO_NO_NARROW_FEED"] = "1"
else:
    os.environ["OAI_TURBO_PACK_THREADS"] = mode
    os.environ["OAI_TURBO_PACK_MIN_GBS"] = "0"
import numpy as np, torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3 * K + 12
g = torch.Generator(); g.manual_seed(1)
y = torch.randint(-16, 17, (B, row), dtype=torch.int16, generator=g).pin_memory()
call = capi.HostBatchCall(y.numpy(), K, 6, 1)
for _ in range(3): call.run()
torch.cuda.synchronize()
ts = []
for _ in range(8):
    t0 = time.perf_counter(); out, st = call.run(); ts.append(time.perf_counter() - t0)
ts.sort()
print("narrow feed %s, %d blocks: median %.2f ms  min %.2f ms -> %.2f Gbit/s (status %s)" % (mode, B, 1e3 * ts[len(ts) // 2], 1e3 * ts[0], B * K / ts[len(ts) // 2] / 1e9, sorted(set(st.tolist()))))
