"""Throughput of the exact saturating policy of k_map16 (inputs beyond the fast-path guard)."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3*K+12; B = 14208
for amp in (16, 200, 700, 1500, 3000, 12000):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    y = torch.randint(-amp, amp + 1, (B, row), dtype=torch.int16, device="cuda", generator=g)
    out = torch.zeros((B, K//8), dtype=torch.uint8, device="cuda"); st = torch.zeros(B, dtype=torch.uint8, device="cuda")
    plan = capi.DevPlan(B, K, 6, 1)
    s = torch.cuda.current_stream().cuda_stream
    plan.decode(y.data_ptr(), row, out.data_ptr(), K//8, st.data_ptr(), s); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.decode(y.data_ptr(), row, out.data_ptr(), K//8, st.data_ptr(), s); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("uniform +-%5d: %.2f ms -> %.0f Mbit/s" % (amp, ms, B*K/ms/1e3))
    plan.close()
