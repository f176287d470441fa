"""Does a concurrent host->device copy slow the decode kernels down?"""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3*K+12; B = 42624
g = torch.Generator(device="cuda"); g.manual_seed(1)
y = torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g)
out = torch.zeros((B, K//8), dtype=torch.uint8, device="cuda"); st = torch.zeros(B, dtype=torch.uint8, device="cuda")
plan = capi.DevPlan(B, K, 6, 1)
s_main = torch.cuda.Stream(); s_copy = torch.cuda.Stream()
host = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); dev = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def decode_ms(with_copy, d2h=False):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if with_copy:
        with torch.cuda.stream(s_copy):
            c0.record()
            for _ in range(6):
                if d2h: host.copy_(dev, non_blocking=True)
                else: dev.copy_(host, non_blocking=True)
            c1.record()
    with torch.cuda.stream(s_main):
        e0.record()
        for _ in range(3):
            plan.decode(y.data_ptr(), row, out.data_ptr(), K//8, st.data_ptr(), s_main.cuda_stream)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3, (6 * 512 / 1024 / (c0.elapsed_time(c1) * 1e-3)) if with_copy else 0.0
for _ in range(2): decode_ms(False)
print("decode alone: %.2f ms" % decode_ms(False)[0])
ms, gbs = decode_ms(True); print("decode with concurrent H2D: %.2f ms (copy %.1f GiB/s)" % (ms, gbs))
ms, gbs = decode_ms(True, d2h=True); print("decode with concurrent D2H: %.2f ms (copy %.1f GiB/s)" % (ms, gbs))
torch.cuda.synchronize(); c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0.record(); [dev.copy_(host, non_blocking=True) for _ in range(6)]; c1.record(); torch.cuda.synchronize()
print("H2D alone: %.1f GiB/s" % (6 * 512 / 1024 / (c0.elapsed_time(c1) * 1e-3)))
