"""isolated-decoder sweep over batch size and block size K (device-resident plan, noise regime:
uniform +-16 LLRs, CRC never passes -> exactly 6 iterations).  Prints time per batch and decoded information Mbit/s."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openair4g_b200 import capi
capi.init_td16()
s = torch.cuda.current_stream().cuda_stream
print("%6s %8s %12s %12s %10s" % ("K", "blocks", "us/batch", "Mbit/s", "ret==7"))
for K in (40, 512, 1024, 2048, 4096, 6144):
    row = 3 * K + 12
    for B in (1, 4, 16, 64, 256, 1024, 4096, 16384, 65536):
        g = torch.Generator(device="cuda"); g.manual_seed(K + B)
        y = torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g)
        out = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda")
        st = torch.zeros(B, dtype=torch.uint8, device="cuda")
        plan = capi.DevPlan(B, K, 6, 1)
        for _ in range(2):
            plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), s)
        torch.cuda.synchronize()
        reps = 20 if B <= 1024 else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        extra = ""
        if B == 65536:                    # per-class kernel times of one decode (CUDA events around every launch)
            plan.profile(True)
            plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), s)
            torch.cuda.synchronize()
            pm, pc = plan.profile(False, fetch=True)
            extra = "   demux %.0f us, map %.0f us (%d), x1 %.0f us, x2 %.0f us" % (pm[0] * 1e3, pm[1] * 1e3, pc[1], pm[2] * 1e3, pm[3] * 1e3)
        print("%6d %8d %12.1f %12.1f %10s%s" % (K, B, ms * 1e3, B * K / ms / 1e3, bool((st == 7).all()), extra))
        plan.close()
        del y, out, st
