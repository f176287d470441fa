"""Throughput in the waterfall / clean regimes as a function of the LLR amplitude A (the fast-path guard depends on it)."""
import sys, importlib.util
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from openair4g_b200 import capi
spec = importlib.util.spec_from_file_location('bench', '/root/repo/bench.py'); bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
capi.init_td16()
K = 6144; row = 3*K+12; B = 28416
for name, sig in (("clean", 0.5), ("waterfall", 1.08)):
    for A in (8, 32, 64, 128, 256, 512):
        ys, info = bench.coded_inputs(K, 64, sig, 4242, A=A)
        idx = (torch.arange(B, device="cuda") * 29) % 64
        y = torch.from_numpy(ys).cuda()[idx].contiguous()
        out = torch.zeros((B, K//8), dtype=torch.uint8, device="cuda"); st = torch.zeros(B, dtype=torch.uint8, device="cuda")
        plan = capi.DevPlan(B, K, 6, 1)
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(2): plan.decode(y.data_ptr(), row, out.data_ptr(), K//8, st.data_ptr(), s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): plan.decode(y.data_ptr(), row, out.data_ptr(), K//8, st.data_ptr(), s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        hist = torch.bincount(st.long(), minlength=8).tolist()
        print("%-9s A=%4d: %6.2f ms -> %6.0f Mbit/s  return values %s" % (name, A, ms, B*K/ms/1e3, {i: h for i, h in enumerate(hist) if h}))
        plan.close()
