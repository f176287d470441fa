import sys, time
sys.path.insert(0, '/root/repo')
import torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3*K+12
def run(nplans, B, steps=5):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    ys = [torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g) for _ in range(nplans)]
    outs = [torch.zeros((B, K//8), dtype=torch.uint8, device="cuda") for _ in range(nplans)]
    sts = [torch.zeros(B, dtype=torch.uint8, device="cuda") for _ in range(nplans)]
    plans = [capi.DevPlan(B, K, 6, 1) for _ in range(nplans)]
    streams = [torch.cuda.Stream() for _ in range(nplans)]
    def step():
        for p, y, o, s, st in zip(plans, ys, outs, sts, streams):
            p.decode(y.data_ptr(), row, o.data_ptr(), K//8, s.data_ptr(), st.cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print("plans %d x %d blocks: %.2f ms/step -> %.0f Mbit/s" % (nplans, B, dt*1e3, nplans*B*K/dt/1e6))
    for p in plans: p.close()
run(1, 42624)
run(2, 21312)
run(3, 14208)
run(4, 10656)
run(6, 7104)
