import sys
sys.path.insert(0, '/root/repo')
import torch
from openair4g_b200 import capi
capi.init_td16()
K = 6144; row = 3*K+12; B = 14208
g = torch.Generator(device="cuda"); g.manual_seed(1)
y = torch.randint(-3000, 3001, (B, row), dtype=torch.int16, device="cuda", generator=g)
out = torch.zeros((B, K//8), dtype=torch.uint8, device="cuda"); st = torch.zeros(B, dtype=torch.uint8, device="cuda")
plan = capi.DevPlan(B, K, 6, 1)
plan.decode(y.data_ptr(), row, out.data_ptr(), K//8, st.data_ptr(), torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
