"""one device-resident decode of 65536 blocks of K=512 (for ncu captures of the short-block regime)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openair4g_b200 import capi
capi.init_td16()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = 65536; row = 3 * K + 12
g = torch.Generator(device="cuda"); g.manual_seed(1)
y = torch.randint(-16, 17, (B, row), dtype=torch.int16, device="cuda", generator=g)
out = torch.zeros((B, K // 8), dtype=torch.uint8, device="cuda"); st = torch.zeros(B, dtype=torch.uint8, device="cuda")
plan = capi.DevPlan(B, K, 6, 1)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2): plan.decode(y.data_ptr(), row, out.data_ptr(), K // 8, st.data_ptr(), s)
torch.cuda.synchronize()
