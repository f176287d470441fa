"""Phase trace (OAI_TURBO_TRACE=1) of one dlsim-shaped subframe through the host-buffer call, both modes."""
import os
import sys

sys.path.insert(0, ".")
os.environ["OAI_TURBO_TRACE"] = "1"
import bench
from openair4g_b200 import capi

capi.init_td16()
for flags in (0, capi.BATCH_SLIDING_WINDOW):
    print("== flags", flags, file=sys.stderr)
    r = bench.subframe_latency(capi, 75376, 90000, 6, 6, reps=3, flags=flags)
    print(flags, r, file=sys.stderr)
