import sys, os
os.environ["OAI_TURBO_DEBUG_T"]="1"
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from openair4g_b200 import capi
from test_golden import iter_td16
capi.init_td16()
def i16(v): return ((int(v) + 32768) % 65536) - 32768
def sat(v): return max(-32768, min(32767, int(v)))
def T_host(tl):
    m11=[sat(int(tl[2*i])+int(tl[2*i+1]))>>1 for i in range(3)]
    m10=[sat(int(tl[2*i])-int(tl[2*i+1]))>>1 for i in range(3)]
    b0=i16(-m11[2]); b1=m11[2]
    b0_2=i16(b0-m11[1]); b1_2=i16(b0+m11[1]); b2_2=i16(b1+m10[1]); b3_2=i16(b1-m10[1])
    t=[i16(b0_2-m11[0]),i16(b0_2+m11[0]),i16(b1_2+m10[0]),i16(b1_2-m10[0]),i16(b2_2-m10[0]),i16(b2_2+m10[0]),i16(b3_2+m11[0]),i16(b3_2-m11[0])]
    bm=max(t)
    return [i16(x-bm) for x in t]
cases = list(iter_td16())
for idx in (28, 42):
    y, out, K, max_it, crc, F, ret = cases[idx]
    print("case", idx, "host T1", T_host(y[3*K:3*K+6]), "T2", T_host(y[3*K+6:3*K+12]), "tails", list(y[3*K:]))
    sys.stdout.flush()
    capi.debug_map16(y, K, 1, 2)
