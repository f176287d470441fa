import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from openair4g_b200 import capi
from oracle import loader, vectors
from test_golden import iter_td16
P = loader.port()
capi.init_td16()

def port_map(y, K, term):
    W = K // 8
    pos = np.arange(K); st = (pos % W) * 8 + pos // W
    s = np.zeros(K + 16, np.int16); p = np.zeros(K + 16, np.int16)
    s[st] = y[0:3*K:3]; p[st] = y[(2 if term else 1):3*K:3]
    t = y[3*K:]
    for i in range(3):
        if term == 0:
            s[K+i] = t[2*i]; p[K+i] = t[2*i+1]
        else:
            s[K+8+i] = t[6+2*i]; p[K+i] = t[7+2*i]
    e = np.zeros(K + 16, np.int16)
    P.orc_log_map16(s, p, e, K, term, None, None)
    return e[:K]

cases = list(iter_td16())
for idx in (0, 23, 28, 41, 42, 72):
    y, out, K, max_it, crc, F, ret = cases[idx]
    for term in (0, 1):
        for pol in (2, 0):
            g = capi.debug_map16(y, K, term, pol)
            w = port_map(y, K, term)
            d = np.nonzero(g != w)[0]
            print("case", idx, "K", K, "term", term, "policy", pol, "diffs", d.size, [(int(i)//8, int(i)%8, int(g[i]), int(w[i])) for i in d[:6]])
