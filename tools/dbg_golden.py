import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from openair4g_b200 import capi
from test_golden import iter_td16
capi.init_td16()
cases = list(iter_td16())
for idx in (23, 28, 41, 42, 48, 63, 72):
    y, out, K, max_it, crc, F, ret = cases[idx]
    r, b = capi.phy_threegpplte_turbo_decoder16(y, K, 0, 0, max_it, crc, F)
    nd = np.unpackbits(b ^ out).sum()
    print("case", idx, "K", K, "alone: ret", r, "want", ret, "bit diffs", nd, "max|y|", np.abs(y.astype(int)).max())
    for mi in (2, 3):
        from oracle import loader
        wb, wr = loader.port_decode16(y, K, mi, crc, F)
        r, b = capi.phy_threegpplte_turbo_decoder16(y, K, 0, 0, mi, crc, F)
        print("   max_it", mi, "ret", r, wr, "bit diffs", np.unpackbits(b ^ wb).sum())
