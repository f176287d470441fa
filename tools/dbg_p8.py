import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from openair4g_b200 import capi
from oracle import vectors
capi.init_td16()
for K in (40, 512, 6144):
    y, _ = vectors.llr_block(K, 1, "waterfall")
    for term in (0, 1):
        a = capi.debug_map16(y, K, term, 1)
        b = capi.debug_map16(y, K, term, 3)
        d = np.nonzero(a != b)[0]
        print(K, term, "diffs", d.size, [(int(i)//8, int(i)%8, int(a[i]), int(b[i])) for i in d[:8]])
