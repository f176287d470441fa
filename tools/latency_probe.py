"""Latency of small batches through the host-buffer C ABI (BASELINE configs[1]: one dlsim subframe = 13 code blocks)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from openair4g_b200 import capi
import bench


def gen(K, seed, regime):
    """soft bits of one block from the product's own TX chain (clean / waterfall) or uniform noise"""
    if regime == "noise":
        return np.random.default_rng(seed).integers(-16, 17, size=3 * K + 12).astype(np.int16)
    return bench.coded_inputs(K, 1, {"clean": 0.5, "waterfall": 1.08}[regime], seed)[0][0]


capi.init_td16()
def run(K, n, regime, iters=6, reps=20, llr8=0):
    ys = [gen(K, 100 + i, regime) for i in range(min(n, 13))]
    pad = 3 * K + 12 + (36 if llr8 else 0)
    pin = capi.PinnedArray((n, 3 * K + 12), np.int16)
    for i in range(n): pin.array[i] = ys[i % len(ys)]
    blocks = [{"y": pin.array[i], "K": K, "max_iterations": iters, "crc_type": 1, "llr8": llr8} for i in range(n)]
    for _ in range(3): capi.decode_batch(blocks)
    t = []
    for _ in range(reps):
        t0 = time.perf_counter(); outs, st = capi.decode_batch(blocks); t.append(time.perf_counter() - t0)
    t.sort()
    print("K=%d n=%d %s llr8=%d: median %.3f ms  min %.3f ms  (status %s)" % (K, n, regime, llr8, 1e3 * t[len(t)//2], 1e3 * t[0], sorted(set(st))))
y = gen(6144, 1, "clean")
capi.phy_threegpplte_turbo_decoder16(y, 6144, 0, 0, 6, 1, 0)
t0 = time.perf_counter()
for _ in range(20): capi.phy_threegpplte_turbo_decoder16(y, 6144, 0, 0, 6, 1, 0)
print("single call K=6144 clean (2 iterations): %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
for regime in ("clean", "waterfall", "noise"):
    run(5824, 13, regime, iters=4)
    run(6144, 13, regime, iters=6)
run(6144, 13 * 8, "noise", iters=6)
run(6144, 13 * 64, "noise", iters=6)
run(3904, 2, "clean", iters=6)
run(5824, 13, "noise", iters=4, llr8=1)
