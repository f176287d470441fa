#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the COMPILED REFERENCE
(oracle/_ref/libref_oai.so = unmodified /root/reference sources)
on seeded inputs.  Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The fixtures carry inputs AND reference outputs, so the GPU box (no /root/reference)
checks the oracle port and the CUDA path against them without regenerating anything.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import loader, vectors  # noqa: E402


def td16_cases():
    cases = []
    blk = 0
    for K in (40, 48, 64, 104, 256, 504, 512, 1024, 2048, 3904, 5824, 6144):
        for regime, A in (("clean", 8), ("waterfall", 8), ("noise", 8), ("full", 8), ("waterfall", 2000)):
            for crc in ((0, 1) if K in (40, 512, 6144) else (1,)):
                cases.append((K, blk, regime, A, crc, 6, 0))
                blk += 1
    cases += [(512, 900, "clean", 8, 0, 6, 16), (6144, 901, "clean", 8, 0, 6, 56),
              (1024, 902, "waterfall", 8, 1, 4, 0), (1024, 903, "clean", 8, 1, 1, 0),
              (1024, 904, "clean", 8, 2, 3, 0), (1024, 905, "clean", 8, 3, 3, 0)]
    return cases


def main():
    R = loader.ref()
    assert R is not None, "compiled reference not available"
    ys, outs, meta = [], [], []
    for (K, blk, regime, A, crc, max_it, F) in td16_cases():
        y, _ = vectors.llr_block(K, blk, regime, A=A, crc_type=min(crc, 1), F=F)
        b, r = loader.ref_decode16(y, K, max_it, crc, F)
        ys.append(y)
        outs.append(b)
        meta.append((K, max_it, crc, F, r))
    np.savez_compressed(os.path.join(HERE, "td16_golden.npz"),
                        y=np.concatenate(ys), out=np.concatenate(outs),
                        meta=np.array(meta, dtype=np.int32))
    print("td16:", len(meta), "blocks; ret histogram",
          {int(k): int(v) for k, v in zip(*np.unique(np.array(meta)[:, 4], return_counts=True))})

    # 8-bit decoder (parity domain K >= 256, K % 16 == 0); amplitudes cover every input-scaling bracket
    ys, outs, meta = [], [], []
    blk = 2000
    for K in (256, 512, 1024, 2048, 5824, 6144):
        for regime, A in (("clean", 8), ("waterfall", 8), ("noise", 8), ("full", 8), ("waterfall", 40),
                          ("waterfall", 100), ("clean", 300), ("waterfall", 2000)):
            crc = blk & 1
            y, _ = vectors.llr_block(K, blk, regime, A=A, crc_type=crc)
            b, r = loader.ref_decode16(y, K, 6, crc, 0, which=8)
            ys.append(y); outs.append(b); meta.append((K, 6, crc, 0, r))
            blk += 1
    np.savez_compressed(os.path.join(HERE, "td8_golden.npz"), y=np.concatenate(ys), out=np.concatenate(outs),
                        meta=np.array(meta, dtype=np.int32))
    print("td8:", len(meta), "blocks; ret histogram",
          {int(k): int(v) for k, v in zip(*np.unique(np.array(meta)[:, 4], return_counts=True))})

    # rate dematching + sub-block deinterleaving chain (dlsim / ulsim shapes + HARQ rounds)
    rng = np.random.default_rng(0xD15)
    rm = {}
    shapes = [("dlsim100_mcs28", 5824, 0, 90000, 13, 6, 1, (0, 1, 12)),
              ("ulsim25_mcs16", 3904, 0, 14400, 2, 4, 1, (0, 1)),
              ("ul100_mcs16", 6144, 0, 57600, 5, 4, 1, (0, 4)),
              ("small_filler", 104, 16, 600, 1, 2, 1, (0,))]
    for name, K, F, G, Cb, Qm, Nl, rs in shapes:
        D = K + 4
        RTC = (D + 31) // 32
        Kpi = 32 * RTC
        for r in rs:
            dw = np.zeros(3 * Kpi, dtype=np.uint8)
            R.ref_generate_dummy_w(D, dw, F if r == 0 else 0)
            w = np.zeros(3 * Kpi, dtype=np.int16)
            E = C.c_uint32(0)
            es, ws, ds = [], [], []
            for rnd, rv in enumerate((0, 2, 1)):
                e = rng.integers(-300, 301, size=G // Cb + 64).astype(np.int16)
                rc = R.ref_lte_rate_matching_turbo_rx(RTC, G, w, dw, e, Cb, 1827072, 8, 1, rv, 1 if rnd == 0 else 0,
                                                      Qm, Nl, r, C.byref(E))
                assert rc == 0
                d = np.zeros(96 + 3 * D + 16, dtype=np.int16)
                R.ref_sub_block_deinterleaving_turbo(D, d.ctypes.data + 96 * 2, w)
                es.append(e[:E.value].copy())
                ws.append(w.copy())
            ds.append(d.copy())            # d after the last round only
            key = "%s_r%d" % (name, r)
            rm[key + "_e"] = np.stack(es)
            rm[key + "_w"] = np.stack(ws)
            rm[key + "_d"] = np.stack(ds)
            rm[key + "_dummy"] = dw
            rm[key + "_par"] = np.array([K, F, G, Cb, Qm, Nl, r, E.value, RTC], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "rm_golden.npz"), **rm)
    print("rm:", len(rm) // 5, "cases")


if __name__ == "__main__":
    main()
