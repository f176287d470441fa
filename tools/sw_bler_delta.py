"""BLER delta of the optional sliding-window mode against the default bit-exact mode on the ulsim / dlsim shaped SNR
sweeps (north_star: the mode "is allowed only if it is separately reported with its BLER delta against the bit-exact mode
on the ulsim SNR sweep").  Both modes decode THE SAME received subframes (same seed -> same data, same noise).
usage: python tools/sw_bler_delta.py [--quick]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openair4g_b200 import capi
from openair4g_b200.sim import linksim


def snr_at(snrs, bler, target):
    """SNR where the (monotonised) curve crosses `target`, by linear interpolation in log10(BLER); None if it does not"""
    b = np.maximum.accumulate(np.asarray(bler, dtype=float)[::-1])[::-1]
    for i in range(len(b) - 1):
        if b[i] >= target > b[i + 1]:
            lo, hi = np.log10(max(b[i], 1e-6)), np.log10(max(b[i + 1], 1e-6))
            return snrs[i] + (snrs[i + 1] - snrs[i]) * (lo - np.log10(target)) / (lo - hi)
    return None


def sweep(title, cfg, snrs, n, max_it, rounds=1, llr_scale=4.0):
    print("# %s, %d iterations max, %d subframes per point and mode (same received subframes in both modes)" % (title, max_it, n))
    print("%7s | %9s %9s %8s | %9s %9s %8s | %s" % ("SNR dB", "BLER", "residual", "avg it", "BLER", "residual", "avg it", "CRC pass on wrong data (exact/sw)"))
    print("%7s | %28s | %28s |" % ("", "bit-exact mode", "sliding-window mode"))
    rows = []
    for s in snrs:
        r = []
        for flags in (0, capi.BATCH_SLIDING_WINDOW):
            sim = linksim.LinkSim(cfg, max_iterations=max_it, seed=int(round(s * 100)) + 7, decoder_flags=flags, llr_scale=llr_scale)
            r.append(sim.run(float(s), n, max_rounds=rounds))
        rows.append(r)
        print("%7.2f | %9.4f %9.4f %8.2f | %9.4f %9.4f %8.2f | %d/%d" % (
            s, r[0]["bler_round0"], r[0]["residual_bler"], r[0]["avg_iterations"] or 0, r[1]["bler_round0"], r[1]["residual_bler"],
            r[1]["avg_iterations"] or 0, r[0]["mismatch_vs_tx"], r[1]["mismatch_vs_tx"]), flush=True)
    for tgt in (0.1, 0.01):
        a = snr_at(list(snrs), [r[0]["bler_round0"] for r in rows], tgt)
        b = snr_at(list(snrs), [r[1]["bler_round0"] for r in rows], tgt)
        if a is not None and b is not None:
            print("SNR at BLER %.0f%%: bit-exact %.2f dB, sliding-window %.2f dB, delta %+.2f dB" % (100 * tgt, a, b, b - a))
    print()


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    t0 = time.time()
    capi.init_td16()
    n = 100 if quick else 1000
    sweep("ulsim 25 PRB MCS16 (2 x K=3904, 61-step windows)", linksim.ULSIM_25PRB_MCS16, np.arange(5.5, 8.01, 0.25), n, 4)
    sweep("ulsim 25 PRB MCS16", linksim.ULSIM_25PRB_MCS16, np.arange(5.5, 8.01, 0.25), n, 6)
    sweep("ulsim 25 PRB MCS16, 2 HARQ rounds", linksim.ULSIM_25PRB_MCS16, np.arange(3.0, 6.01, 0.5), n // 2, 4, rounds=2)
    sweep("dlsim 100 PRB MCS28 TM1 (13 x K=5824, 91-step windows, code rate 0.84)", linksim.DLSIM_100PRB_MCS28, np.arange(18.0, 21.01, 0.25),
          20 if quick else 200, 4)
    sweep("dlsim 100 PRB MCS28 TM1", linksim.DLSIM_100PRB_MCS28, np.arange(18.0, 21.01, 0.25), 20 if quick else 200, 6)
    # the bit-exact mode (= the reference's algorithm) keeps losing blocks at every SNR of the MCS28 sweep.  Not an effect
    # of the soft-bit scale (this harness hands over 4 x LLR): the same sweep with soft bits 8 times smaller
    sweep("dlsim 100 PRB MCS28 TM1, soft bits = 0.5 x LLR instead of 4 x", linksim.DLSIM_100PRB_MCS28, np.arange(18.0, 21.01, 0.25),
          20 if quick else 200, 6, llr_scale=0.5)
    print("# wall time %.0f s" % (time.time() - t0))
