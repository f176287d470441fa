"""BLER-vs-SNR sweeps of the ulsim / dlsim shaped harness on the GPU decoder (BASELINE configs[0], [1]).  Every block of
these runs is bit-identical to the reference chain (tests/test_gpu_linksim.py re-decodes them with the oracle), so the curves
are the reference's curves for this channel model; avg_iterations counts a failed block as max_iterations."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openair4g_b200.sim import linksim
def show(title, rows):
    print("# " + title)
    print("%7s %10s %10s %10s %14s %12s" % ("SNR dB", "BLER rd0", "BLER rd1", "residual", "avg iterations", "mismatch/tx"))
    for r in rows:
        b1 = (r["tb_err"][1] / n) if len(r["tb_err"]) > 1 else float("nan")
        print("%7.2f %10.3f %10.3f %10.3f %14s %12d" % (r["snr_db"], r["bler_round0"], b1, r["residual_bler"],
              "-" if r["avg_iterations"] is None else "%.2f" % r["avg_iterations"], r["mismatch_vs_tx"]))
n = 100
sim = linksim.LinkSim(linksim.ULSIM_25PRB_MCS16, max_iterations=4, seed=1)
show("ulsim 25 PRB MCS16 (2 x K=3904), 4 iterations, %d subframes per point, 2 HARQ rounds" % n,
     [sim.run(s, n, max_rounds=2) for s in np.arange(5.0, 9.01, 0.5)])
sim = linksim.LinkSim(linksim.ULSIM_25PRB_MCS16, max_iterations=4, seed=1, ul_front={"O_ACK": 2, "O_RI": 1, "Or1": 20})
show("the same through the uplink front end (2 ACK bits + RI + 20 CQI bits multiplexed; G = %d instead of 14400)" % sim.G,
     [sim.run(s, n, max_rounds=2) for s in np.arange(5.0, 9.01, 0.5)])
sim = linksim.LinkSim(linksim.ULSIM_25PRB_MCS16, max_iterations=4, seed=1, llr8=1)
show("ulsim 25 PRB MCS16, 8-bit decoder (-L)", [sim.run(s, n, max_rounds=2) for s in np.arange(5.0, 9.01, 0.5)])
n = 30
sim = linksim.LinkSim(linksim.DLSIM_100PRB_MCS28, max_iterations=4, seed=2)
show("dlsim 100 PRB MCS28 TM1 (13 x K=5824), 4 iterations, %d subframes per point" % n,
     [sim.run(s, n) for s in np.arange(17.0, 22.01, 0.5)])
