#!/usr/bin/env python3
"""Generates tests/golden/ul_golden.npz by running the UNMODIFIED reference ulsch_decoding()
(openair1/PHY/LTE_TRANSPORT/ulsch_decoding.c compiled in place behind oracle/ref_tu/shim4, see oracle/Makefile) on seeded
PUSCH allocations: soft bits of a coded transport block multiplexed with HARQ-ACK / RI / CQI positions.  Run in the build
container (where /root/reference is mounted):

    python tests/golden/make_golden_ul.py

Per case the fixture holds the parameters, the demodulator soft bits llr and everything the reference produced: e, q_ACK,
q_RI, q (CQI soft bits), o_ACK, o_RI, the decoded code blocks c[r], the transport block b and the return value -- for up
to two HARQ rounds.  tests/test_golden.py checks the oracle port against it on any box, tests/test_gpu_ul_front.py the
CUDA path."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import chain, loader, ulgen  # noqa: E402

# nb_rb, mcs, TBS, Nsymb, O_ACK, O_RI, Or1, bundling, Nbundled, Ncp, max_iter, llr8, sigma/A, rounds (rv list)
CASES = [
    (25, 16, 7736, 12, 0, 0, 0, 0, 1, 0, 6, 0, 0.5, (0,)),          # ulsim 25 PRB MCS16, data only
    (25, 16, 7736, 12, 1, 0, 0, 0, 1, 0, 6, 0, 0.5, (0,)),          # + 1 HARQ-ACK bit
    (25, 16, 7736, 12, 2, 1, 20, 0, 1, 0, 6, 0, 0.5, (0,)),         # + 2 ACK bits, RI, CQI
    (25, 16, 7736, 12, 2, 1, 20, 0, 1, 0, 4, 0, 0.8, (0, 2)),       # fails in round 0, second round with rv 2
    (100, 16, 30576, 12, 1, 1, 40, 0, 1, 0, 6, 0, 0.6, (0,)),       # 100 PRB MCS16: 5 x K=6144
    (100, 26, 61664, 12, 2, 1, 40, 0, 1, 0, 4, 0, 0.3, (0,)),       # 64QAM
    (10, 5, 872, 12, 1, 1, 11, 0, 1, 0, 6, 0, 0.5, (0,)),           # QPSK, small
    (10, 5, 872, 12, 2, 0, 0, 1, 2, 0, 6, 0, 0.5, (0,)),            # ACK bundling
    (10, 5, 872, 12, 1, 0, 0, 1, 4, 0, 6, 0, 0.5, (0,)),            # 1-bit ACK with bundling
    (25, 16, 6200, 10, 2, 1, 20, 0, 1, 1, 6, 0, 0.5, (0,)),         # extended cyclic prefix (10 data symbols)
    (25, 16, 6200, 10, 1, 1, 20, 1, 3, 1, 6, 0, 0.5, (0,)),
    (1, 5, 72, 12, 2, 1, 0, 0, 1, 0, 6, 0, 0.4, (0,)),              # one PRB: control rows reach the top of the matrix
    (8, 16, 2216, 12, 1, 1, 0, 0, 1, 0, 6, 0, 0.5, (0,)),
    (25, 16, 7736, 12, 1, 0, 0, 0, 1, 0, 6, 1, 0.5, (0,)),          # 8-bit decoder
]


def main():
    assert loader.ref() is not None, "compiled reference not available"
    out = {}
    for ci, case in enumerate(CASES):
        (nb_rb, mcs, TBS, Nsymb, O_ACK, O_RI, Or1, bundling, Nbundled, Ncp, max_it, llr8, sig, rvs) = case
        par = ulgen.params(nb_rb, mcs, TBS, Nsymb, O_ACK, O_RI, Or1, bundling, Nbundled, Ncp, max_it, llr8)
        tb = None
        state = None
        for rnd, rv in enumerate(rvs):
            llr, tb = ulgen.make_llr(par, seed=100 + ci, rv=rv, sigma_over_A=sig, tb=tb)
            p = dict(par["ref"], rvidx=rv, round=rnd)
            r = loader.ref_ulsch_decoding(p, llr, state)
            state = r["state"]
            z = par["sizes"]
            ne = (z["Hprime"] - z["Qprime_CQI"]) * par["Qm"]
            key = "c%02d_r%d_" % (ci, rnd)
            out[key + "par"] = np.array(case[:12] + (rv, rnd), dtype=np.int32)
            out[key + "llr"] = llr
            out[key + "e"] = r["e"][:ne].copy()
            out[key + "qack"] = r["q_ACK"].copy()
            out[key + "qri"] = r["q_RI"].copy()
            out[key + "qcqi"] = r["q_cqi"][:z["Qprime_CQI"] * par["Qm"]].copy()
            out[key + "o"] = np.array([r["o_ACK"][0], r["o_ACK"][1], r["o_RI"][0]], dtype=np.uint8)
            Ks = par["Ks"]
            out[key + "c"] = np.concatenate([r["c"][i, :K // 8] for i, K in enumerate(Ks)])
            out[key + "b"] = r["b"][:(TBS + 24) // 8].copy()
            out[key + "ret"] = np.array([r["ret"]], dtype=np.int32)
            out[key + "txb"] = tb["b"]
            good = r["ret"] <= max_it and np.array_equal(r["b"][:(TBS + 24) // 8], tb["b"])
            print("case %2d round %d: ret %d  TB recovered %s  sizes %s" % (ci, rnd, r["ret"], good, {k: z[k] for k in ("Qprime_RI", "Qprime_ACK", "Qprime_CQI", "G")}))
    np.savez_compressed(os.path.join(HERE, "ul_golden.npz"), **out)
    print("wrote ul_golden.npz:", os.path.getsize(os.path.join(HERE, "ul_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
