"""TEST INFRASTRUCTURE -- input generation for the uplink front-end tests: the soft bits of one PUSCH allocation as the
demodulator would leave them (column by column of the channel interleaver matrix, scrambled), built by INVERTING the
receive-side mapping of ulsch_decoding.c:600-758 around the soft bits e of a coded transport block (oracle/chain.py):
data symbols carry e, the HARQ-ACK / RI / CQI positions carry seeded values.  Only tests/ and tests/golden/ import this."""
import ctypes as C

import numpy as np

from . import chain, loader

CS_RI = ((1, 4, 7, 10), (0, 3, 5, 8))
CS_ACK = ((2, 3, 8, 9), (1, 2, 6, 7))


def params(nb_rb, mcs, TBS, Nsymb, O_ACK, O_RI, Or1, bundling=0, Nbundled=1, Ncp=0, max_it=6, llr8=0, rnti=0x1234, subframe=3,
           Nid_cell=7, beta=(16, 40, 40)):
    """sizes (via the port's orc_ulsch_control_sizes, itself pinned on the compiled reference) + the reference driver's
    parameter dict"""
    P = loader.port()
    Qm = 2 if mcs < 11 else (4 if mcs < 21 else 6)
    seg = chain.segmentation(TBS + 24)
    Ks = chain.block_sizes(seg)
    z = loader.UlSizes()
    rc = P.orc_ulsch_control_sizes(O_RI, O_ACK, Or1, 12 * nb_rb, Nsymb, beta[1], beta[2], beta[0], sum(Ks), nb_rb, Qm, Nsymb, C.byref(z))
    assert rc == 0
    sizes = {k: int(getattr(z, k)) for k, _ in loader.UlSizes._fields_}
    ref = dict(TBS=TBS, nb_rb=nb_rb, Nsymb_pusch=Nsymb, Nsymb_initial=Nsymb, Msc_initial=12 * nb_rb, mcs=mcs, rvidx=0, round=0,
               O_ACK=O_ACK, O_RI=O_RI, Or1=Or1, bundling=bundling, Nbundled=Nbundled, Ncp=Ncp, beta_cqi_x8=beta[0],
               beta_ri_x8=beta[1], beta_ack_x8=beta[2], rnti=rnti, subframe=subframe, Nid_cell=Nid_cell, max_iter=max_it,
               Mdlharq=8, llr8=llr8)
    return {"Qm": Qm, "seg": seg, "Ks": Ks, "sizes": sizes, "z": z, "ref": ref, "TBS": TBS, "c_init": (rnti << 14) + (subframe << 9) + Nid_cell,
            "Ncp": Ncp, "O_ACK": O_ACK, "O_RI": O_RI, "Or1": Or1, "bundling": bundling, "Nbundled": Nbundled, "max_it": max_it,
            "llr8": llr8, "Cmux": Nsymb}


def _placeholder(r, col, n, cs, Rp):
    if col not in cs:
        return -1
    i = 4 * (Rp - 1 - r) + ((4 - cs.index(col)) & 3)
    return i if i < n else -1


def make_llr(par, seed, rv=0, sigma_over_A=0.5, A=8, tb=None):
    """Returns (llr int16[Hpp*Qm], tb): tb = chain.make_tb(...) of the transport block (re-used for further HARQ rounds)."""
    z, Qm = par["sizes"], par["Qm"]
    rng = np.random.default_rng([0x0151, seed, rv])
    if tb is None or tb["rv"] != rv:
        tb = chain.make_tb(par["TBS"], z["G"], Qm, seed=seed, A=A, sigma_over_A=sigma_over_A, rv=rv)
    e = tb["e"]
    assert e.size == z["G"], (e.size, z["G"])
    Rp, Cm = z["Rmux_prime"], z["Cmux"]
    nsym = Rp * Cm
    cs_ri = CS_RI[1 if par["Ncp"] else 0]
    # row-major symbol stream the receiver reads: [L RI symbols][CQI][data ...]; everything else random
    y = rng.integers(-3 * A, 3 * A + 1, size=(nsym, Qm)).astype(np.int64)
    L = 0
    while L < nsym and _placeholder(L // Cm, L % Cm, z["Qprime_RI"], cs_ri, Rp) >= 0:
        L += 1
    d0 = L + z["Qprime_CQI"]
    y[d0:d0 + z["Hprime"] - z["Qprime_CQI"]] = e.reshape(-1, Qm)
    words = np.zeros(nsym * Qm // 32 + 2, dtype=np.uint32)
    loader.port().orc_gold_words(par["c_init"], words.ctypes.data, words.size)
    bits = ((words[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).reshape(-1)[:nsym * Qm].astype(np.int64)
    sign = (2 * bits - 1).reshape(nsym, Qm)                          # indexed by INPUT symbol (col*Rp + r)
    sym = np.arange(nsym)
    r, col = sym // Cm, sym % Cm
    llr = np.zeros((nsym, Qm), dtype=np.int64)
    llr[col * Rp + r] = y * sign[col * Rp + r]                        # data symbols: y = sign * llr  <=>  llr = sign * y
    return np.clip(llr.reshape(-1), -32768, 32767).astype(np.int16), tb
