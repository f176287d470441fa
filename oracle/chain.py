"""TEST INFRASTRUCTURE -- the receive chain of ONE transport block, restated on top of the oracle port:
the per-code-block loops and the transport-block reassembly of

  * dlsch_decoding  (openair1/PHY/LTE_TRANSPORT/dlsch_decoding.c:303-453 loop with the err_flag rule,
                     :455-483 return value, :486-512 reassembly -- only when every block passed)
  * ulsch_decoding  (openair1/PHY/LTE_TRANSPORT/ulsch_decoding.c:1222-1369 loops, :1380-1409 reassembly:
                     failed blocks are SKIPPED without advancing the offset, ret = last passing status unless a
                     block failed)

plus a generator of soft bits for a random transport block (input generation only; it uses the numpy TX chain of
openair4g_b200/sim, which is itself checked against the oracle's TX functions).  Only tests/, smoke() and bench.py's CPU
legs may import this.
"""
import ctypes as C

import numpy as np

from . import loader

NSOFT = 1827072


def segmentation(B):
    P = loader.port()
    v = [C.c_uint32(0) for _ in range(6)]
    rc = P.orc_lte_segmentation(B, *[C.byref(x) for x in v])
    assert rc == 0
    Cn, Cp, Cm, Kp, Km, F = [int(x.value) for x in v]
    return Cn, Cp, Cm, Kp, Km, F


def block_sizes(seg):
    Cn, Cp, Cm, Kp, Km, F = seg
    return [Km if r < Cm else Kp for r in range(Cn)]


def make_tb(tbs, G, Qm, seed, A=8, sigma_over_A=0.5, rv=0, Nl=1, Mdlharq=8, Kmimo=1, noise_blocks=()):
    """Random transport block of `tbs` bits -> dict with the soft bits e (int16[G], positive = bit 1) of one
    transmission, the transmitted code blocks and b (TB + CRC24A bytes).  Blocks listed in noise_blocks carry pure
    noise (they fail their CRC)."""
    from openair4g_b200.sim import txchain as tx
    rng = np.random.default_rng([0xC4A1, tbs, seed])
    a = rng.integers(0, 2, size=(1, tbs)).astype(np.uint8)
    b = np.concatenate([a, tx.crc24a(a)], axis=1)
    seg = segmentation(tbs + 24)
    Cn, Cp, Cm, Kp, Km, F = seg
    Ks = block_sizes(seg)
    L = 24 if Cn > 1 else 0
    cbs, pos = [], 0
    for r, K in enumerate(Ks):                               # lte_segmentation.c:136-170
        cb = np.zeros((1, K), dtype=np.uint8)
        f = F if r == 0 else 0
        take = K - L - f
        cb[:, f:K - L] = b[:, pos:pos + take]
        pos += take
        if Cn > 1:
            cb[:, K - 24:] = tx.crc24b(cb[:, :K - 24])
        cbs.append(cb)
    assert pos == b.shape[1]
    es = []
    for r, (K, cb) in enumerate(zip(Ks, cbs)):
        d = tx.turbo_encode(cb)
        bits, E = tx.rate_match(d, K, F if r == 0 else 0, G, Cn, Qm, Nl, r, rv, Mdlharq, Kmimo)
        bits = bits[0].astype(np.int64)
        if r in noise_blocks:
            e = rng.integers(-A, A + 1, size=E)
        else:
            e = A * (2 * bits - 1) + np.rint(sigma_over_A * A * rng.standard_normal(E)).astype(np.int64)
        es.append(np.clip(e, -32768, 32767).astype(np.int16))
    return {"tbs": tbs, "seg": seg, "Ks": Ks, "G": G, "Qm": Qm, "Nl": Nl, "Mdlharq": Mdlharq, "Kmimo": Kmimo, "rv": rv,
            "e": np.concatenate(es), "E": [x.size for x in es], "cb": [np.packbits(c[0]) for c in cbs],
            "b": np.packbits(b[0])}


def _front_end(tb, r, w, clear, e_slice):
    """generate_dummy_w + lte_rate_matching_turbo_rx + sub_block_deinterleaving_turbo of block r; returns y"""
    P = loader.port()
    K = tb["Ks"][r]
    F = tb["seg"][5] if r == 0 else 0
    D = K + 4
    RTC = (D + 31) // 32
    dw = np.zeros(3 * 32 * RTC, dtype=np.uint8)
    assert P.orc_generate_dummy_w(D, dw, F) == RTC
    E = C.c_uint32(0)
    rc = P.orc_lte_rate_matching_turbo_rx(RTC, tb["G"], w, dw, np.ascontiguousarray(e_slice), tb["seg"][0], NSOFT,
                                          tb["Mdlharq"], tb["Kmimo"], tb["rv"], clear, tb["Qm"], tb["Nl"], r, C.byref(E))
    assert rc == 0 and E.value == e_slice.size
    d = np.zeros(96 + 3 * D + 16, dtype=np.int16)
    P.orc_sub_block_deinterleaving_turbo(D, d.ctypes.data + 96 * 2, w)
    return d[96:96 + 3 * K + 12].copy()


def rx_tb(tb, max_it, downlink, llr8=0, w=None, clear=1, e=None, dec=None):
    """The reference's receive chain for one transport block.  Returns dict(c=[bytes per block], status=[per block or
    None when not decoded], ret, b (bytes or None), w=[HARQ buffers])."""
    seg, Ks = tb["seg"], tb["Ks"]
    Cn, F = seg[0], seg[5]
    e = tb["e"] if e is None else e
    crc_type = 0 if Cn == 1 else 1
    if dec is None:                                          # (tests of the optional sliding-window mode pass its model)
        dec = loader.port_decode8 if llr8 else loader.port_decode16
    if w is None:
        w = [np.zeros(3 * 32 * ((K + 4 + 31) // 32), dtype=np.int16) for K in Ks]
    c, status, off, err = [], [], 0, False
    ys = []
    for r, K in enumerate(Ks):
        ys.append(_front_end(tb, r, w[r], clear, e[off:off + tb["E"][r]]))
        off += tb["E"][r]
    for r, K in enumerate(Ks):
        if downlink and err:                                 # dlsch_decoding.c:400,417: c[r] zeroed, block not decoded
            c.append(np.zeros(K // 8, dtype=np.uint8))
            status.append(None)
            continue
        by, ret = dec(ys[r], K, max_it, crc_type, F if r == 0 else 0)
        if max_it < 2:
            by = np.zeros(K // 8, dtype=np.uint8)
        c.append(by)
        status.append(ret)
        if ret >= 1 + max_it:
            err = True
    strip = 3 if Cn > 1 else 0
    if downlink:
        if err:
            return {"c": c, "status": status, "ret": 1 + max_it, "b": None, "w": w, "y": ys}
        parts = [c[0][F >> 3:Ks[0] // 8 - strip]] + [c[r][:Ks[r] // 8 - strip] for r in range(1, Cn)]
        return {"c": c, "status": status, "ret": status[-1], "b": np.concatenate(parts), "w": w, "y": ys}
    # uplink, ulsch_decoding.c:1380-1409
    nb = sum(K // 8 - strip for K in Ks) - (F >> 3)
    b = np.zeros(nb, dtype=np.uint8)
    offset, ret = 0, 1
    for r, K in enumerate(Ks):
        if status[r] != 1 + max_it:
            if r == 0:
                n0 = K // 8 - (F >> 3) - strip
                b[:n0] = c[0][F >> 3:F // 8 + n0]
                offset = n0
            else:
                n1 = K // 8 - strip
                b[offset:offset + n1] = c[r][:n1]
                offset += n1
            if ret != 1 + max_it:
                ret = status[r]
        else:
            ret = 1 + max_it
    return {"c": c, "status": status, "ret": ret, "b": b, "b_valid": offset, "w": w, "y": ys}
