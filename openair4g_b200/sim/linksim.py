"""ulsim/dlsim-equivalent link-level harness with the GPU decoder switch and a throughput mode
(reference harnesses: openair1/SIMULATION/LTE_PHY/ulsim.c:881-1342,
dlsim.c:3352-3371).  Shapes follow the reference's tables:

  ulsim  25 PRB MCS16 : TBS  7736 -> C=2  x K=3904, G=14400, Qm=4   (BASELINE configs[0])
  dlsim 100 PRB MCS28 : TBS 75376 -> C=13 x K=5824, G=90000, Qm=6   (BASELINE configs[1])
  UL    100 PRB MCS16 : TBS 30576 -> C=5  x K=6144, G=57600, Qm=4   (multi-cell unit, configs[3])

The channel is AWGN per resource element with a max-log soft demapper and a fixed LLR scale
(SC-FDMA / OFDM processing, channel estimation and the UL channel interleaver are upstream
of the hot path and not modelled).  Every subframe goes to the GPU through ONE batched submit
with the fused front end (rate dematching + sub-block deinterleaving + turbo decoding);
HARQ rounds use rv 0,2,3,1 and combine in the caller-owned w buffers like the reference.  On the downlink the codeword
is scrambled (36.211 6.3.1) and the front end descrambles on the GPU (dlsch_unscrambling fused into rate dematching).
"""
import time
from dataclasses import dataclass

import numpy as np

from . import txchain as tx

RV_SEQ = (0, 2, 3, 1)


@dataclass(frozen=True)
class LinkConfig:
    name: str
    tbs: int
    G: int
    Qm: int
    downlink: bool
    Nl: int = 1
    Mdlharq: int = 8
    Kmimo: int = 1
    # downlink scrambling (36.211 6.3.1): c_init = (rnti<<14) + (q<<13) + ((Ns>>1)<<9) + Nid_cell; dlsim defaults:
    # n_rnti 0x1234, subframe 7, Nid_cell 0
    rnti: int = 0x1234
    subframe: int = 7
    nid_cell: int = 0

    @property
    def c_init(self):
        return (self.rnti << 14) + (self.subframe << 9) + self.nid_cell


ULSIM_25PRB_MCS16 = LinkConfig("ulsim 25 PRB MCS16", 7736, 25 * 12 * 12 * 4, 4, False)
DLSIM_100PRB_MCS28 = LinkConfig("dlsim 100 PRB MCS28 TM1", 75376, 90000, 6, True)
UL_100PRB_MCS16 = LinkConfig("UL 100 PRB MCS16", 30576, 100 * 12 * 12 * 4, 4, False)

_PAM = {2: (np.array([1.0, -1.0]) / np.sqrt(2), 1),
        4: (np.array([1.0, 3.0, -1.0, -3.0]) / np.sqrt(10), 2),      # index = b0*2 + b2 (36.211 7.1.3)
        6: (np.array([3.0, 1.0, 5.0, 7.0, -3.0, -1.0, -5.0, -7.0]) / np.sqrt(42), 3)}   # b0*4 + b2*2 + b4


def modulate(bits, Qm):
    """bits (n, E) -> complex symbols (n, E/Qm), 36.211 7.1 Gray mappings (I from even bits, Q from odd)."""
    levels, nb = _PAM[Qm]
    b = bits.reshape(bits.shape[0], -1, Qm).astype(np.int64)
    wi = sum(b[:, :, 2 * j] << (nb - 1 - j) for j in range(nb))
    wq = sum(b[:, :, 2 * j + 1] << (nb - 1 - j) for j in range(nb))
    return levels[wi] + 1j * levels[wq]


def demap_maxlog(sym, Qm, n0, llr_scale):
    """max-log LLR (positive = bit 1) per bit, scaled and rounded to int16, shape (n, Qm*nsym)."""
    levels, nb = _PAM[Qm]
    out = np.empty(sym.shape + (Qm,), dtype=np.float64)
    idx = np.arange(levels.size)
    for part, off in ((sym.real, 0), (sym.imag, 1)):
        d2 = (part[..., None] - levels[None, None, :]) ** 2            # distances to the PAM points
        for j in range(nb):
            bit = (idx >> (nb - 1 - j)) & 1
            l = (d2[..., bit == 0].min(-1) - d2[..., bit == 1].min(-1)) / n0
            out[..., 2 * j + off] = l
    llr = np.rint(out.reshape(sym.shape[0], -1) * llr_scale)
    return np.clip(llr, -32768, 32767).astype(np.int16)


CS_RI = ((1, 4, 7, 10), (0, 3, 5, 8))          # 36.212 table 5.2.2.8-1: RI columns, normal / extended cyclic prefix


def ul_multiplex(e, ctrl, sizes, Qm, c_init, rng, Cmux=12, Ncp=0):
    """Transmit-side mirror of the uplink front end (36.212 5.2.2.7-8 + 36.211 5.3.1 as the reference's receiver reads
    them, ulsch_decoding.c:600-1002): the data soft bits e (G values) are placed into the channel interleaver matrix behind
    the CQI symbols, `ctrl` soft values fill the CQI / RI / HARQ-ACK positions, and the matrix is read out column by column
    and scrambled -- the soft bits a demodulator would leave in lte_eNB_pusch_vars->llr.  Returns int16[Hpp*Qm]."""
    Hprime, Qri, Qcqi = sizes["Hprime"], sizes["Qprime_RI"], sizes["Qprime_CQI"]
    nsym = Hprime + Qri
    Rp = nsym // Cmux
    cs = CS_RI[1 if Ncp else 0]

    def is_ri(sym):
        r, col = divmod(sym, Cmux)
        if col not in cs:
            return False
        return 4 * (Rp - 1 - r) + ((4 - cs.index(col)) & 3) < Qri
    L = 0
    while L < nsym and is_ri(L):
        L += 1
    y = np.array(ctrl, dtype=np.int64).reshape(nsym, Qm).copy()
    y[L + Qcqi:L + Hprime] = np.asarray(e, dtype=np.int64).reshape(-1, Qm)
    sign = 2 * tx.gold_sequence(c_init, nsym * Qm).astype(np.int64).reshape(nsym, Qm) - 1     # by input symbol col*Rp + r
    sym = np.arange(nsym)
    src = (sym % Cmux) * Rp + sym // Cmux
    llr = np.zeros((nsym, Qm), dtype=np.int64)
    llr[src] = y * sign[src]
    return np.clip(llr.reshape(-1), -32767, 32767).astype(np.int16)


class LinkSim:
    def __init__(self, cfg: LinkConfig, max_iterations=4, llr8=0, llr_scale=4.0, seed=1, gpu_tx=False, ul_front=None, decoder_flags=0):
        """gpu_tx: encode / interleave / rate-match on the GPU (oai_turbo_tx_batch) instead of the numpy TX chain.
        ul_front (uplink only): dict(O_ACK, O_RI, Or1) -- every subframe carries that control information, the harness
        hands the GPU the multiplexed, scrambled soft bits of the whole allocation (oai_ul_front_t) and the transport
        block comes back assembled (oai_turbo_submit_tbs)."""
        self.cfg, self.max_it, self.llr8, self.scale, self.gpu_tx = cfg, max_iterations, llr8, llr_scale, gpu_tx
        self.ul_front = ul_front
        self.decoder_flags = decoder_flags               # e.g. capi.BATCH_SLIDING_WINDOW (optional mode, not bit-exact)
        self.rng = np.random.default_rng(seed)
        B = cfg.tbs + 24
        self.C, self.Cp, self.Cm, self.Kp, self.Km, self.F = tx.segmentation(B)
        assert self.Cm == 0, "mixed segment sizes are not exercised by the BASELINE shapes"
        self.K = self.Kp
        self.crc_type = 0 if self.C == 1 else 1
        self.G = cfg.G                                   # soft bits left for the data
        if ul_front is not None:
            assert not cfg.downlink
            from .. import capi
            nb_rb = cfg.G // (144 * cfg.Qm)
            rc, self.ul_sizes = capi.ulsch_control_sizes(ul_front.get("O_RI", 0), ul_front.get("O_ACK", 0), ul_front.get("Or1", 0),
                                                         12 * nb_rb, 12, 40, 40, 16, self.C * self.K, nb_rb, cfg.Qm, 12)
            assert rc == 0
            self.G = self.ul_sizes["G"]

    # ---- transmit n transport blocks; returns code blocks (n, C, K) and the coded bits (n*C, 3K+12) ----
    def make_blocks(self, n):
        cfg, C, K, F = self.cfg, self.C, self.K, self.F
        a = self.rng.integers(0, 2, size=(n, cfg.tbs)).astype(np.uint8)
        b = np.concatenate([a, tx.crc24a(a)], axis=1)                          # B = A + 24
        L = 24 if C > 1 else 0
        per = K - L
        cb = np.zeros((n, C, K), dtype=np.uint8)
        pos = 0
        for r in range(C):
            take = per - (F if r == 0 else 0)
            cb[:, r, (F if r == 0 else 0):per] = b[:, pos:pos + take]
            pos += take
        assert pos == b.shape[1]
        if C > 1:
            flat = cb.reshape(n * C, K)
            flat[:, per:] = tx.crc24b(flat[:, :per])
        if self.gpu_tx:
            return cb, cb                                                      # the GPU chain starts from the code blocks
        d = tx.turbo_encode(cb.reshape(n * C, K))
        return cb, d.reshape(n, C, 3 * K + 12)

    def transmit(self, d, rv, snr_db):
        """coded blocks (n, C, 3K+12) -> list over r of int16 soft bits e (n, E_r) after AWGN + demapping, and the
        position of each block's first bit in the codeword.  Downlink: the codeword is scrambled before modulation
        (dlsch_scrambling) and the soft bits are returned STILL SCRAMBLED in the reference's demodulator convention
        (positive = bit 0), i.e. what dlsch_unscrambling expects: it multiplies by 2c-1."""
        cfg = self.cfg
        n0 = 10.0 ** (-snr_db / 10.0)
        es, offs, off = [], [], 0
        c = tx.gold_sequence(cfg.c_init, cfg.G) if cfg.downlink else None
        if self.gpu_tx:
            from .. import capi
            packed = np.packbits(d, axis=2)
            sent = capi.tx_batch([{"c": packed[i, r], "K": self.K, "F": self.F if r == 0 else 0, "filler_null": 1,
                                   "G": self.G, "C": self.C, "r": r, "rvidx": rv, "Qm": cfg.Qm, "Nl": cfg.Nl,
                                   "Mdlharq": cfg.Mdlharq, "Kmimo": cfg.Kmimo}
                                  for r in range(self.C) for i in range(d.shape[0])])
        for r in range(self.C):
            if self.gpu_tx:
                bits = np.stack(sent[r * d.shape[0]:(r + 1) * d.shape[0]])
                E = bits.shape[1]
            else:
                bits, E = tx.rate_match(d[:, r], self.K, self.F if r == 0 else 0, self.G, self.C, cfg.Qm, cfg.Nl, r, rv,
                                        cfg.Mdlharq, cfg.Kmimo)
            if cfg.downlink:
                bits = bits ^ c[None, off:off + E]
            s = modulate(bits, cfg.Qm)
            s = s + np.sqrt(n0 / 2) * (self.rng.standard_normal(s.shape) + 1j * self.rng.standard_normal(s.shape))
            e = demap_maxlog(s, cfg.Qm, n0, self.scale)
            if cfg.downlink:
                e = (-np.clip(e, -32767, 32767)).astype(np.int16)
            es.append(e)
            offs.append(off)
            off += E
        return es, offs

    def run(self, snr_db, n_subframes, max_rounds=1, capi=None):
        """Monte-Carlo over n_subframes transport blocks (all submitted together each HARQ round).
        Returns per-round TB error counts, block error counts, iteration histogram, and GPU wall time."""
        if capi is None:
            from .. import capi as _c
            capi = _c
        cfg, C, K = self.cfg, self.C, self.K
        cb, d = self.make_blocks(n_subframes)
        Kpi = 32 * ((K + 4 + 31) // 32)
        w = np.zeros((n_subframes, C, 3 * Kpi), dtype=np.int16)                # caller-owned HARQ buffers
        alive = np.ones(n_subframes, dtype=bool)
        res = {"snr_db": snr_db, "tb_err": [], "cb_err": [], "iters": np.zeros(self.max_it + 2, dtype=np.int64),
               "gpu_s": 0.0, "decoded_info_bits": 0, "mismatch_vs_tx": 0, "iters_failed": 0}
        want = np.packbits(cb, axis=2)
        for rnd in range(max_rounds):
            idx = np.nonzero(alive)[0]
            if idx.size == 0:
                res["tb_err"].append(0)
                res["cb_err"].append(0)
                continue
            es, offs = self.transmit(d[idx], RV_SEQ[rnd % 4], snr_db)
            blocks, tbs = [], None
            if self.ul_front is not None:
                # the whole allocation's soft bits, multiplexed with control positions and scrambled, go to the GPU
                tbs = []
                z, uf = self.ul_sizes, self.ul_front
                for ii, sf in enumerate(idx):
                    nbits = (z["Hprime"] + z["Qprime_RI"]) * cfg.Qm
                    ctrl = self.rng.integers(-24, 25, size=nbits)
                    llr = ul_multiplex(np.concatenate([es[r][ii] for r in range(C)]), ctrl, z, cfg.Qm, cfg.c_init, self.rng)
                    tbs.append({"first_cb": ii * C, "C": C, "uplink": 1,
                                "ul_front": {"llr": llr, "c_init": cfg.c_init, "Qm": cfg.Qm, "Ncp": 0, "O_ACK": uf.get("O_ACK", 0),
                                             "O_RI": uf.get("O_RI", 0), "bundling": 0, "Nbundled": 1, "Cmux": 12,
                                             "Qprime_RI": z["Qprime_RI"], "Qprime_ACK": z["Qprime_ACK"],
                                             "Qprime_CQI": z["Qprime_CQI"], "Hprime": z["Hprime"]}})
            for ii, sf in enumerate(idx):
                for r in range(C):
                    blocks.append({"y": None if tbs is not None else np.ascontiguousarray(es[r][ii]), "K": K, "max_iterations": self.max_it,
                                   "crc_type": self.crc_type, "F": self.F if r == 0 else 0, "llr8": self.llr8,
                                   "tb_id": int(sf),
                                   "dematch": {"G": self.G, "C": C, "r": r, "rvidx": RV_SEQ[rnd % 4], "clear": 1 if rnd == 0 else 0,
                                               "Qm": cfg.Qm, "Nl": cfg.Nl, "Mdlharq": cfg.Mdlharq, "Kmimo": cfg.Kmimo,
                                               "w": w[sf, r],
                                               "scr_c_init": cfg.c_init if cfg.downlink else None, "scr_offset": offs[r]}})
            t0 = time.perf_counter()
            if tbs is None:
                outs, status = capi.decode_batch(blocks, flags=self.decoder_flags | (capi.BATCH_DL_STOP_AFTER_FAILURE if cfg.downlink else 0))
            else:
                outs, status, tbo = capi.decode_batch(blocks, tbs=tbs, flags=self.decoder_flags)
                res.setdefault("tb_results", []).append([(t[0], t[1]) for t in tbo])
            res["gpu_s"] += time.perf_counter() - t0
            st = np.array(status).reshape(idx.size, C)
            ok_cb = (st <= self.max_it)                                         # 0xFE (not decoded) counts as failed
            ok_tb = ok_cb.all(axis=1)
            res["iters_failed"] += int((st == self.max_it + 1).sum())
            for ii, sf in enumerate(idx):
                for r in range(C):
                    if ok_cb[ii, r]:
                        res["iters"][st[ii, r]] += 1
                        if not np.array_equal(outs[ii * C + r], want[sf, r]):
                            res["mismatch_vs_tx"] += 1                          # CRC passed on wrong data
            res["tb_err"].append(int((~ok_tb).sum()))
            res["cb_err"].append(int((~ok_cb).sum()))
            res["decoded_info_bits"] += int(ok_tb.sum()) * cfg.tbs
            alive[idx[ok_tb]] = False
        res["bler_round0"] = res["tb_err"][0] / n_subframes
        res["residual_bler"] = int(alive.sum()) / n_subframes
        # mean iterations over every block that was decoded: a block whose CRC never passed ran max_iterations
        # (blocks skipped by the downlink stop-after-failure rule ran none and are not counted); None without a decode
        res["iters"][self.max_it] += res["iters_failed"]
        res["avg_iterations"] = (float((res["iters"] * np.arange(res["iters"].size)).sum() / res["iters"].sum())
                                 if res["iters"].sum() else None)
        return res


def sweep(cfg, snrs, n_subframes=50, max_iterations=4, max_rounds=1, llr8=0, seed=1):
    sim = LinkSim(cfg, max_iterations=max_iterations, llr8=llr8, seed=seed)
    return [sim.run(s, n_subframes, max_rounds=max_rounds) for s in snrs]


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser(description="ulsim/dlsim-equivalent SNR sweep on the GPU decoder")
    ap.add_argument("--config", default="ulsim", choices=["ulsim", "dlsim", "ul100"])
    ap.add_argument("--snr", type=float, nargs=3, default=[6.0, 10.0, 0.5], metavar=("START", "STOP", "STEP"))
    ap.add_argument("-n", type=int, default=100, help="subframes per SNR point")
    ap.add_argument("-I", type=int, default=4, help="max turbo iterations (ulsim -I)")
    ap.add_argument("-L", action="store_true", help="8-bit decoder (ulsim/dlsim -L)")
    ap.add_argument("--rounds", type=int, default=1)
    a = ap.parse_args()
    cfg = {"ulsim": ULSIM_25PRB_MCS16, "dlsim": DLSIM_100PRB_MCS28, "ul100": UL_100PRB_MCS16}[a.config]
    for r in sweep(cfg, np.arange(a.snr[0], a.snr[1] + 1e-9, a.snr[2]), a.n, a.I, a.rounds, 1 if a.L else 0):
        r["iters"] = r["iters"].tolist()
        print(json.dumps(r))
