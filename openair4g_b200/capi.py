"""ctypes mirror of include/oai_turbo_b200.h -- same names, argument meaning and return
codes as the reference's C entry points (openair1/PHY/CODING/defs.h:132,152,239-253,362,
367,470-484,499-513), plus the batched and device-resident calls.

No CPU fallback: if the CUDA library is missing this module raises at import.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

CRC24_A, CRC24_B, CRC16, CRC8 = 0, 1, 2, 3
LTE_NULL = 2
STATUS_NOT_DECODED = 0xFE
BATCH_DL_STOP_AFTER_FAILURE = 1
BATCH_SLIDING_WINDOW = 2          # optional, NOT bit-exact (include/oai_turbo_b200.h)

if not os.path.exists(_build.LIB):
    raise ImportError(
        "openair4g_b200: CUDA library %s is not built (run `python -m openair4g_b200.build` or "
        "__graft_entry__.build()); there is no CPU fallback" % _build.LIB)

lib = C.CDLL(_build.LIB)
LIB_PATH = _build.LIB

EXPORTS = ["init_td16", "free_td16", "init_td8", "free_td8", "phy_threegpplte_turbo_decoder16",
           "phy_threegpplte_turbo_decoder8", "generate_dummy_w", "lte_rate_matching_turbo_rx",
           "sub_block_deinterleaving_turbo", "oai_turbo_submit_batch", "oai_turbo_submit_tbs", "oai_turbo_wait", "oai_ulsch_control_sizes",
           "oai_turbo_dev_plan_create", "oai_turbo_dev_decode", "oai_turbo_dev_plan_destroy", "oai_turbo_dev_plan_set_mode",
           "oai_turbo_dev_plan_profile", "oai_turbo_host_alloc", "oai_turbo_host_free", "oai_lte_segmentation_params",
           "oai_turbo_harq_pool_create", "oai_turbo_harq_pool_read", "oai_turbo_harq_pool_destroy",
           "oai_turbo_b200_version", "oai_turbo_b200_last_error", "oai_turbo_b200_launch_count",
           "oai_turbo_debug_map16", "threegpplte_turbo_encoder", "sub_block_interleaving_turbo",
           "lte_rate_matching_turbo", "oai_turbo_tx_batch"]


class CbDesc(C.Structure):
    """oai_cb_desc_t"""
    _fields_ = [("in_", C.c_void_p), ("decoded_bytes", C.c_void_p), ("status", C.c_void_p),
                ("K", C.c_uint16), ("max_iterations", C.c_uint8), ("crc_type", C.c_uint8),
                ("F", C.c_uint8), ("llr8", C.c_uint8), ("decode_enable", C.c_uint8),
                ("dematch_enable", C.c_uint8), ("w", C.c_void_p), ("G", C.c_uint32),
                ("Nsoft", C.c_uint32), ("C", C.c_uint8), ("r", C.c_uint8), ("rvidx", C.c_uint8),
                ("clear", C.c_uint8), ("Qm", C.c_uint8), ("Nl", C.c_uint8), ("Mdlharq", C.c_uint8),
                ("Kmimo", C.c_uint8), ("tb_id", C.c_uint32), ("harq_pool", C.c_void_p), ("harq_slot", C.c_uint32),
                ("scr_c_init", C.c_uint32), ("scr_offset", C.c_uint32), ("scr_enable", C.c_uint8),
                ("in_fmt", C.c_uint8)]


class UlFront(C.Structure):
    """oai_ul_front_t"""
    _fields_ = [("llr", C.c_void_p), ("llr_fmt", C.c_uint8), ("c_init", C.c_uint32), ("Qm", C.c_uint8), ("Ncp", C.c_uint8),
                ("O_ACK", C.c_uint8), ("O_RI", C.c_uint8), ("bundling", C.c_uint8), ("Nbundled", C.c_uint8), ("Cmux", C.c_uint16),
                ("Qprime_RI", C.c_uint32), ("Qprime_ACK", C.c_uint32), ("Qprime_CQI", C.c_uint32), ("Hprime", C.c_uint32),
                ("q_ACK", C.c_void_p), ("q_RI", C.c_void_p), ("q_cqi", C.c_void_p), ("o_ACK", C.c_void_p), ("o_RI", C.c_void_p)]


class TbDesc(C.Structure):
    """oai_tb_desc_t"""
    _fields_ = [("first_cb", C.c_uint32), ("C", C.c_uint32), ("b", C.c_void_p), ("b_capacity", C.c_uint32),
                ("ret", C.c_void_p), ("valid_bytes", C.c_void_p), ("uplink", C.c_uint8), ("ul_front", C.POINTER(UlFront))]


class TxDesc(C.Structure):
    """oai_tx_desc_t"""
    _fields_ = [("c", C.c_void_p), ("e", C.c_void_p), ("G", C.c_uint32), ("Nsoft", C.c_uint32), ("E", C.c_uint32),
                ("K", C.c_uint16), ("F", C.c_uint8), ("filler_null", C.c_uint8), ("C", C.c_uint8),
                ("Mdlharq", C.c_uint8), ("Kmimo", C.c_uint8), ("rvidx", C.c_uint8), ("Qm", C.c_uint8),
                ("Nl", C.c_uint8), ("r", C.c_uint8), ("reserved", C.c_uint8)]


TX_DEVICE_POINTERS = 1

_stats7 = [C.c_void_p] * 7
for _f in ("phy_threegpplte_turbo_decoder16", "phy_threegpplte_turbo_decoder8"):
    getattr(lib, _f).argtypes = [C.c_void_p, C.c_void_p, C.c_uint16, C.c_uint16, C.c_uint16, C.c_uint8,
                                 C.c_uint8, C.c_uint8] + _stats7
    getattr(lib, _f).restype = C.c_uint8
lib.generate_dummy_w.argtypes = [C.c_uint32, C.c_void_p, C.c_uint8]
lib.generate_dummy_w.restype = C.c_uint32
lib.lte_rate_matching_turbo_rx.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint8,
                                           C.c_uint32] + [C.c_uint8] * 7 + [C.POINTER(C.c_uint32)]
lib.lte_rate_matching_turbo_rx.restype = C.c_int
lib.sub_block_deinterleaving_turbo.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
lib.sub_block_deinterleaving_turbo.restype = None
lib.threegpplte_turbo_encoder.argtypes = [C.c_void_p, C.c_uint16, C.c_void_p, C.c_uint8, C.c_uint16, C.c_uint16]
lib.threegpplte_turbo_encoder.restype = None
lib.sub_block_interleaving_turbo.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
lib.sub_block_interleaving_turbo.restype = C.c_uint32
lib.lte_rate_matching_turbo.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint8, C.c_uint32] + [C.c_uint8] * 8
lib.lte_rate_matching_turbo.restype = C.c_uint32
lib.oai_turbo_tx_batch.argtypes = [C.POINTER(TxDesc), C.c_int, C.c_uint, C.c_int]
lib.oai_turbo_tx_batch.restype = C.c_int
lib.oai_turbo_submit_batch.argtypes = [C.POINTER(CbDesc), C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_void_p)]
lib.oai_turbo_submit_tbs.argtypes = [C.POINTER(CbDesc), C.c_int, C.POINTER(TbDesc), C.c_int, C.c_uint, C.c_int, C.POINTER(C.c_void_p)]
lib.oai_turbo_wait.argtypes = [C.c_void_p]
lib.oai_ulsch_control_sizes.argtypes = [C.c_uint32] * 12 + [C.POINTER(C.c_uint32)] * 6
lib.oai_lte_segmentation_params.argtypes = [C.c_uint32] + [C.POINTER(C.c_uint32)] * 6
lib.oai_lte_segmentation_params.restype = C.c_int
lib.oai_turbo_harq_pool_create.argtypes = [C.c_int, C.c_uint32, C.c_uint16, C.POINTER(C.c_void_p)]
lib.oai_turbo_harq_pool_read.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
lib.oai_turbo_harq_pool_destroy.argtypes = [C.c_void_p]
lib.oai_turbo_harq_pool_destroy.restype = None
lib.oai_turbo_host_alloc.argtypes = [C.c_size_t]
lib.oai_turbo_host_alloc.restype = C.c_void_p
lib.oai_turbo_host_free.argtypes = [C.c_void_p]
lib.oai_turbo_host_free.restype = None
lib.oai_turbo_dev_plan_create.argtypes = [C.c_int, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8, C.POINTER(C.c_void_p)]
lib.oai_turbo_dev_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
lib.oai_turbo_dev_plan_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
lib.oai_turbo_dev_plan_set_mode.argtypes = [C.c_void_p, C.c_uint]
lib.oai_turbo_dev_plan_destroy.argtypes = [C.c_void_p]
lib.oai_turbo_dev_plan_destroy.restype = None
lib.oai_turbo_debug_map16.argtypes = [C.c_void_p, C.c_uint16, C.c_int, C.c_int, C.c_void_p]
lib.oai_turbo_b200_version.restype = C.c_char_p
lib.oai_turbo_b200_last_error.restype = C.c_char_p
lib.oai_turbo_b200_launch_count.restype = C.c_ulonglong


def last_error():
    return lib.oai_turbo_b200_last_error().decode()


def launch_count():
    return int(lib.oai_turbo_b200_launch_count())


def init_td16():
    lib.init_td16()


def init_td8():
    lib.init_td8()


def _decode_one(fn, y, n, f1, f2, max_iterations, crc_type, F, decoded_bytes):
    y = np.ascontiguousarray(y, dtype=np.int16)
    assert y.size >= 3 * n + 12
    if decoded_bytes is None:
        decoded_bytes = np.zeros(n // 8 + 4, dtype=np.uint8)
    ret = fn(y.ctypes.data, decoded_bytes.ctypes.data, n, f1, f2, max_iterations, crc_type, F, *([None] * 7))
    return int(ret), decoded_bytes[: n // 8]


def phy_threegpplte_turbo_decoder16(y, n, f1=0, f2=0, max_iterations=4, crc_type=CRC24_B, F=0, decoded_bytes=None):
    """Returns (ret, decoded_bytes[n/8]); ret as the reference: iterations used, max+1 on
    failure, 255 on illegal arguments."""
    return _decode_one(lib.phy_threegpplte_turbo_decoder16, y, n, f1, f2, max_iterations, crc_type, F, decoded_bytes)


def phy_threegpplte_turbo_decoder8(y, n, f1=0, f2=0, max_iterations=4, crc_type=CRC24_B, F=0, decoded_bytes=None):
    return _decode_one(lib.phy_threegpplte_turbo_decoder8, y, n, f1, f2, max_iterations, crc_type, F, decoded_bytes)


def generate_dummy_w(D, w, F):
    """Marks the NULL positions in the uint8 array w (3*Kpi entries, modified in place); returns RTC."""
    assert w.dtype == np.uint8 and w.flags["C_CONTIGUOUS"]
    return int(lib.generate_dummy_w(D, w.ctypes.data, F))


def lte_rate_matching_turbo_rx(RTC, G, w, dummy_w, soft_input, C_, Nsoft, Mdlharq, Kmimo, rvidx, clear, Qm, Nl, r):
    """Reference call shape; w (int16, in place).  Returns (rc, E)."""
    assert w.dtype == np.int16 and dummy_w.dtype == np.uint8 and soft_input.dtype == np.int16
    E = C.c_uint32(0)
    rc = lib.lte_rate_matching_turbo_rx(RTC, G, w.ctypes.data, dummy_w.ctypes.data, soft_input.ctypes.data, C_, Nsoft,
                                        Mdlharq, Kmimo, rvidx, clear, Qm, Nl, r, C.byref(E))
    return int(rc), int(E.value)


def sub_block_deinterleaving_turbo(D, d_buf, d_offset, w):
    """Writes into d_buf (int16) around element d_offset exactly what the reference writes around its `d`."""
    assert d_buf.dtype == np.int16 and w.dtype == np.int16
    lib.sub_block_deinterleaving_turbo(D, d_buf.ctypes.data + 2 * d_offset, w.ctypes.data)


def threegpplte_turbo_encoder(input_bytes, F=0, f1=0, f2=0):
    """Reference call shape: K/8 info bytes -> uint8[3K+12] coded bits (one per byte)."""
    inp = np.ascontiguousarray(input_bytes, dtype=np.uint8)
    out = np.full(3 * 8 * inp.size + 12, 255, dtype=np.uint8)
    lib.threegpplte_turbo_encoder(inp.ctypes.data, inp.size, out.ctypes.data, F, f1, f2)
    return out


def sub_block_interleaving_turbo(D, d_buf, d_offset, w):
    """d_buf (uint8) holds the reference's d at element d_offset (>= 3*ND bytes in front of it are read, element
    3*D+2 behind it is written); w (uint8[3*Kpi]) receives the three interleaved sub-blocks.  Returns RTC."""
    assert d_buf.dtype == np.uint8 and w.dtype == np.uint8
    return int(lib.sub_block_interleaving_turbo(D, d_buf.ctypes.data + d_offset, w.ctypes.data))


def lte_rate_matching_turbo(RTC, G, w, e, C_, Nsoft, Mdlharq, Kmimo, rvidx, Qm, Nl, r, nb_rb=0, m=0):
    """Reference call shape; writes e (uint8) and returns E."""
    assert w.dtype == np.uint8 and e.dtype == np.uint8
    return int(lib.lte_rate_matching_turbo(RTC, G, w.ctypes.data, e.ctypes.data, C_, Nsoft, Mdlharq, Kmimo, rvidx, Qm, Nl,
                                           r, nb_rb, m))


def tx_batch(blocks, gpu=-1):
    """blocks: list of dicts with c (uint8[K/8]), K, G, C, r, rvidx, Qm and optionally F, filler_null, Nsoft, Mdlharq,
    Kmimo, Nl.  Returns the list of e arrays (uint8, one bit per byte) produced by oai_turbo_tx_batch."""
    n = len(blocks)
    descs = (TxDesc * n)()
    keep, outs = [], []
    for i, b in enumerate(blocks):
        c = np.ascontiguousarray(b["c"], dtype=np.uint8)
        e = np.full(int(b.get("e_cap", b["G"])) + 8, 255, dtype=np.uint8)
        keep.append(c)
        outs.append(e)
        d = descs[i]
        d.c = c.ctypes.data; d.e = e.ctypes.data
        d.K = b["K"]; d.F = b.get("F", 0); d.filler_null = b.get("filler_null", 0)
        d.G = b["G"]; d.Nsoft = b.get("Nsoft", 1827072); d.C = b.get("C", 1); d.Mdlharq = b.get("Mdlharq", 8)
        d.Kmimo = b.get("Kmimo", 1); d.rvidx = b.get("rvidx", 0); d.Qm = b.get("Qm", 2); d.Nl = b.get("Nl", 1)
        d.r = b.get("r", 0)
    rc = lib.oai_turbo_tx_batch(descs, n, 0, gpu)
    if rc:
        raise RuntimeError("oai_turbo_tx_batch failed (%d): %s" % (rc, last_error()))
    return [outs[i][:descs[i].E].copy() for i in range(n)]


def decode_batch(blocks, flags=0, gpu=-1, tbs=None, cb_out=True):
    """blocks: list of dicts {y, K, max_iterations, crc_type, F=0, tb_id=0, llr8=0, decode_enable=1}.
    tbs: optional list of dicts {first_cb, C, uplink} -> oai_turbo_submit_tbs; the function then returns a third value,
    the list of (ret, valid_bytes, b) per transport block (b: uint8 array of the bytes the GPU assembled; a 0xA5 fill
    pattern shows what was left untouched).  cb_out=False leaves decoded_bytes NULL (transport-block outputs only).
    With "dematch": {G, C, r, rvidx, clear, Qm, Nl=1, Mdlharq=8, Kmimo=1, Nsoft=1827072, w=int16 array or None}
    `y` is the block's slice of rate-matched soft bits e and the front end runs on the GPU first.
    One submit + wait; returns (list of uint8 arrays, list of status ints)."""
    n = len(blocks)
    descs = (CbDesc * n)()
    keep, outs = [], []
    status = np.full(n, 255, dtype=np.uint8)
    for i, b in enumerate(blocks):
        y = None if b.get("y") is None else np.ascontiguousarray(b["y"], dtype=np.int8 if b.get("in_fmt") else np.int16)
        out = np.zeros(b["K"] // 8 + 4, dtype=np.uint8)
        keep.append(y)
        outs.append(out)
        d = descs[i]
        d.in_ = None if y is None else y.ctypes.data       # None: the transport block's ul_front supplies the soft bits
        d.decoded_bytes = out.ctypes.data if cb_out else None
        d.status = status.ctypes.data + i
        d.K = b["K"]
        d.max_iterations = b["max_iterations"]
        d.crc_type = b["crc_type"]
        d.F = b.get("F", 0)
        d.llr8 = b.get("llr8", 0)
        d.decode_enable = b.get("decode_enable", 1)
        d.tb_id = b.get("tb_id", 0)
        d.in_fmt = b.get("in_fmt", 0)
        dm = b.get("dematch")
        if dm:
            d.dematch_enable = 1
            w = dm.get("w")
            if w is not None:
                assert w.dtype == np.int16 and w.flags["C_CONTIGUOUS"]
                keep.append(w)
                d.w = w.ctypes.data
            d.G, d.C, d.r, d.rvidx, d.clear, d.Qm = dm["G"], dm["C"], dm["r"], dm["rvidx"], dm["clear"], dm["Qm"]
            d.Nl, d.Mdlharq, d.Kmimo, d.Nsoft = dm.get("Nl", 1), dm.get("Mdlharq", 8), dm.get("Kmimo", 1), dm.get("Nsoft", 1827072)
            if dm.get("scr_c_init") is not None:         # still scrambled soft bits: c_init of the codeword, r_offset
                d.scr_enable, d.scr_c_init, d.scr_offset = 1, dm["scr_c_init"], dm.get("scr_offset", 0)
            if dm.get("harq_pool") is not None:          # device-resident soft buffer: (HarqPool, slot)
                d.harq_pool = dm["harq_pool"].handle
                d.harq_slot = dm["harq_slot"]
    h = C.c_void_p()
    if tbs is None:
        rc = lib.oai_turbo_submit_batch(descs, n, flags, gpu, C.byref(h))
    else:
        nt = len(tbs)
        tds = (TbDesc * max(nt, 1))()
        rets = np.full(nt, 0xEE, dtype=np.uint8)
        valid = np.full(nt, 0xFFFFFFFF, dtype=np.uint32)
        bs = []
        for i, t in enumerate(tbs):
            cap = sum(blocks[j]["K"] // 8 for j in range(t["first_cb"], min(t["first_cb"] + t["C"], n))) + 8
            bb = np.full(cap, 0xA5, dtype=np.uint8)
            bs.append(bb)
            tds[i].first_cb, tds[i].C, tds[i].uplink = t["first_cb"], t["C"], t.get("uplink", 0)
            tds[i].b, tds[i].b_capacity = bb.ctypes.data, cap
            tds[i].ret, tds[i].valid_bytes = rets.ctypes.data + i, valid.ctypes.data + 4 * i
            uf = t.get("ul_front")
            if uf is not None:           # dict: llr (int16 / int8 array), c_init, Qm, Ncp, O_ACK, O_RI, bundling, Nbundled, Cmux, sizes
                f = UlFront()
                llr = np.ascontiguousarray(uf["llr"])
                assert llr.dtype in (np.int16, np.int8)
                f.llr, f.llr_fmt = llr.ctypes.data, 1 if llr.dtype == np.int8 else 0
                for k in ("c_init", "Qm", "Ncp", "O_ACK", "O_RI", "bundling", "Nbundled", "Cmux", "Qprime_RI", "Qprime_ACK",
                          "Qprime_CQI", "Hprime"):
                    setattr(f, k, uf.get(k, 0))
                o = {"q_ACK": np.full(18, 0x7777, dtype=np.int16), "q_RI": np.full(6, 0x7777, dtype=np.int16),
                     "q_cqi": np.full(max(f.Qm * f.Qprime_CQI, 1), 0x55, dtype=np.int8), "o_ACK": np.full(2, 0xEE, dtype=np.uint8),
                     "o_RI": np.full(1, 0xEE, dtype=np.uint8)}
                for k, a in o.items():
                    setattr(f, k, a.ctypes.data)
                keep.extend([llr, f, o])
                uf["out"] = o
                tds[i].ul_front = C.pointer(f)
        rc = lib.oai_turbo_submit_tbs(descs, n, tds, nt, flags, gpu, C.byref(h))
    if rc != 0:
        raise RuntimeError("oai_turbo_submit_batch failed (%d): %s" % (rc, last_error()))
    if h.value:
        rc = lib.oai_turbo_wait(h)
        if rc != 0:
            raise RuntimeError("oai_turbo_wait failed (%d): %s" % (rc, last_error()))
    res = [o[: b["K"] // 8] for o, b in zip(outs, blocks)], [int(s) for s in status]
    if tbs is None:
        return res
    return res[0], res[1], [(int(rets[i]), int(valid[i]), bs[i]) for i in range(len(tbs))]


def ulsch_control_sizes(O_RI, O_ACK, Or1, Msc_initial, Nsymb_initial, beta_ri_x8, beta_ack_x8, beta_cqi_x8, sumKr, nb_rb, Qm, Nsymb_pusch):
    """oai_ulsch_control_sizes: returns (rc, dict(Qprime_RI, Qprime_ACK, Qprime_CQI, G, Hprime, Hpp))."""
    v = [C.c_uint32(0) for _ in range(6)]
    rc = lib.oai_ulsch_control_sizes(O_RI, O_ACK, Or1, Msc_initial, Nsymb_initial, beta_ri_x8, beta_ack_x8, beta_cqi_x8, sumKr,
                                     nb_rb, Qm, Nsymb_pusch, *[C.byref(x) for x in v])
    return rc, dict(zip(("Qprime_RI", "Qprime_ACK", "Qprime_CQI", "G", "Hprime", "Hpp"), [int(x.value) for x in v]))


def debug_map16(y, K, term, policy=0):
    """One MAP pass on the GPU (kernel-level test hook); returns ext[K] in the reference lane layout."""
    y = np.ascontiguousarray(y, dtype=np.int16)
    ext = np.zeros(K, dtype=np.int16)
    rc = lib.oai_turbo_debug_map16(y.ctypes.data, K, term, policy, ext.ctypes.data)
    if rc != 0:
        raise RuntimeError("oai_turbo_debug_map16 failed (%d): %s" % (rc, last_error()))
    return ext


def lte_segmentation_params(B):
    """oai_lte_segmentation_params: returns (rc, dict(C, Cplus, Cminus, Kplus, Kminus, F))."""
    v = [C.c_uint32(0) for _ in range(6)]
    rc = lib.oai_lte_segmentation_params(B, *[C.byref(x) for x in v])
    return rc, dict(zip(("C", "Cplus", "Cminus", "Kplus", "Kminus", "F"), [int(x.value) for x in v]))


class HarqPool:
    """Device-resident HARQ soft buffers (oai_turbo_harq_pool_*): n_slots circular buffers w of 3*Kpi(max_K) int16 in HBM."""

    def __init__(self, n_slots, max_K, gpu=-1):
        self.handle = C.c_void_p()
        rc = lib.oai_turbo_harq_pool_create(gpu, n_slots, max_K, C.byref(self.handle))
        if rc != 0:
            raise RuntimeError("oai_turbo_harq_pool_create failed (%d): %s" % (rc, last_error()))
        self.n_slots = n_slots

    def read(self, slot, n):
        w = np.zeros(n, dtype=np.int16)
        rc = lib.oai_turbo_harq_pool_read(self.handle, slot, w.ctypes.data, n)
        if rc != 0:
            raise RuntimeError("oai_turbo_harq_pool_read failed (%d): %s" % (rc, last_error()))
        return w

    def close(self):
        if self.handle:
            lib.oai_turbo_harq_pool_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedArray:
    """numpy view of page-locked host memory from oai_turbo_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(shape)) * dtype.itemsize, 1)
        self._p = lib.oai_turbo_host_alloc(nbytes)
        if not self._p:
            raise MemoryError("oai_turbo_host_alloc(%d) failed: %s" % (nbytes, last_error()))
        buf = (C.c_uint8 * nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self.array[...] = 0

    def __del__(self):
        if getattr(self, "_p", None) and lib is not None:
            lib.oai_turbo_host_free(self._p)
            self._p = None


class HostBatchCall:
    """A prepared oai_turbo_submit_batch call over equal-parameter blocks laid out back to
    back in ONE host buffer (ideally page-locked): descriptors are built once, run() is
    submit + wait.  y_host: int16 array [n, 3K+12]; returns views of the result buffers."""

    def __init__(self, y_host, K, max_iterations, crc_type, F=0, flags=0, gpu=-1):
        n = y_host.shape[0]
        assert y_host.dtype == np.int16 and y_host.shape[1] == 3 * K + 12 and y_host.flags["C_CONTIGUOUS"]
        self.y, self.n, self.K, self.flags, self.gpu = y_host, n, K, flags, gpu
        self._pin_out = PinnedArray((n, K // 8), np.uint8)     # page-locked: results are copied straight into it
        self.out = self._pin_out.array
        self.status = np.zeros(n, dtype=np.uint8)
        self.descs = (CbDesc * n)()
        base, ob, sb = y_host.ctypes.data, self.out.ctypes.data, self.status.ctypes.data
        row = (3 * K + 12) * 2
        for i in range(n):
            d = self.descs[i]
            d.in_ = base + i * row
            d.decoded_bytes = ob + i * (K // 8)
            d.status = sb + i
            d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable = K, max_iterations, crc_type, F, 1
        self.h2d_bytes = n * row
        self.d2h_bytes = n * (K // 8 + 1)

    def submit(self):
        """oai_turbo_submit_batch: returns the in-flight handle (copies + kernels are enqueued)."""
        h = C.c_void_p()
        rc = lib.oai_turbo_submit_batch(self.descs, self.n, self.flags, self.gpu, C.byref(h))
        if rc != 0:
            raise RuntimeError("oai_turbo_submit_batch failed (%d): %s" % (rc, last_error()))
        return h

    def wait(self, h):
        """oai_turbo_wait: blocks until the batch is done; results are in self.out / self.status."""
        rc = lib.oai_turbo_wait(h)
        if rc != 0:
            raise RuntimeError("oai_turbo_wait failed (%d): %s" % (rc, last_error()))
        return self.out, self.status

    def run(self):
        return self.wait(self.submit())


class DevPlan:
    """Device-resident batch of equal-K code blocks (throughput mode).  Pointers are raw
    device addresses (e.g. torch tensor .data_ptr()), stream a cudaStream_t handle."""

    def __init__(self, ncb, K, max_iterations, crc_type, llr8=0):
        self._h = C.c_void_p()
        rc = lib.oai_turbo_dev_plan_create(ncb, K, max_iterations, crc_type, llr8, C.byref(self._h))
        if rc != 0:
            raise RuntimeError("oai_turbo_dev_plan_create failed (%d): %s" % (rc, last_error()))
        self.ncb, self.K = ncb, K

    def decode(self, y_ptr, y_stride, out_ptr, out_stride, status_ptr, stream=0):
        rc = lib.oai_turbo_dev_decode(self._h, y_ptr, y_stride, out_ptr, out_stride, status_ptr, stream)
        if rc < 0:
            raise RuntimeError("oai_turbo_dev_decode failed (%d): %s" % (rc, last_error()))
        return rc

    def set_mode(self, flags):
        """BATCH_SLIDING_WINDOW: optional sliding-window mode (not bit-exact); 0: default"""
        if lib.oai_turbo_dev_plan_set_mode(self._h, flags) != 0:
            raise RuntimeError("oai_turbo_dev_plan_set_mode failed: " + last_error())

    def profile(self, enable, fetch=False):
        """Switch per-launch event timing on/off; with fetch=True returns and resets
        ({demux,map,x1,x2} ms totals, launch counts).  Synchronise the stream first."""
        ms = (C.c_double * 4)()
        cnt = (C.c_long * 4)()
        rc = lib.oai_turbo_dev_plan_profile(self._h, 1 if enable else 0, ms if fetch else None, cnt if fetch else None)
        if rc != 0:
            raise RuntimeError("oai_turbo_dev_plan_profile failed: " + last_error())
        return (list(ms), list(cnt)) if fetch else None

    def close(self):
        if self._h:
            lib.oai_turbo_dev_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
