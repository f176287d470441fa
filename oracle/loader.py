"""TEST INFRASTRUCTURE -- ctypes access to the two CPU oracles.

* ``port()``  : oracle/_build/liboracle_port.so, this repo's plain-C restatement
  (oracle/port/*.c), built by ``make -C oracle port``.
* ``ref()``   : oracle/_ref/libref_oai.so, the UNMODIFIED reference sources compiled
  in place from /root/reference by ``make -C oracle ref``.
  It only exists where it was built (this container) or where the prebuilt .so
  travelled (the GPU box); ``ref()`` returns None otherwise.

Reference signatures bound here: openair1/PHY/CODING/defs.h:132,152,239-253,362,
470-484,499-513.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_build", "liboracle_port.so")
_REF = os.path.join(_HERE, "_ref", "libref_oai.so")

_port = None
_ref = None
_ref_tried = False

i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u32ref = C.POINTER(C.c_uint32)


def build(ref_too=True):
    """Compile the oracle libraries (called from __graft_entry__.build())."""
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
    if ref_too:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def port():
    global _port
    if _port is not None:
        return _port
    if not os.path.exists(_PORT):
        build(ref_too=False)
    L = C.CDLL(_PORT)
    L.orc_qpp_index.argtypes = [C.c_int]
    L.orc_qpp_f1.argtypes = [C.c_int]
    L.orc_qpp_f2.argtypes = [C.c_int]
    L.orc_qpp_K.argtypes = [C.c_int]
    L.orc_qpp_table.argtypes = [C.c_int, u16p]
    for f in ("orc_crc24a", "orc_crc24b", "orc_crc16", "orc_crc8"):
        getattr(L, f).argtypes = [u8p, C.c_int]
        getattr(L, f).restype = C.c_uint32
    L.orc_lte_segmentation.argtypes = [C.c_uint32] + [u32ref] * 6
    L.orc_lte_gold_generic.argtypes = [u32ref, u32ref, C.c_uint8]
    L.orc_lte_gold_generic.restype = C.c_uint32
    L.orc_gold_words.argtypes = [C.c_uint32, C.c_void_p, C.c_int]
    L.orc_gold_words.restype = None
    L.orc_dlsch_unscrambling.argtypes = [C.c_uint32, i16p, C.c_int]
    L.orc_dlsch_unscrambling.restype = None
    L.orc_generate_dummy_w.argtypes = [C.c_uint32, u8p, C.c_uint8]
    L.orc_generate_dummy_w.restype = C.c_uint32
    L.orc_lte_rate_matching_turbo_rx.argtypes = [C.c_uint32, C.c_uint32, i16p, u8p, i16p, C.c_uint8,
                                                 C.c_uint32] + [C.c_uint8] * 7 + [u32ref]
    L.orc_sub_block_deinterleaving_turbo.argtypes = [C.c_uint32, C.c_void_p, i16p]
    L.orc_sub_block_deinterleaving_turbo.restype = None
    L.orc_turbo_encode.argtypes = [u8p, C.c_int, u8p]
    L.orc_turbo_encode.restype = None
    L.orc_sub_block_interleaving_turbo.argtypes = [C.c_uint32, u8p, u8p]
    L.orc_sub_block_interleaving_turbo.restype = C.c_uint32
    L.orc_lte_rate_matching_turbo.argtypes = [C.c_uint32, C.c_uint32, u8p, u8p, C.c_uint8, C.c_uint32] + [C.c_uint8] * 6
    L.orc_lte_rate_matching_turbo.restype = C.c_uint32
    L.orc_turbo_decoder16.argtypes = [i16p, u8p, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8]
    L.orc_turbo_decoder16.restype = C.c_uint8
    L.orc_turbo_decoder16_sw.argtypes = [i16p, u8p, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p]
    L.orc_turbo_decoder16_sw.restype = C.c_uint8
    L.orc_sw_shift.argtypes = [i16p, C.c_int]
    L.orc_sw_windows.argtypes = [C.c_int]
    L.orc_turbo_decoder8.argtypes = [i16p, u8p, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint8]
    L.orc_turbo_decoder8.restype = C.c_uint8
    L.orc_log_map16.argtypes = [i16p, i16p, i16p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_log_map16.restype = None
    L.orc_turbo_decoder16_batch.argtypes = [i16p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_uint16,
                                            C.c_uint8, C.c_uint8, C.c_int]
    L.orc_turbo_decoder16_batch.restype = None
    L.orc_ulsch_control_sizes.argtypes = [C.c_uint32] * 12 + [C.POINTER(UlSizes)]
    L.orc_ulsch_front.argtypes = [i16p, C.c_uint32, C.c_uint32, C.POINTER(UlSizes)] + [C.c_uint32] * 5 + \
        [i16p, i16p, i16p, np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS"), u8p, u8p]
    _port = L
    return L


class UlSizes(C.Structure):
    """orc_ul_sizes_t"""
    _fields_ = [(k, C.c_uint32) for k in ("Qprime_RI", "Qprime_ACK", "Qprime_CQI", "Q_RI", "Q_CQI", "G", "H", "Hprime", "Hpp",
                                          "Cmux", "Rmux_prime")]


class RefUlParams(C.Structure):
    """ref_ul_params_t of oracle/ref_tu/ulfront_tu.c"""
    _fields_ = [(k, C.c_uint32) for k in ("TBS", "nb_rb", "Nsymb_pusch", "Nsymb_initial", "Msc_initial", "mcs", "rvidx", "round",
                                          "O_ACK", "O_RI", "Or1", "bundling", "Nbundled", "Ncp", "beta_cqi_x8", "beta_ri_x8",
                                          "beta_ack_x8", "rnti", "subframe", "Nid_cell", "max_iter", "Mdlharq", "llr8")]


def ref_ulsch_decoding(params, llr, state=None):
    """Runs the UNMODIFIED reference ulsch_decoding() (compiled behind oracle/ref_tu/shim4) on one subframe's soft bits.
    params: dict of RefUlParams fields; state: the opaque HARQ state of a previous round (or None).
    Returns dict(ret, e, q_ACK, q_RI, q_cqi, o_ACK, o_RI, c (16x768), b, w (16 x 18624), state)."""
    L = ref()
    L.ref_ul_state_size.restype = C.c_size_t
    L.ref_ulsch_decoding_run.restype = C.c_uint
    L.ref_ulsch_decoding_run.argtypes = [C.POINTER(RefUlParams)] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + \
        [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    if state is None:
        state = np.zeros(L.ref_ul_state_size() + 64, dtype=np.uint8)
    p = RefUlParams()
    for k, v in params.items():
        setattr(p, k, v)
    ll = aligned(llr.size + 64, np.int16)
    ll[:llr.size] = llr
    e = np.zeros(14 * 1200 * 6, dtype=np.int16)
    q_ack, q_ri = np.zeros(18, dtype=np.int16), np.zeros(6, dtype=np.int16)
    q_cqi = np.zeros(2560, dtype=np.int8)
    o_ack, o_ri = np.zeros(4, dtype=np.uint8), np.zeros(2, dtype=np.uint8)
    c = np.zeros((16, 768), dtype=np.uint8)
    b = np.zeros(16 * 768, dtype=np.uint8)
    w = np.zeros((16, 3 * (6144 + 64)), dtype=np.int16)
    ret = L.ref_ulsch_decoding_run(C.byref(p), state.ctypes.data, ll.ctypes.data, e.ctypes.data, e.size, q_ack.ctypes.data,
                                   q_ri.ctypes.data, q_cqi.ctypes.data, q_cqi.size, o_ack.ctypes.data, o_ri.ctypes.data,
                                   c.ctypes.data, b.ctypes.data, b.size, w.ctypes.data)
    return {"ret": int(ret), "e": e, "q_ACK": q_ack, "q_RI": q_ri, "q_cqi": q_cqi, "o_ACK": o_ack, "o_RI": o_ri, "c": c, "b": b,
            "w": w, "state": state}


def ref():
    """The compiled reference, initialised (crcTableInit, init_td16, init_td8), or None."""
    global _ref, _ref_tried
    if _ref is not None or _ref_tried:
        return _ref
    _ref_tried = True
    if not os.path.exists(_REF):
        if os.path.isdir("/root/reference/openair1/PHY/CODING"):
            build(ref_too=True)
        if not os.path.exists(_REF):
            return None
    L = C.CDLL(_REF)
    stats = [C.c_void_p] * 7
    for f in ("ref_phy_threegpplte_turbo_decoder16", "ref_phy_threegpplte_turbo_decoder8"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_uint16, C.c_uint16, C.c_uint16,
                                  C.c_uint8, C.c_uint8, C.c_uint8] + stats
        getattr(L, f).restype = C.c_uint8
    L.ref_threegpplte_turbo_encoder.argtypes = [u8p, C.c_uint16, C.c_void_p, C.c_uint8, C.c_uint16, C.c_uint16]
    L.ref_threegpplte_turbo_encoder.restype = None
    for f in ("ref_crc24a", "ref_crc24b", "ref_crc16", "ref_crc8"):
        getattr(L, f).argtypes = [u8p, C.c_int]
        getattr(L, f).restype = C.c_uint32
    L.ref_lte_segmentation.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32] + [u32ref] * 6
    L.ref_lte_gold_generic.argtypes = [u32ref, u32ref, C.c_uint8]
    L.ref_lte_gold_generic.restype = C.c_uint32
    L.ref_generate_dummy_w.argtypes = [C.c_uint32, u8p, C.c_uint8]
    L.ref_generate_dummy_w.restype = C.c_uint32
    L.ref_lte_rate_matching_turbo_rx.argtypes = [C.c_uint32, C.c_uint32, i16p, u8p, i16p, C.c_uint8,
                                                 C.c_uint32] + [C.c_uint8] * 7 + [u32ref]
    L.ref_sub_block_deinterleaving_turbo.argtypes = [C.c_uint32, C.c_void_p, i16p]
    L.ref_sub_block_deinterleaving_turbo.restype = None
    L.ref_sub_block_interleaving_turbo.argtypes = [C.c_uint32, C.c_void_p, u8p]
    L.ref_sub_block_interleaving_turbo.restype = C.c_uint32
    L.ref_lte_rate_matching_turbo.argtypes = [C.c_uint32, C.c_uint32, u8p, u8p, C.c_uint8, C.c_uint32] + [C.c_uint8] * 8
    L.ref_lte_rate_matching_turbo.restype = C.c_uint32
    L.ref_td_batch.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p] + [C.c_int] * 7
    L.ref_td_batch.restype = None
    L.log_map16.argtypes = [C.c_void_p] * 7 + [C.c_ushort, C.c_ubyte, C.c_ubyte, C.c_int] + [C.c_void_p] * 4
    L.log_map16.restype = None
    L.ref_crcTableInit()
    L.ref_init_td16()
    L.ref_init_td8()
    _ref = L
    return L


# ---- helpers shared by the tests -------------------------------------------------

def aligned(n, dtype, align=64):
    """Zeroed 1-D array whose data pointer is `align`-byte aligned (the reference
    decoders use aligned SSE loads on y)."""
    itemsize = np.dtype(dtype).itemsize
    raw = np.zeros(n * itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * itemsize].view(dtype)


_STATS = None


def _stats():
    global _STATS
    if _STATS is None:
        _STATS = [np.zeros(16, dtype=np.int64) for _ in range(7)]
    return [s.ctypes.data for s in _STATS]


def ref_decode16(y, n, max_it, crc_type, F=0, which=16):
    """Run the compiled reference decoder on one block; returns (bytes[n/8], ret)."""
    L = ref()
    yy = aligned(3 * n + 12 + 64, np.int16)      # slack: TD8 reads past the end
    yy[:3 * n + 12] = y[:3 * n + 12]
    out = aligned(n // 8 + 64, np.uint8)
    fn = L.ref_phy_threegpplte_turbo_decoder16 if which == 16 else L.ref_phy_threegpplte_turbo_decoder8
    p = port()
    idx = p.orc_qpp_index(n)
    f1 = p.orc_qpp_f1(idx) if idx >= 0 else 0
    f2 = p.orc_qpp_f2(idx) if idx >= 0 else 0
    r = fn(yy.ctypes.data, out.ctypes.data, n, f1, f2, max_it, crc_type, F, *_stats())
    return out[:n // 8].copy(), int(r)


def port_decode16(y, n, max_it, crc_type, F=0):
    L = port()
    yy = np.ascontiguousarray(y[:3 * n + 12], dtype=np.int16)
    out = np.zeros(n // 8 + 4, dtype=np.uint8)
    r = L.orc_turbo_decoder16(yy, out, n, max_it, crc_type, F)
    return out[:n // 8].copy(), int(r)


def port_decode16_sw(y, n, max_it, crc_type, F=0, llr=False):
    """CPU model of the optional sliding-window mode (oracle/port/td16_sw_port.c)"""
    L = port()
    yy = np.ascontiguousarray(y[:3 * n + 12], dtype=np.int16)
    out = np.zeros(n // 8 + 4, dtype=np.uint8)
    dbg = np.zeros(n, dtype=np.int32) if llr else None
    r = L.orc_turbo_decoder16_sw(yy, out, n, max_it, crc_type, F, dbg.ctypes.data if llr else None)
    return (out[:n // 8].copy(), int(r), dbg) if llr else (out[:n // 8].copy(), int(r))


def ref_decode_batch(y_blocks, n, max_it, crc_type, total=None, threads=1, which=16):
    """Compiled reference over many blocks on `threads` pthreads.  y_blocks: [nd, >=3n+12] int16.
    `total` >= nd decodes are run (cycling over the nd inputs; only the first pass stores results).
    Returns (out[nd, n/8], ret[nd], seconds)."""
    import time
    L = ref()
    nd = y_blocks.shape[0]
    stride = ((3 * n + 12 + 64 + 7) // 8) * 8
    yy = aligned(nd * stride, np.int16).reshape(nd, stride)
    yy[:, :3 * n + 12] = y_blocks[:, :3 * n + 12]
    ostride = ((n // 8 + 64 + 15) // 16) * 16
    out = aligned(nd * ostride, np.uint8).reshape(nd, ostride)
    ret = np.zeros(nd, dtype=np.uint8)
    t0 = time.perf_counter()
    L.ref_td_batch(yy.ctypes.data, stride, out.ctypes.data, ostride, ret.ctypes.data, nd,
                   max(total or nd, nd), n, max_it, crc_type, which, threads)
    dt = time.perf_counter() - t0
    return out[:, :n // 8].copy(), ret, dt


def port_decode8(y, n, max_it, crc_type, F=0):
    L = port()
    yy = np.zeros(3 * n + 12 + 64, dtype=np.int16)
    yy[:3 * n + 12] = y[:3 * n + 12]
    out = np.zeros(n // 8 + 4, dtype=np.uint8)
    r = L.orc_turbo_decoder8(yy, out, n, max_it, crc_type, F)
    return out[:n // 8].copy(), int(r)
