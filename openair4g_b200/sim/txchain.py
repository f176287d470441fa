"""Transmit side of the LTE shared channel for the harness: transport-block CRC, code-block
segmentation, turbo encoding, sub-block interleaving and rate matching (3GPP TS 36.212
5.1.1-5.1.4), vectorised over code blocks with numpy.

Reference counterparts (behaviour mirrored, code independent): lte_segmentation
(openair1/PHY/CODING/lte_segmentation.c:52-170), crc24a/b (crc_byte.c:53-153),
threegpplte_turbo_encoder (3gpplte_sse.c:380-476), sub_block_interleaving_turbo and
lte_rate_matching_turbo (lte_rate_matching.c:51-130, 464-634).
"""
import functools

import numpy as np

NSOFT = 1827072          # openair1/PHY/LTE_TRANSPORT/defs.h:62
LTE_NULL = 2

# 36.212 table 5.1.3-3, (f1, f2) for K = 40..6144
from ._qpp import QPP_F1F2  # noqa: E402


def k_list():
    return (list(range(40, 513, 8)) + list(range(528, 1025, 16)) + list(range(1056, 2049, 32)) +
            list(range(2112, 6145, 64)))


@functools.lru_cache(maxsize=None)
def qpp(K):
    f1, f2 = QPP_F1F2[k_list().index(K)]
    i = np.arange(K, dtype=np.int64)
    return ((f1 * i + f2 * i * i) % K).astype(np.int64)


def _crc_bits(bits, poly, width):
    """MSB-first CRC with zero start value over a (n_blocks, n_bits) 0/1 array; returns (n_blocks, width) bits."""
    reg = np.zeros(bits.shape[0], dtype=np.int64)
    top = 1 << (width - 1)
    mask = (1 << width) - 1
    for j in range(bits.shape[1]):
        fb = ((reg >> (width - 1)) & 1) ^ bits[:, j]
        reg = ((reg << 1) & mask) ^ (fb * poly)
    return ((reg[:, None] >> np.arange(width - 1, -1, -1)[None, :]) & 1).astype(np.uint8)


def crc24a(bits):
    return _crc_bits(bits, 0x864CFB, 24)


def crc24b(bits):
    return _crc_bits(bits, 0x800063, 24)


def segmentation(B):
    """(C, Cplus, Cminus, Kplus, Kminus, F) with the reference's rule (lte_segmentation.c:52-134)."""
    if B <= 6144:
        L, C, Bp = 0, 1, B
    else:
        L = 24
        C = B // (6144 - L)
        if (6144 - L) * C < B:
            C += 1
        Bp = B + C * L
    if C > 16:
        raise ValueError("too many segments")     # MAX_NUM_DLSCH_SEGMENTS, lte_segmentation.c:67-70
    q = Bp // C
    if q <= 40:
        Kp, Km = 40, 0
    elif q <= 512:
        Kp, Km = (q >> 3) << 3, q - 8
    elif q <= 1024:
        Kp = (q >> 4) << 4
        Kp += 16 if Kp < q else 0
        Km = Kp - 16
    elif q <= 2048:
        Kp = (q >> 5) << 5
        Kp += 32 if Kp < q else 0
        Km = Kp - 32
    else:
        Kp = (q >> 6) << 6
        Kp += 64 if Kp < q else 0
        Km = Kp - 64
    if C == 1:
        Cp, Km, Cm = 1, 0, 0
    else:
        Cm = ((C * Kp - Bp) & 0xFFFFFFFF) // (Kp - Km)       # uint32 arithmetic like the reference
        Cp = (C - Cm) & 0xFFFFFFFF
    F = (Cp * Kp + Cm * Km - Bp) & 0xFFFFFFFF
    return C, Cp, Cm, Kp, Km, F


def turbo_encode(c):
    """c: (n, K) info bits -> (n, 3K+12) coded bits in the reference's order (s,p1,p2 triples, then
    x z x z x z of encoder 1 and of encoder 2)."""
    n, K = c.shape
    pi = qpp(K)
    ci = c[:, pi]
    out = np.zeros((n, 3 * K + 12), dtype=np.uint8)
    out[:, 0:3 * K:3] = c

    def rsc(u):
        d1 = np.zeros(n, dtype=np.uint8)
        d2 = np.zeros(n, dtype=np.uint8)
        d3 = np.zeros(n, dtype=np.uint8)
        z = np.zeros((n, K), dtype=np.uint8)
        for k in range(K):                        # g0 = 1+D^2+D^3 (feedback), g1 = 1+D+D^3
            a = u[:, k] ^ d2 ^ d3
            z[:, k] = a ^ d1 ^ d3
            d3, d2, d1 = d2, d1, a
        tail = np.zeros((n, 6), dtype=np.uint8)
        for t in range(3):
            x = d2 ^ d3
            tail[:, 2 * t] = x
            tail[:, 2 * t + 1] = d1 ^ d3
            d3, d2, d1 = d2, d1, np.zeros(n, dtype=np.uint8)
        return z, tail

    z1, t1 = rsc(c)
    z2, t2 = rsc(ci)
    out[:, 1:3 * K:3] = z1
    out[:, 2:3 * K:3] = z2
    out[:, 3 * K:3 * K + 6] = t1
    out[:, 3 * K + 6:] = t2
    return out


@functools.lru_cache(maxsize=64)
def gold_sequence(c_init, n):
    """Scrambling sequence c(0..n-1) of 36.211 7.2 (x1 from 1, x2 from c_init, Nc = 1600) as a 0/1 uint8 array.
    Reference counterpart: lte_gold_generic (openair1/PHY/LTE_REFSIG/lte_gold.c:151-180)."""
    M31 = (1 << 31) - 1
    x1, x2 = 1, c_init & M31
    total = 1600 + n
    # 31-bit Fibonacci LFSRs advanced bit by bit on Python ints, output collected in blocks
    out = np.zeros(total, dtype=np.uint8)
    for i in range(total):
        out[i] = (x1 ^ x2) & 1
        f1 = ((x1 >> 3) ^ x1) & 1
        f2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1
        x1 = (x1 >> 1) | (f1 << 30)
        x2 = (x2 >> 1) | (f2 << 30)
    return out[1600:].copy()


@functools.lru_cache(maxsize=None)
def _interleave_index(K):
    """For w index j (0..3*Kpi): index into the d stream (3*(K+4) entries) or -1 for a NULL slot."""
    D = K + 4
    RTC = (D + 31) // 32
    Kpi = 32 * RTC
    ND = Kpi - D
    brev = np.array([int("{:05b}".format(c)[::-1], 2) for c in range(32)])
    col = np.repeat(np.arange(32), RTC)
    row = np.tile(np.arange(RTC), 32)
    idx = brev[col] + 32 * row                    # position in the padded sub-block, column-permuted order
    src = np.full(3 * Kpi, -1, dtype=np.int64)
    k = np.arange(Kpi)
    ok = idx >= ND
    src[k[ok]] = 3 * (idx[ok] - ND)
    src[Kpi + 2 * k[ok]] = 3 * (idx[ok] - ND) + 1
    idx2 = (idx + 1) % Kpi                        # stream 2 is shifted by one (36.212 5.1.4.1.1)
    ok2 = idx2 >= ND
    src[Kpi + 2 * k[ok2] + 1] = 3 * (idx2[ok2] - ND) + 2
    return src, RTC, Kpi, ND


def rate_match(d, K, F, G, C, Qm, Nl, r, rv, Mdlharq=8, Kmimo=1):
    """d: (n, 3K+12) coded bits of n blocks that share (K,F,r) -> (n, E) transmitted bits."""
    src, RTC, Kpi, ND = _interleave_index(K)
    n = d.shape[0]
    w = np.full((n, 3 * Kpi), LTE_NULL, dtype=np.uint8)
    valid = src >= 0
    w[:, valid] = d[:, src[valid]]
    if F:                                         # filler bits are NULL in the systematic and first parity stream
        fill = valid & (((src % 3) < 2) & (src // 3 < F))
        w[:, fill] = LTE_NULL
    Nir = NSOFT // Kmimo // min(8, Mdlharq)
    Ncb = min(Nir // C, 3 * Kpi)
    Gp = G // Nl // Qm
    E = Nl * Qm * (Gp // C) if r < C - (Gp % C) else Nl * Qm * ((0 if Gp % C == 0 else 1) + Gp // C)
    k0 = RTC * (2 + rv * ((0 if Ncb % (8 * RTC) == 0 else 1) + Ncb // (8 * RTC)) * 2)
    order = np.concatenate([np.arange(k0, Ncb), np.arange(0, k0)]) if k0 < Ncb else np.arange(Ncb)
    keep = order[w[0, order] != LTE_NULL]         # NULL pattern depends only on (K,F)
    reps = (E + keep.size - 1) // keep.size
    sel = np.tile(keep, reps)[:E]
    return w[:, sel], E
