"""TEST INFRASTRUCTURE -- deterministic synthetic code blocks.

info bits: xorshift64* seeded with 0x0A1B2C3D ^ K ^ block index; CRC24A/B appended;
36.212 turbo encoder (oracle port, itself checked against the compiled reference
encoder); BPSK LLR = A*(2b-1) + round(sigma*N(0,1)) as int16.
"""
import numpy as np

from . import loader

MASK64 = (1 << 64) - 1


def xorshift64s_bytes(seed, nbytes):
    x = seed & MASK64
    if x == 0:
        x = 0x9E3779B97F4A7C15
    out = bytearray()
    while len(out) < nbytes:
        x ^= x >> 12
        x ^= (x << 25) & MASK64
        x ^= x >> 27
        out += ((x * 0x2545F4914F6CDD1D) & MASK64).to_bytes(8, "little")
    return np.frombuffer(bytes(out[:nbytes]), dtype=np.uint8).copy()


def info_block(K, blk, crc_type=1, F=0):
    """K/8 bytes: F filler zero bits, payload, CRC24 (A: over bytes[F/8:], B: over all)."""
    L = loader.port()
    nb = K // 8
    b = np.zeros(nb, dtype=np.uint8)
    b[: nb - 3] = xorshift64s_bytes(0x0A1B2C3D ^ K ^ blk, nb - 3)
    if F:
        b[: F // 8] = 0
    if crc_type == 0:
        c = L.orc_crc24a(np.ascontiguousarray(b[F // 8:]), K - 24 - F) >> 8
    else:
        c = L.orc_crc24b(b, K - 24) >> 8
    b[nb - 3] = (c >> 16) & 0xFF
    b[nb - 2] = (c >> 8) & 0xFF
    b[nb - 1] = c & 0xFF
    return b


def encode(info_bytes):
    L = loader.port()
    K = info_bytes.size * 8
    out = np.zeros(3 * K + 12, dtype=np.uint8)
    L.orc_turbo_encode(np.ascontiguousarray(info_bytes), info_bytes.size, out)
    return out


def llr_block(K, blk, regime="clean", A=8, crc_type=1, F=0, sigma_over_A=None):
    """Returns (y int16[3K+12], info bytes).  Regimes follow SURVEY.md 8d config 3:
    clean (sigma/A=0.5), waterfall (1.08), noise (uniform +-16, no signal),
    full (uniform over all of int16: saturation / wrap paths)."""
    rng = np.random.default_rng([0xB200, K, blk, {"clean": 1, "waterfall": 2, "noise": 3, "full": 4}.get(regime, 5)])
    info = info_block(K, blk, crc_type, F)
    n = 3 * K + 12
    if regime == "noise":
        return rng.integers(-16, 17, size=n).astype(np.int16), info
    if regime == "full":
        return rng.integers(-32768, 32768, size=n).astype(np.int16), info
    s = sigma_over_A if sigma_over_A is not None else {"clean": 0.5, "waterfall": 1.08}[regime]
    bits = encode(info).astype(np.int64)
    y = A * (2 * bits - 1) + np.rint(s * A * rng.standard_normal(n)).astype(np.int64)
    return np.clip(y, -32768, 32767).astype(np.int16), info
