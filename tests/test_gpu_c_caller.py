"""The boundary proven from C: tests/c_abi/caller.c is compiled against the reference's own PHY/CODING/defs.h AND
include/oai_turbo_b200.h, linked with liboai_turbo_b200.so, and runs the per-code-block loop of dlsch_decoding.c:303-453
(function pointer `tc`, err_flag rule, reassembly :486-512) plus the batched equivalent on transport blocks whose expected
c[r] / ret / b come from the oracle chain (oracle/chain.py)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import chain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALLER = os.path.join(ROOT, "tests", "c_abi", "_build", "caller")
HAVE_REF = os.path.isdir("/root/reference/openair1/PHY/CODING")


def write_vector(path, tb, max_it, llr8, rx):
    Cn, Cp, Cm, Kp, Km, F = tb["seg"]
    b = rx["b"] if rx["b"] is not None else np.zeros(0, dtype=np.uint8)
    hdr = [Cn, Cm, Kp, Km, F, max_it, llr8, tb["G"], tb["Qm"], tb["Nl"], tb["Mdlharq"], tb["Kmimo"], tb["rv"], 0,
           rx["ret"], b.size]
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 0x0A1C0DE5))
        f.write(struct.pack("<16i", *hdr))
        f.write(tb["e"].astype("<i2").tobytes())
        for c in rx["c"]:
            f.write(c.tobytes())
        f.write(bytes([s if s is not None else 0xFE for s in rx["status"]]))
        f.write(b.tobytes())


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference headers (build container only)")
def test_c_caller_builds_against_reference_headers():
    """both include orders compile (the reference's prototypes and ours are checked against each other by gcc)"""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "c_abi")], check=True)
    assert os.path.exists(CALLER)


CASES = [  # tbs, G, Qm, max_it, llr8, noise_blocks
    (75376, 90000, 6, 4, 0, ()),          # dlsim 100 PRB MCS28: C=13 x K=5824
    (75376, 90000, 6, 4, 0, (4,)),        # block 4 fails: blocks 5..12 not decoded, c zeroed, NACK
    (7736, 14400, 4, 4, 0, ()),           # C=2 x K=3904
    (2216, 3600, 2, 6, 0, ()),            # C=1, K=2240, F=0 ... CRC24A
    (3000, 4800, 2, 4, 0, ()),            # C=1 with filler bits: K=3072, F=48
    (6192, 9000, 2, 4, 0, ()),            # C=2 x K=3136 with F=8
    (40000, 60000, 4, 4, 0, ()),          # mixed sizes: C-=2 x K=5696, C+=5 x K=5760
    (30576, 57600, 4, 4, 1, ()),          # 8-bit decoder through the same tc pointer, C=5 x K=6144
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_c_caller_matches_reference_chain(case, tmp_path):
    assert os.path.exists(CALLER), "tests/c_abi/_build/caller missing: run __graft_entry__.build() in the build container"
    tbs, G, Qm, max_it, llr8, noise = case
    tb = chain.make_tb(tbs, G, Qm, seed=tbs % 97, noise_blocks=noise, sigma_over_A=0.25)   # rate 0.84 at MCS28 needs the margin
    rx = chain.rx_tb(tb, max_it, downlink=True, llr8=llr8)
    assert (rx["b"] is None) == bool(noise)
    if not noise:
        assert np.array_equal(rx["b"], tb["b"])               # the chain recovers the transmitted transport block
    vec = str(tmp_path / "tb.vec")
    write_vector(vec, tb, max_it, llr8, rx)
    p = subprocess.run([CALLER, vec], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "caller: OK" in p.stdout, p.stdout + p.stderr
