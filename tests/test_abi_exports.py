"""CPU-only: the C-ABI library builds, loads and exports every symbol include/*.h declares
(no compute calls -- there is no GPU here and no CPU fallback in the product)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "oai_turbo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_builds_and_exports_every_declared_symbol():
    from openair4g_b200 import build
    build.build()
    from openair4g_b200 import capi
    names = declared_functions()
    assert "phy_threegpplte_turbo_decoder16" in names and "oai_turbo_submit_batch" in names and len(names) >= 17
    for n in names:
        assert hasattr(capi.lib, n), "missing export " + n
    assert set(capi.EXPORTS) == set(names)
    assert b"sm_100a" in capi.lib.oai_turbo_b200_version()


def test_cbdesc_layout_matches_header():
    """ctypes mirror of oai_cb_desc_t has the C layout (checked by compiling a probe)."""
    import ctypes
    import subprocess
    import tempfile
    from openair4g_b200 import capi
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "oai_turbo_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",' \
          'sizeof(oai_tx_desc_t),offsetof(oai_tx_desc_t,E),offsetof(oai_tx_desc_t,K),offsetof(oai_tx_desc_t,C),offsetof(oai_tx_desc_t,r));' \
          'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(oai_cb_desc_t),offsetof(oai_cb_desc_t,K),offsetof(oai_cb_desc_t,w),offsetof(oai_cb_desc_t,C),offsetof(oai_cb_desc_t,tb_id),offsetof(oai_cb_desc_t,harq_pool),offsetof(oai_cb_desc_t,harq_slot),offsetof(oai_cb_desc_t,scr_c_init),offsetof(oai_cb_desc_t,scr_enable));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")], check=True)
        out = subprocess.run([os.path.join(d, "p")], capture_output=True, text=True, check=True).stdout.split()
    T = capi.TxDesc
    assert [int(v) for v in out[:5]] == [ctypes.sizeof(T), T.E.offset, T.K.offset, T.C.offset, T.r.offset]
    out = out[5:]
    D = capi.CbDesc
    assert [int(v) for v in out] == [ctypes.sizeof(D), D.K.offset, D.w.offset, D.C.offset, D.tb_id.offset, D.harq_pool.offset, D.harq_slot.offset, D.scr_c_init.offset, D.scr_enable.offset]


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under openair4g_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "openair4g_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU oracle", ""), os.path.join(dirpath, f)


def test_flag_constants_match_the_header():
    """the ctypes mirror's flag values are the header's #defines"""
    from openair4g_b200 import capi
    txt = open(os.path.join(ROOT, "include", "oai_turbo_b200.h")).read()
    val = lambda name: int(re.search(r"#define\s+" + name + r"\s+(\d+)u", txt).group(1))
    assert capi.BATCH_DL_STOP_AFTER_FAILURE == val("OAI_BATCH_DL_STOP_AFTER_FAILURE")
    assert capi.BATCH_SLIDING_WINDOW == val("OAI_BATCH_SLIDING_WINDOW")
    assert capi.TX_DEVICE_POINTERS == val("OAI_TX_DEVICE_POINTERS")
    assert capi.BATCH_SLIDING_WINDOW & capi.BATCH_DL_STOP_AFTER_FAILURE == 0
