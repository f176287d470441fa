"""The integration artefacts (VERDICT r1 missing #3): the CMake target builds the library, the patch set applies to the
reference tree, the code the patches ADD compiles against include/oai_turbo_b200.h next to the reference's
PHY/CODING/defs.h, and the CMake recipe leaves no CPU copy of a replaced symbol in libPHY.a.

Everything here is CPU-only; the tests that need /root/reference skip on the GPU box."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HAVE_REF = os.path.isdir(os.path.join(REF, "openair1/PHY/CODING"))
PATCHES = sorted(glob.glob(os.path.join(ROOT, "integration", "*.patch")))
REPLACED = ["generate_dummy_w", "lte_rate_matching_turbo_rx", "sub_block_deinterleaving_turbo", "sub_block_interleaving_turbo",
            "lte_rate_matching_turbo", "threegpplte_turbo_encoder", "phy_threegpplte_turbo_decoder16",
            "phy_threegpplte_turbo_decoder8", "init_td16", "init_td8", "free_td16", "free_td8"]


def test_patch_set_is_complete():
    names = [os.path.basename(p) for p in PATCHES]
    assert len(names) == 4, names
    text = "".join(open(p).read() for p in PATCHES)
    for f in ("cmake_targets/CMakeLists.txt", "LTE_TRANSPORT/dlsch_decoding.c", "LTE_TRANSPORT/ulsch_decoding.c",
              "LTE_PHY/ulsim.c", "LTE_PHY/dlsim.c"):
        assert "+++ b/" in text and f in text, f
    assert "case 'G':" in text and "oai_turbo_submit_tbs" in text


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference tree")
@pytest.mark.parametrize("patch", PATCHES, ids=[os.path.basename(p) for p in PATCHES])
def test_patches_apply_to_the_reference(patch):
    p = subprocess.run(["git", "apply", "--check", "-v", patch], cwd=REF, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def _added_block(patch, first_line_re):
    """the '+' lines of one patch from the first line matching first_line_re to the matching '#endif'"""
    out, on = [], False
    for line in open(patch):
        if not line.startswith("+") or line.startswith("+++"):
            continue
        body = line[1:]
        if not on and re.search(first_line_re, body):
            on = True
        if on:
            out.append(body)
            if body.startswith("#endif"):
                break
    assert out and out[-1].startswith("#endif"), "block not found in " + patch
    return "".join(out[:-1])


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference headers")
def test_added_code_compiles_next_to_the_reference_headers(tmp_path):
    dl = _added_block(PATCHES[1], r"if \(llr8_flag&2\) \{")
    ul = _added_block(PATCHES[2], r"if \(llr8_flag&2\) \{")
    src = tmp_path / "hunks.c"
    src.write_text('''
#include "PHY/CODING/defs.h"
#include "oai_turbo_b200.h"
#include "hunk_harness.h"
int opp_enabled;
uint32_t dl_hunk(harq_t *harq_process, sch_t *dlsch, short *dlsch_llr, uint32_t G, uint32_t A, uint8_t harq_pid, uint8_t llr8_flag)
{
  time_stats_t st0, *dlsch_turbo_decoding_stats = &st0;
  uint32_t r, r_offset = 0, Kr, ret, err_flag = 0;
''' + dl + '''
  return 0;
turbo_b200_decoded:
  return ret + err_flag;
}
unsigned int ul_hunk(harq_t *ulsch_harq, sch_t *ulsch, enb_t *phy_vars_eNB, unsigned int G, unsigned int A, uint8_t Q_m, uint8_t harq_pid, uint8_t llr8_flag)
{
  unsigned int r, r_offset = 0, Kr;
''' + ul + '''
  return 0;
}
''')
    cmd = ["gcc", "-std=gnu99", "-fcommon", "-Wall", "-Werror", "-Wno-unused-variable", "-DNO_OPENAIR1", "-DTURBO_B200", "-c", str(src),
           "-include", os.path.join(ROOT, "tests", "c_abi", "ref_prelude.h"),
           "-I" + os.path.join(REF, "openair1"), "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "integration"),
           "-o", str(tmp_path / "hunks.o")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    nm = subprocess.run(["nm", "-u", str(tmp_path / "hunks.o")], capture_output=True, text=True).stdout
    assert "oai_turbo_submit_tbs" in nm and "oai_turbo_wait" in nm


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference sources")
def test_cmake_recipe_leaves_no_cpu_copy_of_a_replaced_symbol(tmp_path):
    """Compiles the two translation units that stay in PHY_SRC with the COMPILE_DEFINITIONS of patch 0001 (recipe of
    oracle/Makefile: shim headers, nothing copied) and checks with nm that they no longer DEFINE any symbol the GPU
    library exports, while still defining the functions it does not replace."""
    text = open(PATCHES[0]).read()
    defs = {}
    for m in re.finditer(r'PHY/CODING/(\w+\.c) PROPERTIES COMPILE_DEFINITIONS\s*\n\+\s*"([^"]+)"', text):
        defs[m.group(1)] = ["-D" + d for d in m.group(2).split(";")]
    assert set(defs) == {"lte_rate_matching.c", "3gpplte_sse.c"}, defs
    removed = re.findall(r"PHY/CODING/(3gpplte_turbo_decoder_sse_\w+\.c)", text.split("list(REMOVE_ITEM")[1].split(")")[0])
    assert sorted(removed) == ["3gpplte_turbo_decoder_sse_16bit.c", "3gpplte_turbo_decoder_sse_8bit.c"]
    orc = os.path.join(ROOT, "oracle")
    subprocess.run(["make", "-s", "-C", orc, "_ref/shim/lte_interleaver.h"], check=True)
    inc = ["-I" + os.path.join(orc, "ref_tu"), "-I" + os.path.join(orc, "_ref", "shim"), "-I" + os.path.join(orc, "_ref"),
           "-I" + os.path.join(REF, "openair1"), "-I" + os.path.join(REF, "openair1/PHY/CODING")]
    flags = ["-O1", "-msse4.1", "-std=gnu99", "-fcommon", "-w", "-DNO_OPENAIR1"]
    defined = set()
    o1 = str(tmp_path / "rm.o")
    subprocess.run(["gcc"] + flags + ["-I" + os.path.join(orc, "ref_tu", "shim2")] + inc + defs["lte_rate_matching.c"] +
                   ["-c", os.path.join(REF, "openair1/PHY/CODING/lte_rate_matching.c"), "-o", o1], check=True)
    o2 = str(tmp_path / "enc.o")
    subprocess.run(["gcc"] + flags + inc + defs["3gpplte_sse.c"] + ["-include", os.path.join(orc, "ref_tu", "prelude.h"),
                   "-c", os.path.join(REF, "openair1/PHY/CODING/3gpplte_sse.c"), "-o", o2], check=True)
    for o in (o1, o2):
        for line in subprocess.run(["nm", "--defined-only", o], capture_output=True, text=True).stdout.splitlines():
            defined.add(line.split()[-1])
    assert not (defined & set(REPLACED)), defined & set(REPLACED)
    assert {"cpu_lte_rate_matching_turbo_rx", "cpu_generate_dummy_w", "cpu_threegpplte_turbo_encoder", "lte_rate_matching_cc",
            "sub_block_interleaving_cc"} <= defined


def test_library_exports_every_replaced_symbol():
    from openair4g_b200 import build
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB], capture_output=True, text=True).stdout
    have = {l.split()[-1] for l in out.splitlines()}
    assert set(REPLACED) <= have, set(REPLACED) - have


@pytest.mark.skipif(shutil.which("cmake") is None, reason="cmake not installed")
def test_cmake_target_configures(tmp_path):
    """configure only (a full nvcc build of the library takes a minute; __graft_entry__.build() is the build check)"""
    p = subprocess.run(["cmake", "-S", os.path.join(ROOT, "cmake_targets", "oai_turbo_b200"), "-B", str(tmp_path / "b"),
                        "-DOAI_TURBO_B200_LIBDIR=" + str(tmp_path / "lib")], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert os.path.exists(tmp_path / "b" / "Makefile") or os.path.exists(tmp_path / "b" / "build.ninja")
