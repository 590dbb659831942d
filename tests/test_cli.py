"""dsv2cu command line vs the reference CLI on the same files and options
(GPU only: the CLI is the product binary and has no emulation)."""
import os
import subprocess

import pytest

import util

CLI = os.path.join(util.PKG, "dsv2cu")


@pytest.mark.gpu
def test_cli_encode_decode_files_match_reference():
    assert os.path.exists(CLI), "run `make`"
    y4m = util.clip("cli", 352, 288, 8, "420")
    ref_dsv = util.ref_encode(y4m, ["-qp=55", "-gop=4"], "cli")
    ref_y4m = util.ref_decode(ref_dsv)
    out_dsv = y4m[:-4] + "_cu.dsv"
    r = subprocess.run([CLI, "e", "-y", "-inp=" + y4m, "-out=" + out_dsv, "-y4m=1", "-qp=55", "-gop=4"])
    assert r.returncode in (0, 254)
    assert open(out_dsv, "rb").read() == open(ref_dsv, "rb").read()
    out_y4m = y4m[:-4] + "_cu_dec.y4m"
    r = subprocess.run([CLI, "d", "-y", "-inp=" + out_dsv, "-out=" + out_y4m, "-y4m=1"])
    assert r.returncode == 0
    assert open(out_y4m, "rb").read() == open(ref_y4m, "rb").read()


@pytest.mark.gpu
def test_cli_chunk_semantics_match_parallel_encode_script():
    """-sfr/-nfr/-noeos as parallel_encode_yuv.sh uses them, and the in-process -chunk= form"""
    y4m = util.clip("cli2", 352, 288, 12, "420")
    cat = b""
    for k in range(3):
        part = y4m[:-4] + "_refpart%d.dsv" % k
        r = subprocess.run([util.REF_BIN, "e", "-y", "-inp=" + y4m, "-out=" + part, "-y4m=1", "-qp=60", "-gop=4",
                            "-sfr=%d" % (4 * k), "-nfr=4", "-noeos=1"], stdout=subprocess.DEVNULL)
        assert r.returncode in (0, 254)
        mine = y4m[:-4] + "_cupart%d.dsv" % k
        r2 = subprocess.run([CLI, "e", "-y", "-inp=" + y4m, "-out=" + mine, "-y4m=1", "-qp=60", "-gop=4",
                             "-sfr=%d" % (4 * k), "-nfr=4", "-noeos=1"])
        assert r2.returncode == r.returncode
        assert open(mine, "rb").read() == open(part, "rb").read()
        cat += open(part, "rb").read()
    one = y4m[:-4] + "_cuchunk.dsv"
    subprocess.run([CLI, "e", "-y", "-inp=" + y4m, "-out=" + one, "-y4m=1", "-qp=60", "-gop=4", "-noeos=1", "-chunk=4",
                    "-threads=3"], check=False)
    got = open(one, "rb").read()
    assert got == cat[:len(got)] and len(cat) - len(got) in (0, 14)


@pytest.mark.gpu
def test_reference_cli_sources_link_against_the_library(tmp_path):
    """INTEGRATION.md route 1: the reference's own dsv_main.c + util.c, compiled
    against include/ and linked with libdsv2cuda.so (oracle/Makefile target
    _ref/dsv2_dropin, built where /root/reference exists), is a working `dsv2`
    whose output equals the reference build's."""
    exe = os.path.join(util.REF_DIR, "dsv2_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dsv2_dropin not built")
    y4m = util.clip("cli3", 352, 288, 6, "420")
    ref_dsv = util.ref_encode(y4m, ["-qp=70", "-gop=3"], "cli3")
    out = str(tmp_path / "o.dsv")
    r = subprocess.run([exe, "e", "-y", "-inp=" + y4m, "-out=" + out, "-y4m=1", "-qp=70", "-gop=3"],
                       stdout=subprocess.DEVNULL)
    assert r.returncode in (0, 254)
    assert open(out, "rb").read() == open(ref_dsv, "rb").read()
    dec = str(tmp_path / "o.y4m")
    subprocess.run([exe, "d", "-y", "-inp=" + out, "-out=" + dec, "-y4m=1", "-postsharp=1"], check=True,
                   stdout=subprocess.DEVNULL)
    refdec = str(tmp_path / "r.y4m")
    subprocess.run([util.REF_BIN, "d", "-y", "-inp=" + ref_dsv, "-out=" + refdec, "-y4m=1", "-postsharp=1"], check=True,
                   stdout=subprocess.DEVNULL)
    assert open(dec, "rb").read() == open(refdec, "rb").read()
