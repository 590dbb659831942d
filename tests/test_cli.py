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
