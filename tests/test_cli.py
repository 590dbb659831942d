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


def _run(cmd, ok=(0,)):
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL)
    assert r.returncode in ok, "%s -> %d" % (" ".join(cmd), r.returncode)
    return r.returncode


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,opt", [("444", "-out420p=1"), ("422", "-out420p=1"), ("420", "-postsharp=1"),
                                     ("444", "-postsharp=1"), ("420", "-drawinfo=7"), ("420", "-drawinfo=2")])
def test_cli_decoder_output_options_match_reference(fmt, opt, tmp_path):
    """-out420p (util.c:78-153 converters), -postsharp (bmc.c:340-361) and the -drawinfo overlay
    (dsv_decoder.c:243-350) of `dsv2 d`, on 352x288 (block grid without overhang)"""
    y4m = util.clip("cliopt", 352, 288, 5, fmt)
    dsv = util.ref_encode(y4m, ["-qp=55", "-gop=4"], "cliopt")
    want = str(tmp_path / "ref.y4m")
    got = str(tmp_path / "cu.y4m")
    _run([util.REF_BIN, "d", "-y", "-inp=" + dsv, "-out=" + want, "-y4m=1", opt])
    _run([CLI, "d", "-y", "-inp=" + dsv, "-out=" + got, "-y4m=1", opt])
    assert open(got, "rb").read() == open(want, "rb").read()


@pytest.mark.gpu
def test_cli_uyvy_input(tmp_path):
    """-fmt=5: packed UYVY 4:2:2 input (dsv.c:142-176), raw .yuv in and out"""
    import numpy as np
    w, h, n = 352, 288, 4
    _, _, fr = util.read_y4m(util.clip("cliuyvy", w, h, n, "422"))
    raw = str(tmp_path / "in.uyvy")
    with open(raw, "wb") as f:
        for Y, U, V in fr:
            y = np.frombuffer(Y, np.uint8).reshape(h, w)
            u = np.frombuffer(U, np.uint8).reshape(h, w // 2)
            v = np.frombuffer(V, np.uint8).reshape(h, w // 2)
            p = np.empty((h, w * 2), np.uint8)
            p[:, 0::4] = u
            p[:, 1::4] = y[:, 0::2]
            p[:, 2::4] = v
            p[:, 3::4] = y[:, 1::2]
            f.write(p.tobytes())
    args = ["-w=%d" % w, "-h=%d" % h, "-fmt=5", "-qp=60", "-gop=3"]
    want, got = str(tmp_path / "ref.dsv"), str(tmp_path / "cu.dsv")
    rc = _run([util.REF_BIN, "e", "-y", "-inp=" + raw, "-out=" + want] + args, ok=(0, 254))
    assert _run([CLI, "e", "-y", "-inp=" + raw, "-out=" + got] + args, ok=(0, 254)) == rc
    assert open(got, "rb").read() == open(want, "rb").read()
    wy, gy = str(tmp_path / "ref.yuv"), str(tmp_path / "cu.yuv")
    _run([util.REF_BIN, "d", "-y", "-inp=" + want, "-out=" + wy])
    _run([CLI, "d", "-y", "-inp=" + got, "-out=" + gy])
    assert open(gy, "rb").read() == open(wy, "rb").read()


@pytest.mark.gpu
def test_cli_streams_long_input_in_bounded_batches(tmp_path):
    """-chunk / -threads: the input is read and the output written batch by batch; a short last
    chunk ends with the end-of-stream packet the reference process coding it would append, and the
    threaded decoder returns the frames of `dsv2 d` on the concatenation"""
    w, h, n, chunk = 352, 288, 22, 4  # 5 full chunks + one of 2 frames; batches of 2 chunks
    y4m = util.clip("clilong", w, h, n, "420")
    cat = b""
    for k in range((n + chunk - 1) // chunk):
        part = str(tmp_path / ("ref%d.dsv" % k))
        _run([util.REF_BIN, "e", "-y", "-inp=" + y4m, "-out=" + part, "-y4m=1", "-qp=60", "-gop=4", "-sfr=%d" % (chunk * k),
              "-nfr=%d" % chunk, "-noeos=1"], ok=(0, 254))
        cat += open(part, "rb").read()
    one = str(tmp_path / "cu.dsv")
    assert _run([CLI, "e", "-y", "-inp=" + y4m, "-out=" + one, "-y4m=1", "-qp=60", "-gop=4", "-noeos=1", "-chunk=%d" % chunk,
                 "-threads=2"], ok=(0, 254)) == 254
    assert open(one, "rb").read() == cat
    refcat = str(tmp_path / "cat.dsv")
    open(refcat, "wb").write(cat)
    want, got = str(tmp_path / "ref.y4m"), str(tmp_path / "cu.y4m")
    _run([util.REF_BIN, "d", "-y", "-inp=" + refcat, "-out=" + want, "-y4m=1"])
    _run([CLI, "d", "-y", "-inp=" + one, "-out=" + got, "-y4m=1", "-threads=2"])
    assert open(got, "rb").read() == open(want, "rb").read()
    # exact multiple of the chunk size: no end-of-stream packet, exit status 254 all the same
    y4m2 = util.clip("clilong2", w, h, 8, "420")
    two = str(tmp_path / "cu2.dsv")
    _run([CLI, "e", "-y", "-inp=" + y4m2, "-out=" + two, "-y4m=1", "-qp=60", "-gop=4", "-noeos=1", "-chunk=4", "-threads=2"],
         ok=(0, 254))
    assert P_split(open(two, "rb").read())[-1][5] != 0x10


def P_split(data):
    return util.pkg().split_packets(data)
