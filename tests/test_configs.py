"""The five BASELINE configurations at (near) full length, odd chroma sizes, and
the second decode oracle.  GPU only except the small odd-chroma streams.

  1  CIF 352x288 4:2:0, 60 frames, -qp=60 -gop=48
  2  1280x720 4:2:0 50 fps, -gop=250 -effort=10            (50 frames)
  3  1920x1080 4:2:0 CRF, psy + EPRM + loop filters        (covered by 5's chunks)
  4  1920x1080 4:4:4 lossless: decode(encode(x)) == x       (8 frames)
  5  1920x1080 4:2:0, 48-frame closed GOPs, sharded         (96 frames = 2 chunks,
     parity target = the per-chunk reference runs, parallel_encode_yuv.sh:34-41;
     includes the natural GOP roll-over, the scene cut at frame 40 and the
     stability refresh at 30 frames)

Every decoded stream is compared with `dsv2 d` AND with the independent
single-header decoder oracle/_ref/dsv28dec (dsv28dec.h:3419)."""
import os
import subprocess

import pytest

import ops
import util

need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def _yuv(y4m):
    _, _, fr = util.read_y4m(y4m)
    return b"".join(ops.yuv_bytes(f) for f in fr), len(fr)


def _d28_decode(dsv_path):
    out = dsv_path[:-4] + "_d28.y4m"
    if not os.path.exists(out):
        subprocess.run([util.REF_D28, "-inp=" + dsv_path, "-out=" + out + ".tmp", "-y4m=1"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        os.replace(out + ".tmp", out)
    return out


def _check_stream(P, name, y4m, args, over, w, h, fmt, fps, emu, lossless=False, threads=1):
    yuv, n = _yuv(y4m)
    sub = {"420": P.SUBSAMP_420, "444": P.SUBSAMP_444, "422": P.SUBSAMP_422}[fmt]
    o = P.enc_opts(w, h, sub, (fps, 1), emu=emu, **over)
    got = P.encode_frames(o, yuv, n, emu=emu)
    ref_path = util.ref_encode(y4m, args, "cfg_" + name)
    ref = open(ref_path, "rb").read()
    assert len(got) == len(ref) and got == ref, "%s: .dsv differs from the reference encoder's" % name
    meta, nfr, dec = P.decode_frames(got, emu=emu, threads=threads)
    assert nfr == n
    want, _ = _yuv(util.ref_decode(ref_path))
    assert dec == want, "%s: decoded frames differ from dsv2 d" % name
    d28, _ = _yuv(_d28_decode(ref_path))
    assert dec == d28, "%s: decoded frames differ from dsv28dec" % name
    if lossless:
        assert dec == yuv, "%s: lossless round trip is not the identity" % name
    return got


@pytest.mark.gpu
def test_config1_cif_60_frames():
    P = util.pkg()
    y4m = util.clip("cfg1", 352, 288, 60, "420")
    _check_stream(P, "c1", y4m, ["-qp=60", "-gop=48"], dict(qp=60, gop=48), 352, 288, "420", 30, False)


@pytest.mark.gpu
def test_config2_720p_gop250_effort10():
    P = util.pkg()
    y4m = util.clip("cfg2", 1280, 720, 50, "420", fps=50)
    _check_stream(P, "c2", y4m, ["-gop=250", "-effort=10"], dict(gop=250, effort=10), 1280, 720, "420", 50, False,
                  threads=2)


@pytest.mark.gpu
def test_config4_1080p_444_lossless_roundtrip():
    P = util.pkg()
    y4m = util.clip("cfg4", 1920, 1080, 8, "444")
    _check_stream(P, "c4", y4m, ["-qp=100"], dict(qp=100), 1920, 1080, "444", 30, False, lossless=True)


@pytest.mark.gpu
def test_config5_1080p_two_48_frame_chunks_sharded():
    """= the bench workload: 96 frames, chunk 48, -qp=60 -gop=48 -noeos=1; each chunk must be
    the reference's bytes for that chunk, the concatenation must decode to what
    dsv2 d / dsv28dec give for the reference's concatenation"""
    P = util.pkg()
    w, h, n, chunk = 1920, 1080, 96, 48
    y4m = util.clip("cfg5", w, h, n, "420")
    yuv, _ = _yuv(y4m)
    parts, procs = [], []
    for k in range(n // chunk):
        part = y4m[:-4] + "_c5ref%d.dsv" % k
        parts.append(part)
        if not os.path.exists(part):
            procs.append((part, subprocess.Popen(
                [util.REF_BIN, "e", "-y", "-inp=" + y4m, "-out=" + part + ".tmp", "-y4m=1", "-qp=60", "-gop=48",
                 "-sfr=%d" % (k * chunk), "-nfr=%d" % chunk, "-noeos=1"], stdout=subprocess.DEVNULL)))
    for part, p in procs:
        p.wait()
        assert p.returncode in (0, 254)
        os.replace(part + ".tmp", part)
    want = [open(p, "rb").read() for p in parts]
    o = P.enc_opts(w, h, P.SUBSAMP_420, (30, 1), qp=60, gop=48, noeos=1)
    got = P.encode_frames(o, yuv, n, chunk=chunk, threads=2)
    cat = b"".join(want)
    # the reference appends an EOS packet to the chunk that reaches the end of the input (dsv_main.c:797)
    assert got == cat[:len(got)] and len(cat) - len(got) in (0, 14)
    assert got[:len(want[0])] == want[0], "chunk 0 differs"
    # 48 pictures per chunk: the hard cut of the clip at frame 40 and the 30-picture stability
    # refresh both fall inside chunk 0, the natural GOP roll-over at its end
    pk = P.split_packets(want[0])
    assert sum(1 for p in pk if p[5] & 0x04) == chunk
    catp = y4m[:-4] + "_c5cat.dsv"
    open(catp, "wb").write(got)
    meta, nfr, dec = P.decode_frames(got, threads=2)
    assert nfr == n
    ref, _ = _yuv(util.ref_decode(catp))
    assert dec == ref
    d28, _ = _yuv(_d28_decode(catp))
    assert dec == d28


ODD = [("oddc_a", 354, 290, 5, "420", ["-qp=60", "-gop=48"], dict(qp=60, gop=48)),
       ("oddc_b", 226, 150, 6, "420", ["-qp=40", "-gop=3"], dict(qp=40, gop=3)),
       ("oddc_ll", 354, 290, 3, "420", ["-qp=100"], dict(qp=100)),
       ("oddc_422", 354, 290, 4, "422", ["-qp=55", "-gop=4"], dict(qp=55, gop=4))]
ODD_BIG = [("oddc_hd", 1282, 722, 4, "420", ["-qp=60", "-gop=48"], dict(qp=60, gop=48))]


def _odd(case, emu):
    name, w, h, n, fmt, args, over = case
    P = util.pkg()
    y4m = util.clip(name, w, h, n, fmt)
    _check_stream(P, name, y4m, args, over, w, h, fmt, 30, emu, lossless=(over.get("qp") == 100))


@need_ref
@pytest.mark.parametrize("case", ODD, ids=[c[0] for c in ODD])
def test_odd_chroma_dimensions_emulated(case):
    """chroma planes with odd width and height (frame.c:41-42 rounds the coefficient
    planes up to even, sbt.c:807 reads one border column)"""
    util.ensure_emu()
    _odd(case, True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ODD + ODD_BIG, ids=[c[0] for c in ODD + ODD_BIG])
def test_odd_chroma_dimensions_gpu(case):
    _odd(case, False)
