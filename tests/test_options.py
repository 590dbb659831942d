"""Encoder option coverage: every hot-path-relevant option of the reference CLI
(effort, psy bits, skip threshold, filters, block sizes,
pyramid depth, rate-control modes, scene-change detection, temporal AQ, quality
extremes, frame rates, chroma formats, odd geometries).  For each: .dsv bytes ==
reference encoder's, decoded frames == reference decoder's."""
import hashlib

import pytest

import ops
import util

need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")

# name, w, h, frames, fmt, fps, reference CLI args, our overrides, clip generator overrides
CASES = [
    ("e0", 352, 288, 6, "420", 30, ["-qp=60", "-effort=0"], dict(qp=60, effort=0), {}),
    ("e4", 352, 288, 6, "420", 30, ["-qp=60", "-effort=4"], dict(qp=60, effort=4), {}),
    ("e7", 352, 288, 6, "420", 30, ["-qp=60", "-effort=7"], dict(qp=60, effort=7), {}),
    ("psy0", 352, 288, 6, "420", 30, ["-qp=60", "-psy=0"], dict(qp=60, psy=0), {}),
    ("psy5", 352, 288, 6, "420", 30, ["-qp=50", "-psy=5"], dict(qp=50, psy=5), {}),
    ("skipoff", 352, 288, 6, "420", 30, ["-qp=60", "-skipthresh=-1"], dict(qp=60, skipthresh=-1), {}),
    ("skip9", 352, 288, 6, "420", 30, ["-qp=60", "-skipthresh=9"], dict(qp=60, skipthresh=9), {}),
    ("nofilt", 352, 288, 6, "420", 30, ["-qp=40", "-ifilter=0", "-pfilter=0"], dict(qp=40, ifilter=0, pfilter=0), {}),
    ("pfilt1", 352, 288, 6, "420", 30, ["-qp=40", "-pfilter=1", "-psharp=0"], dict(qp=40, pfilter=1, psharp=0), {}),
    ("b32", 352, 288, 6, "420", 30, ["-qp=60", "-bszx=1", "-bszy=1"], dict(qp=60, bszx=1, bszy=1), {}),
    ("b32x", 352, 288, 5, "420", 30, ["-qp=60", "-bszx=1", "-bszy=0"], dict(qp=60, bszx=1, bszy=0), {}),
    ("pyr3", 352, 288, 5, "420", 30, ["-qp=60", "-pyrlevels=3"], dict(qp=60, pyrlevels=3), {}),
    ("abr", 352, 288, 10, "420", 30, ["-rc_mode=1", "-kbps=800"], dict(rc_mode=1, kbps=800), {}),
    ("abrgop", 352, 288, 10, "420", 30, ["-rc_mode=1", "-kbps=300", "-rc_pergop=1", "-gop=4"],
     dict(rc_mode=1, kbps=300, rc_pergop=1, gop=4), {}),
    ("noscd", 352, 288, 46, "420", 30, ["-qp=60", "-scd=0", "-gop=60"], dict(qp=60, scd=0, gop=60), {}),
    ("notaq", 352, 288, 8, "420", 30, ["-qp=60", "-tempaq=0", "-gop=3"], dict(qp=60, tempaq=0, gop=3), {}),
    ("dib0", 352, 288, 5, "420", 30, ["-qp=60", "-dib=0"], dict(qp=60, dib=0), {}),
    ("q5", 352, 288, 5, "420", 30, ["-qp=5"], dict(qp=5), {}),
    ("q95", 352, 288, 5, "420", 30, ["-qp=95"], dict(qp=95), {}),
    ("fps60", 352, 288, 6, "420", 60, ["-qp=60"], dict(qp=60), {}),
    ("fps24", 352, 288, 6, "420", 24, ["-qp=60"], dict(qp=60), {}),
    ("static", 352, 288, 8, "420", 30, ["-qp=50"], dict(qp=50), dict(noise=0.0, sensor=0)),
    ("tiny", 64, 48, 5, "420", 30, ["-qp=60"], dict(qp=60), {}),
    ("w180", 180, 100, 5, "420", 30, ["-qp=60"], dict(qp=60), {}),
    ("ll420", 352, 288, 3, "420", 30, ["-qp=100"], dict(qp=100), {}),
    ("c422", 352, 288, 6, "422", 30, ["-qp=60"], dict(qp=60), {}),
    ("c422q", 176, 144, 5, "422", 30, ["-qp=30", "-gop=2"], dict(qp=30, gop=2), {}),
    ("c411", 352, 288, 5, "411", 30, ["-qp=60"], dict(qp=60), {}),
    ("c410", 352, 288, 5, "410", 30, ["-qp=60"], dict(qp=60), {}),
]
BIG = [
    ("fhd_b32", 1920, 1080, 3, "420", 30, ["-qp=50", "-bszx=1", "-bszy=1"], dict(qp=50, bszx=1, bszy=1), {}),
    ("hd_static", 1280, 720, 4, "420", 50, ["-qp=70"], dict(qp=70), dict(noise=0.0, sensor=0)),
    # wider than 1280 and not "mostly square": 32 x 16 blocks (dsv_encoder.c:1203-1209).  Aspect ratios of 8:1
    # and beyond are NOT covered: there a lifting level meets a 1-sample dimension and the reference's
    # DO_SIMPLE_LO / DO_5_TAP_LO read v[s] outside the line (sbt.c:199, :221), i.e. stale scratch memory from
    # earlier calls -- its own encoder and decoder disagree on such input.
    ("wide_32x16", 1536, 384, 3, "420", 30, ["-qp=60"], dict(qp=60), {}),
]
FMT = {"420": 0x5, "444": 0x0, "422": 0x4, "411": 0x8, "410": 0xA}


def _run(case, emu):
    name, w, h, n, fmt, fps, args, over, ckw = case
    P = util.pkg()
    y4m = util.clip("opt_" + name, w, h, n, fmt, fps=fps, **ckw)
    _, _, fr = util.read_y4m(y4m)
    yuv = b"".join(ops.yuv_bytes(f) for f in fr)
    o = P.enc_opts(w, h, FMT[fmt], (fps, 1), emu=emu, **over)
    got = P.encode_frames(o, yuv, n, emu=emu)
    tag = "opt_" + hashlib.md5(" ".join(args).encode()).hexdigest()[:8]
    ref_path = util.ref_encode(y4m, args, tag)
    ref = open(ref_path, "rb").read()
    if got != ref:
        pg, pr = P.split_packets(got), P.split_packets(ref)
        for i, (a, b) in enumerate(zip(pg, pr)):
            if a != b:
                k = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
                raise AssertionError("packet %d (type 0x%02x) differs at byte %d (lengths %d vs %d)" % (
                    i, b[5], k, len(a), len(b)))
        raise AssertionError("packet count %d vs %d" % (len(pg), len(pr)))
    meta, nfr, dec = P.decode_frames(got, emu=emu)
    _, _, rf = util.read_y4m(util.ref_decode(ref_path))
    assert nfr == len(rf)
    assert dec == b"".join(ops.yuv_bytes(f) for f in rf), "decoded frames differ from the reference decoder's"


@need_ref
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_option_emulated(case):
    util.ensure_emu()
    _run(case, True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + BIG, ids=[c[0] for c in CASES + BIG])
def test_option_gpu(case):
    _run(case, False)
