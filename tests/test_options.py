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
    # 8 and 16 times as tall as wide: lifting levels meet ROWS of one sample, where the reference adds
    # the neighbouring coefficient of the plane (sbt.c:199, :221 with v = the plane itself) -- lossless
    # (every plane), intra chroma (CC levels) and P pictures
    ("tall_ll444", 64, 512, 3, "444", 30, ["-qp=100"], dict(qp=100), {}),
    ("tall_420", 64, 512, 4, "420", 30, ["-qp=55", "-gop=3"], dict(qp=55, gop=3), {}),
    ("tall16_422", 32, 512, 3, "422", 30, ["-qp=40"], dict(qp=40), {}),
]
BIG = [
    ("fhd_b32", 1920, 1080, 3, "420", 30, ["-qp=50", "-bszx=1", "-bszy=1"], dict(qp=50, bszx=1, bszy=1), {}),
    ("hd_static", 1280, 720, 4, "420", 50, ["-qp=70"], dict(qp=70), dict(noise=0.0, sensor=0)),
    # wider than 1280 and not "mostly square": 32 x 16 blocks (dsv_encoder.c:1203-1209).  8 times as wide as
    # tall and beyond is NOT covered: see test_wide_pictures_reference_disagrees_with_itself below.
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


@need_ref
def test_wide_pictures_reference_disagrees_with_itself():
    """DESIGN.md section 4, known divergences.  8 or more times as wide as tall, a lifting level meets
    COLUMNS of one sample and the reference's "v[0] op v[s] >> 1" (sbt.c:199, :221) reads the row below in
    its scratch buffer, which another level left there -- a different one in the encoder than in the
    decoder.  Shown here with the reference alone: its lossless 4:4:4 round trip is the identity at 4:1
    and is not at 8:1.  This build reads a zero there, so its own lossless round trip stays the identity;
    bit parity with the reference is neither possible nor meaningful for such pictures."""
    util.ensure_emu()
    P = util.pkg()
    for w, h, consistent in ((512, 128, True), (512, 64, False)):
        y4m = util.clip("aspect", w, h, 3, "444")
        _, _, src = util.read_y4m(y4m)
        want = b"".join(ops.yuv_bytes(f) for f in src)
        dsv = util.ref_encode(y4m, ["-qp=100"], "ref")
        _, _, dec = util.read_y4m(util.ref_decode(dsv))
        assert (b"".join(ops.yuv_bytes(f) for f in dec) == want) == consistent, (w, h)
        o = P.enc_opts(w, h, FMT["444"], (30, 1), emu=True, qp=100)
        ours = P.encode_frames(o, want, 3, emu=True)
        _, n, back = P.decode_frames(ours, emu=True)
        assert n == 3 and back == want, "own lossless round trip %dx%d" % (w, h)
        if consistent:
            assert ours == open(dsv, "rb").read()
