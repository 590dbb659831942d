"""Whole-stream encode parity: the B200 encoder (dsv_enc through the session
layer) must produce byte-identical .dsv to the reference CLI (oracle/_ref/dsv2 e)
on the same input and options; plus the closed-GOP sharded driver vs the
per-chunk reference outputs of parallel_encode_yuv.sh (:31-52)."""
import hashlib
import subprocess

import pytest

import ops
import util

need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")

# name, w, h, frames, fmt, fps, reference CLI args, our option overrides
CASES = [
    ("cif", 352, 288, 12, "420", 30, ["-qp=60", "-gop=48"], dict(qp=60, gop=48)),
    ("cif_cut", 352, 288, 46, "420", 30, ["-qp=60", "-gop=48"], dict(qp=60, gop=48)),
    ("cif_def", 352, 288, 8, "420", 30, [], dict()),
    ("cif_lowq", 352, 288, 8, "420", 30, ["-qp=20", "-gop=4"], dict(qp=20, gop=4)),
    ("cif_cqp", 352, 288, 6, "420", 30, ["-qp=45", "-rc_mode=2", "-effort=5"], dict(qp=45, rc_mode=2, effort=5)),
    ("odd", 200, 136, 6, "420", 30, ["-qp=70", "-gop=3"], dict(qp=70, gop=3)),
    ("cif444", 352, 288, 6, "444", 30, ["-qp=50", "-gop=5"], dict(qp=50, gop=5)),
    ("cif444ll", 352, 288, 4, "444", 30, ["-qp=100"], dict(qp=100)),
    ("cif_intra", 352, 288, 4, "420", 30, ["-qp=60", "-gop=0"], dict(qp=60, gop=0)),
]
BIG = [
    ("hd", 1280, 720, 5, "420", 50, ["-gop=250", "-effort=10"], dict(gop=250, effort=10)),
    ("fhd", 1920, 1080, 5, "420", 30, ["-qp=60", "-gop=48"], dict(qp=60, gop=48)),
    ("fhd444ll", 1920, 1080, 2, "444", 30, ["-qp=100"], dict(qp=100)),
]


def _encode_ours(case, emu, **kw):
    name, w, h, n, fmt, fps, args, over = case
    P = util.pkg()
    y4m = util.clip(name, w, h, n, fmt, fps=fps)
    _, _, fr = util.read_y4m(y4m)
    yuv = b"".join(ops.yuv_bytes(f) for f in fr)
    o = P.enc_opts(w, h, P.SUBSAMP_420 if fmt == "420" else P.SUBSAMP_444, (fps, 1), emu=emu, **over)
    return y4m, P.encode_frames(o, yuv, n, emu=emu, **kw)


def _run(case, emu):
    name, w, h, n, fmt, fps, args, over = case
    y4m, got = _encode_ours(case, emu)
    ref = open(util.ref_encode(y4m, args, "enc_" + hashlib.md5(" ".join(args).encode()).hexdigest()[:8]), "rb").read()
    if got != ref:
        P = util.pkg()
        pg, pr = P.split_packets(got), P.split_packets(ref)
        for i, (a, b) in enumerate(zip(pg, pr)):
            if a != b:
                k = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
                raise AssertionError("packet %d (type 0x%02x) differs at byte %d (lengths %d vs %d)" % (
                    i, b[5], k, len(a), len(b)))
        raise AssertionError("packet count %d vs %d" % (len(pg), len(pr)))


@need_ref
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_encode_emulated(case):
    util.ensure_emu()
    _run(case, True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + BIG, ids=[c[0] for c in CASES + BIG])
def test_encode_gpu(case):
    _run(case, False)


def _ref_sharded(y4m, n, chunk, args):
    """parallel_encode_yuv.sh semantics: one reference process per chunk with
    -sfr/-nfr/-noeos=1, outputs concatenated"""
    out = b""
    for k in range((n + chunk - 1) // chunk):
        part = y4m[:-4] + "_shard%d_%d.dsv" % (chunk, k)
        cmd = [util.REF_BIN, "e", "-y", "-inp=" + y4m, "-out=" + part, "-y4m=1", "-sfr=%d" % (k * chunk),
               "-nfr=%d" % chunk, "-noeos=1"] + args
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL)
        assert r.returncode in (0, 254)
        out += open(part, "rb").read()
    return out


def _run_sharded(emu, w, h, n, chunk, threads):
    P = util.pkg()
    y4m = util.clip("shard", w, h, n, "420")
    _, _, fr = util.read_y4m(y4m)
    yuv = b"".join(ops.yuv_bytes(f) for f in fr)
    o = P.enc_opts(w, h, P.SUBSAMP_420, (30, 1), emu=emu, qp=60, gop=chunk, noeos=1)
    got = P.encode_frames(o, yuv, n, emu=emu, chunk=chunk, threads=threads)
    want = _ref_sharded(y4m, n, chunk, ["-qp=60", "-gop=%d" % chunk])
    # the reference appends an EOS to the chunk that hits the end of the input
    # (dsv_main.c:797); a sharded encode of an exact multiple has none
    assert got == want[:len(got)] and len(want) - len(got) in (0, 14)
    # and the sharded decoder returns the frames the reference decoder gets
    meta, nfr, dec = P.decode_frames(got, emu=emu, threads=threads)
    tmp = y4m[:-4] + "_shardcat.dsv"
    open(tmp, "wb").write(got)
    _, _, ref = util.read_y4m(util.ref_decode(tmp))
    assert nfr == len(ref) == n
    assert dec == b"".join(ops.yuv_bytes(f) for f in ref)


@need_ref
def test_sharded_emulated():
    util.ensure_emu()
    _run_sharded(True, 352, 288, 12, 4, 1)  # the host emulation keeps launch state in globals: one thread


@pytest.mark.gpu
def test_sharded_gpu():
    _run_sharded(False, 352, 288, 24, 6, 4)
    _run_sharded(False, 1920, 1080, 8, 4, 2)
