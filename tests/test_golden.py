"""Committed fixtures (tests/golden/golden.json, made by make_golden.py from the
unmodified reference): our encoder must reproduce the reference's .dsv bytes and
our decoder the reference decoder's frames -- without needing oracle/_ref at
test time."""
import hashlib
import json
import os

import pytest

import ops
import util

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))
OPTS = {"-qp": "qp", "-gop": "gop", "-rc_mode": "rc_mode", "-effort": "effort"}


def _run(name, emu):
    g = G[name]
    P = util.pkg()
    y4m = util.clip("golden_" + name, g["w"], g["h"], g["frames"], g["fmt"], fps=g["fps"], kind="tri")
    # the fixture clips are generated with integer arithmetic only: the input must be
    # the committed one everywhere, a mismatch is a failure (round 1 skipped here)
    assert hashlib.md5(open(y4m, "rb").read()).hexdigest() == g["y4m_md5"], "golden input clip differs"
    _, _, fr = util.read_y4m(y4m)
    yuv = b"".join(ops.yuv_bytes(f) for f in fr)
    kw = {OPTS[a.split("=")[0]]: int(a.split("=")[1]) for a in g["args"]}
    o = P.enc_opts(g["w"], g["h"], P.SUBSAMP_420 if g["fmt"] == "420" else P.SUBSAMP_444, (g["fps"], 1), emu=emu, **kw)
    dsv = P.encode_frames(o, yuv, g["frames"], emu=emu)
    assert len(dsv) == g["dsv_bytes"]
    assert hashlib.md5(dsv).hexdigest() == g["dsv_md5"]
    meta, nfr, dec = P.decode_frames(dsv, emu=emu)
    assert nfr == g["frames"]
    assert hashlib.md5(dec).hexdigest() == g["decoded_frames_md5"]


@pytest.mark.parametrize("name", sorted(G))
def test_golden_emulated(name):
    util.ensure_emu()
    _run(name, True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G))
def test_golden_gpu(name):
    _run(name, False)
