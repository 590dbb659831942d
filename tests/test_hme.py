"""Motion estimation / intra analysis parity: dsvcu_hme and
dsvcu_intra_analysis vs the reference dsv_hme (hme.c:2001-2016) and
dsv_intra_analysis (hme.c:1835-1971) on the same frames."""
import numpy as np
import pytest

import ops
import util

# name, w, h, fmt, frames to test (pairs t-1 -> t), quant, effort
CASES = [
    ("cif", 352, 288, "420", [1, 2, 5], 252, 10),
    ("cif_q", 352, 288, "420", [3], 1000, 10),
    ("cif_e5", 352, 288, "420", [2], 120, 5),
    ("odd", 200, 136, "420", [1, 2], 300, 10),
    ("cif444", 352, 288, "444", [1], 252, 10),
]
BIG = [
    ("hd", 1280, 720, "420", [1], 140, 10),
    ("fhd", 1920, 1080, "420", [1, 2], 252, 10),
]


def _frames(name, w, h, fmt, n, **kw):
    _, _, fr = util.read_y4m(util.clip(name, w, h, n, fmt, **kw))
    return [ops.yuv_bytes(f) for f in fr]


def _recon(name, w, h, fmt, n):
    """decoded frames of a reference encode of the same clip = realistic
    'reconstructed reference' inputs"""
    y4m = util.clip(name, w, h, n, fmt)
    dsv = util.ref_encode(y4m, ["-qp=60", "-gop=48"], "hmerec")
    _, _, fr = util.read_y4m(util.ref_decode(dsv))
    return [ops.yuv_bytes(f) for f in fr]


def _run_hme(case, emu, **kw):
    name, w, h, fmt, ts, quant, effort = case
    n = max(ts) + 1
    src = _frames(name, w, h, fmt, n, **kw)
    rec = _recon(name, w, h, fmt, n) if not kw else src
    cfg = ops.Cfg(w, h, 0x5 if fmt == "420" else 0x0, effort=effort)
    R = ops.Ref()
    prev = None
    for t in ts:
        cfg.fnum = t
        D = ops.Dev(cfg, emu)
        try:
            want, w3 = R.hme(cfg, src[t], rec[t - 1], src[t - 1], prev, quant)
            got, g3 = D.hme(src[t], rec[t - 1], src[t - 1], prev, quant)
        finally:
            D.close()
        bad = ops.mv_diff(got, want)
        assert len(bad) == 0, "t=%d: %d/%d blocks differ, first %d: got %r want %r" % (
            t, len(bad), cfg.nblk, bad[0], got[bad[0]], want[bad[0]])
        assert g3 == w3, "scalars %r vs %r" % (g3, w3)
        prev = want


def _run_ia(case, emu):
    name, w, h, fmt, ts, quant, effort = case
    src = _frames(name, w, h, fmt, max(ts) + 1)
    cfg = ops.Cfg(w, h, 0x5 if fmt == "420" else 0x0, effort=effort, isP=0)
    R = ops.Ref()
    for t in ts:
        D = ops.Dev(cfg, emu)
        try:
            want = R.intra_analysis(cfg, src[t])
            got = D.intra_analysis(src[t])
        finally:
            D.close()
        bad = ops.mv_diff(got, want)
        assert len(bad) == 0, "t=%d: %d blocks differ, first %d: got %r want %r" % (
            t, len(bad), bad[0], got[bad[0]], want[bad[0]])


need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


@need_ref
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_hme_emulated(case):
    util.ensure_emu()
    _run_hme(case, True)


@need_ref
def test_hme_emulated_static():
    """noise-free clip: skip / good-enough / no-transmit paths"""
    util.ensure_emu()
    _run_hme(("cifq", 352, 288, "420", [1, 2], 400, 10), True, noise=0.0, sensor=0)


@need_ref
@pytest.mark.parametrize("case", CASES[:2] + CASES[3:], ids=[c[0] for c in CASES[:2] + CASES[3:]])
def test_intra_analysis_emulated(case):
    util.ensure_emu()
    _run_ia(case, True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + BIG, ids=[c[0] for c in CASES + BIG])
def test_hme_gpu(case):
    _run_hme(case, False)


@pytest.mark.gpu
def test_hme_gpu_static():
    _run_hme(("cifq", 352, 288, "420", [1, 2], 400, 10), False, noise=0.0, sensor=0)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + BIG, ids=[c[0] for c in CASES + BIG])
def test_intra_analysis_gpu(case):
    _run_ia(case, False)
