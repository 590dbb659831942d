"""Encoder-side operator parity (forward SBT, quantiser + symbol list,
prediction/subtract, reconstruct + loop filters) vs the reference operators
(sbt.c:847-886, hzcc.c:585-613, bmc.c:1057-1090)."""
import ctypes as C

import numpy as np
import pytest

import ops
import util

need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")

GEOM = [
    ("cif", 352, 288, "420"),
    ("odd", 200, 136, "420"),
    ("cif444", 352, 288, "444"),
]
BIG = [("hd", 1280, 720, "420"), ("fhd", 1920, 1080, "420"), ("fhd444", 1920, 1080, "444")]


def _frames(name, w, h, fmt, n):
    _, _, fr = util.read_y4m(util.clip(name, w, h, n, fmt))
    return [ops.yuv_bytes(f) for f in fr]


def _cfg(w, h, fmt, **kw):
    return ops.Cfg(w, h, 0x5 if fmt == "420" else 0x0, **kw)


def _rand_blockdata(cfg, seed, isP):
    rng = np.random.default_rng(seed)
    bd = rng.integers(0, 256, cfg.nblk).astype(np.uint8)
    return bd & (0x75 if isP else 0x0b)


def _fwd(geom, emu, isP, lossless):
    name, w, h, fmt = geom
    src = _frames(name, w, h, fmt, 2)[1]
    cfg = _cfg(w, h, fmt, isP=isP, lossless=lossless)
    bd = _rand_blockdata(cfg, 7, isP)
    R = ops.Ref()
    D = ops.Dev(cfg, emu)
    try:
        for p in range(3):
            want = R.fwd_sbt(cfg, p, src, bd)
            got = D.fwd_sbt(p, src, bd)
            assert got.shape == want.shape
            d = np.argwhere(got != want)
            assert len(d) == 0, "plane %d: %d coefs differ, first at (y,x)=%r got %d want %d" % (
                p, len(d), tuple(d[0]), got[tuple(d[0])], want[tuple(d[0])])
    finally:
        D.close()


def _mvs_for(cfg, src_t, rec, src_p, quant=252):
    return ops.Ref().hme(cfg, src_t, rec, src_p, None, quant)[0]


def _blockdata_from_mvs(mvs):
    """what encode_stable_blocks + encode_motion leave in blockdata
    (dsv_encoder.c:838-873, :722-766), minus the neighbour-difference bit"""
    f = mvs["flags"]
    bd = np.zeros(len(mvs), np.uint8)
    bd |= ((f & ops.F_INTRA) != 0).astype(np.uint8) << 4
    skip = ((f & ops.F_SKIP) != 0) & ((f & ops.F_INTRA) == 0)
    bd |= skip.astype(np.uint8) << 2
    bd |= ((f & ops.F_SIMCMPLX) != 0).astype(np.uint8) << 6
    bd |= ((f & ops.F_EPRM) != 0).astype(np.uint8) << 5
    bd |= ((f & ops.F_SKIP) != 0).astype(np.uint8) << 0
    return bd


def _quant(geom, emu, isP, q, lossless=0):
    name, w, h, fmt = geom
    fr = _frames(name, w, h, fmt, 3)
    cfg = _cfg(w, h, fmt, isP=isP, lossless=lossless)
    R = ops.Ref()
    if isP:
        mvs = _mvs_for(cfg, fr[2], fr[1], fr[1])
        bd = _blockdata_from_mvs(mvs)
        _, resd = R.sub_pred(cfg, mvs, fr[2], fr[1])
        src = resd
    else:
        mvs = R.intra_analysis(cfg, fr[2])
        bd = _rand_blockdata(cfg, 3, 0)
        src = fr[2]
    D = ops.Dev(cfg, emu)
    lib = D.lib
    try:
        for p in range(3):
            k = R.fwd_sbt(cfg, p, src, bd)
            want_k, want_bits = R.encode_plane(cfg, p, q, k, bd, mvs)
            got_k, syms, dc = D.quant_plane(p, q, k, bd, mvs)
            d = np.argwhere(got_k != want_k.reshape(got_k.shape))
            assert len(d) == 0, "plane %d: %d dequantised coefs differ, first (y,x)=%r got %d want %d" % (
                p, len(d), tuple(d[0]), got_k[tuple(d[0])], want_k.reshape(got_k.shape)[tuple(d[0])])
            cw, ch = got_k.shape[1], got_k.shape[0]
            out = (C.c_uint8 * (len(want_bits) + len(syms) * 16 + 4096))()
            sy = np.ascontiguousarray(syms, np.int32)
            n = lib.dsv_hzcc_pack_plane(sy.ctypes.data_as(C.c_void_p), len(sy), dc, cw, ch, out, len(out))
            assert n == len(want_bits), "plane %d: %d bytes vs reference %d" % (p, n, len(want_bits))
            assert bytes(out[:n]) == want_bits, "plane %d bytes differ" % p
    finally:
        D.close()


def _mc(geom, emu, q=252, do_filter=1, lossless=0):
    name, w, h, fmt = geom
    fr = _frames(name, w, h, fmt, 3)
    cfg = _cfg(w, h, fmt, isP=1, fnum=3, lossless=lossless)
    R = ops.Ref()
    mvs = _mvs_for(cfg, fr[2], fr[1], fr[1])
    # what encode_stable_blocks does before prediction: skip blocks get a zero vector
    sk = (mvs["flags"] & ops.F_SKIP) != 0
    mvs["x"][sk] = 0
    mvs["y"][sk] = 0
    bd = _blockdata_from_mvs(mvs)
    D = ops.Dev(cfg, emu)
    try:
        wp, wr = R.sub_pred(cfg, mvs, fr[2], fr[1])
        gp, gr = D.sub_pred(mvs, fr[2], fr[1])
        assert gp == wp, "prediction differs: %r" % util.first_diff(gp[:w * h], wp[:w * h], w)
        assert gr == wr, "residual differs: %r" % util.first_diff(gr[:w * h], wr[:w * h], w)
        want = R.add_res(cfg, mvs, bd, q, wr, wp, do_filter)
        got = D.add_res(mvs, bd, q, wr, wp, do_filter)
        assert got == want, "reconstruction differs: luma %r" % util.first_diff(got[:w * h], want[:w * h], w)
    finally:
        D.close()


@need_ref
@pytest.mark.parametrize("geom", GEOM, ids=[g[0] for g in GEOM])
@pytest.mark.parametrize("mode", ["I", "P", "lossless"])
def test_fwd_sbt_emulated(geom, mode):
    util.ensure_emu()
    _fwd(geom, True, mode == "P", mode == "lossless")


@need_ref
@pytest.mark.parametrize("geom", GEOM, ids=[g[0] for g in GEOM])
@pytest.mark.parametrize("mode,q", [("I", 252), ("P", 252), ("P", 1200), ("I", 40)])
def test_quant_emulated(geom, mode, q):
    util.ensure_emu()
    _quant(geom, True, mode == "P", q)


@need_ref
def test_quant_lossless_emulated():
    util.ensure_emu()
    _quant(GEOM[2], True, False, 1, lossless=1)
    _quant(GEOM[0], True, True, 1, lossless=1)


@need_ref
@pytest.mark.parametrize("geom", GEOM, ids=[g[0] for g in GEOM])
def test_mc_emulated(geom):
    util.ensure_emu()
    _mc(geom, True)
    _mc(geom, True, q=1500, do_filter=0)


@need_ref
def test_mc_lossless_emulated():
    util.ensure_emu()
    _mc(GEOM[2], True, q=1, lossless=1)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + BIG, ids=[g[0] for g in GEOM + BIG])
@pytest.mark.parametrize("mode", ["I", "P", "lossless"])
def test_fwd_sbt_gpu(geom, mode):
    _fwd(geom, False, mode == "P", mode == "lossless")


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + BIG, ids=[g[0] for g in GEOM + BIG])
@pytest.mark.parametrize("mode,q", [("I", 252), ("P", 252), ("P", 1200), ("I", 40)])
def test_quant_gpu(geom, mode, q):
    _quant(geom, False, mode == "P", q)


@pytest.mark.gpu
def test_quant_lossless_gpu():
    _quant(BIG[2], False, False, 1, lossless=1)
    _quant(GEOM[0], False, True, 1, lossless=1)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + BIG, ids=[g[0] for g in GEOM + BIG])
def test_mc_gpu(geom):
    _mc(geom, False)
    _mc(geom, False, q=1500, do_filter=0)


def _frame_calls_equal_plane_calls(geom, emu, isP, q=252):
    """dsvcu_fwd_sbt_frame / dsvcu_quant_frame / dsvcu_inv_sbt_frame (all planes
    through shared launches) against the per-plane entry points, which the tests
    above pin to the reference."""
    name, w, h, fmt = geom
    fr = _frames(name, w, h, fmt, 1)
    cfg = _cfg(w, h, fmt, isP=1 if isP else 0, fnum=1)
    bd = np.zeros(cfg.nblk, np.uint8)
    rng = np.random.default_rng(7)
    bd[:] = rng.integers(0, 128, cfg.nblk)
    mvs = np.zeros(cfg.nblk, ops.MV_DTYPE)
    mvs["x"] = rng.integers(-40, 41, cfg.nblk)
    mvs["y"] = rng.integers(-40, 41, cfg.nblk)
    D = ops.Dev(cfg, emu)
    try:
        lib, ctx, fm = D.lib, D.ctx, cfg.fmeta()
        D.set_blockdata(bd)
        D.set_mvs(mvs)
        src = D.frame(fr[0])
        ka, kb = D.coefs(), D.coefs()
        oa, ob = D.frame(), D.frame()
        for p in range(3):
            D.ck(lib.dsvcu_fwd_sbt(ctx, src, p, ka, C.byref(fm)))
        D.ck(lib.dsvcu_fwd_sbt_frame(ctx, src, kb, C.byref(fm), 7))
        for p in range(3):
            assert np.array_equal(D.coefs_download(ka, p), D.coefs_download(kb, p)), "forward transform, plane %d" % p
        want = []
        for p in range(3):
            D.ck(lib.dsvcu_quant_plane(ctx, ka, p, q, C.byref(fm)))
            want.append(_fetch(D, p))
        D.ck(lib.dsvcu_quant_frame(ctx, kb, q, C.byref(fm), 7))
        for p in range(3):
            got = _fetch(D, p)
            assert got[1] == want[p][1] and np.array_equal(got[0], want[p][0]), "symbols, plane %d" % p
            assert np.array_equal(D.coefs_download(ka, p), D.coefs_download(kb, p)), "dequantised coefficients, plane %d" % p
        for p in range(3):
            D.ck(lib.dsvcu_inv_sbt(ctx, oa, p, ka, q, C.byref(fm)))
        D.ck(lib.dsvcu_inv_sbt_frame(ctx, ob, kb, q, C.byref(fm), 7))
        assert D.download(oa) == D.download(ob), "inverse transform"
        # a partial mask only touches the selected planes
        D.ck(lib.dsvcu_frame_clear_plane(ctx, ob, 1, 0))
        D.ck(lib.dsvcu_inv_sbt_frame(ctx, ob, kb, q, C.byref(fm), 5))
        a, b = D.download(oa), D.download(ob)
        ysz = w * h
        csz = cfg.cw * cfg.ch
        assert a[:ysz] == b[:ysz] and a[ysz + csz:] == b[ysz + csz:]
        assert b[ysz:ysz + csz] == bytes(csz)
    finally:
        D.close()


def _fetch(D, plane):
    syms = C.POINTER(D.P.DSVCU_SYMBOL)()
    n, dc = C.c_int(), C.c_int()
    D.ck(D.lib.dsvcu_fetch_symbols(D.ctx, plane, C.byref(syms), C.byref(n), C.byref(dc)))
    arr = np.ctypeslib.as_array(C.cast(syms, C.POINTER(C.c_int32)), shape=(max(n.value, 1), 2))[:n.value].copy()
    return arr, dc.value


@pytest.mark.parametrize("geom", GEOM, ids=[g[0] for g in GEOM])
@pytest.mark.parametrize("mode", ["I", "P"])
def test_frame_level_calls_emulated(geom, mode):
    util.ensure_emu()
    _frame_calls_equal_plane_calls(geom, True, mode == "P")


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + BIG, ids=[g[0] for g in GEOM + BIG])
@pytest.mark.parametrize("mode", ["I", "P"])
def test_frame_level_calls_gpu(geom, mode):
    _frame_calls_equal_plane_calls(geom, False, mode == "P")
