"""Decoder-side and frame-helper operator parity against the reference operators
themselves (oracle/_ref/librefops.so): lossy inverse transform (sbt.c:889-934),
de-quantiser (hzcc.c:450-583 through dsv_decode_plane :616-649), intra filter
(bmc.c:390-457), border extension (frame.c:357-434) and the 2x luma pyramid
(frame.c:210-234 + dsv_encoder.c:493-516).  Round 1 covered these only through
whole streams."""
import ctypes as C

import numpy as np
import pytest

import ops
import util
from test_encops import GEOM, BIG, _frames, _cfg, _rand_blockdata, _mvs_for, _blockdata_from_mvs

need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
ODDC = [("oddchroma", 354, 290, "420"), ("oddchroma2", 226, 150, "420")]  # chroma 177x145, 113x75


def _ref_lib():
    R = ops.Ref()
    return R, R.lib


def _coded_planes(cfg, R, fr, isP, q):
    """reference transform + quantiser output of a real picture: per plane
    (de-quantised coefficients, plane bytes), plus the side information used"""
    if isP:
        mvs = _mvs_for(cfg, fr[2], fr[1], fr[1])
        bd = _blockdata_from_mvs(mvs)
        _, src = R.sub_pred(cfg, mvs, fr[2], fr[1])
    else:
        mvs = R.intra_analysis(cfg, fr[2])
        bd = _rand_blockdata(cfg, 3, 0)
        src = fr[2]
    planes = []
    for p in range(3):
        k = R.fwd_sbt(cfg, p, src, bd)
        planes.append(R.encode_plane(cfg, p, q, k, bd, mvs))
    return planes, bd, mvs


def _inv(geom, emu, isP, q, seed):
    name, w, h, fmt = geom
    fr = _frames(name, w, h, fmt, 3)
    cfg = _cfg(w, h, fmt, isP=1 if isP else 0, fnum=3)
    R, _ = _ref_lib()
    planes, bd, mvs = _coded_planes(cfg, R, fr, isP, q)
    # the inverse transform reads blockdata (ringing / stable / intra flags steer the
    # filtered Haar and the adaptive levels): random flags on top of the real field
    bd = (bd | _rand_blockdata(cfg, seed, isP)).astype(np.uint8)
    D = ops.Dev(cfg, emu)
    try:
        D.set_blockdata(bd)
        fm = cfg.fmeta()
        out = D.frame()
        k = D.coefs()
        for p in range(3):
            D.coefs_upload(k, p, planes[p][0])
            D.ck(D.lib.dsvcu_inv_sbt(D.ctx, out, p, k, q, C.byref(fm)))
        got = D.download(out)
        off = 0
        for p in range(3):
            pw, ph = cfg.plane_dims(p)
            want = R.inv_sbt(cfg, p, q, planes[p][0], bd)
            g = got[off:off + pw * ph]
            assert g == want, "plane %d: %r" % (p, util.first_diff(g, want, pw))
            off += pw * ph
    finally:
        D.close()


def _dequant(geom, emu, isP, q):
    name, w, h, fmt = geom
    fr = _frames(name, w, h, fmt, 3)
    cfg = _cfg(w, h, fmt, isP=1 if isP else 0, fnum=3)
    R, rl = _ref_lib()
    planes, bd, mvs = _coded_planes(cfg, R, fr, isP, q)
    D = ops.Dev(cfg, emu)
    lib = D.lib
    try:
        D.set_blockdata(bd)
        D.set_mvs(mvs)
        fm = cfg.fmeta()
        k = D.coefs()
        for p in range(3):
            deq, bits = planes[p]
            cw, ch = D.coef_dims(k, p)
            # reference decoder on the reference's bytes
            want = np.zeros(cw * ch, np.int32)
            rc = cfg.ref()
            ok = rl.refop_decode_plane(C.byref(rc), p, q, ops._buf(bits), len(bits), ops._buf(bytes(bd)),
                                       mvs.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
            assert ok == 1
            # our host parser on the same bytes -> ordered symbols -> device de-quantiser
            cap = C.c_int()
            st = lib.dsvcu_symbol_staging(D.ctx, p, C.byref(cap))
            lstart = (C.c_int * 5)()
            dc = C.c_int()
            n = lib.dsv_hzcc_unpack_plane(ops._buf(bits), len(bits), st, cap.value - 1, cw, ch, lstart, C.byref(dc))
            assert n >= 0
            D.ck(lib.dsvcu_dequant_plane(D.ctx, k, p, q, C.byref(fm), n, lstart, dc.value))
            got = D.coefs_download(k, p).reshape(-1)
            d = np.nonzero(got != want)[0]
            assert len(d) == 0, "plane %d: %d coefficients differ, first at %d (x=%d y=%d) got %d want %d" % (
                p, len(d), d[0], d[0] % cw, d[0] // cw, got[d[0]], want[d[0]])
    finally:
        D.close()


def _intra_filter(geom, emu, q):
    name, w, h, fmt = geom
    fr = _frames(name, w, h, fmt, 3)
    cfg = _cfg(w, h, fmt, isP=0, fnum=2)
    R, _ = _ref_lib()
    # a decoded-looking picture: reference inverse transform of a coarsely quantised I picture
    planes, bd, mvs = _coded_planes(cfg, R, fr, False, q)
    pic = b"".join(R.inv_sbt(cfg, p, q, planes[p][0], bd) for p in range(3))
    bd = _rand_blockdata(cfg, 11, 0)
    D = ops.Dev(cfg, emu)
    try:
        D.set_blockdata(bd)
        fm = cfg.fmeta()
        for do_filter in (1, 0):
            want = R.intra_filter(cfg, q, bd, pic, do_filter)
            f = D.frame(pic, extend=False)
            D.ck(D.lib.dsvcu_intra_filter(D.ctx, q, C.byref(fm), 0, f, do_filter))
            got = D.download(f)
            assert got == want, "do_filter=%d: %r" % (do_filter, util.first_diff(got[:w * h], want[:w * h], w))
    finally:
        D.close()


def _bordered(D, f, plane):
    w, h, s = C.c_int(), C.c_int(), C.c_int()
    D.lib.dsvcu_frame_plane_dims(f, plane, C.byref(w), C.byref(h), C.byref(s))
    out = (C.c_uint8 * (s.value * (h.value + 64)))()
    D.ck(D.lib.dsvcu_frame_download_bordered(D.ctx, f, plane, out))
    D.ck(D.lib.dsvcu_sync(D.ctx))
    return np.frombuffer(bytes(out), np.uint8).reshape(h.value + 64, s.value), w.value, h.value, s.value


def _extend_and_pyramid(geom, emu):
    name, w, h, fmt = geom
    fr = _frames(name, w, h, fmt, 2)
    cfg = _cfg(w, h, fmt)
    R, rl = _ref_lib()
    rc = cfg.ref()
    D = ops.Dev(cfg, emu)
    try:
        f = D.frame(fr[1])  # upload + dsvcu_extend_frame
        for p in range(3):
            got, gw, gh, gs = _bordered(D, f, p)
            buf = (C.c_uint8 * (gs * (gh + 64)))()
            pw, ph = C.c_int(), C.c_int()
            rs = rl.refop_extend_plane(C.byref(rc), ops._buf(fr[1]), p, buf, C.byref(pw), C.byref(ph))
            assert (rs, pw.value, ph.value) == (gs, gw, gh), "plane %d geometry" % p
            want = np.frombuffer(bytes(buf), np.uint8).reshape(gh + 64, gs)
            # compare the visible area and the full 32-px border (not the stride padding)
            a, b = got[:, :gw + 64], want[:, :gw + 64]
            d = np.argwhere(a != b)
            assert len(d) == 0, "plane %d: %d border/pixel bytes differ, first at row %d col %d" % (
                p, len(d), d[0][0] - 32, d[0][1] - 32)
        # the fused form (border of every plane + all levels in two launches) on a frame
        # that has NOT been extended must give the same borders and the same pyramid
        f2 = D.frame(fr[1], extend=False)
        pyr = C.c_void_p()
        D.ck(D.lib.dsvcu_pyramid_create(D.ctx, C.byref(pyr), cfg.pyr))
        D._pyr.append(pyr)
        D.ck(D.lib.dsvcu_extend_pyramid(D.ctx, f2, pyr))
        for p in range(3):
            a, gw, gh, gs = _bordered(D, f, p)
            b, _, _, _ = _bordered(D, f2, p)
            assert np.array_equal(a[:, :gw + 64], b[:, :gw + 64]), "fused extension, plane %d" % p
        for lvl in range(1, cfg.pyr + 1):
            pf = D.lib.dsvcu_pyramid_level(pyr, lvl)
            got, gw, gh, gs = _bordered(D, C.c_void_p(pf), 0)
            buf = (C.c_uint8 * (gs * (gh + 64) + 4096))()
            pw, ph = C.c_int(), C.c_int()
            rs = rl.refop_pyramid_level(C.byref(rc), ops._buf(fr[1]), lvl, buf, C.byref(pw), C.byref(ph))
            assert (rs, pw.value, ph.value) == (gs, gw, gh), "pyramid level %d geometry" % lvl
            want = np.frombuffer(bytes(buf)[:gs * (gh + 64)], np.uint8).reshape(gh + 64, gs)
            a, b = got[:, :gw + 64], want[:, :gw + 64]
            d = np.argwhere(a != b)
            assert len(d) == 0, "pyramid level %d: %d bytes differ, first at row %d col %d" % (
                lvl, len(d), d[0][0] - 32, d[0][1] - 32)
    finally:
        D.close()


CASES = [("I", 252), ("P", 252), ("P", 1200), ("I", 40), ("I", 1700)]


@need_ref
@pytest.mark.parametrize("geom", GEOM + ODDC[:1], ids=[g[0] for g in GEOM + ODDC[:1]])
@pytest.mark.parametrize("mode,q", CASES[:4])
def test_inv_sbt_lossy_emulated(geom, mode, q):
    util.ensure_emu()
    _inv(geom, True, mode == "P", q, 5)


@need_ref
@pytest.mark.parametrize("geom", GEOM + ODDC[:1], ids=[g[0] for g in GEOM + ODDC[:1]])
@pytest.mark.parametrize("mode,q", CASES[:3])
def test_dequant_emulated(geom, mode, q):
    util.ensure_emu()
    _dequant(geom, True, mode == "P", q)


@need_ref
@pytest.mark.parametrize("geom", GEOM[:2] + ODDC[:1], ids=[g[0] for g in GEOM[:2] + ODDC[:1]])
def test_intra_filter_emulated(geom):
    util.ensure_emu()
    _intra_filter(geom, True, 400)


@need_ref
@pytest.mark.parametrize("geom", GEOM + ODDC, ids=[g[0] for g in GEOM + ODDC])
def test_extend_and_pyramid_emulated(geom):
    util.ensure_emu()
    _extend_and_pyramid(geom, True)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + ODDC + BIG, ids=[g[0] for g in GEOM + ODDC + BIG])
@pytest.mark.parametrize("mode,q", CASES)
def test_inv_sbt_lossy_gpu(geom, mode, q):
    _inv(geom, False, mode == "P", q, 9)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + ODDC + BIG, ids=[g[0] for g in GEOM + ODDC + BIG])
@pytest.mark.parametrize("mode,q", CASES)
def test_dequant_gpu(geom, mode, q):
    _dequant(geom, False, mode == "P", q)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + ODDC + BIG[:2], ids=[g[0] for g in GEOM + ODDC + BIG[:2]])
@pytest.mark.parametrize("q", [252, 900])
def test_intra_filter_gpu(geom, q):
    _intra_filter(geom, False, q)


@pytest.mark.gpu
@pytest.mark.parametrize("geom", GEOM + ODDC + BIG, ids=[g[0] for g in GEOM + ODDC + BIG])
def test_extend_and_pyramid_gpu(geom):
    _extend_and_pyramid(geom, False)
