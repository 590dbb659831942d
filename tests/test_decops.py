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

P = util.pkg()
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
        wants, counts = [], []
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
            wants.append(want)
            counts.append(n)
        # the same bytes through the device parser (k_hzcc_parse): a batch of two "pictures"
        # (the three planes twice), de-quantised without the symbols leaving the device
        keep = [ops._buf(planes[p][1]) for p in range(3)]
        pl = (P.DSVCU_PLANE_BITS * 6)()
        for i in range(6):
            cw, ch = D.coef_dims(k, i % 3)
            pl[i].bits = C.cast(keep[i % 3], C.c_void_p)
            pl[i].len = len(planes[i % 3][1])
            pl[i].w, pl[i].h = cw, ch
        okv = (C.c_int * 6)()
        # two batches in flight at once (the decoder parses one batch ahead): the first one
        # holds the planes in reverse order
        plr = (P.DSVCU_PLANE_BITS * 3)()
        for i in range(3):
            plr[i].bits, plr[i].len, plr[i].w, plr[i].h = pl[2 - i].bits, pl[2 - i].len, pl[2 - i].w, pl[2 - i].h
        set_r = lib.dsvcu_parse_begin(D.ctx, plr, 3, 3, None, 0, 0)
        set_f = lib.dsvcu_parse_begin(D.ctx, pl, 6, 2, None, 0, 0)  # in two parts: planes 0-1, planes 2-5
        assert sorted((set_r, set_f)) == [0, 1], lib.dsvcu_last_error()
        assert lib.dsvcu_parse_begin(D.ctx, pl, 6, 6, None, 0, 0) < 0  # a third one has nowhere to go
        okr = (C.c_int * 3)()
        D.ck(lib.dsvcu_parse_end(D.ctx, set_r, 0, okr, None))
        D.ck(lib.dsvcu_parse_end(D.ctx, set_f, 1, okv, None))
        assert list(okv) == [0, 0, 1, 1, 1, 1]  # (only the entries of the part are written)
        assert lib.dsvcu_dequant_parsed(D.ctx, k, q, C.byref(fm), set_f, 0) != 0  # part 0 not collected yet
        D.ck(lib.dsvcu_parse_end(D.ctx, set_f, 0, okv, None))
        assert list(okv) == [1] * 6 and list(okr) == [1] * 3
        assert [lib.dsvcu_parsed_count(D.ctx, set_f, i) for i in range(6)] == counts * 2
        assert [lib.dsvcu_parsed_count(D.ctx, set_r, i) for i in range(3)] == counts[::-1]
        for first in (3, 0):
            D.ck(lib.dsvcu_dequant_parsed(D.ctx, k, q, C.byref(fm), set_f, first))
            for p in range(3):
                cw, ch = D.coef_dims(k, p)
                got = D.coefs_download(k, p).reshape(-1)
                d = np.nonzero(got != wants[p])[0]
                assert len(d) == 0, "device-parsed plane %d: %d coefficients differ, first at %d got %d want %d" % (
                    p, len(d), d[0], got[d[0]], wants[p][d[0]])
        if D.coef_dims(k, 0) != D.coef_dims(k, 2):  # planes of another geometry are refused
            assert lib.dsvcu_dequant_parsed(D.ctx, k, q, C.byref(fm), set_r, 0) != 0
        # damaged planes are handed back to the host parser: cut short, end marker gone, length word wrong
        bits = planes[0][1]
        cw, ch = D.coef_dims(k, 0)
        bad = [bits[:len(bits) // 2], bits[:-1] + b"\x54", bits[:3] + bytes([bits[3] ^ 1]) + bits[4:],
               bits[:12] + bytes(len(bits) - 12)]
        keep2 = [ops._buf(b) for b in bad]
        pl2 = (P.DSVCU_PLANE_BITS * len(bad))()
        for i, b in enumerate(bad):
            pl2[i].bits = C.cast(keep2[i], C.c_void_p)
            pl2[i].len = len(b)
            pl2[i].w, pl2[i].h = cw, ch
        ok2 = (C.c_int * len(bad))()
        assert lib.dsvcu_parse_planes(D.ctx, pl2, len(bad), ok2) >= 0, lib.dsvcu_last_error()
        assert list(ok2) == [0] * len(bad)
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


def _parser_fuzz(emu, rounds, w=96, h=64):
    """Random symbol lists of every flavour (dense / sparse, tiny / huge magnitudes, so that the
    Rice parameter wanders through and beyond the rows of the prefix table, runs longer than the
    table knows) written by the host coder; the device parser (k_hzcc_parse) and the host parser
    must put the same coefficients into the planes (lossless de-quantiser: coefficient = value)."""
    cfg = _cfg(w, h, "444", isP=0, fnum=0, lossless=1)
    D = ops.Dev(cfg, emu)
    lib = D.lib
    rng = np.random.default_rng(1234)
    try:
        fm = cfg.fmeta()
        assert fm.lossless == 1
        k = D.coefs()
        cw, ch = D.coef_dims(k, 0)
        part = (C.c_int * 5)()
        total = lib.dsvcu_scan_layout(cw, ch, part)
        for rnd in range(rounds):
            blobs, syms_all = [], []
            for p in range(3):
                while True:
                    density = [0.002, 0.05, 0.4, 0.95][int(rng.integers(0, 4))]
                    npos = max(1, int(total * density))
                    pos = np.sort(rng.choice(np.arange(1, total), size=min(npos, total - 1), replace=False)).astype(np.uint32)
                    scale = [1, 3, 40, 300, 3000][int(rng.integers(0, 5))]
                    mag = 1 + np.floor(rng.exponential(scale, size=len(pos))).astype(np.int64)
                    val = (mag * rng.choice([-1, 1], size=len(pos))).astype(np.int32)
                    if rnd % 3 == 0 and len(pos) > 4:   # a stretch of small values between large ones
                        val[len(val) // 3:len(val) // 2] = np.sign(val[len(val) // 3:len(val) // 2])
                    sy = np.zeros(len(pos), dtype=[("pos", np.uint32), ("v", np.int32)])
                    sy["pos"], sy["v"] = pos, val
                    dc = int(rng.integers(-500, 500))
                    out = (C.c_uint8 * (len(pos) * 600 + 4096))()
                    n = lib.dsv_hzcc_pack_plane(sy.ctypes.data_as(C.c_void_p), len(sy), dc, cw, ch, out, len(out))
                    assert n > 0
                    if n - 4 < cw * ch * 8:   # (longer planes are refused by every parser: hzcc.c:600)
                        break
                blobs.append(bytes(out[:n]))
                syms_all.append((sy, dc))
            keep = [ops._buf(b) for b in blobs]
            pl = (P.DSVCU_PLANE_BITS * 3)()
            for i in range(3):
                pl[i].bits = C.cast(keep[i], C.c_void_p)
                pl[i].len = len(blobs[i])
                pl[i].w, pl[i].h = cw, ch
            okv = (C.c_int * 3)()
            st = lib.dsvcu_parse_planes(D.ctx, pl, 3, okv)
            assert st >= 0 and list(okv) == [1, 1, 1], (rnd, list(okv))
            assert [lib.dsvcu_parsed_count(D.ctx, st, i) for i in range(3)] == [len(s[0]) for s in syms_all]
            D.ck(lib.dsvcu_dequant_parsed(D.ctx, k, 60, C.byref(fm), st, 0))
            got = [D.coefs_download(k, p).reshape(-1).copy() for p in range(3)]
            for p in range(3):
                cap = C.c_int()
                stg = lib.dsvcu_symbol_staging(D.ctx, p, C.byref(cap))
                lstart = (C.c_int * 5)()
                dcv = C.c_int()
                n = lib.dsv_hzcc_unpack_plane(ops._buf(blobs[p]), len(blobs[p]), stg, cap.value - 1, cw, ch, lstart, C.byref(dcv))
                assert n == len(syms_all[p][0]) and dcv.value == syms_all[p][1]
                host_syms = np.ctypeslib.as_array(C.cast(stg, C.POINTER(C.c_int32)), shape=(n, 2)).copy()
                assert (host_syms[:, 0].astype(np.uint32) == syms_all[p][0]["pos"]).all()
                assert (host_syms[:, 1] == syms_all[p][0]["v"]).all()
                D.ck(lib.dsvcu_dequant_plane(D.ctx, k, p, 60, C.byref(fm), n, lstart, dcv.value))
                want = D.coefs_download(k, p).reshape(-1)
                d = np.nonzero(got[p] != want)[0]
                assert len(d) == 0, "round %d plane %d: %d coefficients differ, first at %d got %d want %d" % (
                    rnd, p, len(d), d[0], got[p][d[0]], want[d[0]])
    finally:
        D.close()


def test_device_parser_fuzz_emulated():
    util.ensure_emu()
    _parser_fuzz(True, 12)


@pytest.mark.gpu
def test_device_parser_fuzz_gpu():
    _parser_fuzz(False, 12)
