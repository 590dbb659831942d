"""Operator-level drivers used by the parity tests.

`Ref` calls the unmodified reference operators through oracle/_ref/librefops.so
(test infrastructure); `Dev` calls the product C ABI (dsv_cuda.h) -- the CUDA
library on the GPU box, or the test-only host emulation of the kernel sources in
the CPU suite.  Both take/return tightly packed planar YUV and DSV_MV arrays.
"""
import ctypes as C
import os

import numpy as np

import util

MV_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("flags", "<u4"), ("err", "<u2"),
                     ("dc", "<u2"), ("submask", "u1"), ("pad", "u1", 3)])
assert MV_DTYPE.itemsize == 16

F_INTRA, F_EPRM, F_MAINTAIN, F_SKIP, F_RINGING, F_NOXMITY, F_NOXMITC, F_SIMCMPLX = [1 << i for i in range(8)]


class RefCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("w", "h", "subsamp", "fps_num", "fps_den", "effort", "do_psy", "blk_w", "blk_h",
                 "temporal_mc", "lossless", "inter_sharpen", "skip_thresh", "pyramid_levels", "isP")] + \
               [("fnum", C.c_uint)]


def lb2(n):
    l, i = 0, 1
    while i < n:
        i <<= 1
        l += 1
    return l


class Cfg:
    """Per-picture parameters as encode_one_frame derives them
    (reference dsv_encoder.c:1193-1241)."""

    def __init__(self, w, h, subsamp=0x5, effort=10, do_psy=255, isP=1, fnum=1, lossless=0,
                 skip_thresh=0, fps=(30, 1), inter_sharpen=1, blk=None):
        self.w, self.h, self.subsamp = w, h, subsamp
        bw = 32 if w > 1280 else 16
        bh = 32 if h > 1280 else 16
        if abs(w - h) < min(w, h):
            bw = bh = min(bw, bh)
        if blk:
            bw, bh = blk
        self.blk_w, self.blk_h = bw, bh
        self.nbh = (w + bw - 1) // bw
        self.nbv = (h + bh - 1) // bh
        self.nblk = self.nbh * self.nbv
        lv = lb2(min(w, h))
        while (1 << lv) > max(self.nbh, self.nbv):
            lv -= 1
        self.pyr = min(max(lv, 3), 5)
        self.effort, self.do_psy, self.isP, self.fnum = effort, do_psy, isP, fnum
        self.lossless, self.skip_thresh, self.fps, self.inter_sharpen = lossless, skip_thresh, fps, inter_sharpen
        self.hs = (subsamp >> 2) & 3
        self.vs = subsamp & 3
        self.cw = (w + (1 << self.hs) - 1) >> self.hs
        self.ch = (h + (1 << self.vs) - 1) >> self.vs

    def plane_dims(self, p):
        return (self.w, self.h) if p == 0 else (self.cw, self.ch)

    def frame_bytes(self):
        return self.w * self.h + 2 * self.cw * self.ch

    def ref(self):
        return RefCfg(self.w, self.h, self.subsamp, self.fps[0], self.fps[1], self.effort, self.do_psy,
                      self.blk_w, self.blk_h, self.fnum % 2 if self.isP else 0, self.lossless,
                      self.inter_sharpen, self.skip_thresh, self.pyr, self.isP, self.fnum)

    def fmeta(self):
        m = util.pkg().DSVCU_FMETA()
        m.isP, m.lossless, m.do_psy = self.isP, self.lossless, self.do_psy
        m.blk_w, m.blk_h, m.nblocks_h, m.nblocks_v = self.blk_w, self.blk_h, self.nbh, self.nbv
        m.temporal_mc = self.fnum % 2 if self.isP else 0
        m.inter_sharpen, m.effort, m.fnum = self.inter_sharpen, self.effort, self.fnum
        return m


def yuv_bytes(frame):
    """(Y,U,V) bytes tuple -> one packed buffer."""
    return b"".join(frame)


def _buf(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(b)


class Ref:
    def __init__(self):
        path = os.path.join(util.REF_DIR, "librefops.so")
        self.lib = C.CDLL(path)
        assert self.lib.refop_sizeof_mv() == 16

    def hme(self, cfg, src, ref, ogr, prev_mvs, quant):
        out = np.zeros(cfg.nblk, MV_DTYPE)
        o3 = (C.c_int * 3)()
        rc = cfg.ref()
        pm = prev_mvs.ctypes.data_as(C.c_void_p) if prev_mvs is not None else None
        self.lib.refop_hme(C.byref(rc), _buf(src), _buf(ref), _buf(ogr), pm, quant,
                           out.ctypes.data_as(C.c_void_p), o3)
        return out, tuple(o3)

    def intra_analysis(self, cfg, src):
        out = np.zeros(cfg.nblk, MV_DTYPE)
        rc = cfg.ref()
        self.lib.refop_intra_analysis(C.byref(rc), _buf(src), out.ctypes.data_as(C.c_void_p))
        return out

    def coef_dims(self, cfg, plane):
        w, h = C.c_int(), C.c_int()
        rc = cfg.ref()
        self.lib.refop_coef_dims(C.byref(rc), plane, C.byref(w), C.byref(h))
        return w.value, h.value

    def fwd_sbt(self, cfg, plane, yuv, blockdata):
        w, h = self.coef_dims(cfg, plane)
        out = np.zeros(w * h, np.int32)
        rc = cfg.ref()
        self.lib.refop_fwd_sbt(C.byref(rc), plane, _buf(yuv), _buf(bytes(blockdata)), out.ctypes.data_as(C.c_void_p))
        return out.reshape(h, w)

    def encode_plane(self, cfg, plane, q, coefs, blockdata, mvs):
        k = np.ascontiguousarray(coefs, np.int32).copy()
        bits = (C.c_uint8 * (k.size * 8 + 1024))()
        n = C.c_int()
        rc = cfg.ref()
        self.lib.refop_encode_plane(C.byref(rc), plane, q, k.ctypes.data_as(C.c_void_p), _buf(bytes(blockdata)),
                                    mvs.ctypes.data_as(C.c_void_p), bits, C.byref(n))
        return k, bytes(bits[:n.value])

    def inv_sbt(self, cfg, plane, q, coefs, blockdata):
        w, h = cfg.plane_dims(plane)
        out = (C.c_uint8 * (w * h))()
        k = np.ascontiguousarray(coefs, np.int32)
        rc = cfg.ref()
        self.lib.refop_inv_sbt(C.byref(rc), plane, q, k.ctypes.data_as(C.c_void_p), _buf(bytes(blockdata)), out)
        return bytes(out)

    def sub_pred(self, cfg, mvs, src, ref):
        n = cfg.frame_bytes()
        pred, resd = (C.c_uint8 * n)(), (C.c_uint8 * n)()
        rc = cfg.ref()
        self.lib.refop_sub_pred(C.byref(rc), mvs.ctypes.data_as(C.c_void_p), _buf(src), _buf(ref), pred, resd)
        return bytes(pred), bytes(resd)

    def add_res(self, cfg, mvs, blockdata, q, resd, pred, do_filter):
        r = _buf(resd)
        rc = cfg.ref()
        self.lib.refop_add_res(C.byref(rc), mvs.ctypes.data_as(C.c_void_p), _buf(bytes(blockdata)), q, r, _buf(pred),
                               do_filter)
        return bytes(r)

    def intra_filter(self, cfg, q, blockdata, yuv, do_filter):
        r = _buf(yuv)
        rc = cfg.ref()
        self.lib.refop_intra_filter(C.byref(rc), q, _buf(bytes(blockdata)), r, do_filter)
        return bytes(r)


class Dev:
    """One dsvcu context of a fixed geometry."""

    def __init__(self, cfg, emu):
        self.P = util.pkg()
        self.lib = self.P.load(emu)
        self.cfg = cfg
        ctx = C.c_void_p()
        if self.lib.dsvcu_ctx_create(C.byref(ctx), 0, cfg.w, cfg.h, cfg.subsamp):
            raise RuntimeError(self.lib.dsvcu_last_error().decode())
        self.ctx = ctx
        self._frames, self._pyr, self._coefs = [], [], []

    def ck(self, r):
        if r:
            raise RuntimeError(self.lib.dsvcu_last_error().decode())

    def close(self):
        for f in self._frames:
            self.lib.dsvcu_frame_destroy(self.ctx, f)
        for p in self._pyr:
            self.lib.dsvcu_pyramid_destroy(self.ctx, p)
        for k in self._coefs:
            self.lib.dsvcu_coefs_destroy(self.ctx, k)
        self.lib.dsvcu_ctx_destroy(self.ctx)

    def frame(self, yuv=None, extend=True):
        f = C.c_void_p()
        self.ck(self.lib.dsvcu_frame_create(self.ctx, C.byref(f)))
        self._frames.append(f)
        if yuv is not None:
            self.upload(f, yuv)
            if extend:
                self.ck(self.lib.dsvcu_extend_frame(self.ctx, f, 0))
        return f

    def upload(self, f, yuv):
        off = 0
        keep = _buf(yuv)
        for p in range(3):
            w, h = self.cfg.plane_dims(p)
            self.ck(self.lib.dsvcu_frame_upload(self.ctx, f, p, C.byref(keep, off), w))
            off += w * h
        self.ck(self.lib.dsvcu_sync(self.ctx))

    def download(self, f):
        n = self.cfg.frame_bytes()
        out = (C.c_uint8 * n)()
        off = 0
        for p in range(3):
            w, h = self.cfg.plane_dims(p)
            self.ck(self.lib.dsvcu_frame_download(self.ctx, f, p, C.byref(out, off), w))
            off += w * h
        self.ck(self.lib.dsvcu_sync(self.ctx))
        return bytes(out)

    def pyramid(self, base):
        p = C.c_void_p()
        self.ck(self.lib.dsvcu_pyramid_create(self.ctx, C.byref(p), self.cfg.pyr))
        self._pyr.append(p)
        self.ck(self.lib.dsvcu_pyramid_build(self.ctx, p, base))
        return p

    def coefs(self):
        k = C.c_void_p()
        self.ck(self.lib.dsvcu_coefs_create(self.ctx, C.byref(k)))
        self._coefs.append(k)
        return k

    def set_blockdata(self, bd):
        self.ck(self.lib.dsvcu_set_blockdata(self.ctx, _buf(bytes(bd)), len(bd)))
        self.ck(self.lib.dsvcu_sync(self.ctx))

    def set_mvs(self, mvs):
        self.ck(self.lib.dsvcu_set_mvs(self.ctx, mvs.ctypes.data_as(C.c_void_p), len(mvs)))
        self.ck(self.lib.dsvcu_sync(self.ctx))

    def hme(self, src, ref, ogr, prev_mvs, quant):
        cfg = self.cfg
        fs, fr, fo = self.frame(src), self.frame(ref), self.frame(ogr)
        ps, pr, po = self.pyramid(fs), self.pyramid(fr), self.pyramid(fo)
        hp = self.P.DSVCU_HME_PARAMS(quant, cfg.skip_thresh, cfg.pyr, 1 if prev_mvs is not None else 0)
        if prev_mvs is not None:
            self.ck(self.lib.dsvcu_set_prev_mvs(self.ctx, prev_mvs.ctypes.data_as(C.c_void_p), cfg.nblk))
        fm = cfg.fmeta()
        self.ck(self.lib.dsvcu_hme(self.ctx, C.byref(fm), C.byref(hp), fs, ps, fr, pr, fo, po))
        out = np.zeros(cfg.nblk, MV_DTYPE)
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.ck(self.lib.dsvcu_hme_fetch(self.ctx, out.ctypes.data_as(C.c_void_p), cfg.nblk, C.byref(a), C.byref(b),
                                         C.byref(c)))
        return out, (a.value, b.value, c.value)

    def intra_analysis(self, src):
        cfg = self.cfg
        fs = self.frame(src)
        out = np.zeros(cfg.nblk, MV_DTYPE)
        fm = cfg.fmeta()
        self.ck(self.lib.dsvcu_intra_analysis(self.ctx, C.byref(fm), fs, out.ctypes.data_as(C.c_void_p), cfg.nblk))
        return out

    def coef_dims(self, k, plane):
        w, h = C.c_int(), C.c_int()
        self.lib.dsvcu_coefs_plane_dims(k, plane, C.byref(w), C.byref(h))
        return w.value, h.value

    def coefs_download(self, k, plane):
        w, h = self.coef_dims(k, plane)
        out = np.zeros(w * h, np.int32)
        self.ck(self.lib.dsvcu_coefs_download(self.ctx, k, plane, out.ctypes.data_as(C.c_void_p)))
        self.ck(self.lib.dsvcu_sync(self.ctx))
        return out.reshape(h, w)

    def coefs_upload(self, k, plane, arr):
        a = np.ascontiguousarray(arr, np.int32)
        self.ck(self.lib.dsvcu_coefs_upload(self.ctx, k, plane, a.ctypes.data_as(C.c_void_p)))
        self.ck(self.lib.dsvcu_sync(self.ctx))

    def fwd_sbt(self, plane, yuv, blockdata):
        f = self.frame(yuv)
        k = self.coefs()
        self.set_blockdata(blockdata)
        fm = self.cfg.fmeta()
        self.ck(self.lib.dsvcu_fwd_sbt(self.ctx, f, plane, k, C.byref(fm)))
        return self.coefs_download(k, plane)

    def quant_plane(self, plane, q, coefs, blockdata, mvs):
        """-> (dequantised coefs, [(pos, v)...], dc)"""
        k = self.coefs()
        self.coefs_upload(k, plane, coefs)
        self.set_blockdata(blockdata)
        self.set_mvs(mvs)
        fm = self.cfg.fmeta()
        self.ck(self.lib.dsvcu_quant_plane(self.ctx, k, plane, q, C.byref(fm)))
        syms = C.POINTER(self.P.DSVCU_SYMBOL)()
        n, dc = C.c_int(), C.c_int()
        self.ck(self.lib.dsvcu_fetch_symbols(self.ctx, plane, C.byref(syms), C.byref(n), C.byref(dc)))
        arr = np.ctypeslib.as_array(C.cast(syms, C.POINTER(C.c_int32)), shape=(max(n.value, 1), 2))[:n.value].copy()
        return self.coefs_download(k, plane), arr, dc.value

    def sub_pred(self, mvs, src, ref):
        fr = self.frame(ref)
        resd = self.frame(src)
        pred = self.frame()
        self.set_mvs(mvs)
        fm = self.cfg.fmeta()
        self.ck(self.lib.dsvcu_sub_pred(self.ctx, C.byref(fm), pred, resd, fr))
        return self.download(pred), self.download(resd)

    def add_res(self, mvs, blockdata, q, resd, pred, do_filter):
        fr = self.frame(resd, extend=False)
        fp = self.frame(pred, extend=False)
        self.set_mvs(mvs)
        self.set_blockdata(blockdata)
        fm = self.cfg.fmeta()
        self.ck(self.lib.dsvcu_add_res(self.ctx, C.byref(fm), q, fr, fp, do_filter))
        return self.download(fr)


def mv_diff(a, b):
    """indices where two DSV_MV arrays differ in any meaningful field"""
    bad = np.zeros(len(a), bool)
    for f in ("x", "y", "flags", "err", "dc", "submask"):
        bad |= a[f] != b[f]
    return np.nonzero(bad)[0]
