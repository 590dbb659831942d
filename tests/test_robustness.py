"""Damaged streams: the decoder must behave like the reference's (log, keep
going, same frames) -- flipped bytes inside coefficient planes and motion
data, a destroyed end-of-plane marker (hzcc.c:636-639), a truncated stream --
and must reject a bad packet start code (dsv_decoder.c:32-35)."""
import ctypes as C
import os
import subprocess

import pytest

import ops
import util

need_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")


def _mutations(P, data):
    pk = P.split_packets(bytes(data))
    off1 = len(pk[0])
    off2 = off1 + len(pk[1])

    def flip(at, mask):
        def f(d):
            d[at] ^= mask
        return f

    def trunc(d):
        del d[off2 + len(pk[2]) + 100:]

    return [("flip_plane", flip(off1 + len(pk[1]) - 200, 0xff)), ("flip_p", flip(off2 + len(pk[2]) // 2, 0x55)),
            ("flip_motion", flip(off2 + 40, 0x10)), ("bad_eop", flip(off1 + len(pk[1]) - 1, 0x55)), ("trunc", trunc)]


def _run(emu):
    P = util.pkg()
    y4m = util.clip("cif", 352, 288, 8, "420")
    dsv = util.ref_encode(y4m, ["-qp=60", "-gop=4"], "corrupt")
    data = bytearray(open(dsv, "rb").read())
    for name, mutate in _mutations(P, data):
        d = bytearray(data)
        mutate(d)
        path = dsv[:-4] + "_" + name + ".dsv"
        open(path, "wb").write(d)
        out = path[:-4] + "_dec.y4m"
        if os.path.exists(out):
            os.remove(out)
        subprocess.run([util.REF_BIN, "d", "-y", "-inp=" + path, "-out=" + out, "-y4m=1"], stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        _, _, rf = util.read_y4m(out)
        lib = P.load(emu)
        lib.dsv_set_log_level(0)
        meta, nfr, dec = P.decode_frames(bytes(d), emu=emu)
        lib.dsv_set_log_level(1)
        assert nfr == len(rf), name
        assert dec == b"".join(ops.yuv_bytes(f) for f in rf), name


@need_ref
def test_damaged_streams_emulated():
    util.ensure_emu()
    _run(True)


@pytest.mark.gpu
def test_damaged_streams_gpu():
    _run(False)


def test_bad_start_code_is_an_error():
    util.ensure_emu()
    P = util.pkg()
    lib = P.load(True)
    lib.dsv_set_log_level(0)
    dec = P.DSV_DECODER()
    buf = P.DSV_BUF()
    pkt = b"XSV2" + bytes(20)
    lib.dsv_mk_buf(C.byref(buf), len(pkt))
    C.memmove(buf.data, pkt, len(pkt))
    fr = C.POINTER(P.DSV_FRAME)()
    fno = C.c_uint32()
    assert lib.dsv_dec(C.byref(dec), C.byref(buf), C.byref(fr), C.byref(fno)) == P.DEC_ERROR
    assert not fr
    # a picture packet before any metadata is skipped, not an error (dsv_decoder.c:436-440)
    pkt = b"DSV2" + bytes([8, 0x06]) + bytes(8) + bytes(64)
    lib.dsv_mk_buf(C.byref(buf), len(pkt))
    C.memmove(buf.data, pkt, len(pkt))
    assert lib.dsv_dec(C.byref(dec), C.byref(buf), C.byref(fr), C.byref(fno)) == P.DEC_OK
    assert not fr
    lib.dsv_set_log_level(1)
