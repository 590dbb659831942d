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


# ---- regressions for the round-1 advisor findings (host logic, run in emulation)

def _tiny_stream(P, w, h, n, **kw):
    import numpy as np
    rng = np.random.default_rng(w * 131 + h)
    fsz = w * h * 3 // 2
    yuv = rng.integers(0, 256, n * fsz, dtype=np.uint8).tobytes()
    o = P.enc_opts(w, h, P.SUBSAMP_420, (30, 1), emu=True, qp=60, gop=4, **kw)
    return P.encode_frames(o, yuv, n, emu=True), yuv


def test_geometry_change_inside_a_stream_is_refused():
    """dsv_pipe.c sizes its output from the stream's metadata: a stream whose later
    metadata packet announces a bigger picture must be refused, not decoded into the
    first size (heap overflow in round 1)."""
    util.ensure_emu()
    P = util.pkg()
    lib = P.load(True)
    a, _ = _tiny_stream(P, 128, 96, 3, noeos=1)
    b, _ = _tiny_stream(P, 256, 192, 3)
    lib.dsv_set_log_level(0)
    try:
        with pytest.raises(RuntimeError):
            P.decode_frames(a + b, emu=True)
        with pytest.raises(RuntimeError):
            P.decode_frames(a + b, emu=True, threads=2)
        # the two halves on their own are fine, and so is the same geometry twice
        assert P.decode_frames(a, emu=True)[1] == 3
        assert P.decode_frames(a + a, emu=True, threads=2)[1] == 6
    finally:
        lib.dsv_set_log_level(1)


def test_pool_recycles_device_state_only_for_the_same_block_geometry():
    """one persistent pool, same resolution, block-size override changed between
    calls: the second encoder must not inherit buffers sized for fewer blocks"""
    util.ensure_emu()
    P = util.pkg()
    lib = P.load(True)
    import numpy as np
    w, h, n = 256, 192, 4
    fsz = w * h * 3 // 2
    yuv = np.random.default_rng(5).integers(0, 256, n * fsz, dtype=np.uint8).tobytes()
    buf = (C.c_uint8 * len(yuv)).from_buffer_copy(yuv)
    devs = (C.c_int * 1)(0)
    pool = lib.dsv_pool_create(1, devs, 1)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    try:
        got = {}
        for bs in (1, 0, 1, -1):
            o = P.enc_opts(w, h, P.SUBSAMP_420, (30, 1), emu=True, qp=60, gop=2, noeos=1, bszx=bs, bszy=bs)
            out, outn = C.c_void_p(), C.c_size_t()
            assert lib.dsv_pool_encode(pool, C.byref(o), buf, n, 2, C.byref(out), C.byref(outn)) == 0
            got[bs] = C.string_at(out, outn.value)
            libc.free(out)
            # same bytes as a fresh one-shot sharded encode with these options
            assert got[bs] == P.encode_frames(o, yuv, n, emu=True, chunk=2)
    finally:
        lib.dsv_pool_destroy(pool)


def test_corrupt_side_information_lengths_are_contained():
    """every byte of the picture header / first sub-stream lengths forced to 0x00, 0x01
    and 0xff: the decoder must come back (error or damaged picture) without touching
    memory outside the packet.  Run under the emulation library: a wild read shows
    up as a crash of the test process."""
    util.ensure_emu()
    P = util.pkg()
    lib = P.load(True)
    data, _ = _tiny_stream(P, 128, 96, 3)
    pk = P.split_packets(data)
    pics = [i for i, p in enumerate(pk) if p[5] & 0x04]
    assert len(pics) == 3
    lib.dsv_set_log_level(0)
    try:
        for pi in (pics[0], pics[1]):  # an intra and a predicted picture
            base = sum(len(p) for p in pk[:pi])
            for at in range(14, 14 + 48):
                for val in (0x00, 0x01, 0xff):
                    d = bytearray(data)
                    d[base + at] = val
                    try:
                        P.decode_frames(bytes(d), emu=True)
                    except RuntimeError:
                        pass
        # metadata with absurd dimensions is refused as well
        d = bytearray(data)
        for at in range(14, 20):
            d[at] = 0x00
        with pytest.raises(RuntimeError):
            P.decode_frames(bytes(d), emu=True)
    finally:
        lib.dsv_set_log_level(1)


def _device_vs_host_parser(emu, rounds, seed):
    """Whatever is in a picture packet, the whole-stream decoder must produce the same frames
    with the coefficient planes and side information parsed on the device (k_hzcc_parse, which
    refuses what it does not trust) and with everything parsed by the host threads: random
    damage to the payload of random pictures of a 150-picture stream (single bits, bytes, short
    stretches; the packet headers and links stay intact)."""
    import random
    P = util.pkg()
    lib = P.load(emu)
    lib.dsv_set_log_level(0)
    dsv = util.ref_encode(util.clip("manybatches", 96, 64, 150, "420"), ["-qp=40", "-gop=35"], "ref")
    data = bytearray(open(dsv, "rb").read())
    pk = P.split_packets(bytes(data))
    offs, o = [], 0
    for p in pk:
        offs.append(o)
        o += len(p)
    rng = random.Random(seed)
    try:
        for it in range(rounds):
            d = bytearray(data)
            for _ in range(rng.randint(1, 6)):
                k = rng.randrange(1, len(pk) - 1)
                if len(pk[k]) < 40:
                    continue
                pos = offs[k] + rng.randrange(14, len(pk[k]))
                mode = rng.randrange(3)
                if mode == 0:
                    d[pos] ^= 1 << rng.randrange(8)
                elif mode == 1:
                    d[pos] = rng.randrange(256)
                else:
                    for j in range(pos, min(pos + rng.randint(1, 12), offs[k] + len(pk[k]))):
                        d[j] = rng.randrange(256)
            got = []
            for device_entropy in (0, 1):
                try:
                    _, nfr, frames = P.decode_frames(bytes(d), emu=emu, device_entropy=device_entropy)
                    got.append((nfr, frames))
                except RuntimeError as e:
                    got.append(("error", str(e)))
            assert got[0] == got[1], "round %d: host parser %r frames, device parser %r" % (it, got[0][0], got[1][0])
    finally:
        lib.dsv_set_log_level(1)


@need_ref
def test_device_and_host_parser_agree_on_damaged_streams_emulated():
    util.ensure_emu()
    _device_vs_host_parser(True, 40, 11)


@pytest.mark.gpu
def test_device_and_host_parser_agree_on_damaged_streams_gpu():
    _device_vs_host_parser(False, 40, 12)
