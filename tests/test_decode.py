"""Whole-stream decode parity: dsv_dec() of the B200 library vs the reference
decoder (oracle/_ref/dsv2 d) on streams produced by the reference encoder."""
import ctypes as C

import pytest

import util

CASES = [
    # name, w, h, frames, fmt, encoder args  (BASELINE.json configs, shortened)
    ("cif", 352, 288, 20, "420", ["-qp=60", "-gop=48"]),
    ("cif_lowq", 352, 288, 8, "420", ["-qp=20", "-gop=4"]),
    ("odd", 200, 136, 6, "420", ["-qp=70", "-gop=3"]),
    ("hd", 1280, 720, 4, "420", ["-gop=250", "-effort=10"]),
    ("fhd", 1920, 1080, 4, "420", ["-qp=60", "-gop=48"]),
    ("fhd444ll", 1920, 1080, 2, "444", ["-qp=100"]),
    ("cif444", 352, 288, 6, "444", ["-qp=50", "-gop=5"]),
]
SMALL = [c for c in CASES if c[1] <= 352]


def _run(case, emu):
    name, w, h, n, fmt, args = case
    y4m = util.clip(name, w, h, n, fmt)
    dsv = util.ref_encode(y4m, args, "ref")
    _, _, ref = util.read_y4m(util.ref_decode(dsv))
    meta, got = util.pkg().decode_stream(open(dsv, "rb").read(), emu=emu)
    assert meta["width"] == w and meta["height"] == h
    util.assert_same_frames(got, ref, w, h)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
def test_decode_emulated_kernels(case):
    """CPU-only: host logic + kernel arithmetic (test-only host emulation)."""
    util.ensure_emu()
    _run(case, emu=True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_decode_gpu(case):
    _run(case, emu=False)


def _whole_stream_vs_packetwise(emu):
    """dsv_decode_sharded hands the coefficient planes to the device parser (k_hzcc_parse)
    in batches of up to 64 pictures, one batch ahead of the reconstruction; dsv_dec() packet by
    packet parses on the host.  Both against the reference decoder, on a segment of several
    batches (150 pictures, one metadata packet) with intra pictures in the middle of a batch."""
    w, h, n = 96, 64, 150
    y4m = util.clip("manybatches", w, h, n, "420")
    dsv = util.ref_encode(y4m, ["-qp=40", "-gop=35"], "ref")
    data = open(dsv, "rb").read()
    _, _, ref = util.read_y4m(util.ref_decode(dsv))
    P = util.pkg()
    _, packetwise = P.decode_stream(data, emu=emu)
    util.assert_same_frames(packetwise, ref, w, h)
    lib = P.load(emu)
    sizes = sorted(len(p) for p in P.split_packets(data)[1:])
    try:
        # the size model; limits that send the larger half of the pictures into the second part
        # of their batch and keep the largest on the host; everything in the first part
        for limits in ((0, 0, 0), (sizes[len(sizes) // 2], sizes[-8], 2), (1 << 30, 0, 0)):
            lib.dsv_set_device_entropy_limits(*limits)
            for device_entropy in (1, 0, -1):
                meta, nfr, whole = P.decode_frames(data, emu=emu, device_entropy=device_entropy)
                assert nfr == n
                assert whole == b"".join(b"".join(f) for f in packetwise), "device_entropy=%d limits %r" % (
                    device_entropy, limits)
        if emu:
            # the host takes over pictures whose part of the batch is "not through yet"
            # (emulation-only hook: every n-th question is answered that way; 1 = always, so
            # parts stay uncollected until their slot is needed again)
            lib.dsvcu_emu_parse_not_ready.argtypes = [C.c_int]
            for limits in ((0, 0, 0), (sizes[len(sizes) // 2], sizes[-8], 2), (sizes[len(sizes) // 4], sizes[-2], 0)):
                lib.dsv_set_device_entropy_limits(*limits)
                for every in (1, 2, 3, 5, -1):  # (-1: the second part of every batch is never through)
                    lib.dsvcu_emu_parse_not_ready(every)
                    meta, nfr, whole = P.decode_frames(data, emu=emu, device_entropy=1)
                    assert nfr == n and whole == b"".join(b"".join(f) for f in packetwise), "not ready every %d, limits %r" % (
                        every, limits)
    finally:
        lib.dsv_set_device_entropy_limits(0, 0, 0)
        if emu:
            lib.dsvcu_emu_parse_not_ready(0)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_whole_stream_device_parser_emulated():
    util.ensure_emu()
    _whole_stream_vs_packetwise(True)


@pytest.mark.gpu
def test_whole_stream_device_parser_gpu():
    _whole_stream_vs_packetwise(False)


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref not built")
def test_uncollected_part_does_not_cost_the_next_batches_their_slot():
    """A part of a batch that nobody waited for (all its pictures were parsed on the host, because
    they were reached first) is still pending when its slot comes round again, two batches later --
    while the batch in the other slot is the one being decoded.  Batches of 64 pictures; with
    fixed limits (second part: intra-sized packets at least 50 pictures into their batch) only the
    second batch of this clip has a second part, which the emulation hook keeps "not through"."""
    util.ensure_emu()
    w, h, n = 96, 64, 200
    y4m = util.clip("slots", w, h, n, "420")
    dsv = util.ref_encode(y4m, ["-qp=40", "-gop=48"], "ref")
    data = open(dsv, "rb").read()
    P = util.pkg()
    lib = P.load(True)
    _, packetwise = P.decode_stream(data, emu=True)
    want = b"".join(b"".join(f) for f in packetwise)
    pk = [p for p in P.split_packets(data)[1:] if len(p) > 20]
    intra = min(len(p) for p in pk if not (p[5] & 1))
    assert max(len(p) for p in pk if p[5] & 1) < intra
    late = [[i for i in range(b, min(b + 64, len(pk))) if len(pk[i]) >= intra and i - b >= 50] for b in range(0, len(pk), 64)]
    assert len(late) >= 4 and not late[0] and late[1] and not late[2], late
    lib.dsvcu_emu_parse_not_ready.argtypes = [C.c_int]
    try:
        lib.dsv_set_device_entropy_limits(intra - 1, 1 << 30, 50)
        lib.dsvcu_emu_parse_not_ready(-1)
        meta, nfr, whole = P.decode_frames(data, emu=True, device_entropy=1)
        assert nfr == n and whole == want
    finally:
        lib.dsv_set_device_entropy_limits(0, 0, 0)
        lib.dsvcu_emu_parse_not_ready(0)
