"""Persistent worker pool (include/dsv_session.h): frames given in DEVICE
memory and decoded into DEVICE memory (unified addressing) must give the same
bytes as the host-memory path; several chunks in flight on one GPU."""
import ctypes as C

import pytest

import ops
import util


@pytest.mark.gpu
def test_pool_device_memory_paths_match_host_paths():
    torch = pytest.importorskip("torch")
    P = util.pkg()
    lib = P.load()
    w, h, n, chunk = 352, 288, 24, 6
    _, _, fr = util.read_y4m(util.clip("pool", w, h, n, "420"))
    yuv = b"".join(ops.yuv_bytes(f) for f in fr)
    o = P.enc_opts(w, h, P.SUBSAMP_420, (30, 1), qp=60, gop=chunk, noeos=1)
    want = P.encode_frames(o, yuv, n, chunk=chunk, threads=1)

    host = torch.frombuffer(bytearray(yuv), dtype=torch.uint8)
    dev = host.cuda()
    pinned = host.pin_memory()
    devs = (C.c_int * 1)(0)
    pool = lib.dsv_pool_create(4, devs, 1)
    assert pool and lib.dsv_pool_threads(pool) == 4
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    try:
        for ptr in (dev.data_ptr(), pinned.data_ptr()):
            for _ in range(2):  # the pool (contexts, device buffers) is reused between calls
                out, outn = C.c_void_p(), C.c_size_t()
                assert lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(ptr), n, chunk, C.byref(out), C.byref(outn)) == 0
                got = C.string_at(out, outn.value)
                libc.free(out)
                assert got == want
        meta, nfr, ref_frames = P.decode_frames(want, threads=1)
        dsv = (C.c_uint8 * len(want)).from_buffer_copy(want)
        ddev = torch.zeros(len(yuv), dtype=torch.uint8, device="cuda")
        m, k = P.DSV_META(), C.c_int()
        assert lib.dsv_pool_decode(pool, dsv, len(want), C.c_void_p(ddev.data_ptr()), len(yuv), C.byref(k), C.byref(m)) == 0
        torch.cuda.synchronize()
        assert k.value == n and m.width == w and m.height == h
        assert bytes(ddev.cpu().numpy().tobytes()) == ref_frames
        # too small a destination is refused
        assert lib.dsv_pool_decode(pool, dsv, len(want), C.c_void_p(ddev.data_ptr()), len(yuv) - 1, C.byref(k), C.byref(m)) != 0
    finally:
        lib.dsv_pool_destroy(pool)
