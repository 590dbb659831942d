"""dsv2_b200 -- Python (ctypes) view of the B200 DSV2 pixel-path library.

The product is `libdsv2cuda.so` (CUDA, sm_100a + host C).  This module only
binds its C ABI (include/dsv.h, dsv_decoder.h, dsv_encoder.h, dsv_cuda.h) so
tests and bench.py can drive it; it contains no codec arithmetic.

`load()` returns the product library and raises if it is missing -- there is
no CPU fallback.  Tests that run without a GPU may ask for the test-only host
emulation of the kernel sources with `load(emu=True)` (tests/_emu/, built by
`make emu`); the product path never does.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

SUBSAMP_444 = 0x0
SUBSAMP_422 = 0x4
SUBSAMP_420 = 0x5
SUBSAMP_411 = 0x8

DEC_OK, DEC_ERROR, DEC_EOS, DEC_GOT_META, DEC_NEED_NEXT = 0, 1, 2, 3, 4
PACKET_HDR_SIZE = 14


class DSV_META(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("subsamp", C.c_int),
                ("fps_num", C.c_int), ("fps_den", C.c_int),
                ("aspect_num", C.c_int), ("aspect_den", C.c_int),
                ("inter_sharpen", C.c_int), ("reserved", C.c_int)]


class DSV_PLANE(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_int), ("format", C.c_int),
                ("stride", C.c_int), ("w", C.c_int), ("h", C.c_int)]


class DSV_COEFS(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_int32)), ("width", C.c_int), ("height", C.c_int)]


class DSV_FRAME(C.Structure):
    _fields_ = [("alloc", C.POINTER(C.c_uint8)), ("planes", DSV_PLANE * 3),
                ("refcount", C.c_int), ("format", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("border", C.c_int)]


class DSV_MV(C.Structure):
    _fields_ = [("x", C.c_int16), ("y", C.c_int16), ("flags", C.c_uint32),
                ("err", C.c_uint16), ("dc", C.c_uint16), ("submask", C.c_uint8)]


class DSV_PARAMS(C.Structure):
    _fields_ = [("vidmeta", C.POINTER(DSV_META)), ("effort", C.c_int), ("do_psy", C.c_int),
                ("is_ref", C.c_int), ("has_ref", C.c_int),
                ("blk_w", C.c_int), ("blk_h", C.c_int),
                ("nblocks_h", C.c_int), ("nblocks_v", C.c_int),
                ("temporal_mc", C.c_int), ("lossless", C.c_int), ("reserved", C.c_int)]


class DSV_FMETA(C.Structure):
    _fields_ = [("params", C.POINTER(DSV_PARAMS)), ("mvs", C.POINTER(DSV_MV)),
                ("blockdata", C.POINTER(C.c_uint8)), ("cur_plane", C.c_uint8),
                ("isP", C.c_uint8), ("fnum", C.c_uint32)]


class DSV_BUF(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_uint)]


class DSV_IMAGE(C.Structure):
    _fields_ = [("params", DSV_PARAMS), ("out_frame", C.POINTER(DSV_FRAME)),
                ("ref_frame", C.POINTER(DSV_FRAME)), ("blockdata", C.POINTER(C.c_uint8)),
                ("refcount", C.c_int)]


class DSV_DECODER(C.Structure):
    _fields_ = [("vidmeta", DSV_META), ("ref", C.POINTER(DSV_IMAGE)),
                ("draw_info", C.c_int), ("got_metadata", C.c_int)]


class DSVCU_FMETA(C.Structure):
    _fields_ = [("isP", C.c_int), ("lossless", C.c_int), ("do_psy", C.c_int),
                ("blk_w", C.c_int), ("blk_h", C.c_int),
                ("nblocks_h", C.c_int), ("nblocks_v", C.c_int),
                ("temporal_mc", C.c_int), ("inter_sharpen", C.c_int),
                ("effort", C.c_int), ("fnum", C.c_uint)]


class DSV_ENC_OPTS(C.Structure):
    """dsv_enc_opts (include/dsv_session.h): the reference CLI's parameter table"""
    _fields_ = [(n, C.c_int) for n in
                ("w", "h", "fmt", "fps_num", "fps_den", "aspect_num", "aspect_den", "qp", "effort", "gop",
                 "rc_mode", "rc_pergop", "kbps", "minqstep", "maxqstep", "minqp", "maxqp", "iminqp",
                 "stabref", "scd", "tempaq", "bszx", "bszy", "scpct", "skipthresh", "varint", "psy", "dib",
                 "ifilter", "pfilter", "psharp", "ipct", "pyrlevels", "noeos")]


class DSVCU_HME_PARAMS(C.Structure):
    _fields_ = [("quant", C.c_int), ("skip_block_thresh", C.c_int),
                ("pyramid_levels", C.c_int), ("use_prev_mvs", C.c_int)]


class DSVCU_SYMBOL(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("v", C.c_int32)]


class DSVCU_PLANE_BITS(C.Structure):
    _fields_ = [("bits", C.c_void_p), ("len", C.c_uint32), ("w", C.c_int), ("h", C.c_int)]


def lib_path(emu=False):
    if emu:
        return os.path.join(_ROOT, "tests", "_emu", "libdsv2cuda_emu.so")
    return os.path.join(_HERE, "libdsv2cuda.so")


_cache = {}


def load(emu=False):
    """dlopen the library (product by default) and declare prototypes."""
    path = lib_path(emu)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `make%s` (there is no CPU fallback)"
                           % (path, " emu" if emu else ""))
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp, ip = C.c_void_p, C.c_int
    P = C.POINTER
    sig = {
        "dsv_alloc": (vp, [ip]),
        "dsv_free": (None, [vp]),
        "dsv_mk_buf": (None, [P(DSV_BUF), ip]),
        "dsv_buf_free": (None, [P(DSV_BUF)]),
        "dsv_mk_frame": (P(DSV_FRAME), [ip, ip, ip, ip]),
        "dsv_frame_ref_dec": (None, [P(DSV_FRAME)]),
        "dsv_set_log_level": (None, [ip]),
        "dsv_dec": (ip, [P(DSV_DECODER), P(DSV_BUF), P(P(DSV_FRAME)), P(C.c_uint32)]),
        "dsv_dec_free": (None, [P(DSV_DECODER)]),
        "dsv_hzcc_pack_plane": (ip, [vp, ip, ip, ip, ip, vp, ip]),
        "dsv_hzcc_unpack_plane": (ip, [vp, ip, vp, ip, ip, ip, P(ip), P(ip)]),
        "dsv_enc_opts_default": (None, [P(DSV_ENC_OPTS), ip, ip, ip, ip, ip]),
        "dsv_set_thread_device": (None, [ip]),
        "dsv_pinned_alloc": (vp, [C.c_size_t]),
        "dsv_pinned_free": (None, [vp]),
        "dsv_encode_buffer": (ip, [P(DSV_ENC_OPTS), vp, ip, ip, P(vp), P(C.c_size_t)]),
        "dsv_encode_sharded": (ip, [P(DSV_ENC_OPTS), vp, ip, ip, ip, P(ip), ip, P(vp), P(C.c_size_t)]),
        "dsv_decode_buffer": (ip, [vp, C.c_size_t, ip, P(vp), P(C.c_size_t), P(ip), P(DSV_META)]),
        "dsv_decode_sharded": (ip, [vp, C.c_size_t, ip, P(ip), ip, ip, P(vp), P(C.c_size_t), P(ip), P(DSV_META)]),
        "dsv_pool_create": (vp, [ip, P(ip), ip]),
        "dsv_pool_destroy": (None, [vp]),
        "dsv_pool_threads": (ip, [vp]),
        "dsv_pool_encode": (ip, [vp, P(DSV_ENC_OPTS), vp, ip, ip, P(vp), P(C.c_size_t)]),
        "dsv_pool_decode": (ip, [vp, vp, C.c_size_t, vp, C.c_size_t, P(ip), P(DSV_META)]),
        "dsvcu_host_alloc": (vp, [C.c_size_t]),
        "dsvcu_host_free": (None, [vp]),
        "dsvcu_device_count": (ip, []),
        "dsvcu_last_error": (C.c_char_p, []),
        "dsvcu_ctx_create": (ip, [P(vp), ip, ip, ip, ip]),
        "dsvcu_ctx_destroy": (None, [vp]),
        "dsvcu_ctx_stream": (vp, [vp]),
        "dsvcu_sync": (ip, [vp]),
        "dsvcu_frame_create": (ip, [vp, P(vp)]),
        "dsvcu_frame_create_luma": (ip, [vp, P(vp), ip, ip]),
        "dsvcu_frame_destroy": (None, [vp, vp]),
        "dsvcu_frame_plane_dims": (ip, [vp, ip, P(ip), P(ip), P(ip)]),
        "dsvcu_frame_upload": (ip, [vp, vp, ip, vp, ip]),
        "dsvcu_frame_download": (ip, [vp, vp, ip, vp, ip]),
        "dsvcu_frame_clear_plane": (ip, [vp, vp, ip, ip]),
        "dsvcu_frame_upload_bordered": (ip, [vp, vp, ip, vp]),
        "dsvcu_frame_download_bordered": (ip, [vp, vp, ip, vp]),
        "dsvcu_coefs_create": (ip, [vp, P(vp)]),
        "dsvcu_coefs_destroy": (None, [vp, vp]),
        "dsvcu_coefs_plane_dims": (ip, [vp, ip, P(ip), P(ip)]),
        "dsvcu_coefs_upload": (ip, [vp, vp, ip, vp]),
        "dsvcu_coefs_download": (ip, [vp, vp, ip, vp]),
        "dsvcu_set_blockdata": (ip, [vp, vp, ip]),
        "dsvcu_set_mvs": (ip, [vp, vp, ip]),
        "dsvcu_fwd_sbt": (ip, [vp, vp, ip, vp, P(DSVCU_FMETA)]),
        "dsvcu_inv_sbt": (ip, [vp, vp, ip, vp, ip, P(DSVCU_FMETA)]),
        "dsvcu_fwd_sbt_frame": (ip, [vp, vp, vp, P(DSVCU_FMETA), ip]),
        "dsvcu_inv_sbt_frame": (ip, [vp, vp, vp, ip, P(DSVCU_FMETA), ip]),
        "dsvcu_quant_plane": (ip, [vp, vp, ip, ip, P(DSVCU_FMETA)]),
        "dsvcu_quant_frame": (ip, [vp, vp, ip, P(DSVCU_FMETA), ip]),
        "dsvcu_fetch_symbols": (ip, [vp, ip, P(P(DSVCU_SYMBOL)), P(ip), P(ip)]),
        "dsvcu_symbol_staging": (P(DSVCU_SYMBOL), [vp, ip, P(ip)]),
        "dsvcu_dequant_plane": (ip, [vp, vp, ip, ip, P(DSVCU_FMETA), ip, P(ip), ip]),
        "dsv_set_device_entropy_decode": (ip, [ip]),
        "dsv_set_device_entropy_limits": (None, [C.c_long, C.c_long, ip]),
        "dsvcu_parse_begin": (ip, [vp, P(DSVCU_PLANE_BITS), ip, ip, vp, ip, ip]),
        "dsvcu_parse_end": (ip, [vp, ip, ip, P(ip), P(ip)]),
        "dsvcu_set_side_parsed": (ip, [vp, ip, ip, ip]),
        "dsvcu_parse_ready": (ip, [vp, ip, ip]),
        "dsvcu_parse_planes": (ip, [vp, P(DSVCU_PLANE_BITS), ip, P(ip)]),
        "dsvcu_parsed_count": (ip, [vp, ip, ip]),
        "dsvcu_dequant_parsed": (ip, [vp, vp, ip, P(DSVCU_FMETA), ip, ip]),
        "dsvcu_scan_layout": (ip, [ip, ip, P(ip)]),
        "dsvcu_sub_pred": (ip, [vp, P(DSVCU_FMETA), vp, vp, vp]),
        "dsvcu_add_pred": (ip, [vp, P(DSVCU_FMETA), ip, vp, vp, vp, ip]),
        "dsvcu_add_res": (ip, [vp, P(DSVCU_FMETA), ip, vp, vp, ip]),
        "dsvcu_intra_filter": (ip, [vp, ip, P(DSVCU_FMETA), ip, vp, ip]),
        "dsvcu_post_process": (ip, [vp, vp]),
        "dsvcu_extend_frame": (ip, [vp, vp, ip]),
        "dsvcu_ds2x_luma": (ip, [vp, vp, vp]),
        "dsvcu_frame_copy": (ip, [vp, vp, vp]),
        "dsvcu_pyramid_create": (ip, [vp, P(vp), ip]),
        "dsvcu_pyramid_destroy": (None, [vp, vp]),
        "dsvcu_pyramid_build": (ip, [vp, vp, vp]),
        "dsvcu_pyramid_level": (vp, [vp, ip]),
        "dsvcu_extend_pyramid": (ip, [vp, vp, vp]),
        "dsvcu_set_side": (ip, [vp, vp, vp, ip]),
        "dsvcu_hme_counters": (ip, [vp, P(C.c_longlong)]),
        "dsvcu_mvs_swap_prev": (ip, [vp, ip]),
        "dsvcu_sub_pred_from": (ip, [vp, P(DSVCU_FMETA), vp, vp, vp, vp]),
        "dsvcu_set_prev_mvs": (ip, [vp, vp, ip]),
        "dsvcu_mvs_to_prev": (ip, [vp, ip]),
        "dsvcu_hme": (ip, [vp, P(DSVCU_FMETA), P(DSVCU_HME_PARAMS), vp, vp, vp, vp, vp, vp]),
        "dsvcu_hme_fetch": (ip, [vp, vp, ip, P(ip), P(ip), P(ip)]),
        "dsvcu_intra_analysis": (ip, [vp, P(DSVCU_FMETA), vp, vp, ip]),
        "dsvcu_frame_luma_avg": (ip, [vp, vp, P(C.c_uint)]),
        "dsvcu_timer_start": (ip, [vp]),
        "dsvcu_timer_stop_ms": (ip, [vp, P(C.c_float)]),
        "dsvcu_launch_count": (C.c_longlong, [vp]),
        "dsvcu_total_launches": (C.c_longlong, []),
        "dsvcu_isqrt": (C.c_uint, [C.c_uint]),
    }
    for name, (res, args) in sig.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue
        fn.restype = res
        fn.argtypes = args
    _cache[path] = lib
    return lib


def split_packets(data):
    """Split a .dsv byte string into packets using the next-link field
    (reference dsv_main.c:912-957)."""
    out, off = [], 0
    n = len(data)
    while off + PACKET_HDR_SIZE <= n:
        if data[off:off + 4] != b"DSV2":
            raise ValueError("bad 4cc at %d" % off)
        size = int.from_bytes(data[off + 10:off + 14], "big")
        if size == 0:
            size = PACKET_HDR_SIZE
        if size < PACKET_HDR_SIZE or off + size > n:
            break
        out.append(data[off:off + size])
        off += size
    return out


def decode_stream(data, emu=False, loglevel=1):
    """Decode a whole .dsv byte string through dsv_dec().  Returns
    (meta dict, [ (Y,U,V) bytes per frame ])."""
    lib = load(emu)
    lib.dsv_set_log_level(loglevel)
    dec = DSV_DECODER()
    frames = []
    for pkt in split_packets(data):
        buf = DSV_BUF()
        lib.dsv_mk_buf(C.byref(buf), len(pkt))
        C.memmove(buf.data, pkt, len(pkt))
        fr = C.POINTER(DSV_FRAME)()
        fno = C.c_uint32()
        code = lib.dsv_dec(C.byref(dec), C.byref(buf), C.byref(fr), C.byref(fno))
        if code == DEC_EOS:
            break
        if code != DEC_OK or not fr:
            continue
        planes = []
        f = fr.contents
        for c in range(3):
            p = f.planes[c]
            rows = [C.string_at(C.addressof(p.data.contents) + y * p.stride, p.w) for y in range(p.h)]
            planes.append(b"".join(rows))
        frames.append(tuple(planes))
        lib.dsv_frame_ref_dec(fr)
    meta = {k: getattr(dec.vidmeta, k) for k, _ in DSV_META._fields_}
    lib.dsv_dec_free(C.byref(dec))
    return meta, frames


def enc_opts(w, h, fmt=SUBSAMP_420, fps=(30, 1), **kw):
    """dsv_enc_opts with the reference CLI defaults; keyword names are the
    reference's option names (qp, gop, effort, rc_mode, ...)."""
    lib_ = load(kw.pop("emu", False))
    o = DSV_ENC_OPTS()
    lib_.dsv_enc_opts_default(C.byref(o), w, h, fmt, fps[0], fps[1])
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


def _take(lib, ptr, n, pinned=False):
    data = C.string_at(ptr, n.value)
    if pinned:
        lib.dsv_pinned_free(ptr)
    else:
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(ptr)
    return data


def encode_frames(opts, yuv, nframes, emu=False, exhausted=True, chunk=0, threads=1, devices=None):
    """Encode packed planar frames (bytes) -> .dsv bytes.  chunk > 0 selects the
    closed-GOP sharded driver (parallel_encode_yuv.sh semantics)."""
    lib = load(emu)
    out, n = C.c_void_p(), C.c_size_t()
    buf = (C.c_uint8 * len(yuv)).from_buffer_copy(yuv) if not isinstance(yuv, C.Array) else yuv
    if chunk > 0:
        devs = devices or [0]
        arr = (C.c_int * len(devs))(*devs)
        r = lib.dsv_encode_sharded(C.byref(opts), buf, nframes, chunk, threads, arr, len(devs), C.byref(out), C.byref(n))
    else:
        r = lib.dsv_encode_buffer(C.byref(opts), buf, nframes, 1 if exhausted else 0, C.byref(out), C.byref(n))
    if r:
        raise RuntimeError("encode failed: %s" % lib.dsvcu_last_error().decode())
    return _take(lib, out, n)


def decode_frames(data, emu=False, threads=1, devices=None, device_entropy=1):
    """Decode .dsv bytes -> (DSV_META, nframes, packed planar frames bytes).
    device_entropy: dsv_set_device_entropy_decode() for this call (1 = coefficient planes are
    entropy-decoded by k_hzcc_parse, 0 = by the host threads, -1 = the library's default, which is 1)."""
    lib = load(emu)
    lib.dsv_set_device_entropy_decode(device_entropy)
    out, n, nfr, meta = C.c_void_p(), C.c_size_t(), C.c_int(), DSV_META()
    devs = devices or [0]
    arr = (C.c_int * len(devs))(*devs)
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    r = lib.dsv_decode_sharded(buf, len(data), threads, arr, len(devs), 0, C.byref(out), C.byref(n), C.byref(nfr),
                               C.byref(meta))
    if r:
        raise RuntimeError("decode failed: %s" % lib.dsvcu_last_error().decode())
    return meta, nfr.value, _take(lib, out, n)


def rank_range(nchunks, rank, world):
    """Closed-GOP chunks [first, last) owned by `rank` of `world` ranks (one rank
    per GPU): contiguous ranges, sizes differing by at most one.  Chunks are
    independent (parallel_encode_yuv.sh semantics), so ranks never exchange
    data; the host concatenates the per-rank byte strings in rank order."""
    base, extra = divmod(nchunks, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def encode_rank_shard(opts, yuv, nframes, chunk, rank, world, emu=False, threads=1, device=0):
    """This rank's part of a sharded encode: bytes of its chunk range."""
    nchunks = (nframes + chunk - 1) // chunk
    first, last = rank_range(nchunks, rank, world)
    if first >= last:
        return b""
    fsz = len(yuv) // nframes
    f0, f1 = first * chunk, min(nframes, last * chunk)
    return encode_frames(opts, yuv[f0 * fsz:f1 * fsz], f1 - f0, emu=emu, chunk=chunk, threads=threads, devices=[device])
