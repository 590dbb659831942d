"""The C-ABI library loads and exports every function include/*.h declares
(no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re

import util

INC = os.path.join(util.ROOT, "include")
DECL = re.compile(r"^\s*(?:extern\s+)?(?:const\s+)?(?:unsigned\s+|long\s+long\s+)?[A-Za-z_][A-Za-z_0-9]*\s*\**\s*\*?\s*((?:dsvcu|dsv)_[a-z0-9_]+)\s*\(",
                  re.M)


def declared():
    names = set()
    for h in sorted(os.listdir(INC)):
        if not h.endswith(".h"):
            continue
        txt = open(os.path.join(INC, h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
        names |= set(DECL.findall(txt))
    return names


def test_product_library_exports_every_declared_symbol():
    so = util.pkg().lib_path()
    assert os.path.exists(so), "run `make` first: the product library is missing"
    lib = C.CDLL(so)
    names = declared()
    assert len(names) > 60, names
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, "declared in include/*.h but not exported: %r" % missing


def test_scan_layout_is_pure_host_code():
    lib = util.pkg().load()
    part = (C.c_int * 5)()
    total = lib.dsvcu_scan_layout(1920, 1080, part)
    assert total == 1920 * 1080 and list(part) == [0, 240 * 135, 240 * 135 * 4, 240 * 135 * 4 + 3 * 480 * 270, total]


def test_no_cpu_fallback_without_device():
    """without a CUDA device context creation must fail loudly (there is no CPU path)"""
    lib = util.pkg().load()
    if lib.dsvcu_device_count() > 0:
        return
    ctx = C.c_void_p()
    assert lib.dsvcu_ctx_create(C.byref(ctx), 0, 352, 288, 5) != 0
    assert b"no CUDA device" in lib.dsvcu_last_error()


def test_isqrt_matches_reference_form():
    """me_isqrt (hardware sqrt + exact fix-up) == the reference's digit-by-digit
    iisqrt (hme.c:99-124) == floor(sqrt(n)): every perfect square +-2 up to 2^32
    (sampled) and the extremes"""
    import math
    lib = util.pkg().load()

    def ref(n):  # restatement of the reference loop
        if n == 0:
            return 0
        pos, res, rem = 1 << 30, 0, n
        while pos > rem:
            pos >>= 2
        while pos:
            dif = res + pos
            res >>= 1
            if rem >= dif:
                rem -= dif
                res += pos
            pos >>= 2
        return res

    vals = {0, 1, 2, 3, 0xffffffff, 0xfffffffe, 0x7fffffff, 0x80000000}
    for r in list(range(0, 3000)) + list(range(3000, 65536, 37)) + [65535]:
        for d in (-2, -1, 0, 1, 2):
            n = r * r + d
            if 0 <= n <= 0xffffffff:
                vals.add(n)
    for n in sorted(vals):
        got = lib.dsvcu_isqrt(n)
        assert got == math.isqrt(n) == ref(n), n
