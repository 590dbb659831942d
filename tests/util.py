"""Shared helpers of the test-suite.

The checker is always the reference (oracle/_ref, built from /root/reference by
oracle/Makefile and shipped to the GPU box as a prebuilt binary) or the CPU
restatement in oracle/; the thing under test is always the product library
reached through its C ABI.
"""
import ctypes as C
import hashlib
import importlib.util
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "digital-subband-video-2_b200")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BIN = os.path.join(REF_DIR, "dsv2")
REF_LIB = os.path.join(REF_DIR, "libdsvref.so")
REF_D28 = os.path.join(REF_DIR, "dsv28dec")
CACHE = os.environ.get("DSV2_TEST_CACHE", "/tmp/dsv2_b200_cache")

sys.path.insert(0, os.path.join(ROOT, "tools"))


def pkg():
    name = "dsv2_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, "__init__.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def have_ref():
    return os.path.exists(REF_BIN) and os.path.exists(REF_LIB)


def ensure_emu():
    """Build the test-only host emulation of the kernel sources if needed."""
    so = pkg().lib_path(emu=True)
    subprocess.run(["make", "-s", "emu"], cwd=ROOT, check=True,
                   stdout=subprocess.DEVNULL)
    assert os.path.exists(so)
    return so


def clip(name, w, h, n, fmt="420", **kw):
    """Synthetic y4m (tools/synth_y4m.py), cached on disk."""
    import synth_y4m
    os.makedirs(CACHE, exist_ok=True)
    tag = "_".join("%s%s" % (k, v) for k, v in sorted(kw.items()))
    path = os.path.join(CACHE, "%s_%dx%d_%d_%s_%s.y4m" % (name, w, h, n, fmt, tag))
    if not os.path.exists(path):
        synth_y4m.write_y4m(path + ".tmp", w, h, n, fmt, **kw)
        os.replace(path + ".tmp", path)
    return path


def ref_encode(y4m, args, tag):
    """Run the reference encoder (oracle/_ref/dsv2 e) -> .dsv path."""
    out = y4m[:-4] + "_" + tag + ".dsv"
    if not os.path.exists(out):
        cmd = [REF_BIN, "e", "-y", "-inp=" + y4m, "-out=" + out + ".tmp", "-y4m=1"] + list(args)
        # exit status 254 (-2) = "input exhausted", the normal end (dsv_main.c:904)
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL)
        assert r.returncode in (0, 254), "reference encoder failed: %d" % r.returncode
        os.replace(out + ".tmp", out)
    return out


def ref_decode(dsv):
    out = dsv[:-4] + "_refdec.y4m"
    if not os.path.exists(out):
        subprocess.run([REF_BIN, "d", "-y", "-inp=" + dsv, "-out=" + out + ".tmp", "-y4m=1"],
                       check=True, stdout=subprocess.DEVNULL)
        os.replace(out + ".tmp", out)
    return out


def read_y4m(path):
    d = open(path, "rb").read()
    e = d.index(b"\n")
    toks = d[:e].decode().split()
    w = int([t for t in toks if t[0] == "W"][0][1:])
    h = int([t for t in toks if t[0] == "H"][0][1:])
    c = [t for t in toks if t[0] == "C"][0][1:]
    if c.startswith("444"):
        cw, ch = w, h
    elif c.startswith("422"):
        cw, ch = (w + 1) // 2, h
    elif c.startswith("411"):
        cw, ch = (w + 3) // 4, h
    elif c.startswith("410"):
        cw, ch = (w + 3) // 4, (h + 3) // 4
    else:
        cw, ch = (w + 1) // 2, (h + 1) // 2
    fsz = w * h + 2 * cw * ch
    off, frames = e + 1, []
    while off < len(d):
        assert d[off:off + 6] == b"FRAME\n"
        off += 6
        frames.append((d[off:off + w * h], d[off + w * h:off + w * h + cw * ch],
                       d[off + w * h + cw * ch:off + fsz]))
        off += fsz
    return w, h, frames


def frames_md5(frames):
    m = hashlib.md5()
    for f in frames:
        for p in f:
            m.update(p)
    return m.hexdigest()


def first_diff(a, b, w):
    x = np.frombuffer(a, np.uint8).astype(int)
    y = np.frombuffer(b, np.uint8).astype(int)
    d = np.nonzero(x != y)[0]
    if len(d) == 0:
        return None
    return dict(n=len(d), x=int(d[0] % w), y=int(d[0] // w), maxabs=int(abs(x - y).max()))


def assert_same_frames(got, ref, w, h):
    assert len(got) == len(ref), "frame count %d vs %d" % (len(got), len(ref))
    for i, (a, b) in enumerate(zip(got, ref)):
        for c in range(3):
            if a[c] != b[c]:
                pw = w if len(b[c]) == w * h else (w + 1) // 2
                assert len(a[c]) == len(b[c]), "frame %d plane %d size" % (i, c)
                raise AssertionError("frame %d plane %d differs: %r" % (i, c, first_diff(a[c], b[c], pw)))
