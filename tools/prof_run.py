#!/usr/bin/env python
"""Workload for profilers: encode N frames with one encoder instance, then
decode the result.  usage: prof_run.py W H N [enc|dec|both]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import util  # noqa: E402
import ops  # noqa: E402


def main():
    w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    what = sys.argv[4] if len(sys.argv) > 4 else "both"
    P = util.pkg()
    _, _, fr = util.read_y4m(util.clip("perf", w, h, n, "420"))
    yuv = b"".join(ops.yuv_bytes(f) for f in fr)
    o = P.enc_opts(w, h, P.SUBSAMP_420, (30, 1), qp=60, gop=48)
    cache = "/tmp/prof_%dx%d_%d.dsv" % (w, h, n)
    if what in ("enc", "both") or not os.path.exists(cache):
        t0 = time.time()
        dsv = P.encode_frames(o, yuv, n)
        print("encode %d frames: %.3fs" % (n, time.time() - t0))
        open(cache, "wb").write(dsv)
    dsv = open(cache, "rb").read()
    if what in ("dec", "both"):
        t0 = time.time()
        meta, nfr, dec = P.decode_frames(dsv)
        print("decode %d frames: %.3fs" % (nfr, time.time() - t0))


if __name__ == "__main__":
    main()
