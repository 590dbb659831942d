#!/usr/bin/env python
"""Does the motion search slow down because of OTHER motion searches, or because
of the other kernels of the encoder?  N host threads, each with its own context
and stream, run ONLY dsvcu_hme in a loop on resident 1080p pictures.
usage: me_concurrency.py 1,8,16,32 [reps]"""
import ctypes as C
import os
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import ops  # noqa: E402
import util  # noqa: E402

W, H = 1920, 1080


def main():
    counts = [int(t) for t in sys.argv[1].split(",")]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    _, _, fr = util.read_y4m(util.clip("perf", W, H, 3, "420"))
    yuv = [ops.yuv_bytes(f) for f in fr]
    cfg = ops.Cfg(W, H)
    devs = []
    for _ in range(max(counts)):
        d = ops.Dev(cfg, False)
        fs, frf, fo = d.frame(yuv[2]), d.frame(yuv[1]), d.frame(yuv[1])
        d.args = (fs, d.pyramid(fs), frf, d.pyramid(frf), fo, d.pyramid(fo))
        d.hp = d.P.DSVCU_HME_PARAMS(400, cfg.skip_thresh, cfg.pyr, 0)
        d.fm = cfg.fmeta()
        d.out = np.zeros(cfg.nblk, ops.MV_DTYPE)
        devs.append(d)

    def work(d, n):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        fs, ps, frf, pr, fo, po = d.args
        for _ in range(n):
            d.ck(d.lib.dsvcu_hme(d.ctx, C.byref(d.fm), C.byref(d.hp), fs, ps, frf, pr, fo, po))
            d.ck(d.lib.dsvcu_hme_fetch(d.ctx, d.out.ctypes.data_as(C.c_void_p), cfg.nblk, C.byref(a), C.byref(b), C.byref(c)))

    for d in devs:  # first use of a context allocates its search scratch (cudaMalloc synchronises the device)
        work(d, 1)
    for n in counts:
        th = [threading.Thread(target=work, args=(devs[k], reps)) for k in range(n)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        print("%2d concurrent searches: %.2f ms per search, %.0f searches/s" % (n, 1000 * dt / reps, n * reps / dt), flush=True)


if __name__ == "__main__":
    main()
