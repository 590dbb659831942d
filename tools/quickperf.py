#!/usr/bin/env python
"""Quick throughput probe (not the bench): encode/decode fps vs host threads."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import util  # noqa: E402
import ops  # noqa: E402


def main():
    w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    threads = [int(t) for t in sys.argv[4].split(",")]
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    P = util.pkg()
    lib = P.load()
    _, _, fr = util.read_y4m(util.clip("perf", w, h, n, "420"))
    one = b"".join(ops.yuv_bytes(f) for f in fr)
    o = P.enc_opts(w, h, P.SUBSAMP_420, (30, 1), qp=60, gop=n, noeos=1)
    for t in threads:
        nch = t * reps
        yuv = one * nch
        buf = (C.c_uint8 * len(yuv)).from_buffer_copy(yuv)
        t0 = time.time()
        dsv = P.encode_frames(o, buf, n * nch, chunk=n, threads=t)
        dt = time.time() - t0
        t0 = time.time()
        meta, nfr, dec = P.decode_frames(dsv, threads=t)
        dd = time.time() - t0
        print("threads %2d: encode %4d frames %.3fs = %.1f fps (%d bytes/frame) | decode %.3fs = %.1f fps" % (
            t, n * nch, dt, n * nch / dt, len(dsv) // (n * nch), dd, nfr / dd), flush=True)


if __name__ == "__main__":
    main()
