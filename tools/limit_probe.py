#!/usr/bin/env python
"""Which stage caps the many-instance throughput?  Same 1080p chunks, T instances,
encoder options that switch stages off: loop filters off, intra only (no motion
search, no filters), low effort.  usage: limit_probe.py THREADS[,THREADS..]"""
import ctypes as C, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, util
import torch
P = util.pkg(); lib = P.load()
data = bench.synth_chunks(2)
GOPN = int(os.environ.get("GOPN", "12"))
libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
VARIANTS = [("default", {}), ("pfilter=0", dict(pfilter=0)), ("intra only (gop=0)", dict(gop=0)),
            ("effort=3 (no sub-pel)", dict(effort=3)), ("pfilter=0 effort=3", dict(pfilter=0, effort=3))]
for threads in [int(t) for t in sys.argv[1].split(",")]:
    nfr = threads * GOPN
    host = torch.empty(nfr * bench.FRAME_BYTES, dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    for c in range(threads):
        k = c % 2
        hv[c*GOPN*bench.FRAME_BYTES:(c+1)*GOPN*bench.FRAME_BYTES] = data[k*48*bench.FRAME_BYTES:(k*48+GOPN)*bench.FRAME_BYTES]
    dev = host.cuda(); torch.cuda.synchronize()
    devs = (C.c_int * 1)(0)
    pool = lib.dsv_pool_create(threads, devs, 1)
    for name, kw in VARIANTS:
        args = dict(qp=60, gop=48, noeos=1); args.update(kw)
        o = P.enc_opts(bench.W, bench.H, P.SUBSAMP_420, (30, 1), **args)
        out, outn = C.c_void_p(), C.c_size_t()
        best = 0
        for rep in range(3):
            t0 = time.perf_counter()
            lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(dev.data_ptr()), nfr, GOPN, C.byref(out), C.byref(outn))
            dt = time.perf_counter() - t0
            libc.free(out)
            best = max(best, nfr / dt)
        print("threads %2d  %-24s %7.1f fps" % (threads, name, best), flush=True)
    lib.dsv_pool_destroy(pool)
