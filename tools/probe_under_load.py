#!/usr/bin/env python
"""Diagnostics (needs a -DME_TIMING build): one probe warp per SM measures the
cost of a dependent ALU step, a dependent L2 load and a dependent shared-memory
load, first on an idle GPU, then while N encoder instances run.
usage: probe_under_load.py 32"""
import ctypes as C, os, sys, threading, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, util
import torch
P = util.pkg(); lib = P.load()
lib.dsvcu_debug_probe.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
threads = int(sys.argv[1]) if len(sys.argv) > 1 else 32
GOPN = 24


def probe(tag):
    r = []
    for mode, it in ((0, 200000), (1, 20000), (2, 200000)):
        v = C.c_double()
        best = []
        for _ in range(5):
            lib.dsvcu_debug_probe(mode, it, C.byref(v))
            best.append(v.value)
        r.append(sorted(best)[len(best) // 2])
    print("%-28s ALU step %.2f cyc   L2 load %.0f cyc   shared load %.1f cyc" % (tag, r[0], r[1], r[2]), flush=True)


data = bench.synth_chunks(2)
nfr = threads * GOPN
host = torch.empty(nfr * bench.FRAME_BYTES, dtype=torch.uint8, pin_memory=True)
hv = host.numpy()
for c in range(threads):
    k = c % 2
    hv[c*GOPN*bench.FRAME_BYTES:(c+1)*GOPN*bench.FRAME_BYTES] = data[k*48*bench.FRAME_BYTES:(k*48+GOPN)*bench.FRAME_BYTES]
dev = host.cuda(); torch.cuda.synchronize()
probe("idle GPU")
devs = (C.c_int * 1)(0)
pool = lib.dsv_pool_create(threads, devs, 1)
o = P.enc_opts(bench.W, bench.H, P.SUBSAMP_420, (30, 1), qp=60, gop=48, noeos=1)
out, outn = C.c_void_p(), C.c_size_t()
libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
stop = False


def load():
    while not stop:
        lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(dev.data_ptr()), nfr, GOPN, C.byref(out), C.byref(outn))
        libc.free(out)


t = threading.Thread(target=load); t.start()
time.sleep(3.0)
for k in range(3):
    probe("%d encoder instances" % threads)
    time.sleep(0.5)
stop = True
t.join()
lib.dsv_pool_destroy(pool)
