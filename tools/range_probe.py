#!/usr/bin/env python
"""One dsv_pool_encode step with T concurrent encoder instances between
cudaProfilerStart/Stop, for `ncu --replay-mode app-range` (counters over the whole
range with every stream running, which single-kernel captures cannot give).
usage: range_probe.py THREADS [FRAMES_PER_GOP]"""
import ctypes as C, os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, util
import torch
P = util.pkg(); lib = P.load()
threads = int(sys.argv[1]); GOPN = int(sys.argv[2]) if len(sys.argv) > 2 else 4
data = bench.synth_chunks(2)
nfr = threads * GOPN
host = torch.empty(nfr * bench.FRAME_BYTES, dtype=torch.uint8, pin_memory=True)
hv = host.numpy()
for c in range(threads):
    k = c % 2
    hv[c*GOPN*bench.FRAME_BYTES:(c+1)*GOPN*bench.FRAME_BYTES] = data[k*48*bench.FRAME_BYTES:(k*48+GOPN)*bench.FRAME_BYTES]
dev = host.cuda(); torch.cuda.synchronize()
devs = (C.c_int * 1)(0)
pool = lib.dsv_pool_create(threads, devs, 1)
o = P.enc_opts(bench.W, bench.H, P.SUBSAMP_420, (30, 1), qp=60, gop=48, noeos=1)
out, outn = C.c_void_p(), C.c_size_t()
libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(dev.data_ptr()), nfr, GOPN, C.byref(out), C.byref(outn))
libc.free(out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(dev.data_ptr()), nfr, GOPN, C.byref(out), C.byref(outn))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
libc.free(out)
lib.dsv_pool_destroy(pool)
print("done", nfr)
