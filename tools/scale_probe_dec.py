#!/usr/bin/env python
"""decode throughput vs host threads.  usage: scale_probe_dec.py THREADS[,THREADS..] [MODES]
MODES: 1 = entropy decode on the device (default), 0 = on the host, e.g. 0,1"""
import ctypes as C, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, util
import torch
P = util.pkg()
if os.environ.get("DSV2CUDA_LIB"):  # a variant build (tools/gpu_filtvar.sh)
    P.lib_path = lambda emu=False, _p=os.environ["DSV2CUDA_LIB"]: _p
lib = P.load()
data = bench.synth_chunks(2)
GOPN = 48
NCH = 32
nfr = NCH * GOPN
host = torch.empty(nfr * bench.FRAME_BYTES, dtype=torch.uint8, pin_memory=True)
hv = host.numpy()
for c in range(NCH):
    k = c % 2
    hv[c*GOPN*bench.FRAME_BYTES:(c+1)*GOPN*bench.FRAME_BYTES] = data[k*48*bench.FRAME_BYTES:(k*48+GOPN)*bench.FRAME_BYTES]
dev = host.cuda(); torch.cuda.synchronize()
devs = (C.c_int * 1)(0)
pool = lib.dsv_pool_create(32, devs, 1)
o = P.enc_opts(bench.W, bench.H, P.SUBSAMP_420, (30, 1), qp=60, gop=48, noeos=1)
out, outn = C.c_void_p(), C.c_size_t()
t0 = time.perf_counter()
lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(dev.data_ptr()), nfr, GOPN, C.byref(out), C.byref(outn))
print("encode 32 threads: %.1f fps" % (nfr / (time.perf_counter() - t0)), flush=True)
dsv = (C.c_uint8 * outn.value).from_buffer_copy(C.string_at(out, outn.value))
lib.dsv_pool_destroy(pool)
nf, meta = C.c_int(), P.DSV_META()
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
for threads, mode in [(int(t), m) for t in sys.argv[1].split(",") for m in modes]:
    lib.dsv_set_device_entropy_decode(mode)
    pool = lib.dsv_pool_create(threads, devs, 1)
    for rep in range(3):
        t0 = time.perf_counter()
        lib.dsv_pool_decode(pool, dsv, len(dsv), C.c_void_p(dev.data_ptr()), nfr * bench.FRAME_BYTES, C.byref(nf), C.byref(meta))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    for rep in range(2):
        t0 = time.perf_counter()
        lib.dsv_pool_decode(pool, dsv, len(dsv), C.c_void_p(host.data_ptr()), nfr * bench.FRAME_BYTES, C.byref(nf), C.byref(meta))
        dth = time.perf_counter() - t0
    print("threads %2d, entropy decode on the %s: decode %7.1f fps to HBM, %7.1f fps to pinned host (%.2f ms/frame/stream)" % (threads, "device" if mode else "host", nfr / dt, nfr / dth, 1000 * dt * threads / nfr), flush=True)
    lib.dsv_pool_destroy(pool)
