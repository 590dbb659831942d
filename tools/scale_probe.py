#!/usr/bin/env python
"""encode throughput vs host threads, with the per-phase profile of dsv_enc"""
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
GOPN = int(os.environ.get("GOPN", "24"))
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
    o = P.enc_opts(bench.W, bench.H, P.SUBSAMP_420, (30, 1), qp=60, gop=48, noeos=1)
    out, outn = C.c_void_p(), C.c_size_t()
    libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
    import resource
    for rep in range(3):
        ru0 = resource.getrusage(resource.RUSAGE_SELF)
        t0 = time.perf_counter()
        lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(dev.data_ptr()), nfr, GOPN, C.byref(out), C.byref(outn))
        dt = time.perf_counter() - t0
        libc.free(out)
    ru1 = resource.getrusage(resource.RUSAGE_SELF)
    cpu = (ru1.ru_utime - ru0.ru_utime) + (ru1.ru_stime - ru0.ru_stime)
    print("threads %2d: %6.1f fps  (%.1f ms/frame/stream)  host CPU %.2f ms/frame (user %.2f sys %.2f), %.1f cores busy" % (
        threads, nfr / dt, 1000 * dt / GOPN, 1000 * cpu / nfr, 1000 * (ru1.ru_utime - ru0.ru_utime) / nfr,
        1000 * (ru1.ru_stime - ru0.ru_stime) / nfr, cpu / dt), flush=True)
    if hasattr(lib, "dsvcu_debug_me_counters"):
        d = (C.c_ulonglong * 4)()
        lib.dsvcu_debug_me_counters(d, 1)
        if d[2]:
            print("   level-0 search: %.1f us in me_block, %.1f us waiting, per block (all reps)" % (d[0] / d[2] / 1965.0, d[1] / d[2] / 1965.0), flush=True)
    lib.dsv_pool_destroy(pool)
