#!/usr/bin/env python
"""bench.py -- DSV2 pixel hot path on B200: 1080p 4:2:0 closed-GOP encode (and
decode) throughput, bit-exact with the reference.

Workload (BASELINE.json configs[4], the one the metric is quoted on): synthetic
1920x1080 4:2:0 frames (tools/synth_y4m.py), 48-frame closed GOPs, each GOP coded
by a fresh encoder exactly like the reference's parallel_encode_yuv.sh
(`-qp=60 -gop=48 -noeos=1`, CRF, all psy options, loop filters on).  One "step"
= every worker thread of the rank encodes one GOP chunk (weak scaling: the
number of chunks per step is fixed per GPU).

  value  frames/s, source frames already resident in HBM when the clock starts
  e2e    the same through the public C ABI with HOST (pinned) frame buffers:
         the host->device copy of every frame and the device->host read of the
         coded symbols are inside the timed region (the .dsv bytes end up in
         host memory in both modes)
  decode the decoder on the stream just produced, frames left in HBM (value)
         and copied to pinned host memory (e2e)

`--impl reference` times the UNMODIFIED reference (oracle/_ref/dsv2, built from
/root/reference with cc -O3) on the host cores, one process per core on chunked
input like parallel_encode_yuv.sh.

Launch: python bench.py --gpus N --steps K --warmup W   (N>1: under torchrun)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# one CUDA stream per encoder/decoder instance: give them separate hardware queues
# (must be set before the CUDA context is created; the library does the same)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# stdout carries exactly one JSON line: everything any library prints on fd 1
# (NCCL's version banner, for one) is sent to stderr; emit() writes the line to
# the real stdout
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, GOP, QP, FPS = 1920, 1080, 48, 60, 30
FRAME_BYTES = W * H * 3 // 2
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dsv2")
METRIC = "1080p 4:2:0 encode fps (48-frame closed GOPs, bit-exact .dsv)"
WORKLOAD = "1920x1080 4:2:0, -qp=60 -gop=48 -noeos=1 closed-GOP chunks of 48 frames (BASELINE configs[4])"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synth_chunks(ndistinct):
    """ndistinct different 48-frame chunks as one uint8 array (cached on tmpfs)."""
    import numpy as np
    import synth_y4m
    cache = "/dev/shm/dsv2_bench_%dx%d_%d_%d.npy" % (W, H, GOP, ndistinct)
    if os.path.exists(cache):
        try:
            a = np.load(cache, mmap_mode="r")
            if a.size == ndistinct * GOP * FRAME_BYTES:
                return np.ascontiguousarray(a).reshape(-1)
        except Exception:
            pass
    out = np.empty((ndistinct * GOP, FRAME_BYTES), np.uint8)
    for i, (Y, U, V) in enumerate(synth_y4m.frames(W, H, ndistinct * GOP, "420", cut=40)):
        out[i, :W * H] = Y.reshape(-1)
        out[i, W * H:W * H + W * H // 4] = U.reshape(-1)
        out[i, W * H + W * H // 4:] = V.reshape(-1)
    try:
        np.save(cache + ".tmp.npy", out)
        os.replace(cache + ".tmp.npy", cache)
    except Exception:
        pass
    return out.reshape(-1)


class ClockSampler:
    """SM clocks / throttle reasons of the GPUs in use while the timed region
    runs, read through NVML (no nvidia-smi processes competing with the job)."""

    def __init__(self, indices, period=0.5):
        self.indices, self.period = list(indices), period
        self.sm, self.mx, self.reasons, self.stop = [], [], set(), False
        self.th = threading.Thread(target=self.run, daemon=True)
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
        except Exception:
            self.nv = None

    def sample(self):
        nv = self.nv
        if nv is None:
            return self.sample_smi()
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        for h in self.handles:
            self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            self.mx.append(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)

    def sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        o = subprocess.run(["nvidia-smi", "-i", ",".join(map(str, self.indices)), "--query-gpu=" + q,
                            "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in o.strip().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) >= 6 and t[0].isdigit():
                self.sm.append(int(t[0]))
                self.mx.append(int(t[1]))
                for k, n in enumerate(names):
                    if t[2 + k] == "Active":
                        self.reasons.add(n)

    def run(self):
        while not self.stop:
            try:
                self.sample()
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm), "gpus_sampled": self.indices,
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


class NoSampler:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def summary(self):
        return None


def cpu_count():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------- reference arm

def write_y4m(path, frames_u8, nframes):
    with open(path, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F%d:1 A1:1 Ip C420\n" % (W, H, FPS)).encode())
        for i in range(nframes):
            f.write(b"FRAME\n")
            f.write(frames_u8[i * FRAME_BYTES:(i + 1) * FRAME_BYTES].tobytes())


def ref_encode_procs(y4m, jobs, tag):
    """one reference encoder process per job; job = (first frame, frames):
    `dsv2 e -sfr=first -nfr=frames -noeos=1`, parallel_encode_yuv.sh:34-41"""
    ps, outs = [], []
    for k, (first, n) in enumerate(jobs):
        out = "/dev/shm/dsv2_bench_ref_%s_%d.dsv" % (tag, k)
        outs.append(out)
        ps.append(subprocess.Popen([REF_BIN, "e", "-y", "-inp=" + y4m, "-out=" + out, "-y4m=1", "-qp=%d" % QP,
                                    "-gop=%d" % GOP, "-sfr=%d" % first, "-nfr=%d" % n, "-noeos=1"],
                                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in ps:
        p.wait()
        if p.returncode not in (0, 254):
            raise RuntimeError("reference encoder exit %d" % p.returncode)
    return outs


DISTINCT = 2  # different 48-frame chunks in the workload (chunk c of rank r is chunk (c + r) % DISTINCT)


def ref_input():
    """the bench workload's distinct chunks as one y4m on tmpfs"""
    y4m = "/dev/shm/dsv2_bench_ref_in.y4m"
    want = len("YUV4MPEG2 W%d H%d F%d:1 A1:1 Ip C420\n" % (W, H, FPS)) + DISTINCT * GOP * (FRAME_BYTES + 6)
    if not (os.path.exists(y4m) and os.path.getsize(y4m) == want):
        write_y4m(y4m + ".tmp", synth_chunks(DISTINCT), DISTINCT * GOP)
        os.replace(y4m + ".tmp", y4m)
    return y4m


def run_reference(args):
    """the UNMODIFIED reference on the host cores: one `dsv2 e` process per core, each
    coding one WHOLE 48-frame chunk of the workload per step (same chunks, same I:P
    mix as the repo arm; parallel_encode_yuv.sh semantics)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(REF_BIN):
        emit({"impl": "reference", "unavailable": "oracle/_ref/dsv2 not built (needs /root/reference)"})
        return 0
    cores = cpu_count()
    y4m = ref_input()
    jobs = [((k % DISTINCT) * GOP, GOP) for k in range(cores)]
    for _ in range(args.warmup):
        ref_encode_procs(y4m, jobs[:max(1, cores)], "w")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_encode_procs(y4m, jobs, "t")
    dt = time.perf_counter() - t0
    fps = len(jobs) * GOP * args.steps / dt
    sample = "%d processes x one whole %d-frame closed-GOP chunk per step (1 I + 47 P, cut at 40 -> scene-change I)" % (
        len(jobs), GOP)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(fps, 3), "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000 * dt / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "reference": "oracle/_ref/dsv2 (unmodified reference, cc -O3), one process per host core, "
                                "whole chunks", "frames_per_step": len(jobs) * GOP, "host_cores": cores},
        "cpu_baseline": {"value": round(fps, 3), "unit": "frames/s", "cores": len(jobs), "kind": "reference",
                         "sample": sample},
        "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def md5(b):
    import hashlib
    return hashlib.md5(b).hexdigest()


def reference_chunks(compute):
    """.dsv bytes and decoded frames of every distinct chunk of the workload from the
    unmodified reference (`dsv2 e -qp=60 -gop=48 -sfr=48k -nfr=48 -noeos=1`, `dsv2 d`).
    Checker only: runs outside every timed region.  compute=False reads what another
    rank left on tmpfs."""
    names = [("/dev/shm/dsv2_bench_ref_par_%d.dsv" % k, "/dev/shm/dsv2_bench_ref_par_%d.yuv" % k) for k in range(DISTINCT)]
    if compute:
        ref_encode_procs(ref_input(), [(k * GOP, GOP) for k in range(DISTINCT)], "par")
        ps = [subprocess.Popen([REF_BIN, "d", "-y", "-inp=" + o, "-out=" + d], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL) for o, d in names]
        for p in ps:
            p.wait()
    return [open(o, "rb").read() for o, _ in names], [open(d, "rb").read() for _, d in names]


# --------------------------------------------------------------------- own arm

def kernel_rooflines(P, lib, peak_gbs):
    """Per-kernel-family device times (CUDA events on the context's stream) for
    the HBM-bound operators, over a ring of frames larger than L2, and for the
    motion search.  Returns (list of dicts, dominant-kernel dict)."""
    import numpy as np
    import ops
    cfg = ops.Cfg(W, H, P.SUBSAMP_420, isP=1, fnum=1)
    D = ops.Dev(cfg, False)
    lib, ctx = D.lib, D.ctx
    data = synth_chunks(2)
    RING = 12  # 12 x (12.4 MB coefs + 3.1 MB frame) > 126 MB L2
    src = [D.frame(bytes(data[i * FRAME_BYTES:(i + 1) * FRAME_BYTES])) for i in range(RING)]
    dst = [D.frame() for _ in range(RING)]
    coefs = [D.coefs() for _ in range(RING)]
    bd = np.zeros(cfg.nblk, np.uint8)
    D.set_blockdata(bd)
    mvs = np.zeros(cfg.nblk, ops.MV_DTYPE)
    rng = np.random.default_rng(1)
    mvs["x"] = rng.integers(-24, 25, cfg.nblk)
    mvs["y"] = rng.integers(-24, 25, cfg.nblk)
    D.set_mvs(mvs)
    fmP, fmI = cfg.fmeta(), cfg.fmeta()
    fmI.isP = 0
    q = 252
    Pb = FRAME_BYTES
    ms = C.c_float()

    def timed(fn, reps=3):
        for i in range(RING):
            fn(i)
        lib.dsvcu_sync(ctx)
        best = 1e9
        for _ in range(reps):
            lib.dsvcu_timer_start(ctx)
            for i in range(RING):
                fn(i)
            lib.dsvcu_timer_stop_ms(ctx, C.byref(ms))
            best = min(best, ms.value / RING)
        return best

    out = []

    def add(name, ms_, alg_bytes, launches):
        gbs = alg_bytes / (ms_ * 1e-3) / 1e9
        out.append({"kernel": name, "ms_per_frame": round(ms_, 4), "algorithmic_bytes": alg_bytes,
                    "launches_per_frame": launches, "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak_gbs, 4)})

    def all_planes(f):
        def g(i):
            for p in range(3):
                f(i, p)
        return g

    l0 = lib.dsvcu_launch_count(ctx)
    t = timed(lambda i: lib.dsvcu_fwd_sbt_frame(ctx, src[i], coefs[i], C.byref(fmP), 7))
    nl = (lib.dsvcu_launch_count(ctx) - l0) // (RING * 4)
    add("fwd_sbt (P picture, 3 planes per launch: k_sbt_fwd)", t, 5 * Pb, nl)
    l0 = lib.dsvcu_launch_count(ctx)
    t = timed(lambda i: lib.dsvcu_quant_frame(ctx, coefs[i], q, C.byref(fmP), 7))
    nl = (lib.dsvcu_launch_count(ctx) - l0) // (RING * 4)
    add("quantise + symbol compaction (k_quant_*, k_compact_*)", t, 8 * Pb, nl)
    l0 = lib.dsvcu_launch_count(ctx)
    t = timed(lambda i: lib.dsvcu_inv_sbt_frame(ctx, dst[i], coefs[i], q, C.byref(fmP), 7))
    nl = (lib.dsvcu_launch_count(ctx) - l0) // (RING * 4)
    add("inv_sbt (P picture, 3 planes per launch: k_sbt_inv)", t, 5 * Pb, nl)
    t = timed(lambda i: lib.dsvcu_inv_sbt_frame(ctx, dst[i], coefs[i], q, C.byref(fmI), 7))
    add("inv_sbt (I picture, 3 planes per launch)", t, 5 * Pb, nl)
    t = timed(lambda i: lib.dsvcu_sub_pred(ctx, C.byref(fmP), dst[i], dst[(i + 1) % RING], src[i]))
    add("predict + subtract (k_predict)", t, 4 * Pb, 1)
    t = timed(lambda i: lib.dsvcu_add_res(ctx, C.byref(fmP), q, dst[i], src[i], 0))
    add("reconstruct + filter traversal, do_filter=0 (k_reconstruct + k_filter_skew; random vectors: sharpening cells active)", t, 3 * Pb, 2)
    t_rec = t
    t = timed(lambda i: lib.dsvcu_add_res(ctx, C.byref(fmP), q, dst[i], src[i], 1))
    add("loop filters, extra cost of do_filter=1 (k_filter_skew; random vectors: every cell active, worst case)", max(t - t_rec, 1e-4), 2 * Pb, 1)
    t = timed(lambda i: lib.dsvcu_extend_frame(ctx, dst[i], 0))
    add("border extension (k_extend)", t, 2 * 64 * (W + H) * 3 // 2, 1)

    # motion search: one picture pair, true previous-picture input
    fs, fr = src[1], src[0]
    ps, pr = D.pyramid(fs), D.pyramid(fr)
    hp = P.DSVCU_HME_PARAMS(q, 0, cfg.pyr, 0)
    best = 1e9
    mvs_out = np.zeros(cfg.nblk, ops.MV_DTYPE)
    a3 = [C.c_int(), C.c_int(), C.c_int()]
    cnt = (C.c_longlong * 2)()
    l0 = lib.dsvcu_launch_count(ctx)
    for _ in range(3):
        lib.dsvcu_timer_start(ctx)
        lib.dsvcu_hme(ctx, C.byref(fmP), C.byref(hp), fs, ps, fr, pr, fr, pr)
        lib.dsvcu_timer_stop_ms(ctx, C.byref(ms))
        best = min(best, ms.value)
    me_launches = (lib.dsvcu_launch_count(ctx) - l0) // 3
    lib.dsvcu_hme_fetch(ctx, mvs_out.ctypes.data_as(C.c_void_p), cfg.nblk, C.byref(a3[0]), C.byref(a3[1]), C.byref(a3[2]))
    lib.dsvcu_hme_counters(ctx, cnt)
    # algorithmic bytes of the search: source, reconstructed and original reference luma + their
    # pyramids (1/3 extra) read once, the vector fields written once
    me_bytes = int(3 * W * H * 4 / 3) + cfg.nblk * 16 * 2
    gbs = me_bytes / (best * 1e-3) / 1e9
    me = {"kernel": "dsvcu_hme: k_me_prepass + k_me_level x 6 pyramid levels (speculative prepass, thin wavefront)",
          "ms_per_frame": round(best, 4), "algorithmic_bytes": me_bytes, "launches_per_frame": int(me_launches),
          "achieved_gbs": round(gbs, 2), "frac": round(gbs / peak_gbs, 6),
          "block_metric_evals_per_frame": int(cnt[0]), "subpel_position_metrics_per_frame": int(cnt[1]),
          "block_metric_evals_per_s": round((cnt[0] + cnt[1]) / (best * 1e-3), 0)}
    out.append(me)
    D.close()
    return out, me


def batched_rooflines(P, lib, peak_gbs, nstreams=16, ring=6):
    """The HBM-bound operator families with >= 64 independent pictures in flight
    (SURVEY 8d): nstreams contexts (one CUDA stream each, as the encoder instances
    run) x ring pictures each.  One 1080p picture sits in L2, so bandwidth is only
    meaningful over a working set like this one (nstreams x ring x 15.5 MB >> L2).
    Timed on the device: an event on every stream before and after, elapsed =
    latest end - earliest start."""
    import numpy as np
    import ops
    import torch
    cfg = ops.Cfg(W, H, P.SUBSAMP_420, isP=1, fnum=1)
    data = synth_chunks(2)
    Ds, src, dst, coefs = [], [], [], []
    rng = np.random.default_rng(1)
    mvs = np.zeros(cfg.nblk, ops.MV_DTYPE)
    mvs["x"] = rng.integers(-24, 25, cfg.nblk)
    mvs["y"] = rng.integers(-24, 25, cfg.nblk)
    for s_ in range(nstreams):
        D = ops.Dev(cfg, False)
        Ds.append(D)
        D.set_blockdata(np.zeros(cfg.nblk, np.uint8))
        D.set_mvs(mvs)
        src.append([D.frame(bytes(data[((s_ * ring + i) % 96) * FRAME_BYTES:((s_ * ring + i) % 96 + 1) * FRAME_BYTES]))
                    for i in range(ring)])
        dst.append([D.frame() for _ in range(ring)])
        coefs.append([D.coefs() for _ in range(ring)])
    fmP = cfg.fmeta()
    q = 252
    Pb = FRAME_BYTES
    npics = nstreams * ring
    streams = [torch.cuda.ExternalStream(lib.dsvcu_ctx_stream(D.ctx)) for D in Ds]
    out = []

    def run(name, fn, alg_bytes, reps=3):
        def once():
            for i in range(ring):
                for s_ in range(nstreams):
                    fn(Ds[s_], s_, i)
        once()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            ev0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
            ev1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
            for e, st in zip(ev0, streams):
                e.record(st)
            once()
            for e, st in zip(ev1, streams):
                e.record(st)
            torch.cuda.synchronize()
            # all start events are recorded on idle streams within microseconds of each other: the job's
            # device time is from the first start to the last end
            t = max(ev0[0].elapsed_time(e1) for e1 in ev1)
            best = min(best, t)
        gbs = alg_bytes * npics / (best * 1e-3) / 1e9
        out.append({"kernel": name, "pictures_in_flight": npics, "streams": nstreams, "ms_per_picture": round(best / npics, 5),
                    "algorithmic_bytes_per_picture": alg_bytes, "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak_gbs, 4)})

    run("fwd_sbt (k_sbt_fwd)", lambda D, s_, i: lib.dsvcu_fwd_sbt_frame(D.ctx, src[s_][i], coefs[s_][i], C.byref(fmP), 7), 5 * Pb)
    run("quantise + compaction (k_quant_*, k_compact_*)",
        lambda D, s_, i: lib.dsvcu_quant_frame(D.ctx, coefs[s_][i], q, C.byref(fmP), 7), 8 * Pb)
    run("inv_sbt (k_sbt_inv)", lambda D, s_, i: lib.dsvcu_inv_sbt_frame(D.ctx, dst[s_][i], coefs[s_][i], q, C.byref(fmP), 7), 5 * Pb)
    run("predict + subtract (k_predict)",
        lambda D, s_, i: lib.dsvcu_sub_pred(D.ctx, C.byref(fmP), dst[s_][i], dst[s_][(i + 1) % ring], src[s_][i]), 4 * Pb)
    run("reconstruct (k_reconstruct, filters off)",
        lambda D, s_, i: lib.dsvcu_add_res(D.ctx, C.byref(fmP), q, dst[s_][i], src[s_][i], 0), 3 * Pb)
    for D in Ds:
        D.close()
    return out


def cpu_baseline_sample():
    """single-core reference encoder + decoder on ONE WHOLE chunk of the workload
    (rank 0, N=1): about 10 s of CPU work"""
    if not os.path.exists(REF_BIN):
        return None
    y4m = ref_input()
    t0 = time.perf_counter()
    out = ref_encode_procs(y4m, [(0, GOP)], "cpu")[0]
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    subprocess.run([REF_BIN, "d", "-y", "-inp=" + out, "-out=/dev/shm/dsv2_bench_cpu_dec.yuv"],
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dd = time.perf_counter() - t0
    try:
        os.unlink("/dev/shm/dsv2_bench_cpu_dec.yuv")
    except OSError:
        pass
    return {"value": round(GOP / dt, 3), "unit": "frames/s", "cores": 1, "kind": "reference",
            "sample": "one whole %d-frame chunk of the workload (chunk 0), oracle/_ref/dsv2 e -qp=60 -gop=48 -nfr=48 "
                      "-noeos=1, 1 process; decode of the same stream with dsv2 d: %.2f frames/s" % (GOP, GOP / dd),
            "decode_value": round(GOP / dd, 3)}


def single_stream_configs(P, lib):
    """BASELINE configs 1-4 (frame-serial: one encoder / decoder instance, they cannot be sharded
    bit-exactly), bounded lengths: our single-instance frames/s through the session API on a
    persistent one-thread pool (contexts and device buffers exist, as in any long-running job), frames
    in pinned HOST memory on both sides, next to the single-core reference on the same clip, with
    byte parity.  The whole clip is one chunk, i.e. exactly one `dsv2 e` run without its final
    end-of-stream packet."""
    import torch
    import util
    import ops
    out = []
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    cases = [("1: CIF 352x288 4:2:0, -qp=60 -gop=48", 352, 288, 60, "420", 30, ["-qp=60", "-gop=48"], dict(qp=60, gop=48)),
             ("2: 1280x720 4:2:0 50 fps, -gop=250 -effort=10", 1280, 720, 50, "420", 50, ["-gop=250", "-effort=10"],
              dict(gop=250, effort=10)),
             ("3: 1920x1080 4:2:0 CRF, -qp=60 -gop=48", 1920, 1080, 48, "420", 30, ["-qp=60", "-gop=48"], dict(qp=60, gop=48)),
             ("4: 1920x1080 4:4:4 lossless, -qp=100", 1920, 1080, 8, "444", 30, ["-qp=100"], dict(qp=100))]
    devs = (C.c_int * 1)(torch.cuda.current_device())
    for name, w, h, n, fmt, fps, rargs, over in cases:
        pool = None
        try:
            y4m = util.clip("bench_cfg%s" % name[0], w, h, n, fmt, fps=fps)
            _, _, fr = util.read_y4m(y4m)
            yuv = b"".join(ops.yuv_bytes(f) for f in fr)
            fsz = len(yuv) // n
            host = torch.frombuffer(bytearray(yuv), dtype=torch.uint8).pin_memory()
            o = P.enc_opts(w, h, P.SUBSAMP_420 if fmt == "420" else P.SUBSAMP_444, (fps, 1), noeos=1, **over)
            pool = lib.dsv_pool_create(1, devs, 1)
            op, on = C.c_void_p(), C.c_size_t()

            def enc(k):
                if lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(host.data_ptr()), k, k, C.byref(op), C.byref(on)):
                    raise RuntimeError("encode failed")
                b = C.string_at(op, on.value)
                libc.free(op)
                return b
            enc(min(n, 3))  # contexts, device buffers, clocks
            t0 = time.perf_counter()
            dsv = enc(n)
            te = time.perf_counter() - t0
            dbuf = (C.c_uint8 * len(dsv)).from_buffer_copy(dsv)
            dst = torch.empty(n * fsz, dtype=torch.uint8).pin_memory()
            nf, meta = C.c_int(), P.DSV_META()

            def dec():
                if lib.dsv_pool_decode(pool, dbuf, len(dsv), C.c_void_p(dst.data_ptr()), n * fsz, C.byref(nf), C.byref(meta)):
                    raise RuntimeError("decode failed")
            dec()
            t0 = time.perf_counter()
            dec()
            td = time.perf_counter() - t0
            rout = "/dev/shm/dsv2_bench_cfg%s.dsv" % name[0]
            t0 = time.perf_counter()
            r = subprocess.run([REF_BIN, "e", "-y", "-inp=" + y4m, "-out=" + rout, "-y4m=1"] + rargs, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
            tre = time.perf_counter() - t0
            ryuv = "/dev/shm/dsv2_bench_cfg%s.yuv" % name[0]
            t0 = time.perf_counter()
            subprocess.run([REF_BIN, "d", "-y", "-inp=" + rout, "-out=" + ryuv], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            trd = time.perf_counter() - t0
            ref = open(rout, "rb").read()
            # the reference run reached the end of its input and appended the 14-byte end-of-stream packet
            ok_e = ref[:len(dsv)] == dsv and len(ref) - len(dsv) in (0, 14)
            ok_d = open(ryuv, "rb").read() == dst.numpy().tobytes()
            os.unlink(ryuv)
            out.append({"config": name, "frames": n, "encode_fps": round(n / te, 2), "decode_fps": round(nf.value / td, 2),
                        "reference_1core_encode_fps": round(n / tre, 2), "reference_1core_decode_fps": round(n / trd, 2),
                        "encode_speedup": round(tre / te, 1), "decode_speedup": round(trd / td, 1),
                        "parity": {"encode": ok_e, "decode": ok_d}})
        except Exception as e:
            out.append({"config": name, "error": str(e)})
        finally:
            if pool:
                lib.dsv_pool_destroy(pool)
    return out


def run_own(args):
    import numpy as np
    import torch
    import util
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = util.pkg()
    lib = P.load()
    cores = cpu_count()
    # instances spend most of their time blocked in event waits (measured: < 3 ms of host CPU per
    # picture), so many share a core; at most one per hardware queue (32)
    threads = args.threads or max(2, min(32, 8 * cores // max(1, world)))
    chunks = args.chunks or threads
    nframes = chunks * GOP
    # the decoder parses its coefficient planes on the device (dsv_session.h; the library's default)
    dev_entropy = 1
    lib.dsv_set_device_entropy_decode(dev_entropy)

    distinct = DISTINCT
    # rank 0 generates (and caches on tmpfs) the synthetic chunks, the others read the cache
    if rank == 0:
        data = synth_chunks(distinct)
    if dist is not None:
        dist.barrier()
    if rank != 0:
        data = synth_chunks(distinct)
    host = torch.empty(nframes * FRAME_BYTES, dtype=torch.uint8, pin_memory=True)
    hv = host.numpy()
    for c in range(chunks):
        k = (c + rank) % distinct
        hv[c * GOP * FRAME_BYTES:(c + 1) * GOP * FRAME_BYTES] = data[k * GOP * FRAME_BYTES:(k + 1) * GOP * FRAME_BYTES]
    dev = host.cuda()
    torch.cuda.synchronize()

    devs = (C.c_int * 1)(local)
    pool = lib.dsv_pool_create(threads, devs, 1)
    o = P.enc_opts(W, H, P.SUBSAMP_420, (FPS, 1), qp=QP, gop=GOP, noeos=1)
    out, outn = C.c_void_p(), C.c_size_t()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    def encode(ptr):
        r = lib.dsv_pool_encode(pool, C.byref(o), C.c_void_p(ptr), nframes, GOP, C.byref(out), C.byref(outn))
        if r:
            raise RuntimeError("encode failed: " + lib.dsvcu_last_error().decode())
        n = outn.value
        return n

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    stream_bytes = 0

    def enc_dev():
        nonlocal stream_bytes
        stream_bytes = encode(dev.data_ptr())
        libc.free(out)

    def enc_host():
        encode(host.data_ptr())
        libc.free(out)

    for _ in range(args.warmup):
        enc_dev()
    l0 = lib.dsvcu_total_launches()
    # rank 0 samples the clocks of every GPU of the job (NVML, 2 Hz)
    with (ClockSampler(range(world)) if rank == 0 else NoSampler()) as clk:
        dt = timed(enc_dev, args.steps)
    launches = lib.dsvcu_total_launches() - l0
    enc_host()
    dt_e2e = timed(enc_host, args.steps)

    # decoder on the stream of the last step
    encode(dev.data_ptr())
    dsv = (C.c_uint8 * outn.value).from_buffer_copy(C.string_at(out, outn.value))
    libc.free(out)
    ddev = torch.empty(nframes * FRAME_BYTES, dtype=torch.uint8, device="cuda")
    dhost = torch.empty(nframes * FRAME_BYTES, dtype=torch.uint8, pin_memory=True)
    nfr, meta = C.c_int(), P.DSV_META()

    def dec(ptr):
        r = lib.dsv_pool_decode(pool, dsv, len(dsv), C.c_void_p(ptr), nframes * FRAME_BYTES, C.byref(nfr), C.byref(meta))
        if r or nfr.value != nframes:
            raise RuntimeError("decode failed (%d frames)" % nfr.value)

    for _ in range(max(1, args.warmup)):
        dec(ddev.data_ptr())
    ddt = timed(lambda: dec(ddev.data_ptr()), args.steps)
    dec(dhost.data_ptr())
    ddt_e2e = timed(lambda: dec(dhost.data_ptr()), args.steps)
    chk = int(dhost[:FRAME_BYTES].to(torch.int64).sum().item())
    lib.dsv_pool_destroy(pool)

    # ---- parity of exactly what was timed (outside every timed region): the stream of
    # the last encode step and the frames of the last decode step against the unmodified
    # reference on the same chunks.  Every rank checks its own output.
    parity = None
    if os.path.exists(REF_BIN) and not args.no_parity:
        ref_dsv, ref_yuv = (None, None)
        if rank == 0:
            ref_dsv, ref_yuv = reference_chunks(True)
        if dist is not None:
            dist.barrier()  # rank 0 has left the reference outputs on tmpfs
            if rank != 0:
                ref_dsv, ref_yuv = reference_chunks(False)
        order = [(c + rank) % DISTINCT for c in range(chunks)]
        want_dsv = b"".join(ref_dsv[k] for k in order)
        ok_e = bytes(dsv) == want_dsv
        got = dhost.numpy()
        ok_d = True
        for c, k in enumerate(order):
            a = got[c * GOP * FRAME_BYTES:(c + 1) * GOP * FRAME_BYTES]
            if a.tobytes() != ref_yuv[k]:
                ok_d = False
                break
        dd = ddev.cpu().numpy()
        ok_d = ok_d and bool((dd == got).all())
        parity = {"encode": ok_e, "decode": ok_d, "chunks": DISTINCT, "chunks_checked": chunks,
                  "reference": "oracle/_ref/dsv2 e -qp=60 -gop=48 -sfr=48k -nfr=48 -noeos=1 per distinct chunk, "
                               "dsv2 d of each; compared: every byte of the last timed step's .dsv and every decoded frame",
                  "dsv_md5": md5(bytes(dsv)), "dsv_bytes": len(dsv)}
        if dist is not None:
            t = torch.tensor([int(ok_e), int(ok_d)], device="cuda", dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            parity["encode"], parity["decode"] = bool(t[0].item()), bool(t[1].item())
            parity["ranks"] = world
        if not (parity["encode"] and parity["decode"]):
            log("PARITY FAILURE: encode %s decode %s" % (parity["encode"], parity["decode"]))

    total_frames = nframes * world * args.steps
    value = total_frames / dt
    e2e = total_frames / dt_e2e
    dvalue = total_frames / ddt
    de2e = total_frames / ddt_e2e
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks = {}
    src_peak = "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        src_peak = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern, me, batched, single = [], None, [], []
    cpu = None
    if world == 1 and not args.no_micro:
        try:
            kern, me = kernel_rooflines(P, lib, peak)
        except Exception as e:  # the headline numbers stand on their own
            log("kernel microbench failed:", e)
        try:
            batched = batched_rooflines(P, lib, peak)
        except Exception as e:
            log("batched microbench failed:", e)
        cpu = cpu_baseline_sample()
        try:
            single = single_stream_configs(P, lib) if os.path.exists(REF_BIN) else []
        except Exception as e:
            log("single-stream configs failed:", e)
            single = []
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1000 * dt / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chunks_per_gpu_per_step": chunks,
                   "frames_per_step": nframes * world, "host_threads_per_gpu": threads, "host_cores": cores,
                   "l2": "inputs larger than L2: %.0f MB of source frames per step" % (nframes * FRAME_BYTES / 1e6),
                   "timer": "host clock around a device-synchronised, barrier-bracketed region (work spans %d CUDA "
                            "streams; per-kernel times below use CUDA events on the launching stream)" % threads,
                   "stream_bytes_per_frame": stream_bytes // max(1, nframes)},
        "e2e": {"value": round(e2e, 3), "unit": "frames/s", "h2d_bytes_per_step": nframes * FRAME_BYTES,
                "d2h_bytes_per_step": int(stream_bytes)},
        "decode": {"value": round(dvalue, 3), "e2e": round(de2e, 3), "unit": "frames/s",
                   "d2h_bytes_per_step_e2e": nframes * FRAME_BYTES, "first_frame_checksum": chk,
                   "entropy_decode": "device (k_hzcc_parse; intra pictures on the host)" if dev_entropy else "host threads"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "parity": parity,
    }
    if me is not None:
        ncu = {}
        try:  # per-launch counters of the SAME kernels from the committed ncu --set full captures
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")))
        except Exception:
            pass
        l0n = ncu.get("me_level_L0", {})
        prn = ncu.get("me_prepass_L0", {})
        spn = ncu.get("me_subpel", {})
        inst = (l0n.get("warp_instructions") or 0) + (prn.get("warp_instructions") or 0) + (spn.get("warp_instructions") or 0)
        sm_mhz = (clk.summary() or {}).get("sm_mhz") or 1965
        issue_peak = 148 * 4 * sm_mhz * 1e6  # warp instructions per second the GPU can issue
        line["roofline"] = {
            "bound": "issue", "kernel": me["kernel"],
            "achieved": round(me["block_metric_evals_per_s"] / 1e6, 2), "unit": "M block-metric evaluations/s (one instance)",
            "peak": None, "frac": None,
            "issue_slot_pct": round(100.0 * inst / max(me["ms_per_frame"] * 1e-3 * issue_peak, 1e-9), 2) if inst else None,
            "issue_slot_note": "warp instructions of the level-0 prepass + sub-pel + wavefront launches (ncu, profiles/r2_ncu_summary.json) / "
                               "(live CUDA-event time of dsvcu_hme x 148 SMs x 4 schedulers x sampled SM clock)",
            "ms_per_picture_live": me["ms_per_frame"],
            "hbm": {"achieved": me["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": me["frac"],
                    "traffic": int(sum(k.get("dram_bytes_read", 0) + k.get("dram_bytes_write", 0) for k in (l0n, prn, spn))) or None,
                    "peak_source": src_peak},
            "note": "dominant kernel family by device time is the motion search: the prepass is issue / latency bound "
                    "(no HBM pressure: one picture and its pyramids sit in L2), the wavefront is bound by the dependency "
                    "chain between blocks.  The HBM-bound operator families are under 'kernels' (one picture per launch, "
                    "one stream: launch-latency bound) and 'kernels_batched' (>= 64 pictures in flight: bandwidth)"}
        line["kernels"] = kern
        if batched:
            tmap = {"fwd_sbt": "sbt_fwd_L1", "inv_sbt": "sbt_inv_L1", "predict": "predict", "reconstruct": "reconstruct",
                    "quantise": "quant_hf_L2"}
            for k in batched:
                for key, nm in tmap.items():
                    if k["kernel"].startswith(key) and nm in ncu:
                        k["traffic_dominant_launch"] = int(ncu[nm].get("dram_bytes_read", 0) + ncu[nm].get("dram_bytes_write", 0))
                        k["traffic_note"] = "dram bytes of the largest launch of the family (ncu, profiles/r2_ncu_%s.txt)" % nm
            line["kernels_batched"] = batched
            best_b = max(batched, key=lambda k: k["frac"])
            line["roofline"]["hbm_family_best"] = {"kernel": best_b["kernel"], "achieved": best_b["achieved_gbs"], "peak": peak,
                                                    "unit": "GB/s", "frac": best_b["frac"]}
    if single:
        line["single_stream_configs"] = single
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not (parity["encode"] and parity["decode"]):
        return 3  # a fast encoder whose bytes differ from the reference's is not a result
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--threads", type=int, default=0, help="host threads (CUDA streams) per GPU")
    ap.add_argument("--chunks", type=int, default=0, help="48-frame chunks per GPU and step")
    ap.add_argument("--no-micro", action="store_true", help="skip the per-kernel microbench / CPU sample")
    ap.add_argument("--no-parity", action="store_true", help="skip the reference comparison of the timed output")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
