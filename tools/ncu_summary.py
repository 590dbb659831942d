#!/usr/bin/env python
"""Condenses `ncu --set full` captures (gpurun_out/<tag>_<name>.ncu-rep, one launch each) into
profiles/<round>_ncu_summary.json (numbers bench.py and DESIGN.md quote) and one text file per
kernel: key counters + the source lines with the largest share of warp-stall samples.
usage: ncu_summary.py TAG ROUND [INPUT_DIR [OUTPUT_DIR]]   e.g.  ncu_summary.py r2f r2 /tmp/prof gpurun_out/prof"""
import collections, glob, json, os, subprocess, sys
sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.environ.get("NCU_SRC") or os.path.join(ROOT, "digital-subband-video-2_b200", "csrc")
tag, rnd = sys.argv[1], sys.argv[2]
IN = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out")
OUT = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
KEYS = {
    "gpu__time_duration.sum": ("duration_ms", 1e-6),
    "dram__bytes_read.sum": ("dram_bytes_read", 1), "dram__bytes_write.sum": ("dram_bytes_write", 1),
    "lts__t_bytes.sum": ("l2_bytes", 1),
    "smsp__inst_executed.sum": ("warp_instructions", 1),
    "smsp__issue_active.avg.pct_of_peak_sustained_active": ("issue_active_pct", 1),
    "sm__warps_active.avg.per_cycle_active": ("warps_active_per_sm", 1),
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": ("sm_throughput_pct", 1),
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": ("dram_throughput_pct", 1),
    "launch__registers_per_thread": ("registers_per_thread", 1), "launch__grid_size": ("grid", 1),
    "launch__block_size": ("block", 1), "sm__icc_request_hit_rate.pct": ("icache_hit_pct", 1),
    "l1tex__t_sector_hit_rate.pct": ("l1_hit_pct", 1), "lts__t_sector_hit_rate.pct": ("l2_hit_pct", 1),
    "launch__occupancy_limit_registers": ("occ_limit_regs_ctas", 1), "launch__occupancy_limit_shared_mem": ("occ_limit_smem_ctas", 1),
    "sm__warps_active.avg.pct_of_peak_sustained_active": ("achieved_occupancy_pct", 1),
}
summary = {}
for rep in sorted(glob.glob(os.path.join(IN, tag + "_*.ncu-rep"))):
    name = os.path.basename(rep)[len(tag) + 1:-8]
    try:
        act = ncu_report.load_report(rep).range_by_idx(0).action_by_idx(0)
    except Exception as e:
        print("skip", rep, e); continue
    d = {"kernel": act.name(), "source": "ncu --set full --clock-control none, one launch, 1080p 4:2:0 (%s)" % os.path.basename(rep)}
    for k, (nm, sc) in KEYS.items():
        try:
            d[nm] = act.metric_by_name(k).as_double() * sc
        except Exception:
            pass
    # stall reasons (share of sampled warp states)
    st = collections.Counter()
    for sn in act.metric_names():
        if sn.startswith("smsp__pcsamp_warps_issue_stalled_") and not sn.endswith("not_issued"):
            m = act.metric_by_name(sn)
            st[sn.replace("smsp__pcsamp_warps_issue_stalled_", "")] = sum(m.as_uint64(i) for i in range(m.num_instances()))
    tot = sum(st.values()) or 1
    d["stall_pct"] = {k: round(100 * v / tot, 1) for k, v in st.most_common(6)}
    summary[name] = d
    # per-line shares
    lines = []
    try:
        m = act.metric_by_name("inst_executed"); cid = m.correlation_ids()
        samp = collections.Counter()
        for sn in act.metric_names():
            if sn.startswith("smsp__pcsamp_warps_issue_stalled") and not sn.endswith("not_issued"):
                sm = act.metric_by_name(sn); c = sm.correlation_ids()
                for i in range(sm.num_instances()): samp[c.as_uint64(i)] += sm.as_uint64(i)
        li, ls = collections.Counter(), collections.Counter()
        for i in range(m.num_instances()):
            pc = cid.as_uint64(i); si = act.source_info(pc)
            key = (si.file_name().split("/")[-1], si.line()) if si else ("?", 0)
            li[key] += m.as_uint64(i); ls[key] += samp.get(pc, 0)
        ti, ts = sum(li.values()) or 1, sum(ls.values()) or 1
        src = {}
        for key, v in ls.most_common(14):
            f, ln = key
            if f not in src:
                p = os.path.join(CSRC, f); src[f] = open(p).read().split("\n") if os.path.exists(p) else None
            txt = src[f][ln - 1].strip()[:96] if src[f] and 0 < ln <= len(src[f]) else ""
            lines.append("%-16s %5d  samples %5.1f%%  instructions %5.1f%%  %s" % (f, ln, 100 * v / ts, 100 * li[key] / ti, txt))
    except Exception as e:
        lines.append("(no source correlation: %s)" % e)
    with open(os.path.join(OUT, "%s_ncu_%s.txt" % (rnd, name)), "w") as f:
        f.write("%s  -- %s\n" % (d["kernel"], d["source"]))
        for k in sorted(d):
            if k not in ("kernel", "source", "stall_pct"): f.write("  %-26s %s\n" % (k, ("%.4f" % d[k]) if isinstance(d[k], float) else d[k]))
        f.write("  warp-stall samples: %s\n" % ", ".join("%s %.1f%%" % kv for kv in d["stall_pct"].items()))
        f.write("top source lines by warp-stall samples:\n")
        for l in lines: f.write("  " + l + "\n")
        # SASS evidence of wide memory access
        try:
            sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, timeout=120).stdout
            ops = collections.Counter()
            for l in sass.split("\n"):
                for op in ("LDG.E.128", "STG.E.128", "LDG.E.64", "STG.E.64", "LDG.E.U8", "STG.E.U8", "LDG.E ", "STG.E ", "LDS.128", "REDUX", "VABSDIFF4", "IDP.4A"):
                    if op in l: ops[op.strip()] += 1
            f.write("SASS memory / SIMD mnemonics (static count): %s\n" % ", ".join("%s x%d" % kv for kv in sorted(ops.items())))
        except Exception:
            pass
    print("%-18s %-16s %8.3f ms  inst %6.1fM  issue %5.1f%%  dram %6.2f MB  regs %3d  grid %5d x %4d" % (
        name, d["kernel"], d.get("duration_ms", 0), d.get("warp_instructions", 0) / 1e6, d.get("issue_active_pct", 0),
        (d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)) / 1e6, int(d.get("registers_per_thread", 0)),
        int(d.get("grid", 0)), int(d.get("block", 0))))
json.dump(summary, open(os.path.join(OUT, "%s_ncu_summary.json" % rnd), "w"), indent=1, sort_keys=True)
