#!/usr/bin/env python
"""Per-source-function shares (executed warp instructions, stall samples) of one
kernel in an .ncu-rep, by mapping each SASS pc's source line to the enclosing
function of csrc/*.cuh.  usage: ncu_funcs.py report.ncu-rep [top N]"""
import collections, re, sys, os
sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "digital-subband-video-2_b200", "csrc")
def func_map(path):
    """line -> function name (very small C parser: a definition starts at a line whose
    previous non-blank lines hold DSVCU_DEV/DSVCU_HD/DSVCU_KERNEL/static and that has NAME( at column 0)"""
    m, cur = {}, None
    lines = open(path).read().split("\n")
    for i, l in enumerate(lines, 1):
        g = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if g and i >= 2 and re.search(r"DSVCU_(DEV|HD|KERNEL)|static|template|__global__", "\n".join(lines[max(0, i - 4):i - 1])):
            cur = g.group(1)
        g2 = re.match(r"^DSVCU_(?:DEV|HD)\s+\S+\s+([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if g2:
            cur = g2.group(1)
        m[i] = cur
    return m
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
act = ncu_report.load_report(rep).range_by_idx(0).action_by_idx(0)
m = act.metric_by_name("inst_executed"); cid = m.correlation_ids()
samp = collections.Counter()
for sn in act.metric_names():
    if sn.startswith("smsp__pcsamp_warps_issue_stalled") and not sn.endswith("not_issued"):
        sm = act.metric_by_name(sn); c = sm.correlation_ids()
        for i in range(sm.num_instances()):
            samp[c.as_uint64(i)] += sm.as_uint64(i)
maps = {}
fi, fs = collections.Counter(), collections.Counter()
li, ls = collections.Counter(), collections.Counter()
for i in range(m.num_instances()):
    pc = cid.as_uint64(i); si = act.source_info(pc)
    if si:
        f = si.file_name().split("/")[-1]; ln = si.line()
        if f not in maps:
            p = os.path.join(ROOT, f)
            maps[f] = func_map(p) if os.path.exists(p) else {}
        fn = maps[f].get(ln) or f
    else:
        f, ln, fn = "?", 0, "?"
    fi[fn] += m.as_uint64(i); fs[fn] += samp.get(pc, 0)
    li[(f, ln)] += m.as_uint64(i); ls[(f, ln)] += samp.get(pc, 0)
ti, ts = sum(fi.values()), sum(fs.values())
dur = act.metric_by_name("gpu__time_duration.sum").as_double() / 1e6
print("%s: %.3f ms, %d warp instructions, %d samples" % (act.name(), dur, ti, ts))
for fn, v in fi.most_common(top):
    print("%-28s inst %5.1f%%  samp %5.1f%%" % (fn, 100 * v / ti, 100 * fs[fn] / max(ts, 1)))
print("--- top lines by samples")
src = {}
for key, v in ls.most_common(25):
    f, ln = key
    p = os.path.join(ROOT, f)
    if f not in src:
        src[f] = open(p).read().split("\n") if os.path.exists(p) else None
    txt = src[f][ln - 1].strip()[:80] if src[f] and 0 < ln <= len(src[f]) else ""
    print("%-14s %5d samp %5.1f%% inst %5.1f%%  %s" % (f, ln, 100 * v / max(ts, 1), 100 * li[key] / ti, txt))
