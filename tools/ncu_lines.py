#!/usr/bin/env python
"""Per-source-line sample / instruction shares of one kernel in an .ncu-rep
(needs -lineinfo at compile time and --import-source on at capture time).
usage: ncu_lines.py report.ncu-rep [source file to print lines from] [top N]"""
import collections
import sys

sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report  # noqa: E402

rep = sys.argv[1]
srcf = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
act = ncu_report.load_report(rep).range_by_idx(0).action_by_idx(0)
m = act.metric_by_name("inst_executed")
cid = m.correlation_ids()
samp = collections.Counter()
for sn in act.metric_names():
    if sn.startswith("smsp__pcsamp_warps_issue_stalled") and not sn.endswith("not_issued"):
        sm = act.metric_by_name(sn)
        c = sm.correlation_ids()
        for i in range(sm.num_instances()):
            samp[c.as_uint64(i)] += sm.as_uint64(i)
li, ls = collections.Counter(), collections.Counter()
for i in range(m.num_instances()):
    pc = cid.as_uint64(i)
    si = act.source_info(pc)
    key = (si.file_name().split("/")[-1], si.line()) if si else ("?", 0)
    li[key] += m.as_uint64(i)
    ls[key] += samp.get(pc, 0)
ti, ts = sum(li.values()), sum(ls.values())
dur = act.metric_by_name("gpu__time_duration.sum").as_double() / 1e6
print("%s: %.3f ms, %d warp instructions, %d samples" % (act.name(), dur, ti, ts))
lines = open(srcf).read().split("\n") if srcf else None
for key, v in ls.most_common(top):
    txt = ""
    if lines and key[0] == srcf.split("/")[-1].replace("_r1d", "").replace("k_hme_r1d.cuh", "k_hme.cuh") or (lines and key[0].endswith(".cuh")):
        try:
            txt = lines[key[1] - 1].strip()[:90]
        except Exception:
            pass
    print("%-22s %5d  samp %5.1f%%  inst %5.1f%%  %s" % (key[0], key[1], 100 * v / ts, 100 * li[key] / ti, txt))
