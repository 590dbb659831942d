#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
TAG=${1:-ll}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_$TAG.csv python tools/prof_run.py 1920 1080 4 both > $O/launches_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_$TAG.csv", errors="ignore")))
hdr = None; data = []
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
tot = collections.Counter(); cnt = collections.Counter()
for d in data:
    if d.get("Metric Name") == "gpu__time_duration.sum":
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v = v / 1e3 if u.startswith("ns") or u == "nsecond" else (v if u.startswith("us") else v * 1e3)
        n = d["Kernel Name"].split("(")[0]
        tot[n] += v; cnt[n] += 1
T = sum(tot.values())
print("total %.1f us over %d launches (4 frames encode: 1 I + 3 P, then decode)" % (T, sum(cnt.values())))
for n, v in tot.most_common(): print("%-22s %5d launches %9.1f us %5.1f%%  avg %7.1f us" % (n, cnt[n], v, 100 * v / T, v / cnt[n]))
PY
