#!/bin/bash
# round-2 checkpoint c: speculative prepass -- tests, bench, ME captures, scaling probe
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
TAG=${1:-r2c}
timeout 1500 python -m pytest tests -x -q -m gpu > $O/gputest_$TAG.log 2>&1
echo "pytest rc=$?"
tail -4 $O/gputest_$TAG.log
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value",d.get("value"),"e2e",d.get("e2e",{}).get("value"),"decode",d.get("decode"),"parity",{k:d["parity"][k] for k in ("encode","decode")} if d.get("parity") else None)
for k in d.get("kernels",[]): print("  ",k["kernel"][:60],k["ms_per_frame"],k["achieved_gbs"])
PY
DSV_PROFILE=2 timeout 300 python tools/scale_probe.py 1,48 2>&1 | grep -v "^$" | tail -12
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_me_level -s 5 -c 1 -f -o $O/me_l0_$TAG python tools/prof_run.py 1920 1080 3 enc > $O/ncu_me_$TAG.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_me_prepass -s 5 -c 1 -f -o $O/me_pre_$TAG python tools/prof_run.py 1920 1080 3 enc > $O/ncu_pre_$TAG.log 2>&1
