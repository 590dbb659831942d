#!/bin/bash
# full GPU checkpoint: test-suite, smoke, bench (own + reference arm), launch list
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
TAG=${1:-ck}
timeout 1500 python -m pytest tests -x -q -m gpu > $O/gputest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $O/gputest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
python - <<PY
import json
for f in ("gpurun_out/bench_$TAG.json","gpurun_out/bench_ref_$TAG.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d.get("value"), "e2e", d.get("e2e",{}).get("value"), "decode", d.get("decode"), "parity", d.get("parity") and {k:d["parity"][k] for k in ("encode","decode")})
        for k in d.get("kernels",[]): print("   ", k["kernel"][:70], k["ms_per_frame"], k["achieved_gbs"], k["frac"])
        print("   roofline", d.get("roofline")); print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable", e)
PY
