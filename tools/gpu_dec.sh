#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
TAG=${1:-dec}
timeout 900 python -m pytest tests/test_decode.py tests/test_robustness.py tests/test_configs.py tests/test_pool.py tests/test_cli.py tests/test_decops.py -x -q -m gpu > $O/gputest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $O/gputest_$TAG.log
timeout 300 python tools/scale_probe_dec.py 1,16,32 2>&1 | grep -E "threads" | cut -c1-200
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err
echo "bench rc=$?"; tail -3 $O/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value",d.get("value"),"e2e",d.get("e2e",{}).get("value"),"decode",d.get("decode"),"parity",{k:d["parity"][k] for k in ("encode","decode")} if d.get("parity") else None)
print("roofline", json.dumps(d.get("roofline"))[:900])
for k in d.get("kernels_batched",[]): print("  B ",k["kernel"][:50],k["ms_per_picture"],k["achieved_gbs"],k["frac"])
PY
