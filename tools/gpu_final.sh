#!/bin/bash
# decode probe + the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
TAG=${1:-f}
bash tools/gpu_dec.sh $TAG "tests/test_decops.py tests/test_decode.py tests/test_robustness.py tests/test_pool.py"
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.log
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","decode","parity","gpu_launches","clocks")})
for s in d.get("single_stream_configs",[]): print(s)
PY
