#!/bin/bash
# the bench line on N GPUs of one box, launched the way the driver launches it
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out
nproc; nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.log
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step","e2e","decode","parity","clocks")})
print(d["config"])
PY
