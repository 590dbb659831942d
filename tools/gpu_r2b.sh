#!/bin/bash
# round-2 checkpoint b: GPU test-suite + bench with parity + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/gputest_r2b.log 2>&1
echo "pytest rc=$?"
tail -5 $O/gputest_r2b.log
timeout 900 python bench.py > $O/bench_r2b.json 2> $O/bench_r2b.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_r2b.json 2> $O/bench_ref_r2b.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_r2b.json","gpurun_out/bench_ref_r2b.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("e2e"), d.get("decode"), d.get("parity"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 $O/bench_r2b.err
