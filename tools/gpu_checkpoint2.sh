#!/bin/bash
# round-2 checkpoint: GPU tests (all, or the files given), the default bench line, the launch list
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
TAG=${1:-ck}
TESTS=${2:-tests}
timeout 1500 python -m pytest $TESTS -x -q -m gpu > $O/gputest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $O/gputest_$TAG.log
timeout 900 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.log
echo "bench rc=$?"; tail -c 600 $O/bench_$TAG.json | head -c 300; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_$TAG.csv python tools/prof_run.py 1920 1080 6 both > $O/launches_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l $O/launches_$TAG.csv
