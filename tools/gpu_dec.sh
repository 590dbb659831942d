#!/bin/bash
# decoder iteration: decoder-side parity tests, decode throughput vs instances with the
# host-thread phase profile (also on 4 cores = the host share of one GPU on an 8-GPU box),
# entropy decode on the host vs on the device, launch durations of the entropy-decode kernel
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
TAG=${1:-d}
TESTS=${2:-tests/test_decops.py tests/test_decode.py tests/test_golden.py tests/test_robustness.py tests/test_pool.py tests/test_configs.py}
timeout 900 python -m pytest $TESTS -x -q -m gpu > $O/gputest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 $O/gputest_$TAG.log
filt() { grep -v "dsv_enc profile" | awk '/dsv_dec profile/{n++; if(n%32==1)print; next}{print}' | grep -E "threads|dsv_dec profile" | tail -${1:-12}; }
DSV_PROFILE=1 timeout 600 python tools/scale_probe_dec.py 1,8,32 0,1 > $O/probe_$TAG.log 2>&1; filt 20 < $O/probe_$TAG.log
echo "--- 4 cores"
DSV_PROFILE=1 timeout 600 taskset -c 0-3 python tools/scale_probe_dec.py 32 0,1 > $O/probe4_$TAG.log 2>&1; filt 8 < $O/probe4_$TAG.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_hzcc_parse -c 2 python tools/scale_probe_dec.py 1 1 2>&1 | grep -E "k_hzcc|gpu__time|inst_executed" | head -12
