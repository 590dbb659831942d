#!/bin/bash
# entropy-decode kernel: duration + instruction count on the bench stream, one full ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
O=gpurun_out
TAG=${1:-hz}
timeout 300 python -m pytest tests/test_decops.py tests/test_decode.py -x -q -m gpu 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_hzcc_parse -c 1 python tools/scale_probe_dec.py 1 2>&1 | grep -E "k_hzcc|gpu__time|inst_executed|threads" | head -12
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_hzcc_parse -c 1 -f -o $O/hzcc_$TAG python tools/scale_probe_dec.py 1 > $O/hzcc_$TAG.log 2>&1
echo "capture rc=$?"; ls -la $O/hzcc_$TAG.ncu-rep
