#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export CUDA_DEVICE_MAX_CONNECTIONS=32
touch digital-subband-video-2_b200/csrc/dsvcu_api.cu
make -s -j4 all EXTRA="-DDSVCU_DIAG" 2>&1 | grep -E "error" | head
for v in "$@"; do
  printf "%-44s " "$v"
  env $v GOPN=12 timeout 120 python tools/scale_probe.py 32 2>&1 | grep -E "^threads" | cut -c1-60
done
